import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
from oracle import head as ohead
dev = "cuda:0"
def rel(a, b): return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
for (O, C, H, W, N, sigma) in [(16, 128, 20, 20, 2, 0.3), (19, 256, 16, 24, 1, 0.3), (19, 256, 16, 24, 1, 1.0), (19, 256, 16, 24, 1, 0.1), (5, 32, 9, 12, 3, 1.0), (24, 96, 16, 16, 1, 0.2)]:
    P, A = synth.head_params(O, C, seed=13, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    g = torch.Generator().manual_seed(3)
    dl = torch.randn((N, O, H, W), generator=g) * 1e-3
    du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
    args = (u.to(dev), P.to(dev), A.to(dev), 1.0, dl.to(dev))
    os.environ["HALO_BWD_CUDA_CORE"] = "0"
    du, dP, dA = halo_b200.head_backward(*args)
    os.environ["HALO_BWD_CUDA_CORE"] = "1"
    du2, dP2, dA2 = halo_b200.head_backward(*args)
    print((O, C, H, W, N, sigma), "TC du %.2e dP %.2e dA %.2e | CC du %.2e dP %.2e dA %.2e" % (rel(du, du_ref), rel(dP, dP_ref), rel(dA, dA_ref), rel(du2, du_ref), rel(dP2, dP_ref), rel(dA2, dA_ref)))
    e = (dP.cpu().double() - dP_ref).abs().amax(dim=1) / dP_ref.abs().max()
    print("   dP err per class:", [float("%.1e" % v) for v in e][:10])
