import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
from oracle import head as ohead
dev = "cuda:0"
def rel(a, b): return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
shapes = [(16, 128, 20, 20, 2, 0.3), (16, 128, 20, 20, 2, 0.1), (16, 128, 16, 24, 1, 0.3), (16, 256, 20, 20, 2, 0.3),
          (19, 128, 20, 20, 2, 0.3), (12, 128, 20, 20, 2, 0.3), (16, 64, 20, 20, 2, 0.3), (16, 128, 32, 32, 1, 0.3),
          (16, 128, 64, 64, 3, 0.3)]
for (O, C, H, W, N, sigma) in shapes:
    P, A = synth.head_params(O, C, seed=13, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    g = torch.Generator().manual_seed(3)
    dl = torch.randn((N, O, H, W), generator=g) * 1e-3
    du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
    args = (u.to(dev), P.to(dev), A.to(dev), 1.0, dl.to(dev))
    os.environ["HALO_BWD_CUDA_CORE"] = "0"
    du, dP, dA = halo_b200.head_backward(*args)
    du_b, _, _ = halo_b200.head_backward(*args)
    os.environ["HALO_BWD_CUDA_CORE"] = "1"
    du2, dP2, dA2 = halo_b200.head_backward(*args)
    print((O, C, H, W, N, sigma), "TC du %.2e dP %.2e dA %.2e | CC du %.2e dP %.2e dA %.2e | rerun-equal %s" % (
        rel(du, du_ref), rel(dP, dP_ref), rel(dA, dA_ref), rel(du2, du_ref), rel(dP2, dP_ref), rel(dA2, dA_ref),
        bool((du == du_b).all())))
    d = (du.cpu().double() - du_ref.double()).abs() / du_ref.abs().max()      # (N,C,H,W)
    if d.max() > 2e-5:
        per_px = d.amax(dim=1).flatten(1)            # (N, HW)
        bad = (per_px > 2e-5)
        print("   bad pixels per image:", bad.sum(dim=1).tolist(), "of", H * W)
        for n in range(N):
            idx = bad[n].nonzero().flatten().tolist()
            print("   img", n, "bad px idx (first 40):", idx[:40])
        per_ch = d.amax(dim=(0, 2, 3))
        print("   err per channel (first 32):", [float("%.1e" % v) for v in per_ch[:32]])
        nrm = u.double().pow(2).sum(1).sqrt().flatten(1)
        n0 = bad.nonzero()[0]
        print("   |u| at first bad px: %.3f ; median |u| %.3f ; max |u| %.3f" % (nrm[n0[0], n0[1]], nrm.median(), nrm.max()))
        # is the error at a bad pixel proportional to u (alpha wrong) or not (D2 wrong)?
        n_, p_ = int(n0[0]), int(n0[1])
        e = (du.cpu().double() - du_ref.double()).flatten(2)[n_, :, p_]
        uu = u.double().flatten(2)[n_, :, p_]
        coef = (e @ uu) / (uu @ uu)
        print("   at that px: |err| %.3e, |err - coef*u| %.3e (coef %.3e)  -> alpha error if small" % (e.norm(), (e - coef * uu).norm(), coef))
