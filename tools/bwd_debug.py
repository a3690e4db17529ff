import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
from oracle import head as ohead
dev = "cuda:0"
O, C, H, W, N = 19, 32, 8, 16, 1
P, A = synth.head_params(O, C, seed=13, dtype=torch.float64)
u = torch.stack([synth.image_features(i, C, H, W, sigma=0.1) for i in range(N)])
g = torch.Generator().manual_seed(3)
dl = torch.randn((N, O, H, W), generator=g) * 1e-3
du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
args = (u.to(dev), P.to(dev), A.to(dev), 1.0, dl.to(dev))
os.environ["HALO_BWD_CUDA_CORE"] = "1"
du_cc, dP_cc, dA_cc = halo_b200.head_backward(*args)
os.environ["HALO_BWD_CUDA_CORE"] = "0"
def rel(a, b): return ((a.double().cpu() - b.double()).abs().max() / b.double().abs().max()).item()
for desc in [""]:
    if desc: os.environ["HALO_BWD_DESC"] = desc
    du, dP, dA = halo_b200.head_backward(*args)
    torch.cuda.synchronize()
    print("desc", desc or "default", "du", "%.3e" % rel(du, du_ref), "dP", "%.3e" % rel(dP, dP_ref), "dA", "%.3e" % rel(dA, dA_ref),
          "| du - du_cc", "%.3e" % rel(du, du_cc.cpu()))
# structure of the error for the default descriptor
os.environ.pop("HALO_BWD_DESC", None)
du, _, _ = halo_b200.head_backward(*args)
d = (du.cpu().double() - du_ref.double())[0]          # (C,H,W)
print("err per channel (max abs):", [float("%.2e" % v) for v in d.abs().amax(dim=(1, 2))][:16])
print("ref per channel (max abs):", [float("%.2e" % v) for v in du_ref[0].abs().amax(dim=(1, 2))][:16])
# hypothesis test: is the tensor-core D2 contribution simply missing?  G planes live in the workspace after the pack.
from halo_b200 import _native as nat
ws = nat.workspace.get(torch.device(dev), "head_bwd", 0)
OP = (O + 3) // 4 * 4; KP = 2 * OP; CPAD = (C + 3) // 4 * 4; HWp = H * W
off = ((CPAD * KP + 4 * OP) * 4 + 255) // 256 * 256
G = ws[off: off + N * KP * HWp * 4].view(torch.float32).view(N, KP, HWp).cpu().double()
Pf, Af = P.float().double(), A.float().double()
ahat = Af / Af.norm(dim=1, keepdim=True)
Wrows = torch.zeros(KP, C, dtype=torch.float64); Wrows[:O] = -Pf; Wrows[OP:OP + O] = ahat
D2_ref = torch.einsum("nkp,kc->ncp", G, Wrows).view(N, C, H, W)
miss = (du_ref.double() - du.cpu().double())
print("|D2_ref| max %.3e ; |du_ref - du_tc - D2_ref| max %.3e ; |du_ref - du_tc| max %.3e" % (D2_ref.abs().max(), (miss - D2_ref).abs().max(), miss.abs().max()))
d2_tc = du.cpu().double() - (du_cc.cpu().double() - D2_ref)      # what the tensor core actually contributed
print("tc contribution max %.3e ; ratio to D2_ref (first pixels, ch0):" % d2_tc.abs().max(), (d2_tc[0, 0].flatten()[:6] / D2_ref[0, 0].flatten()[:6]).tolist())
print("tc contribution ch0..5 px0:", d2_tc[0, :6, 0, 0].tolist()); print("D2_ref      ch0..5 px0:", D2_ref[0, :6, 0, 0].tolist())
