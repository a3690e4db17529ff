"""Phase breakdown of the selection kernel (needs a -DHALO_SEL_PROFILE build) + CUDA-event times of K2/K3 at batch B."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth, _native as nat
from halo_b200.floating_region import score_planes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
dev = "cuda:0"
H, W = 640, 1280
cfg = halo_b200.AcquisitionConfig(budget=0.05)
g = torch.Generator(device=dev).manual_seed(1)
pixunc = torch.rand((B, H, W), generator=g, device=dev) * 0.9
radius = torch.rand((B, H, W), generator=g, device=dev) * 3 + 1
gt = torch.randint(0, 19, (B, H, W), generator=g, device=dev, dtype=torch.int64).to(torch.uint8)
def run_once():
    active = torch.zeros((B, H, W), dtype=torch.uint8, device=dev); sel = torch.zeros_like(active); msk = torch.full_like(active, 255)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    score, _, _ = score_planes(pixunc, radius, None, None, active, unc_mode=nat.UNC_BOXSUM, pur_mode=nat.PUR_NORM, normalize=True, k=3, pk=3, n_bins=19, want_impurity=False, want_maps=False)
    e[1].record()
    n_picked, _ = halo_b200.select_planes(score, active, sel, msk, gt, 4552, 1, 5, keep_score=True)
    e[2].record(); torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]), int(n_picked.min())
for _ in range(2): run_once()
ws = nat.workspace.get(torch.device(dev), "select", 0)
need = nat.load().halo_select_workspace_bytes(B, H, W, 4552)
prof = ws[need - 128: need - 128 + 48].view(torch.int64)
prof.zero_()
t = [run_once() for _ in range(3)]
print("B=%d score %.3f ms  select %.3f ms  picks %d" % (B, sum(x[0] for x in t) / 3, sum(x[1] for x in t) / 3, t[0][2]))
p = prof.cpu().tolist()
tot = sum(p) or 1
names = ["descent(hist)", "gather", "sort", "greedy", "loop-end", "replay"]
print(" ".join("%s=%.1f%%(%.0f kcyc/img/run)" % (n, 100.0 * v / tot, v / 3 / B / 1e3) for n, v in zip(names, p)))
