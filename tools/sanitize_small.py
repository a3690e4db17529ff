"""Small invocations of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth, pool
dev = "cuda:0"
C, O, H, W, B = 128, 19, 32, 40, 2
P, A = synth.head_params(O, C, seed=0, device=dev)
d = synth.batch(0, B, C, O, H, W, device=dev)
cfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.05)
res = halo_b200.acquire_batch(d["feat"], P, A, cfg, d["gt"], d["active"], d["selected"], d["active_mask"], want_picks=True)
lab = pool.pack_round_delta(res["picks"], res["n_picked"], d["gt"], cfg.radius_k)
rep = torch.full((B, H, W), 255, dtype=torch.uint8, device=dev)
pool.apply_round_delta(rep, torch.arange(B, dtype=torch.int32, device=dev), res["picks"], res["n_picked"], lab, cfg.radius_k)
assert torch.equal(rep, d["active_mask"])
lg = halo_b200.head_forward(d["feat"], P, A, 1.0, want_logits=True)["logits"]
dl = torch.randn_like(lg) * 1e-3
du, dP, dA = halo_b200.head_backward(d["feat"], P, A, 1.0, dl)
torch.cuda.synchronize()
print("ok", float(du.abs().max()), float(dP.abs().max()), int(res["n_picked"].min()))
