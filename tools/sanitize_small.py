"""Small invocations of every kernel for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth, pool
from halo_b200.losses import fused_seg_loss
from halo_b200.hfr import reduce_hfr
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
ex = pool.RoundExchange(B, H, W, cfg.regions_per_image(H, W), cfg.radius_k, dev)
ex.pack(0, res["picks"], res["n_picked"], d["gt"])
rep2 = torch.full((B, H, W), 255, dtype=torch.uint8, device=dev)
cnt = torch.zeros((B,), dtype=torch.int32, device=dev)
ex.exchange(rep2, cnt); ex.wait()
assert ex.verify(rep2, cnt)[1] and torch.equal(rep2, d["active_mask"])
hcfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.02, purity="hyper")
d2 = synth.batch(0, B, C, O, H, W, device=dev)
halo_b200.acquire_batch(d2["feat"], P, A, hcfg, d2["gt"], d2["active"], d2["selected"], d2["active_mask"])
dus = []
for Cb, Hb, Wb, Nb in ((128, 32, 40, 2), (256, 36, 40, 3), (64, 30, 44, 5)):   # ragged tiles, 1 and 2 channel blocks, several drains
    Pb, Ab = synth.head_params(O, Cb, seed=0, device=dev)
    u = torch.stack([synth.image_features(i, Cb, Hb, Wb, device=dev) for i in range(Nb)])
    r = halo_b200.head_forward(u, Pb, Ab, 1.0, want_logits=True, want_saved=True)
    dl = torch.randn_like(r["logits"]) * 1e-3
    du, dP, dA = halo_b200.head_backward(u, Pb, Ab, 1.0, dl, saved=r["saved"])     # streaming kernel
    du2, _, _ = halo_b200.head_backward(u, Pb, Ab, 1.0, dl)                         # recompute + streaming kernel
    assert torch.equal(du, du2)
    dus.append(float(du.abs().max()))
os.environ["HALO_BWD_TWO_KERNEL"] = "1"
halo_b200.head_backward(d["feat"], P, A, 1.0, torch.randn(B, O, H, W, device=dev) * 1e-3)
del os.environ["HALO_BWD_TWO_KERNEL"]
x = (torch.randn(2, O, 8, 10, device=dev) * 3).requires_grad_(True)
lbl = torch.randint(0, O, (2, 30, 37), device=dev)
fused_seg_loss(x, lbl, (30, 37))[0].backward()
conv = torch.nn.Conv2d(96, 32, 1).to(dev).eval()
mlp = torch.nn.Sequential(torch.nn.Linear(32, 32), torch.nn.BatchNorm1d(32), torch.nn.ReLU(), torch.nn.Linear(32, 32)).to(dev).eval()
with torch.no_grad():
    z = reduce_hfr(torch.randn(2, 96, 9, 11, device=dev), conv, mlp)
conv.train(); mlp.train()
zt = reduce_hfr(torch.randn(2, 96, 9, 11, device=dev, requires_grad=True), conv, mlp)    # training mode: batch statistics + backward
zt.square().sum().backward()
torch.cuda.synchronize()
print("ok", dus, float(dP.abs().max()), int(res["n_picked"].min()), float(x.grad.abs().max()), float(z.abs().max()))
