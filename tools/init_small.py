import os, sys, torch
sys.path.insert(0, "/root/repo")
import halo_b200
from halo_b200 import synth
dev = "cuda:0"
C, B, H, W, O = 256, 1, 37, 100, 19     # HW = 3700: ragged (3700 % 128 = 116)
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.stack([synth.image_features(i, C, H, W, device=dev) for i in range(B)])
dl = torch.randn((B, O, H, W), device=dev) * 1e-3
r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
out = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
torch.cuda.synchronize()
print("ok", float(out[1].abs().max()))
