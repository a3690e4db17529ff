import os, sys, torch
sys.path.insert(0, os.getcwd())
import halo_b200
from halo_b200 import synth
dev = "cuda:0"
B, C, O, H, W = 64, 64, 19, 640, 1280
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.stack([synth.image_features(i % 8, C, H, W, device=dev) for i in range(B)])
for _ in range(3):
    halo_b200.head_forward(feat, P, A, 1.0, want_logits=False, want_radius=True, want_pixunc=True)
torch.cuda.synchronize()
