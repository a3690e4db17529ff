"""Per-iteration times of the fused head kernel over a long back-to-back run (is the rate time-dependent?)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
B = int(sys.argv[1]); iters = int(sys.argv[2])
dev = "cuda:0"
C, H, W, O = 256, 640, 1280, 19
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.empty((B, C, H, W), device=dev)
for i in range(B):
    feat[i] = synth.image_features(i, C, H, W, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
torch.cuda.synchronize()
ev[0].record()
for i in range(iters):
    halo_b200.head_forward(feat, P, A, 1.0, want_logits=False, want_radius=True, want_pixunc=True)
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
gbs = [4.0 * C * B * H * W / m / 1e6 for m in ms]
print("B=%d" % B, " ".join("%.0f" % g for g in gbs[:: max(1, iters // 24)]))
