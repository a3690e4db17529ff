"""BASELINE.json configs[4]: fused head forward+backward training step (batch 8 x 1280x640 x C-d, 19 classes, fp32).

    python tools/bwd_time.py [batch] [C]     # env HALO_BWD_TWO_KERNEL=1 times the round-1 two-kernel path
Training forward = logits + saved contractions; backward = the streaming kernel fed with them (features read once)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200  # noqa: E402
from halo_b200 import _native as nat  # noqa: E402
from halo_b200 import synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = "cuda:0"
H, W, O = 640, 1280, 19
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.empty((B, C, H, W), device=dev)
for i in range(B):
    feat[i] = synth.image_features(i, C, H, W, device=dev)
dl = torch.randn((B, O, H, W), device=dev) * 1e-3
two_kernel = os.environ.get("HALO_BWD_TWO_KERNEL") == "1"


def fwd():
    return halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=not two_kernel)


for _ in range(3):
    r = fwd()
    halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
path = nat.last_path()
reps = 5
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
torch.cuda.synchronize()
e[0].record()
for _ in range(reps):
    r = fwd()
e[1].record()
for _ in range(reps):
    halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
e[2].record()
for _ in range(reps):   # the step as autograd runs it: forward then backward, back to back
    r = fwd()
    halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
e[3].record()
torch.cuda.synchronize()
f, b, s = e[0].elapsed_time(e[1]) / reps, e[1].elapsed_time(e[2]) / reps, e[2].elapsed_time(e[3]) / reps
px = B * H * W
alg = (12 * C + 8 * O) * px
print(json.dumps({"batch": B, "C": C, "path": path, "fwd_ms": round(f, 3), "bwd_ms": round(b, 3), "step_ms": round(s, 3),
                  "Mpixel/s": round(px / s / 1e3, 1), "algorithmic_GB": round(alg / 1e9, 2), "GB/s": round(alg / s / 1e6, 1),
                  "frac_of_6547.5": round(alg / s / 1e6 / 6547.5, 4)}))
