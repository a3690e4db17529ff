"""BASELINE.json configs[4]: fused head forward+backward training step (batch 8 x 1280x640 x 256-d, 19 classes, fp32)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = "cuda:0"
C, H, W, O = 256, 640, 1280, 19
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.empty((B, C, H, W), device=dev)
for i in range(B):
    feat[i] = synth.image_features(i, C, H, W, device=dev)
dl = torch.randn((B, O, H, W), device=dev) * 1e-3
def step():
    lg = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True)["logits"]
    return halo_b200.head_backward(feat, P, A, 1.0, dl)
for _ in range(2): step()
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
torch.cuda.synchronize(); e[0].record()
for _ in range(3): halo_b200.head_forward(feat, P, A, 1.0, want_logits=True)
e[1].record()
for _ in range(3): halo_b200.head_backward(feat, P, A, 1.0, dl)
e[2].record(); torch.cuda.synchronize()
fwd, bwd = e[0].elapsed_time(e[1]) / 3, e[1].elapsed_time(e[2]) / 3
px = B * H * W
alg = (12 * C + 8 * O) * px
print(json.dumps({"batch": B, "fwd_ms": round(fwd, 3), "bwd_ms": round(bwd, 3), "step_ms": round(fwd + bwd, 3),
                  "Mpixel/s": round(px / (fwd + bwd) / 1e3, 1), "algorithmic_GB": round(alg / 1e9, 2),
                  "GB/s": round(alg / (fwd + bwd) / 1e6, 1), "frac_of_6547.5": round(alg / (fwd + bwd) / 1e6 / 6547.5, 4)}))
