"""Time the fused head kernel alone (CUDA events) on a resident batch; prints GB/s of algorithmic bytes."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
O = int(sys.argv[2]) if len(sys.argv) > 2 else 19
dev = "cuda:0"
C = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 256
H, W = 640, 1280
P, A = synth.head_params(O, C, seed=0, device=dev)
feat = torch.empty((B, C, H, W), device=dev)
for i in range(B):
    feat[i] = synth.image_features(i, C, H, W, device=dev)
for tc in (True, False) if "--cc" in sys.argv else (True,):
    for _ in range(3):
        halo_b200.head_forward(feat, P, A, 1.0, want_logits=False, want_radius=True, want_pixunc=True, tensor_cores=tc)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        halo_b200.head_forward(feat, P, A, 1.0, want_logits=False, want_radius=True, want_pixunc=True, tensor_cores=tc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gbs = 4.0 * C * B * H * W / ms / 1e6
    print(json.dumps({"lib": os.environ.get("HALO_B200_LIB", "default"), "tensor_cores": tc, "batch": B, "classes": O, "C": C, "ns_per_px": round(ms * 1e6 / (B * H * W), 4), "ms": round(ms, 3), "GB/s": round(gbs, 1), "frac_of_6547.5": round(gbs / 6547.5, 4)}))
