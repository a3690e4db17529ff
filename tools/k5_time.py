"""K5 (fused bilinear up-sampling in front of the score) at the real Cityscapes pipeline shape:
head at 160x320 (stride 4... of a 640x1280 crop) -> label size 1024x2048, 19 logits + 256-d embedding."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
dev = "cuda:0"
C, O, h, w, H, W = 256, 19, 160, 320, 1024, 2048
P, A = synth.head_params(O, C, seed=0, device=dev)
u = synth.image_features(0, C, h, w, device=dev).unsqueeze(0)
mapper = halo_b200.HyperMapper(c=1.0)
mlr = halo_b200.HyperMLR(C, O, c=1.0).to(dev)
with torch.no_grad():
    emb = mapper.expmap(u, dim=1)
    logits = mlr(emb)
frs = halo_b200.FloatingRegionScore(in_channels=O, size=3, purity_type="radius", curvature=1.0)
def run():
    return frs.forward_upsampled(logits, emb, (H, W), unc_type="entropy", pur_type="radius", normalize=True)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
# the reference materialises (1,19,H,W) fp32 + (1,256,H,W) fp64 = 159 MB + 4.3 GB before scoring
print(json.dumps({"ms_per_image_upsample_plus_score": round(ms, 3), "Mpixel/s_label_res": round(H * W / ms / 1e3, 1),
                  "flops_fp64_interp": 2 * 4 * C * H * W, "reference_temporaries_GB": round((O * 4 + C * 8) * H * W / 1e9, 2)}))
