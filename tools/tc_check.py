"""Quick parity probe of the tcgen05 head kernel against the CUDA-core kernel and the fp64 oracle."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200  # noqa: E402
from halo_b200 import synth  # noqa: E402
from oracle import head as ohead  # noqa: E402

dev = "cuda:0"
cases = [(19, 32, 8, 16, 1, 0.1), (19, 256, 16, 64, 2, 0.1), (16, 64, 24, 40, 3, 0.3), (19, 256, 24, 40, 2, 1.0),
         (3, 64, 10, 10, 1, 0.2), (32, 128, 9, 12, 2, 0.05), (19, 256, 640, 1280, 1, 0.1)]
for O, C, H, W, N, sigma in cases:
    P, A = synth.head_params(O, C, seed=3, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    ud, Pd, Ad = u.to(dev), P.to(dev), A.to(dev)
    kw = dict(want_logits=True, want_radius=True, want_pixunc=True, want_label=True, want_stats=True)
    tc = halo_b200.head_forward(ud, Pd, Ad, 1.0, tensor_cores=True, **kw)
    torch.cuda.synchronize()
    cc = halo_b200.head_forward(ud, Pd, Ad, 1.0, tensor_cores=False, **kw)
    torch.cuda.synchronize()
    d = (tc["logits"] - cc["logits"]).abs().max().item() / cc["logits"].abs().max().item()
    msg = "O=%d C=%d %dx%d N=%d sigma=%g: tc-vs-cc logits %.2e radius %.2e pixunc %.2e label-agree %.5f" % (
        O, C, H, W, N, sigma, d, (tc["radius"] - cc["radius"]).abs().max().item(),
        (tc["pixunc"] - cc["pixunc"]).abs().max().item(), (tc["label"] == cc["label"]).float().mean().item())
    if H * W <= 100000:
        lo, x, rad = ohead.head_forward(u, P, A, 1.0)
        e_tc = (tc["logits"].cpu().double() - lo).abs().max().item() / lo.abs().max().item()
        e_cc = (cc["logits"].cpu().double() - lo).abs().max().item() / lo.abs().max().item()
        msg += " | vs oracle: tc %.2e cc %.2e" % (e_tc, e_cc)
    print(msg, flush=True)
print("stats", tc["stats"][0, :2].tolist(), cc["stats"][0, :2].tolist())
