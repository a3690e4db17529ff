"""Race hunt for the streaming backward: many launches at full size, every result compared bitwise with the first one
(the kernel is deterministic by construction; any difference or hang is a synchronisation bug).  Run under `timeout`."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200  # noqa: E402
from halo_b200 import synth  # noqa: E402

dev = "cuda:0"
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
bad = 0
for C, B, H, W in ((256, 8, 640, 1280), (128, 8, 640, 1280), (64, 8, 640, 1280), (256, 3, 333, 500), (64, 16, 160, 320)):
    O = 19
    P, A = synth.head_params(O, C, seed=0, device=dev)
    feat = torch.stack([synth.image_features(i, C, H, W, device=dev) for i in range(B)])
    dl = torch.randn((B, O, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 1e-3
    r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
    ref = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
    torch.cuda.synchronize()
    n_bad = 0
    for it in range(iters):
        r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
        out = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
        if not all(torch.equal(a, b) for a, b in zip(out, ref)):
            n_bad += 1
    torch.cuda.synchronize()
    print("C=%d B=%d %dx%d: %d / %d launches differ" % (C, B, H, W, n_bad, iters), flush=True)
    bad += n_bad
print("RACE" if bad else "clean")
