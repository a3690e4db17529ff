"""Determinism / parity probe of the streaming backward on ragged images (H*W % 128 != 0): prints, per configuration, how
many of `iters` launches deviate from the two-kernel path by more than 1e-4 (a synchronisation bug shows as a whole CTA's
partial going wrong: ~1e-2 on dA)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
dev = "cuda:0"
O = 19
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 12
res = []
for C, B, H, W in ((256, 3, 333, 500), (256, 1, 333, 500), (128, 3, 333, 500), (64, 4, 333, 500), (256, 3, 320, 520)):
    P, A = synth.head_params(O, C, seed=0, device=dev)
    feat = torch.stack([synth.image_features(i, C, H, W, device=dev) for i in range(B)])
    dl = torch.randn((B, O, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 1e-3
    os.environ["HALO_BWD_TWO_KERNEL"] = "1"
    ref = [t.clone() for t in halo_b200.head_backward(feat, P, A, 1.0, dl)]
    del os.environ["HALO_BWD_TWO_KERNEL"]
    bad = 0
    for it in range(iters):
        try:
            r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
            out = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])
        except TypeError:
            out = halo_b200.head_backward(feat, P, A, 1.0, dl)
        err = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(out, ref))
        bad += err > 1e-4
    res.append("%dx%dx%d:%d/%d" % (C, B, H * W % 128, bad, iters))
print(os.environ.get("HALO_B200_LIB", "default").split("/")[-1], " ".join(res))
