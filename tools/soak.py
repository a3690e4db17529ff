"""Soak: repeat the tensor-core forward and the streaming backward at full size and compare every launch bit for bit with the
first (fixed-order reductions => any difference is a race).   python tools/soak.py [launches]"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
dev = torch.device("cuda", 0)
NL = int(sys.argv[1]) if len(sys.argv) > 1 else 500
O, H, W = 19, 640, 1280
for C, B in ((256, 8), (64, 8), (128, 8)):
    P, A = synth.head_params(O, C, seed=0, device=dev)
    feat = torch.stack([synth.image_features(i, C, H, W, device=dev) for i in range(B)])
    dl = torch.randn((B, O, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 1e-3
    t0 = time.time()
    r0 = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_radius=True, want_pixunc=True, want_saved=True)
    keys = ("logits", "radius", "pixunc", "saved")
    bad_f = torch.zeros((), dtype=torch.int32, device=dev)
    for _ in range(NL):
        r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_radius=True, want_pixunc=True, want_saved=True)
        for k in keys:
            bad_f += (r[k] != r0[k]).any()
        del r
    b0 = [t.clone() for t in halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r0["saved"])]
    bad_b = torch.zeros((), dtype=torch.int32, device=dev)
    for _ in range(NL):
        out = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r0["saved"])
        for a, b in zip(out, b0):
            bad_b += (a != b).any()
    torch.cuda.synchronize()
    print("C=%d batch %d: %d launches each: forward outputs differing %d, backward outputs differing %d (%.0f s)" % (
        C, B, NL, int(bad_f), int(bad_b), time.time() - t0), flush=True)
    del feat, dl, r0, b0
