import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200
from halo_b200 import synth
dev = "cuda:0"
O = 19
first = sys.argv[1] if len(sys.argv) > 1 else "stream"
for C, B, H, W in ((256, 3, 333, 500), (256, 1, 333, 500), (128, 3, 333, 500), (256, 3, 320, 520)):
    P, A = synth.head_params(O, C, seed=0, device=dev)
    feat = torch.stack([synth.image_features(i, C, H, W, device=dev) for i in range(B)])
    dl = torch.randn((B, O, H, W), device=dev, generator=torch.Generator(device=dev).manual_seed(1)) * 1e-3
    def two():
        os.environ["HALO_BWD_TWO_KERNEL"] = "1"
        r = [t.clone() for t in halo_b200.head_backward(feat, P, A, 1.0, dl)]
        del os.environ["HALO_BWD_TWO_KERNEL"]
        return r
    ref = two() if first == "two" else None
    outs = []
    for it in range(4):
        r = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
        outs.append([t.clone() for t in halo_b200.head_backward(feat, P, A, 1.0, dl, saved=r["saved"])])
        if first == "sync":
            torch.cuda.synchronize()
    if ref is None:
        ref = two()
    torch.cuda.synchronize()
    line = []
    for it in range(4):
        line.append("it%d " % it + " ".join("%s %.1e" % (name, float((a - b).abs().max() / b.abs().max())) for name, a, b in zip(("du", "dP", "dA"), outs[it], ref)))
    print("C=%d B=%d %dx%d | " % (C, B, H, W) + " | ".join(line), flush=True)
