"""Accuracy of training-mode conv_reduce + HFR at the reference's training shape: this library and the torch fp32 sequence,
both against the same sequence in float64 (max error relative to the max of the reference).  With batch statistics the
max-norm error of dfeat / dW1 is set by the few pixels whose hidden pre-activation sits on the ReLU kink: fp32 rounding decides
the side, and one flipped pixel is 1e-3 of max|dfeat| at this shape (the mean error is 1e-10)."""
import os, sys, copy, torch, torch.nn as nn, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo_b200.hfr import reduce_hfr
dev = "cuda:0"
N, Cin, C, h, w = 4, 512, 64, 160, 320
torch.manual_seed(5)
conv = nn.Conv2d(Cin, C, 1).to(dev).train()
mlp = nn.Sequential(nn.Linear(C, C), nn.BatchNorm1d(C), nn.ReLU(), nn.Linear(C, C)).to(dev).train()
g = torch.Generator(device=dev).manual_seed(12)
feats = torch.randn((N, Cin, h, w), device=dev, generator=g) * 0.3
dz = torch.randn((N, C, h, w), device=dev, generator=g)
torch.backends.cudnn.allow_tf32 = False
def seq(conv, mlp, x, dzz):
    y = conv(x)
    if mlp is not None:
        t = mlp(y.permute(0, 2, 3, 1).contiguous().view(-1, C)).view(-1, h * w, C)
        wt = torch.clamp(torch.mean(t, dim=1).view(-1, C, 1, 1), min=1e-5)
        y = F.normalize(y.reshape(-1, C, h * w), dim=-1).reshape(-1, C, h, w) * wt
    y.backward(dzz)
    extra = [p.grad for p in mlp.parameters()] if mlp is not None else []
    return [y.detach(), x.grad, conv.weight.grad] + extra
for label, use_mlp, train in (("no HFR", False, True), ("HFR, BatchNorm in eval mode", True, False), ("HFR, batch statistics", True, True)):
    def mk(dtype):
        c, m = copy.deepcopy(conv).to(dtype), (copy.deepcopy(mlp).to(dtype) if use_mlp else None)
        if m is not None: m.train(train)
        return c, m
    c64, m64 = mk(torch.float64)
    ref = seq(c64, m64, feats.double().requires_grad_(True), dz.double())
    c32, m32 = mk(torch.float32)
    t32 = seq(c32, m32, feats.clone().requires_grad_(True), dz)
    co, mo = mk(torch.float32)
    x = feats.clone().requires_grad_(True)
    z = reduce_hfr(x, co, mo); z.backward(dz)
    ours = [z.detach(), x.grad, co.weight.grad] + ([p.grad for p in mo.parameters()] if mo is not None else [])
    for name, a in (("torch fp32", t32), ("halo_b200 ", ours)):
        errs = [float((p.double() - q).abs().max() / max(float(q.abs().max()), 1e-30)) for p, q in zip(a, ref)]
        print("%-28s %s  z %.2e  dfeat %.2e  dWr %.2e   mlp (W1 b1 gamma beta W2 b2): %s" % (label, name, errs[0], errs[1], errs[2], " ".join("%.1e" % e for e in errs[3:])))
        if name.startswith("halo") and use_mlp:
            print("      magnitudes of the reference mlp grads:", " ".join("%.1e" % float(q.abs().max()) for q in ref[3:]), " dfeat max %.2e" % float(ref[1].abs().max()))
