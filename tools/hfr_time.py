"""Training-mode conv_reduce + HFR at the reference's training shape, forward + backward (for ncu launch lists / timing)."""
import os, sys, torch, torch.nn as nn
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo_b200.hfr import reduce_hfr
dev = "cuda:0"
N, Cin, C, h, w = 4, 512, 64, 160, 320
torch.manual_seed(5)
conv = nn.Conv2d(Cin, C, 1).to(dev).train()
mlp = nn.Sequential(nn.Linear(C, C), nn.BatchNorm1d(C), nn.ReLU(), nn.Linear(C, C)).to(dev).train()
feats = torch.randn((N, Cin, h, w), device=dev) * 0.3
dz = torch.randn((N, C, h, w), device=dev)
def step():
    x = feats.detach().requires_grad_(True)
    z = reduce_hfr(x, conv, mlp)
    z.backward(dz)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    step()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("reduce_hfr train fwd+bwd: %.3f ms" % (e0.elapsed_time(e1) / 5))
import torch.nn.functional as F
def tstep():
    x = feats.detach().requires_grad_(True)
    y = conv(x)
    t = mlp(y.permute(0, 2, 3, 1).contiguous().view(-1, C)).view(-1, h * w, C)
    wt = torch.clamp(torch.mean(t, dim=1).view(-1, C, 1, 1), min=1e-5)
    z = F.normalize(y.reshape(-1, C, h * w), dim=-1).reshape(-1, C, h, w) * wt
    z.backward(dz)
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    for _ in range(3): tstep()
    e0.record()
    for _ in range(5): tstep()
    e1.record(); torch.cuda.synchronize()
    print("torch eager (cudnn tf32 %s): %.3f ms" % (tf32, e0.elapsed_time(e1) / 5))
