"""Fused loss forward + backward at the reference's training shape (CUDA events)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo_b200.losses import fused_seg_loss
dev = "cuda:0"
N, O, h, w, H, W = 8, 19, 160, 320, 640, 1280
g = torch.Generator(device=dev).manual_seed(11)
logits = torch.randn((N, O, h, w), device=dev, generator=g) * 3.0
labels = torch.randint(0, O, (N, H, W), device=dev, generator=g)
labels[torch.rand((N, H, W), device=dev, generator=g) > 0.05] = 255
lab8 = labels.to(torch.uint8)
def ours():
    x = logits.detach().requires_grad_(True)
    loss, _, _ = fused_seg_loss(x, lab8, (H, W), 1.0)
    loss.backward()
for _ in range(3): ours()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ours()
e1.record(); torch.cuda.synchronize()
print("%s: fused loss fwd+bwd %.3f ms" % (os.environ.get("HALO_B200_LIB", "default"), e0.elapsed_time(e1) / 10))
