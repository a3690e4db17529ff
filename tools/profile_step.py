"""Tiny driver for ncu: a few acquisition steps on a resident batch (same kernels/arguments as bench.py)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200  # noqa: E402
from halo_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--classes", type=int, default=19)
ap.add_argument("--mode", default="acquire", choices=["acquire", "head", "train"])
args = ap.parse_args()
dev = "cuda:0"
C, O, H, W = 256, args.classes, 640, 1280
cfg = halo_b200.AcquisitionConfig(num_classes=O, budget=0.05)
P, A = synth.head_params(O, C, seed=0, device=dev)
d = synth.batch(0, args.batch, C, O, H, W, device=dev)
for _ in range(args.steps):
    if args.mode == "acquire":
        d["active"].zero_(); d["selected"].zero_(); d["active_mask"].fill_(255)
        halo_b200.acquire_batch(d["feat"], P, A, cfg, d["gt"], d["active"], d["selected"], d["active_mask"])
    elif args.mode == "head":
        halo_b200.head_forward(d["feat"], P, A, 1.0, want_logits=True, want_radius=True)
    else:
        dl = torch.randn((args.batch, O, H, W), device=dev) * 1e-3
        r = halo_b200.head_forward(d["feat"], P, A, 1.0, want_logits=True, want_saved=True)   # the training step of bench.py
        halo_b200.head_backward(d["feat"], P, A, 1.0, dl, saved=r["saved"])
torch.cuda.synchronize()
print("done")
