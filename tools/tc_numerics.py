"""Numerics probe (needs a -DHALO_TC_VARIANTS build): accuracy of the raw contraction T_k = <u, a_hat_k> on the
tensor-core path for several accumulation schemes, against the exact fp64 value and the CUDA-core fp32 chain."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import halo_b200  # noqa: E402
from halo_b200 import synth  # noqa: E402
from oracle import head as ohead  # noqa: E402

dev = "cuda:0"
O, C, H, W, N = 19, 256, 48, 64, 2
P, A = synth.head_params(O, C, seed=3, dtype=torch.float64)
Pf, Af = P.float(), A.float()
a_hat = (Af.double() / Af.double().norm(dim=1, keepdim=True)).float().double()  # what the pack kernel rounds to
variants = {"cc": None, "tc main/corr split": "0"}
for sigma in (0.1, 0.3, 1.0, 3.0):
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    exact = torch.einsum("bchw,oc->bohw", u.double(), a_hat)
    lo_ref, _, _ = ohead.head_forward(u, Pf.double(), Af.double(), 1.0)
    scale = exact.abs().max().item()
    ud = u.to(dev)
    for name, v in variants.items():
        os.environ["HALO_TC_DEBUG"] = (v or "0") + "r"
        raw = halo_b200.head_forward(ud, Pf.to(dev), Af.to(dev), 1.0, tensor_cores=v is not None)["logits"].cpu().double()
        os.environ["HALO_TC_DEBUG"] = (v or "0")
        lg = halo_b200.head_forward(ud, Pf.to(dev), Af.to(dev), 1.0, tensor_cores=v is not None)["logits"].cpu().double()
        err = (raw - exact).abs()
        print("sigma=%.2f %-14s dot err max %.2e rms %.2e (of max|T|=%.2f) | logits err/max %.2e" % (
            sigma, name, err.max().item() / scale, err.pow(2).mean().sqrt().item() / scale, scale,
            (lg - lo_ref).abs().max().item() / lo_ref.abs().max().item()), flush=True)
