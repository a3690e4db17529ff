"""Seeded synthetic inputs shaped like the reference's workloads (no datasets, no checkpoints).

Common recipe: features u ~ N(0, sigma^2) fp32 NCHW with sigma = 0.1 (|u| ~ 1.6, mostly
interior points); head parameters U(+-1/sqrt(C)) (what kaiming_uniform_(a=sqrt(5)) gives an (O,C) matrix,
hyperbolic.py:117-118); ground truth randint(0,O) with 5 % ignore (255); active/selected all False and
active_mask all 255 for round 1.  Generators are per image (seed + image index) so any shard of a pool can
be produced independently on any rank and on either device type.
"""
import math

import torch

CONFIGS = {
    # BASELINE.json configs[0..4]
    "cfg1_cpu_reference": dict(n_images=4, H=320, W=640, C=256, O=19, radius_k=1, mask_radius_k=5, budget=0.022, n_rounds=1),
    "cfg2_gtav_cityscapes": dict(n_images=2975, H=640, W=1280, C=256, O=19, radius_k=1, mask_radius_k=5, budget=0.05, n_rounds=1),
    "cfg4_synthia_pixel": dict(n_images=2975, H=640, W=1280, C=256, O=16, radius_k=2, mask_radius_k=5, budget=0.022, n_rounds=1),
    "cfg5_train_step": dict(n_images=8, H=640, W=1280, C=256, O=19),
}


def head_params(num_classes, channels, seed=0, device="cpu", dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / math.sqrt(channels)
    P = (torch.rand((num_classes, channels), generator=g, dtype=torch.float64) * 2 - 1) * bound
    A = (torch.rand((num_classes, channels), generator=g, dtype=torch.float64) * 2 - 1) * bound
    return P.to(device=device, dtype=dtype), A.to(device=device, dtype=dtype)


def image_features(index, C, H, W, sigma=0.1, seed=1234, device="cpu"):
    """(C,H,W) fp32 features of pool image `index`."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed + index)
    return torch.randn((C, H, W), generator=g, device=dev, dtype=torch.float32) * sigma


def image_labels(index, O, H, W, seed=1234, device="cpu", ignore_frac=0.05):
    """(H,W) uint8 ground truth with `ignore_frac` of the pixels set to 255."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed + 7919 * (index + 1))
    gt = torch.randint(0, O, (H, W), generator=g, device=dev, dtype=torch.int64).to(torch.uint8)
    hole = torch.rand((H, W), generator=g, device=dev) < ignore_frac
    gt[hole] = 255
    return gt


def batch(lo, hi, C, O, H, W, sigma=0.1, seed=1234, device="cpu"):
    """Round-1 state for pool images [lo, hi): dict(feat, gt, active, selected, active_mask)."""
    n = hi - lo
    feat = torch.empty((n, C, H, W), dtype=torch.float32, device=device)  # filled in place: a pool batch can be > 100 GB
    gt = torch.empty((n, H, W), dtype=torch.uint8, device=device)
    for k, i in enumerate(range(lo, hi)):
        feat[k] = image_features(i, C, H, W, sigma, seed, device)
        gt[k] = image_labels(i, O, H, W, seed, device)
    return dict(
        feat=feat, gt=gt,
        active=torch.zeros((n, H, W), dtype=torch.uint8, device=device),
        selected=torch.zeros((n, H, W), dtype=torch.uint8, device=device),
        active_mask=torch.full((n, H, W), 255, dtype=torch.uint8, device=device),
    )
