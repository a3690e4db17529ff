"""Oracle for the floating-region acquisition score (ORACLE -- test infrastructure, CPU only).

Restates core/active/floating_region.py:
  * init_conv_layer / __init__   :12-19, :26-68   all-ones k x k box filters, zero padding,
                                                   'hyper' purity forces a 3x3 purity window (:54-55)
  * normalize_map                :22-23
  * compute_region_uncertainty   :70-92
  * quantize_uncert_map          :94-110
  * compute_region_impurity      :112-121
  * compute_pixel_entropy        :123-127
  * forward                      :129-217
Quirks kept on purpose: entropy is divided by log(19) whatever the class count (:74-76,:126);
the box filter is a zero-padded SUM; `count` is the true window population only in the
ripu/oracle_ripu/hyper branches and 1 elsewhere (:184-198,:204).
"""
import math

import torch
import torch.nn.functional as F

from . import head as _head

_LOG19 = math.log(19)


def _box_sum(x, size):
    """x: (1,Cn,H,W) float32 -> depthwise all-ones size x size conv, zero padded (:12-19,:42-66)."""
    cn = x.shape[1]
    w = torch.ones((cn, 1, size, size), dtype=torch.float32)
    return F.conv2d(x, w, bias=None, stride=1, padding=size // 2, groups=cn)


def minmax_normalize(x):
    """:22-23 -- python-float extrema, (x - min) / (max - min)."""
    lo = x.min().item()
    hi = x.max().item()
    return (x - lo) / (hi - lo)


def pixel_entropy(p):
    """:72-76 / :124-126 -- p: (O,H,W) softmax; returns (1,1,H,W)."""
    e = torch.sum(-p * torch.log(p + 1e-6), dim=0)
    return e[None, None] / _LOG19


def region_uncertainty(unc_type, p, size, ground_truth=None):
    """:70-92 and the pixel_entropy shortcut at :158-159."""
    h, w = p.shape[1:]
    if unc_type == "pixel_entropy":
        return pixel_entropy(p)
    if unc_type == "entropy":
        r = pixel_entropy(p)
    elif unc_type == "oracle_acc":
        gt = ground_truth.clone()
        hole = ground_truth == 255
        gt[hole] = p.argmax(dim=0)[hole]
        r = (1 - torch.gather(p, 0, gt[None]))[None]
    else:  # "none" and every unknown string (e.g. "hyperbolic", "certainty") -> zeros (:85-87)
        r = torch.zeros((1, 1, h, w), dtype=torch.float32)
    if unc_type != "none":
        r = _box_sum(r.float(), size)
    return r


def quantize_radius(decoder_out, K, c):
    """:94-110 -- radius -> min-max -> 1-x -> min-max -> K bins, round half to even."""
    eps = 1e-5
    r = _head.radius(decoder_out, c, dim=1).squeeze(0)
    r = (r - r.min().item()) / (r.max().item() - r.min().item())
    r = 1 - r
    lo, hi = r.min(), r.max()
    r = (r - lo) / (hi - lo)
    b = r * K - 0.5
    b = torch.clamp(b, min=-0.5 + eps, max=K - 0.5 - eps)
    return torch.round(b).long()


def region_impurity(predict, n_bins, size):
    """:112-121 -- k x k label histogram entropy; returns (impurity, count), both (1,1,H,W) fp32."""
    one_hot = F.one_hot(predict, num_classes=n_bins).float().permute(2, 0, 1)[None]
    summary = _box_sum(one_hot, size)
    count = summary.sum(dim=1, keepdim=True)
    dist = summary / count
    imp = torch.sum(-dist * torch.log(dist + 1e-6), dim=1, keepdim=True) / math.log(n_bins)
    return imp, count


def floating_region_score(
    logit,
    decoder_out=None,
    unc_type=None,
    pur_type=None,
    normalize=False,
    ground_truth=None,
    *,
    in_channels=19,
    size=3,
    ctor_purity_type=None,
    K=100,
    c=1.0,
):
    """forward, :129-217.  logit (1,O,H,W) fp32; decoder_out (1,C,H,W) on the ball (fp64 in the reference).

    `size` is the constructor's window; `ctor_purity_type`/`K` mirror the constructor arguments that
    decide the purity window (3x3 when the module was BUILT for 'hyper', :54-55).
    Returns (score, impurity, uncertainty), each (H,W)."""
    assert size % 2 == 1, "error size"
    purity_size = 3 if ctor_purity_type == "hyper" else size
    logit = logit.squeeze(0)
    h, w = logit.shape[1:]
    p = torch.softmax(logit, dim=0)
    unc = region_uncertainty(unc_type, p, size, ground_truth)

    ones = torch.ones((1, 1, h, w), dtype=torch.float32)
    if pur_type == "ripu":
        imp, count = region_impurity(torch.argmax(p, dim=0), in_channels, purity_size)
    elif pur_type == "oracle_ripu":
        pred = ground_truth.clone()
        hole = ground_truth == 255
        pred[hole] = p.argmax(dim=0)[hole]
        imp, count = region_impurity(pred, in_channels, purity_size)
    elif pur_type == "hyper":
        imp, count = region_impurity(quantize_radius(decoder_out, K, c), K, purity_size)
    elif pur_type == "none":
        imp, count = torch.zeros((1, 1, h, w), dtype=torch.float32), ones
    elif pur_type == "radius":
        imp, count = _head.radius(decoder_out, c, dim=1).unsqueeze(0), ones
    elif pur_type == "euc_norm":
        imp, count = decoder_out.norm(dim=1).unsqueeze(0), ones
    else:
        raise NotImplementedError("Error: purity type '{}' not implemented".format(pur_type))

    unc = unc / count  # :204
    if normalize:  # :206-208
        unc = minmax_normalize(unc)
        imp = minmax_normalize(imp)
    score = imp * unc  # :210
    return score[0, 0], imp[0, 0], unc[0, 0]
