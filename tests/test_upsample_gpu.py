"""GPU: fused bilinear up-sampling in front of the score (halo_upsample_score_inputs) against the
reference sequence F.interpolate(logits) / F.interpolate(fp64 embedding) -> FloatingRegionScore (core/active/build.py:122-144)
run by the oracle on the CPU, and the drop-in RegionSelection end to end on a fake loader."""
import math
import os

import numpy as np
import pytest
import torch

import halo_b200
from halo_b200 import synth
from oracle import acquire as oacquire
from oracle import head as ohead
from oracle import select as oselect
from tests.util import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("case", [
    # (C, O, lh, lw, eh, ew, H, W, unc, pur, normalize)
    (64, 19, 20, 40, 20, 40, 64, 128, "entropy", "radius", True),      # DeepLab v2 head: both at decoder size
    (64, 19, 80, 160, 20, 40, 128, 256, "entropy", "radius", True),    # v3+ head: logits already up-sampled once
    (32, 16, 17, 23, 17, 23, 50, 71, "entropy", "euc_norm", False),    # odd sizes
    (32, 19, 16, 24, 16, 24, 48, 72, "entropy", "ripu", False),
    (32, 19, 16, 24, 16, 24, 48, 72, "pixel_entropy", "radius", True),
    (32, 19, 16, 24, 16, 24, 48, 72, "entropy", "hyper", True),        # K=100 radius bins in fp64 (floating_region.py:94-110)
    (64, 19, 40, 80, 20, 40, 96, 160, "entropy", "hyper", False),
])
@pytest.mark.parametrize("emb_kind", ["lazy", "ball64"])
def test_upsampled_score_matches_reference_sequence(case, emb_kind):
    C, O, lh, lw, eh, ew, H, W, unc, pur, norm = case
    P, A = synth.head_params(O, C, seed=4, dtype=torch.float64)
    g = torch.Generator().manual_seed(5)
    u = torch.randn((1, C, eh, ew), generator=g) * 0.15
    logits_e, x, _ = ohead.head_forward(u, P, A, 1.0)          # logits at the embedding resolution
    logits_lr = torch.nn.functional.interpolate(logits_e, size=(lh, lw), mode="bilinear", align_corners=True)
    ref_s, ref_i, ref_u = oacquire.upsampled_score(logits_lr, x, (H, W), unc_type=unc, pur_type=pur, normalize=norm,
                                                   ground_truth=None, in_channels=O, size=3, ctor_purity_type=pur, K=100,
                                                   c=1.0)
    frs = halo_b200.FloatingRegionScore(in_channels=O, size=3, purity_type=pur, curvature=1.0)
    if emb_kind == "lazy":
        emb = halo_b200.HyperMapper(1.0).expmap(u.to(DEV), dim=1)
    else:
        emb = x.to(DEV)
    s, imp, un = frs.forward_upsampled(logits_lr.to(DEV), emb, (H, W), unc_type=unc, pur_type=pur, normalize=norm)
    tol = TOL if pur != "ripu" else 1e-3   # an arg-max flip between near-equal interpolated logits moves a 3x3 histogram
    frac = ((s.double().cpu() - ref_s.double()).abs() <= tol * ref_s.abs().max()).float().mean().item()
    assert frac >= (1.0 if pur != "ripu" else 0.995)
    assert rel_err(un, ref_u) <= TOL
    if pur != "ripu":
        assert rel_err(imp, ref_i) <= TOL


class _Head(torch.nn.Module):
    """Stand-in for ASPP_Classifier_V2_Hyper.forward (core/models/classifier.py:365-379) on precomputed features."""

    def __init__(self, C, O):
        super().__init__()
        self.mapper = halo_b200.HyperMapper(c=1.0)
        self.conv_seg = halo_b200.HyperMLR(C, O, c=1.0)

    def forward(self, x, size=None):
        embed = self.mapper.expmap(x["out"], dim=1)
        out = self.conv_seg(embed.double()).float()
        return out, embed


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def test_region_selection_dropin_end_to_end(tmp_path):
    C, O, h, w, H, W = 32, 19, 24, 48, 96, 192
    cfg = _Cfg(ACTIVE=_Cfg(RADIUS_K=1, MASK_RADIUS_K=5, BUDGET=0.05, SELECT_ITER=[0], UNCERTAINTY="entropy", PURITY="radius",
                           K=100, NORMALIZE=True, VIZ_MASK=False),
               MODEL=_Cfg(NUM_CLASSES=O, CURVATURE=1.0, HYPER=True))
    head = _Head(C, O).to(DEV)
    fe = torch.nn.Identity()
    items, refs = [], []
    for i in range(2):
        u = synth.image_features(i, C, h, w, sigma=0.15)
        gt = synth.image_labels(i, O, H, W).long()
        items.append({
            "img": {"out": u[None]}, "path_to_mask": [os.path.join(tmp_path, "m%d.png" % i)],
            "path_to_indicator": [os.path.join(tmp_path, "i%d.pth" % i)], "origin_mask": [torch.full((H, W), 255, dtype=torch.long)],
            "origin_label": [gt], "size": [(H, W)], "active": [torch.zeros((H, W), dtype=torch.bool)],
            "selected": [torch.zeros((H, W), dtype=torch.bool)],
        })
        # oracle: reference sequence on the CPU
        lo, x, _ = ohead.head_forward(u[None], head.conv_seg.P_MLR.detach().cpu().double(), head.conv_seg.A_MLR.detach().cpu().double(), 1.0)
        s, _, _ = oacquire.upsampled_score(lo, x, (H, W), unc_type="entropy", pur_type="radius", normalize=True,
                                           ground_truth=None, in_channels=O, size=3, ctor_purity_type="radius", K=100, c=1.0)
        n_regions = math.ceil(H * W * 0.05 / 9)
        s_np = s.numpy().copy()
        a_np, sel_np = np.zeros((H, W), bool), np.zeros((H, W), bool)
        m_np = np.full((H, W), 255, np.int64)
        oselect.select_numpy(s_np, n_regions, 1, 5, a_np, sel_np, m_np, gt.numpy())
        refs.append(m_np.astype(np.uint8))

    class _Img(dict):
        """a batch of backbone features standing in for the image tensor (the backbone is out of scope)"""
        shape = (1, 3, h * 4, w * 4)
        device = torch.device("cpu")

        def cuda(self, non_blocking=False):
            r = _Img({k: v.cuda() for k, v in self.items()})
            r.device = torch.device(DEV)
            return r

    for it in items:
        it["img"] = _Img(it["img"])

    halo_b200.RegionSelection(cfg, fe, head, items, round_number=1)
    from PIL import Image

    for i in range(2):
        got = np.array(Image.open(items[i]["path_to_mask"][0]))
        assert got.shape == (H, W) and got.dtype == np.uint8
        assert (got == refs[i]).mean() >= 0.999
        ind = torch.load(items[i]["path_to_indicator"][0])
        assert ind["active"].dtype == torch.bool and ind["selected"].sum().item() > 0
        assert ((got != 255) == ind["selected"].numpy()).mean() >= 0.94   # selected pixels carry labels (255 = ignore in gt)
