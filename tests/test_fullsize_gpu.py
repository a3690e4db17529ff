"""GPU: VALUE parity at the BASELINE.json image sizes (not only size-independent properties).

The oracle (fp64 head, box-filter score, numpy selection) does one 1280x640x256 image in a few seconds, so the
configurations the headline number is quoted on are compared value by value:
  * configs[0] literally: 4 x 640x320 px, 256-d, 19 classes, radius_K=1, 2.2 % budget -> 501 picks per image;
  * configs[1]: 1280x640 px, 256-d, 19 classes, 3x3 regions, 5 % budget single-shot -> 4 552 picks per image;
  * configs[3]: 16 classes, 5x5 regions, 2.2 % budget -> 721 picks per image;
  * K1 on the tensor cores vs K1 on the CUDA cores vs the oracle over 8 full-size images in ONE launch (more than
    170 tiles per converter warpgroup of every persistent CTA: all mbarrier phases and TMEM buffers wrap many times).
Reference: core/utils/hyperbolic.py:28-39,74-83,120-188, core/active/floating_region.py:129-217,
core/active/build.py:137-160.
"""
import numpy as np
import pytest
import torch

import halo_b200
from halo_b200 import synth
from oracle import acquire as oacquire
from oracle import head as ohead
from tests.util import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _oracle_image(feat, gt, P, A, cfg):
    H, W = feat.shape[-2:]
    return oacquire.acquire_image(
        feat[None], P.double(), A.double(), gt.long(), torch.zeros((H, W), dtype=torch.bool),
        torch.zeros((H, W), dtype=torch.bool), torch.full((H, W), 255, dtype=torch.int64), c=cfg.curvature,
        radius_k=cfg.radius_k, mask_radius_k=cfg.mask_radius_k, budget=cfg.budget, n_rounds=cfg.n_rounds,
        unc_type=cfg.uncertainty, pur_type=cfg.purity, normalize=cfg.normalize, K=cfg.K, fast_select=True)


@pytest.mark.parametrize("name,cfg,shape,expect", [
    ("configs0_cpu_reference", halo_b200.AcquisitionConfig(budget=0.022), (4, 256, 19, 320, 640), 501),
    ("configs1_gtav_cityscapes", halo_b200.AcquisitionConfig(budget=0.05), (2, 256, 19, 640, 1280), 4552),
    ("configs3_synthia_5x5", halo_b200.AcquisitionConfig(num_classes=16, radius_k=2, budget=0.022),
     (2, 256, 16, 640, 1280), 721),
])
def test_acquire_full_size_matches_oracle(name, cfg, shape, expect):
    n, C, O, H, W = shape
    P, A = synth.head_params(O, C, seed=0)
    b = synth.batch(0, n, C, O, H, W)
    d = {k: v.to(DEV) for k, v in b.items()}
    res = halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"],
                                  d["active_mask"], want_score=True, want_picks=True)
    assert cfg.regions_per_image(H, W) == expect
    for i in range(n):
        ref = _oracle_image(b["feat"][i], b["gt"][i], P, A, cfg)
        assert ref["n_regions"] == expect
        assert rel_err(res["score"][i], ref["score"]) <= TOL, (name, i)
        assert int(res["n_picked"][i]) == len(ref["picks"]) == expect, (name, i)
        m = d["active_mask"][i].cpu().numpy()
        agree = (m == ref["active_mask"].numpy().astype(np.uint8)).mean()
        assert agree >= 0.999, (name, i, agree)             # north_star: >= 99.9 % mask agreement end to end
        assert (d["active"][i].cpu().bool() == ref["active"]).float().mean().item() >= 0.999
        assert (d["selected"][i].cpu().bool() == ref["selected"]).float().mean().item() >= 0.999
        # the fp32 score can swap near-equal neighbours in the pick ORDER; the pick SET must agree to >= 99.9 %
        got = set(res["picks"][i][:expect].cpu().tolist())
        want = set(int(h) * W + int(w) for h, w in ref["picks"])
        assert len(got & want) >= 0.999 * expect, (name, i, len(got & want))


@pytest.mark.parametrize("C", [256, 64], ids=["c256-one-epilogue-warpgroup", "c64-two-epilogue-warpgroups"])
def test_head_tensor_core_eight_full_size_images_one_launch(C):
    """K1-TC vs K1-CUDA-core vs the fp64 oracle: logits, radius, pixel entropy, label and per-image min/max over
    8 x 1280x640 px in a single launch (51 200 tiles over 148 persistent CTAs) -- at the BASELINE channel count and at 64,
    the channel count of every shipped HALO config (core/configs/defaults.py:14), where the kernel runs its 640-thread form."""
    N, O, H, W = 8, 19, 640, 1280
    P, A = synth.head_params(O, C, seed=0, dtype=torch.float64)
    u = torch.empty((N, C, H, W), dtype=torch.float32, device=DEV)
    for i in range(N):
        u[i] = synth.image_features(i, C, H, W, device=DEV)
    Pd, Ad = P.to(DEV), A.to(DEV)
    tc = halo_b200.head_forward(u, Pd, Ad, 1.0, want_logits=True, want_radius=True, want_pixunc=True, want_label=True,
                                want_stats=True)
    cc = halo_b200.head_forward(u, Pd, Ad, 1.0, want_logits=True, want_radius=True, want_pixunc=True, want_label=True,
                                want_stats=True, tensor_cores=False)
    again = halo_b200.head_forward(u, Pd, Ad, 1.0, want_logits=True, want_radius=True)
    assert torch.equal(again["logits"], tc["logits"]) and torch.equal(again["radius"], tc["radius"])  # run-to-run bitwise
    assert rel_err(tc["logits"], cc["logits"]) <= TOL
    assert rel_err(tc["radius"], cc["radius"]) <= 2e-6
    assert rel_err(tc["pixunc"], cc["pixunc"]) <= 1e-4
    assert (tc["label"] == cc["label"]).float().mean().item() >= 0.999
    assert torch.allclose(tc["stats"][:, :2], cc["stats"][:, :2])
    log19 = torch.log(torch.tensor(19.0, dtype=torch.float64))
    for i in range(N):   # the oracle image by image (its fp64 embedding is 1.7 GB per image)
        logits, _, rad = ohead.head_forward(u[i:i + 1].cpu(), P, A, 1.0)
        assert rel_err(tc["logits"][i:i + 1], logits) <= TOL, i
        assert rel_err(cc["logits"][i:i + 1], logits) <= TOL, i
        assert rel_err(tc["radius"][i:i + 1], rad) <= TOL, i
        p = torch.softmax(logits.double(), dim=1)
        ent = torch.sum(-p * torch.log(p + 1e-6), dim=1) / log19
        assert rel_err(tc["pixunc"][i:i + 1], ent) <= TOL, i
        assert (tc["label"][i:i + 1].cpu().long() == p.argmax(dim=1)).float().mean().item() >= 0.999, i
        r32 = rad.float()
        assert abs(float(tc["stats"][i, 0]) - float(r32.min())) <= 1e-5 * float(r32.max())
        assert abs(float(tc["stats"][i, 1]) - float(r32.max())) <= 1e-5 * float(r32.max())
        del logits, rad, p, ent


def test_head_tensor_core_boundary_regime_full_size():
    """Two full-size images of near-boundary features (sigma 0.3: MLR-projection branch, lambda ~ 1000) -- the
    ill-conditioned regime of profiles/r1_tc_numerics.md -- at the BASELINE tile count."""
    N, C, O, H, W = 2, 256, 19, 640, 1280
    P, A = synth.head_params(O, C, seed=4, dtype=torch.float64)
    u = torch.stack([synth.image_features(50 + i, C, H, W, sigma=0.3, device=DEV) for i in range(N)])
    tc = halo_b200.head_forward(u, P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True)
    for i in range(N):
        logits, _, rad = ohead.head_forward(u[i:i + 1].cpu(), P, A, 1.0)
        assert rel_err(tc["logits"][i:i + 1], logits) <= TOL, i
        assert rel_err(tc["radius"][i:i + 1], rad) <= TOL, i
