"""GPU: the whole acquisition step (K1 -> K2 -> K3, halo_b200.acquire_batch) against the oracle's per-image
restatement of RegionSelection (core/active/build.py:137-160), plus size-independent properties at the
BASELINE image size."""
import numpy as np
import pytest
import torch

import halo_b200
from halo_b200 import synth
from oracle import acquire as oacquire
from tests.util import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def oracle_round(feat, gt, P, A, cfg, active=None):
    outs = []
    for i in range(feat.shape[0]):
        H, W = feat.shape[-2:]
        act = torch.zeros((H, W), dtype=torch.bool) if active is None else active[i].bool().clone()
        outs.append(oacquire.acquire_image(
            feat[i:i + 1], P.double(), A.double(), gt[i].long(), act, torch.zeros((H, W), dtype=torch.bool),
            torch.full((H, W), 255, dtype=torch.int64), c=cfg.curvature, radius_k=cfg.radius_k,
            mask_radius_k=cfg.mask_radius_k, budget=cfg.budget, n_rounds=cfg.n_rounds, unc_type=cfg.uncertainty,
            pur_type=cfg.purity, normalize=cfg.normalize, K=cfg.K, fast_select=True))
    return outs


@pytest.mark.parametrize("name,cfg,shape", [
    ("cfg1", halo_b200.AcquisitionConfig(budget=0.022), (2, 256, 19, 160, 320)),
    ("halo_default_round", halo_b200.AcquisitionConfig(budget=0.05, n_rounds=5), (2, 64, 19, 128, 256)),
    ("ripu", halo_b200.AcquisitionConfig(budget=0.022, purity="ripu", normalize=False), (2, 64, 19, 96, 160)),
    ("synthia_5x5", halo_b200.AcquisitionConfig(num_classes=16, radius_k=2, budget=0.022), (2, 64, 16, 96, 160)),
    ("pixel_mode", halo_b200.AcquisitionConfig(radius_k=0, mask_radius_k=0, budget=0.01), (1, 64, 19, 96, 160)),
    # the reference's default purity (defaults.py:68): K = 100 radius bins, 3x3 purity window, bins formed in fp64
    ("hyper_default", halo_b200.AcquisitionConfig(budget=0.022, purity="hyper"), (2, 64, 19, 96, 160)),
    ("hyper_c256", halo_b200.AcquisitionConfig(budget=0.022, purity="hyper", normalize=False, K=50), (1, 256, 19, 128, 256)),
])
def test_acquire_batch_matches_oracle(name, cfg, shape):
    n, C, O, H, W = shape
    P, A = synth.head_params(O, C, seed=0)
    b = synth.batch(0, n, C, O, H, W)
    ref = oracle_round(b["feat"], b["gt"], P, A, cfg)
    d = {k: v.to(DEV) for k, v in b.items()}
    res = halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"],
                                  d["active_mask"], want_score=True, want_picks=True)
    for i in range(n):
        assert rel_err(res["score"][i], ref[i]["score"]) <= (TOL if cfg.purity != "ripu" else 1e-3), name
        agree = (d["active_mask"][i].cpu().numpy() == ref[i]["active_mask"].numpy().astype(np.uint8)).mean()
        assert agree >= 0.999, (name, agree)   # north_star: >= 99.9 % mask agreement end to end
        assert int(res["n_picked"][i]) == len(ref[i]["picks"])


def test_selection_bit_exact_on_reference_scores():
    """north_star: the selection mask is bit-exact when the kernel is fed the reference's (fp64) score tensor."""
    C, O, H, W = 64, 19, 128, 256
    cfg = halo_b200.AcquisitionConfig(budget=0.05)
    P, A = synth.head_params(O, C, seed=0)
    b = synth.batch(0, 1, C, O, H, W)
    ref = oracle_round(b["feat"], b["gt"], P, A, cfg)[0]
    score = ref["score"].clone().to(DEV)
    assert score.dtype == torch.float64
    active = torch.zeros((H, W), dtype=torch.bool)
    selected = torch.zeros((H, W), dtype=torch.bool)
    mask = torch.full((H, W), 255, dtype=torch.int64, device=DEV)
    halo_b200.select_pixels_to_label(score, ref["n_regions"], cfg.radius_k, cfg.mask_radius_k, active, selected, mask,
                                     b["gt"][0].long().to(DEV))
    assert torch.equal(mask.cpu(), ref["active_mask"])
    assert torch.equal(active, ref["active"]) and torch.equal(selected, ref["selected"])
    assert torch.equal(score.cpu(), ref["score_after"])


def test_second_round_respects_existing_labels():
    C, O, H, W = 64, 19, 96, 160
    cfg = halo_b200.AcquisitionConfig(budget=0.05, n_rounds=5)
    P, A = synth.head_params(O, C, seed=0)
    d = {k: v.to(DEV) for k, v in synth.batch(0, 2, C, O, H, W).items()}
    halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"], d["active_mask"])
    act1, sel1 = d["active"].clone(), d["selected"].clone()
    feat2 = torch.stack([synth.image_features(100 + i, C, H, W).to(DEV) for i in range(2)])
    res = halo_b200.acquire_batch(feat2, P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"], d["active_mask"],
                                  want_picks=True)
    picks = res["picks"]
    for i in range(2):
        p = picks[i][: int(res["n_picked"][i])].long()
        assert not act1[i].flatten()[p].any()              # never re-pick an already-active pixel
        assert (d["active"][i] >= act1[i]).all() and (d["selected"][i] >= sel1[i]).all()


@pytest.mark.parametrize("name,cfg,O,expect", [
    ("cfg2_gtav_cityscapes", halo_b200.AcquisitionConfig(budget=0.05), 19, 4552),
    ("cfg2_reference_default_round", halo_b200.AcquisitionConfig(budget=0.05, n_rounds=5), 19, 911),
    ("cfg4_synthia_5x5", halo_b200.AcquisitionConfig(num_classes=16, radius_k=2, budget=0.022), 16, 721),
])
def test_full_size_properties(name, cfg, O, expect):
    """BASELINE config-2 / config-4 image size: properties that do not need the oracle."""
    C, H, W = 256, 640, 1280
    P, A = synth.head_params(O, C, seed=0)
    d = synth.batch(0, 2, C, O, H, W, device=DEV)
    res = halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"],
                                  d["active_mask"], want_picks=True)
    n_regions = cfg.regions_per_image(H, W)
    assert n_regions == expect
    ar, mr = cfg.radius_k, cfg.mask_radius_k
    for i in range(2):
        k = int(res["n_picked"][i])
        assert k == n_regions
        p = res["picks"][i][:k].long()
        assert p.unique().numel() == k
        hh, ww = p // W, p % W
        # no two picks within the suppression radius: dilating picks by m never covers another pick
        grid = torch.zeros((H, W), dtype=torch.int32, device=DEV)
        grid[hh, ww] = 1
        cnt = torch.nn.functional.conv2d(grid[None, None].float(), torch.ones(1, 1, 2 * mr + 1, 2 * mr + 1, device=DEV),
                                         padding=mr)[0, 0]
        assert int(cnt[hh, ww].max()) == 1
        sel = d["selected"][i].bool()
        assert torch.equal(d["active_mask"][i][sel], d["gt"][i][sel])
        assert (d["active_mask"][i][~sel] == 255).all()
        assert sel.sum().item() <= (2 * ar + 1) ** 2 * k and d["active"][i].sum().item() <= (2 * max(mr, 0) + 1) ** 2 * k
        if mr >= ar:
            assert (d["active"][i].bool() | ~sel).all()    # selected implies active
    # idempotence of a zero budget
    before = d["active_mask"].clone()
    zero = halo_b200.AcquisitionConfig(budget=0.0)
    r0 = halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), zero, d["gt"], d["active"], d["selected"], d["active_mask"])
    assert r0["n_picked"].sum().item() == 0 and torch.equal(before, d["active_mask"])


def test_runs_on_a_side_stream_and_accepts_strided_inputs():
    """Every entry point enqueues on the caller's current stream; non-contiguous features are made contiguous."""
    C, O, H, W = 64, 19, 64, 96
    cfg = halo_b200.AcquisitionConfig(budget=0.022)
    P, A = synth.head_params(O, C, seed=0)
    b = synth.batch(0, 2, C, O, H, W)
    d = {k: v.to(DEV) for k, v in b.items()}
    ref = halo_b200.acquire_batch(d["feat"], P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"].clone(), d["selected"].clone(),
                                  d["active_mask"].clone(), want_score=True)
    side = torch.cuda.Stream()
    feat_cl = d["feat"].to(memory_format=torch.channels_last)          # NHWC strides: must not be mis-read as NCHW
    assert not feat_cl.is_contiguous()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        out = halo_b200.acquire_batch(feat_cl, P.to(DEV), A.to(DEV), cfg, d["gt"], d["active"], d["selected"],
                                      d["active_mask"], want_score=True)
    side.synchronize()
    assert torch.equal(out["score"], ref["score"]) and torch.equal(out["n_picked"], ref["n_picked"])



def test_round_delta_reproduces_the_masks_two_rounds():
    """SURVEY 8e, compact exchange: picks + window labels of a round, replayed onto a replica, give the mask planes the
    round wrote -- bit-exact, over two consecutive rounds (the second one starts from the first one's state)."""
    from halo_b200 import pool
    from oracle import delta as odelta

    C, O, H, W, B = 32, 19, 96, 160, 3
    cfg = halo_b200.AcquisitionConfig(num_classes=O, radius_k=1, budget=0.02, n_rounds=2)
    P, A = synth.head_params(O, C, seed=1, device=DEV)
    d = synth.batch(5, 5 + B, C, O, H, W, device=DEV)
    replica = torch.full((B + 2, H, W), 255, dtype=torch.uint8, device=DEV)    # a pool of B+2 images, this shard owns 1..B
    row_image = torch.arange(1, B + 1, dtype=torch.int32, device=DEV)
    for rnd in range(2):
        res = halo_b200.acquire_batch(d["feat"], P, A, cfg, d["gt"], d["active"], d["selected"], d["active_mask"],
                                      want_picks=True)
        lab = pool.pack_round_delta(res["picks"], res["n_picked"], d["gt"], cfg.radius_k)
        # the CUDA kernels agree with the oracle's restatement (oracle/delta.py)
        lab_cpu = odelta.pack_round_delta(res["picks"].cpu(), res["n_picked"].cpu(), d["gt"].cpu(), cfg.radius_k)
        assert torch.equal(lab.cpu(), lab_cpu)
        ref = odelta.apply_round_delta(replica.cpu(), row_image.cpu(), res["picks"].cpu(), res["n_picked"].cpu(), lab_cpu,
                                       cfg.radius_k)
        pool.apply_round_delta(replica, row_image, res["picks"], res["n_picked"], lab, cfg.radius_k)
        assert torch.equal(replica[1:B + 1], d["active_mask"]), rnd
        assert torch.equal(replica.cpu(), ref), rnd
        assert bool((replica[0] == 255).all()) and bool((replica[B + 1] == 255).all())
        assert int(res["n_picked"].min()) > 0
    out = pool.gather_round_delta(res["n_picked"], res["picks"], d["gt"], torch.full((B, H, W), 255, dtype=torch.uint8, device=DEV),
                                  B, cfg.radius_k)
    assert torch.equal(out["n_picked"], res["n_picked"])
