"""GPU: halo_select_{f32,f64} -- bit-exact against the golden outputs of the reference's
select_pixels_to_label (core/active/build.py:27-64) and against the oracle on larger seeded cases."""
import numpy as np
import pytest
import torch

import halo_b200
from oracle import select as oselect
from tests.golden.make_golden import SELECT_CASES
from tests.util import t

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("i", range(len(SELECT_CASES)))
def test_select_bit_exact_vs_golden(golden, i):
    g = golden["select"]
    n, ar, mr = (int(v) for v in g["k%d_args" % i])
    score = t(g["k%d_score_in" % i].copy()).to(DEV)
    active = t(g["k%d_active_in" % i].copy())          # CPU bool, like the reference's indicator tensors
    selected = torch.zeros_like(active)
    mask = torch.full(active.shape, 255, dtype=torch.int64, device=DEV)
    gt = t(g["k%d_gt" % i].astype(np.int64)).to(DEV)
    s2, a2, sel2, m2 = halo_b200.select_pixels_to_label(score, n, ar, mr, active, selected, mask, gt)
    assert s2 is score and a2 is active and sel2 is selected and m2 is mask  # in place
    assert np.array_equal(score.cpu().numpy(), g["k%d_score_out" % i])
    assert np.array_equal(active.numpy(), g["k%d_active_out" % i])
    assert np.array_equal(selected.numpy(), g["k%d_selected_out" % i])
    assert np.array_equal(mask.cpu().numpy().astype(np.uint8), g["k%d_mask_out" % i])


def run_both(score, act, gt, n, ar, mr):
    """oracle (numpy) and device selection on the same inputs; returns (oracle outs, device outs, picks)."""
    H, W = score.shape
    s_np, a_np = score.numpy().copy(), act.numpy().copy()
    sel_np = np.zeros((H, W), dtype=bool)
    m_np = np.full((H, W), 255, dtype=np.int64)
    _, _, _, _, picks = oselect.select_numpy(s_np, n, ar, mr, a_np, sel_np, m_np, gt.numpy().astype(np.int64))
    sd = score.to(DEV).contiguous().view(1, H, W)
    ad = act.to(DEV).to(torch.uint8).view(1, H, W).contiguous()
    seld = torch.zeros((1, H, W), dtype=torch.uint8, device=DEV)
    md = torch.full((1, H, W), 255, dtype=torch.uint8, device=DEV)
    gd = gt.to(DEV).to(torch.uint8).view(1, H, W).contiguous()
    n_picked, dpicks = halo_b200.select_planes(sd, ad, seld, md, gd, n, ar, mr, want_picks=True)
    return (s_np, a_np, sel_np, m_np, picks), (sd[0].cpu().numpy(), ad[0].cpu().numpy().astype(bool),
                                               seld[0].cpu().numpy().astype(bool), md[0].cpu().numpy()), \
        (int(n_picked[0]), dpicks[0].cpu().numpy())


def check(o, d, p, W):
    s_np, a_np, sel_np, m_np, picks = o
    assert p[0] == len(picks)
    assert [int(v) for v in p[1][:p[0]]] == [h * W + w for h, w in picks]   # same picks in the same order
    assert np.array_equal(d[0], s_np)
    assert np.array_equal(d[1], a_np)
    assert np.array_equal(d[2], sel_np)
    assert np.array_equal(d[3], m_np.astype(np.uint8))


@pytest.mark.parametrize("case", [
    # (H, W, dtype, quant, p_active, n, ar, mr)
    (320, 640, torch.float32, 0, 0.0, 501, 1, 5),        # BASELINE config 1 shape
    (640, 1280, torch.float32, 0, 0.0, 4552, 1, 5),      # BASELINE config 2 shape, 5 % single shot
    (640, 1280, torch.float32, 0, 0.3, 911, 1, 5),       # 1 %/round with a third of the image already labelled
    (640, 1280, torch.float64, 0, 0.0, 911, 1, 5),       # fp64 scores (the reference's radius-mode dtype)
    (640, 1280, torch.float32, 0, 0.0, 18023, 0, 0),     # pixel mode, pure top-k
    (640, 1280, torch.float32, 0, 0.0, 721, 2, 5),       # BASELINE config 4 region variant
    (640, 1280, torch.float32, 16, 0.0, 4552, 1, 5),     # 17 distinct values: massive ties -> index descent
    (200, 300, torch.float32, 1, 0.0, 100000, 1, 2),     # two values, budget exceeds what can be picked
    (200, 300, torch.float64, 3, 0.1, 3000, 2, 7),
    (64, 2000, torch.float32, 0, 0.0, 300, 3, 20),       # wide suppression window (> 32 rows)
    (1024, 2048, torch.float32, 0, 0.0, 2000, 1, 5),     # Cityscapes label size: bitmap in global memory
])
def test_select_bit_exact_vs_oracle(case):
    H, W, dtype, quant, p_act, n, ar, mr = case
    g = torch.Generator().manual_seed(H * 7 + W + n)
    score = torch.rand((H, W), generator=g, dtype=dtype)
    if quant:
        score = (score * quant).round() / quant
    act = torch.rand((H, W), generator=g) < p_act
    gt = torch.randint(0, 19, (H, W), generator=g).to(torch.uint8)
    score[act] = -float("inf")
    o, d, p = run_both(score, act, gt, n, ar, mr)
    check(o, d, p, W)


def test_smooth_score_map_long_dependency_chains():
    """A monotone ramp is the worst case for parallel local-max formulations: every pick depends on the previous."""
    H, W = 96, 700
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    score = -(xx * 1.0 + yy * 0.001)
    o, d, p = run_both(score, torch.zeros(H, W, dtype=torch.bool), torch.zeros(H, W, dtype=torch.uint8), 500, 1, 5)
    check(o, d, p, W)


def test_special_values():
    H, W = 40, 60
    g = torch.Generator().manual_seed(3)
    score = torch.randn((H, W), generator=g)
    score[5, 7] = float("inf")
    score[9, 9] = -0.0
    score[9, 30] = 0.0
    score[20:30, 20:40] = -float("inf")
    o, d, p = run_both(score, torch.zeros(H, W, dtype=torch.bool), torch.zeros(H, W, dtype=torch.uint8), 60, 1, 3)
    check(o, d, p, W)


def test_batched_images_are_independent():
    N, H, W = 5, 64, 96
    g = torch.Generator().manual_seed(9)
    score = torch.rand((N, H, W), generator=g)
    gt = torch.randint(0, 19, (N, H, W), generator=g).to(torch.uint8)
    sd = score.to(DEV)
    ad = torch.zeros((N, H, W), dtype=torch.uint8, device=DEV)
    seld, md = torch.zeros_like(ad), torch.full_like(ad, 255)
    n_picked, _ = halo_b200.select_planes(sd, ad, seld, md, gt.to(DEV), 25, 1, 4)
    for i in range(N):
        s_np = score[i].numpy().copy()
        a_np, sel_np = np.zeros((H, W), bool), np.zeros((H, W), bool)
        m_np = np.full((H, W), 255, np.int64)
        oselect.select_numpy(s_np, 25, 1, 4, a_np, sel_np, m_np, gt[i].numpy().astype(np.int64))
        assert np.array_equal(md[i].cpu().numpy(), m_np.astype(np.uint8))
        assert np.array_equal(sd[i].cpu().numpy(), s_np)
    assert n_picked.tolist() == [25] * N
