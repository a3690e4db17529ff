"""CPU: the oracle restatement against the golden vectors frozen from the reference's own modules
(tests/golden/make_golden.py), and -- when the reference checkout is present -- against the live reference."""
import numpy as np
import pytest
import torch

from oracle import head as ohead
from oracle import ref_import
from oracle import score as oscore
from oracle import select as oselect
from tests.golden.make_golden import SCORE_COMBOS, SELECT_CASES


def t(a):
    return torch.from_numpy(np.asarray(a))


def test_head_matches_golden(golden):
    g = golden["head"]
    for k in range(int(g["n_cases"])):
        tag = "h%d_" % k
        c = float(g[tag + "c"])
        logits, x, rad = ohead.head_forward(t(g[tag + "u"]), t(g[tag + "P"]), t(g[tag + "A"]), c)
        assert torch.equal(x, t(g[tag + "x"]))
        assert torch.equal(rad, t(g[tag + "radius"]))
        ref = t(g[tag + "logits"])
        assert (logits - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def test_head_grads_match_golden(golden):
    g = golden["head"]
    for k in range(8):
        tag = "h%d_" % k
        c = float(g[tag + "c"])
        du, dP, dA = ohead.head_grads(t(g[tag + "u"]), t(g[tag + "P"]), t(g[tag + "A"]), t(g[tag + "dlogits"]), c)
        for got, name in ((du, "du"), (dP, "dP"), (dA, "dA")):
            ref = t(g[tag + name])
            assert got.dtype == ref.dtype
            assert (got - ref).abs().max().item() <= 1e-9 * max(ref.abs().max().item(), 1e-30)


def test_score_matches_golden(golden):
    g = golden["score"]
    logits, x, gt, c = t(g["logits"]), t(g["x"]), t(g["gt"]), float(g["c"])
    for i, (ctor, size, unc, pur, norm) in enumerate(SCORE_COMBOS):
        s, imp, un = oscore.floating_region_score(
            logits.clone(), decoder_out=x, unc_type=unc, pur_type=pur, normalize=norm, ground_truth=gt,
            in_channels=19, size=size, ctor_purity_type=ctor, K=100, c=c)
        for got, name in ((s, "score"), (imp, "impurity"), (un, "uncertainty")):
            ref = t(g["s%d_%s" % (i, name)])
            assert got.dtype == ref.dtype, (i, name)
            assert torch.equal(torch.isnan(got), torch.isnan(ref))
            assert torch.nan_to_num(got - ref).abs().max().item() <= 1e-6 * max(torch.nan_to_num(ref).abs().max().item(), 1e-30), (i, name)


@pytest.mark.parametrize("impl", ["sequential", "numpy"])
def test_select_matches_golden(golden, impl):
    g = golden["select"]
    for i in range(int(g["n_cases"])):
        n, ar, mr = (int(v) for v in g["k%d_args" % i])
        sc = g["k%d_score_in" % i].copy()
        act = g["k%d_active_in" % i].copy()
        gt = g["k%d_gt" % i].astype(np.int64)
        sel = np.zeros_like(act)
        am = np.full(act.shape, 255, dtype=np.int64)
        if impl == "sequential":
            sc_t, act_t, sel_t, am_t = t(sc), t(act), t(sel), t(am)
            oselect.select_sequential(sc_t, n, ar, mr, act_t, sel_t, am_t, t(gt))
            sc, act, sel, am = sc_t.numpy(), act_t.numpy(), sel_t.numpy(), am_t.numpy()
        else:
            oselect.select_numpy(sc, n, ar, mr, act, sel, am, gt)
        assert np.array_equal(sc, g["k%d_score_out" % i]), i
        assert np.array_equal(act, g["k%d_active_out" % i]), i
        assert np.array_equal(sel, g["k%d_selected_out" % i]), i
        assert np.array_equal(am.astype(np.uint8), g["k%d_mask_out" % i]), i


def test_select_cases_cover_edge_conditions(golden):
    g = golden["select"]
    assert len(SELECT_CASES) == int(g["n_cases"])
    # case 9: everything already active -> no picks; case 2: budget larger than pickable set
    assert not g["k9_selected_out"].any()
    assert np.isneginf(g["k2_score_out"]).all()


@pytest.mark.skipif(not ref_import.available(), reason="reference checkout not present (GPU box)")
def test_oracle_vs_live_reference():
    ref = ref_import.load()
    torch.manual_seed(7)
    C, O, H, W = 40, 19, 18, 26
    for c in (1.0, 0.7):
        ref.cfg.MODEL.CURVATURE = c
        for sigma in (0.05, 0.4, 2.0):
            u = torch.randn(1, C, H, W) * sigma
            mlr = ref.HyperMLR(C, O, c=c)
            mapper = ref.HyperMapper(c=c)
            x_ref = mapper.expmap(u, dim=1)
            out_ref = mlr(x_ref.double()).float()
            logits, x, rad = ohead.head_forward(u, mlr.P_MLR.data, mlr.A_MLR.data, c)
            assert torch.equal(x, x_ref)
            assert (logits - out_ref).abs().max().item() <= 1e-6 * out_ref.abs().max().item()
            gt = torch.randint(0, O, (H, W))
            for ctor, size, unc, pur, norm in SCORE_COMBOS:
                with ref_import.cpu_only():
                    frs = ref.FloatingRegionScore(in_channels=O, size=size, purity_type=ctor, K=100)
                    s_ref, _, _ = frs(out_ref.detach().clone(), decoder_out=x_ref, unc_type=unc, pur_type=pur,
                                      normalize=norm, ground_truth=gt)
                s, _, _ = oscore.floating_region_score(
                    out_ref.detach().clone(), decoder_out=x_ref, unc_type=unc, pur_type=pur, normalize=norm,
                    ground_truth=gt, in_channels=O, size=size, ctor_purity_type=ctor, K=100, c=c)
                assert s.dtype == s_ref.dtype
                assert torch.nan_to_num(s - s_ref).abs().max().item() <= 1e-6 * max(torch.nan_to_num(s_ref).abs().max().item(), 1e-30)
