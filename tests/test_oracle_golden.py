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


def _hfr_modules(g, dtype=torch.float64):
    """conv_reduce / wn_mlp rebuilt from the golden parameters of the reference's own classifier instance."""
    import torch.nn as nn

    Wr = t(g["hfr_Wr"])
    C, Cin = Wr.shape[0], Wr.shape[1]
    conv = nn.Conv2d(Cin, C, kernel_size=1)
    mlp = nn.Sequential(nn.Linear(C, C), nn.BatchNorm1d(C, eps=float(g["hfr_bn_eps"])), nn.ReLU(), nn.Linear(C, C))
    with torch.no_grad():
        conv.weight.copy_(Wr); conv.bias.copy_(t(g["hfr_br"]))
        mlp[0].weight.copy_(t(g["hfr_W1"])); mlp[0].bias.copy_(t(g["hfr_b1"]))
        mlp[1].weight.copy_(t(g["hfr_bn_w"])); mlp[1].bias.copy_(t(g["hfr_bn_b"]))
        mlp[1].running_mean.copy_(t(g["hfr_bn_mean"])); mlp[1].running_var.copy_(t(g["hfr_bn_var"]))
        mlp[3].weight.copy_(t(g["hfr_W2"])); mlp[3].bias.copy_(t(g["hfr_b2"]))
    return conv.to(dtype).eval(), mlp.to(dtype).eval()


def test_hfr_and_head_call_site_match_the_reference_classifier(golden):
    """oracle/hfr.py + oracle/head.py against what the reference's REAL classifier class returned
    (DepthwiseSeparableASPP_Hyper.forward, core/models/classifier.py:526-558, frozen by tests/golden/make_golden.py)."""
    from oracle import hfr as ohfr

    g = golden["train"]
    conv, mlp = _hfr_modules(g, torch.float32)      # the reference runs this block in float32
    with torch.no_grad():
        z = ohfr.reduce_hfr(t(g["hfr_f"]), conv, mlp)
    logits, x, _ = ohead.head_forward(z, t(g["hfr_P"]), t(g["hfr_A"]), float(g["hfr_c"]))
    assert torch.equal(x, t(g["hfr_emb"]))
    ref = t(g["hfr_logits"])
    assert (logits - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def _hfr_train_modules(g, dtype=torch.float32):
    """conv_reduce / wn_mlp of the reference's TRAINING-mode classifier instance, as they were before its step."""
    import torch.nn as nn

    Wr = t(g["hfrt_Wr"])
    C, Cin = Wr.shape[0], Wr.shape[1]
    conv = nn.Conv2d(Cin, C, kernel_size=1)
    mlp = nn.Sequential(nn.Linear(C, C), nn.BatchNorm1d(C, eps=float(g["hfrt_bn_eps"]), momentum=float(g["hfrt_bn_momentum"])),
                        nn.ReLU(), nn.Linear(C, C))
    with torch.no_grad():
        conv.weight.copy_(Wr); conv.bias.copy_(t(g["hfrt_br"]))
        mlp[0].weight.copy_(t(g["hfrt_W1"])); mlp[0].bias.copy_(t(g["hfrt_b1"]))
        mlp[1].weight.copy_(t(g["hfrt_bn_w"])); mlp[1].bias.copy_(t(g["hfrt_bn_b"]))
        mlp[1].running_mean.copy_(t(g["hfrt_bn_mean0"])); mlp[1].running_var.copy_(t(g["hfrt_bn_var0"]))
        mlp[3].weight.copy_(t(g["hfrt_W2"])); mlp[3].bias.copy_(t(g["hfrt_b2"]))
    return conv.to(dtype).train(), mlp.to(dtype).train()


HFRT_PARAMS = (("dWr", lambda c, m: c.weight), ("dbr", lambda c, m: c.bias), ("dW1", lambda c, m: m[0].weight),
               ("db1", lambda c, m: m[0].bias), ("dbn_w", lambda c, m: m[1].weight), ("dbn_b", lambda c, m: m[1].bias),
               ("dW2", lambda c, m: m[3].weight), ("db2", lambda c, m: m[3].bias))


def check_hfr_train_grads(g, conv, mlp, df, dP, dA, tol):
    """Gradients of one training step of the block against the golden vectors of the reference's own classifier class; every
    parameter gradient is measured against the larger of its own magnitude and 1e-2 of the largest one (the bias in front of
    a batch-statistics BatchNorm has an exactly zero gradient)."""
    refs = {k: t(g["hfrt_" + k]) for k, _ in HFRT_PARAMS}
    floor = 1e-2 * max(float(v.abs().max()) for v in refs.values())
    for k, get in HFRT_PARAMS:
        got = get(conv, mlp).grad.detach().cpu().double().reshape(refs[k].shape)
        err = float((got - refs[k].double()).abs().max()) / max(float(refs[k].abs().max()), floor)
        assert err <= tol, (k, err)
    for got, key in ((df, "hfrt_df"), (dP, "hfrt_dP"), (dA, "hfrt_dA")):
        ref = t(g[key]).double()
        assert float((got.detach().cpu().double() - ref).abs().max()) <= tol * float(ref.abs().max()), key
    bn = mlp[1]
    for got, key in ((bn.running_mean, "hfrt_bn_mean1"), (bn.running_var, "hfrt_bn_var1")):
        ref = t(g[key]).double()
        assert float((got.detach().cpu().double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max()), key


def test_hfr_training_step_matches_the_reference_classifier(golden):
    """oracle/hfr.py + oracle/head.py in TRAINING mode against one forward + backward of the reference's real classifier
    class (DepthwiseSeparableASPP_Hyper, classifier.py:526-554; BatchNorm1d on batch statistics): logits, the gradient that
    reaches the decoder features, every parameter gradient of conv_reduce / wn_mlp / conv_seg, and the running statistics."""
    from oracle import hfr as ohfr

    g = golden["train"]
    conv, mlp = _hfr_train_modules(g)
    f = t(g["hfrt_f"]).clone().requires_grad_(True)
    P, A = t(g["hfrt_P"]).clone().requires_grad_(True), t(g["hfrt_A"]).clone().requires_grad_(True)
    z = ohfr.reduce_hfr(f, conv, mlp)
    logits = ohead.mlr_logits(ohead.expmap(z, float(g["hfrt_c"]), dim=1).double(), P, A, float(g["hfrt_c"])).float()
    ref = t(g["hfrt_logits"])
    assert (logits - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()
    (logits * t(g["hfrt_R"])).sum().backward()
    check_hfr_train_grads(g, conv, mlp, f.grad, P.grad, A.grad, 1e-5)


def test_loss_oracle_matches_the_reference_sequence(golden):
    """oracle/loss.py against the learner's sequence run with the reference's own NegativeLearningLoss class."""
    from oracle import loss as oloss

    g = golden["train"]
    for i in range(int(g["n_loss_cases"])):
        tag = "loss%d_" % i
        size = tuple(int(v) for v in g[tag + "size"])
        loss, sup, neg, grad = oloss.seg_loss(t(g[tag + "logits"]), t(g[tag + "labels"]).long(), size, neg_weight=float(g[tag + "weight"]))
        assert abs(float(loss) - float(g[tag + "loss"])) <= 1e-12 * max(1.0, abs(float(g[tag + "loss"])))
        assert abs(float(sup) - float(g[tag + "sup"])) <= 1e-12 and abs(float(neg) - float(g[tag + "neg"])) <= 1e-12
        ref = t(g[tag + "grad"])
        assert (grad - ref).abs().max().item() <= 1e-12 * ref.abs().max().item()
