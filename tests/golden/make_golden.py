"""Freeze golden vectors from the reference's OWN modules (run in the container that has the
reference checkout; the fixtures travel, the checkout does not).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz

Every array below is produced by code imported UNCHANGED from the reference
(core/utils/hyperbolic.py, core/active/floating_region.py, core/active/build.py) behind the
stubs of oracle/ref_import.py -- never by the oracle restatement or by the CUDA path.
Inputs are stored next to the outputs so the fixtures do not depend on RNG stability.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import ref_import  # noqa: E402

SCORE_COMBOS = [
    # (ctor_purity, size, unc_type, pur_type, normalize)
    ("radius", 3, "entropy", "radius", True),      # the shipped HALO config (configs/gtav/source_free.yaml:22-25)
    ("radius", 3, "entropy", "radius", False),
    ("radius", 5, "entropy", "radius", True),
    ("radius", 1, "entropy", "radius", True),      # pixel mode, RADIUS_K=0
    ("radius", 3, "pixel_entropy", "radius", True),
    ("radius", 3, "none", "radius", False),
    ("radius", 3, "hyperbolic", "radius", False),  # unknown string -> zeros (floating_region.py:85-90)
    ("radius", 3, "oracle_acc", "radius", True),
    ("ripu", 3, "entropy", "ripu", False),          # RIPU baseline
    ("ripu", 3, "entropy", "ripu", True),
    ("ripu", 5, "entropy", "ripu", False),
    ("ripu", 3, "entropy", "oracle_ripu", False),
    ("ripu", 3, "entropy", "none", False),
    ("ripu", 3, "entropy", "euc_norm", True),
    ("hyper", 3, "entropy", "hyper", True),         # defaults.py:68 PURITY="hyper"
    ("hyper", 5, "entropy", "hyper", False),        # purity window stays 3x3 (floating_region.py:54-55)
    ("hyper", 3, "pixel_entropy", "hyper", True),
]

SELECT_CASES = [
    # (seed, H, W, dtype, quantize, p_active, n_regions, active_radius, mask_radius)
    (0, 37, 53, "f32", 0, 0.2, 10, 1, 5),
    (1, 37, 53, "f64", 0, 0.2, 50, 0, 0),
    (2, 37, 53, "f32", 8, 0.2, 5000, 2, 3),     # heavy ties, budget > pickable (loop exits on -inf)
    (3, 37, 53, "f64", 8, 0.0, 30, 1, 2),
    (4, 64, 96, "f32", 0, 0.0, 40, 1, 5),
    (5, 64, 96, "f32", 2, 0.5, 400, 1, 1),
    (6, 16, 16, "f32", 1, 0.0, 100, 0, 3),      # all scores in {0,1}: pure tie-break lattice
    (7, 48, 80, "f64", 0, 0.1, 64, 2, 5),
    (8, 5, 300, "f32", 4, 0.0, 50, 1, 5),       # window taller than the image
    (9, 40, 40, "f32", 0, 1.0, 5, 1, 5),        # everything already active -> zero picks
]


def head_fixtures(ref):
    out = {}
    C, O, H, W = 48, 19, 12, 20
    k = 0
    for c in (1.0, 0.5):
        for sigma in (0.01, 0.1, 0.3, 1.0):
            g = torch.Generator().manual_seed(100 + k)
            u = torch.randn((2, C, H, W), generator=g) * sigma
            torch.manual_seed(200 + k)
            mlr = ref.HyperMLR(C, O, c=c)
            mapper = ref.HyperMapper(c=c)
            u_req = u.clone().requires_grad_(True)
            x = mapper.expmap(u_req, dim=1)                    # classifier.py:553
            logits = mlr(x.double()).float()                   # classifier.py:554
            rad = mapper.poincare_distance_origin(x, dim=1)    # floating_region.py:188
            dlog = torch.randn(logits.shape, generator=g) * 1e-3
            du, dP, dA = torch.autograd.grad(logits, (u_req, mlr.P_MLR, mlr.A_MLR), grad_outputs=dlog)
            tag = "h%d" % k
            out[tag + "_c"] = np.float64(c)
            out[tag + "_u"] = u.numpy()
            out[tag + "_P"] = mlr.P_MLR.detach().numpy()
            out[tag + "_A"] = mlr.A_MLR.detach().numpy()
            out[tag + "_x"] = x.detach().numpy()
            out[tag + "_logits"] = logits.detach().numpy()
            out[tag + "_radius"] = rad.detach().numpy()
            out[tag + "_dlogits"] = dlog.numpy()
            out[tag + "_du"] = du.numpy()
            out[tag + "_dP"] = dP.numpy()
            out[tag + "_dA"] = dA.numpy()
            k += 1
    # O=16 (SYNTHIA-shaped) and a P with ||p|| > 1/sqrt(c) (B_k <= 0 branch)
    g = torch.Generator().manual_seed(300)
    u = torch.randn((1, 32, 9, 17), generator=g) * 0.2
    torch.manual_seed(301)
    mlr = ref.HyperMLR(32, 16, c=1.0)
    with torch.no_grad():
        mlr.P_MLR[3] *= 8.0
        mlr.P_MLR[7] *= 3.5
    mapper = ref.HyperMapper(c=1.0)
    x = mapper.expmap(u, dim=1)
    out["h8_c"] = np.float64(1.0)
    out["h8_u"] = u.numpy()
    out["h8_P"] = mlr.P_MLR.detach().numpy()
    out["h8_A"] = mlr.A_MLR.detach().numpy()
    out["h8_x"] = x.numpy()
    out["h8_logits"] = mlr(x.double()).float().detach().numpy()
    out["h8_radius"] = mapper.poincare_distance_origin(x, dim=1).numpy()
    out["n_cases"] = np.int64(9)
    return out


def score_fixtures(ref):
    out = {}
    C, O, H, W = 32, 19, 20, 28
    c = 1.0
    ref.cfg.MODEL.CURVATURE = c
    g = torch.Generator().manual_seed(400)
    u = torch.randn((1, C, H, W), generator=g) * 0.1
    torch.manual_seed(401)
    mlr = ref.HyperMLR(C, O, c=c)
    mapper = ref.HyperMapper(c=c)
    x = mapper.expmap(u, dim=1)
    logits = mlr(x.double()).float().detach()
    gt = torch.randint(0, O, (H, W), generator=g)
    gt[torch.rand((H, W), generator=g) < 0.05] = 255
    out["u"], out["P"], out["A"] = u.numpy(), mlr.P_MLR.detach().numpy(), mlr.A_MLR.detach().numpy()
    out["x"], out["logits"], out["gt"], out["c"] = x.numpy(), logits.numpy(), gt.numpy(), np.float64(c)
    for i, (ctor, size, unc, pur, norm) in enumerate(SCORE_COMBOS):
        with ref_import.cpu_only():
            frs = ref.FloatingRegionScore(in_channels=O, size=size, purity_type=ctor, K=100)
            s, imp, un = frs(logits.clone(), decoder_out=x, unc_type=unc, pur_type=pur, normalize=norm,
                             ground_truth=gt)
        out["s%d_score" % i] = s.numpy()
        out["s%d_impurity" % i] = imp.numpy()
        out["s%d_uncertainty" % i] = un.numpy()
    return out


def select_inputs(seed, H, W, dtype, quant, p_active):
    g = torch.Generator().manual_seed(500 + seed)
    sc = torch.rand((H, W), generator=g, dtype=torch.float64 if dtype == "f64" else torch.float32)
    if quant:
        sc = (sc * quant).round() / quant
    act = torch.rand((H, W), generator=g) < p_active
    gt = torch.randint(0, 19, (H, W), generator=g)
    gt[torch.rand((H, W), generator=g) < 0.05] = 255
    sc[act] = -float("inf")
    sel = torch.zeros((H, W), dtype=torch.bool)
    am = torch.full((H, W), 255, dtype=torch.int64)
    return sc, act, sel, am, gt


def select_fixtures(ref):
    out = {}
    for i, (seed, H, W, dtype, quant, p_act, n, ar, mr) in enumerate(SELECT_CASES):
        sc, act, sel, am, gt = select_inputs(seed, H, W, dtype, quant, p_act)
        out["k%d_score_in" % i] = sc.numpy().copy()
        out["k%d_active_in" % i] = act.numpy().copy()
        out["k%d_gt" % i] = gt.numpy().astype(np.uint8)
        out["k%d_args" % i] = np.array([n, ar, mr], dtype=np.int64)
        s2, a2, sel2, am2 = ref.select_pixels_to_label(sc, n, ar, mr, act, sel, am, gt)  # build.py:27-64
        out["k%d_score_out" % i] = s2.numpy()
        out["k%d_active_out" % i] = a2.numpy()
        out["k%d_selected_out" % i] = sel2.numpy()
        out["k%d_mask_out" % i] = am2.numpy().astype(np.uint8)
    out["n_cases"] = np.int64(len(SELECT_CASES))
    return out


def train_fixtures():
    """Rows next to the hot path, from the reference's own classes:
    * `DepthwiseSeparableASPP_Hyper` (core/models/classifier.py:388-558) built small, in eval mode, run end to end on the
      CPU: the input of its conv_reduce (captured by a hook), the parameters of conv_reduce / wn_mlp / conv_seg, and what
      its forward returns (logits, float64 embedding) -- pins conv_reduce + HFR (:526-550) and the head call site (:552-554)
      under the reference's real classifier class;
    * the learner's loss sequence (core/train_learners.py:343-356) with the reference's NegativeLearningLoss
      (core/loss/negative_learning_loss.py) in float64, and its gradient w.r.t. the low-resolution logits."""
    import torch.nn as nn
    import torch.nn.functional as F

    side = ref_import.load_training_side()
    out = {}
    torch.manual_seed(11)
    C, O = 32, 19
    with ref_import.cpu_only():
        clf = side.classifier.DepthwiseSeparableASPP_Hyper(inplanes=48, dilation_series=[1, 2], padding_series=[1, 2],
                                                           num_classes=O, norm_layer=nn.BatchNorm2d, reduced_channels=C, hfr=True)
    with torch.no_grad():   # statistics / parameters as after some training, not the initial identity
        for mod in clf.modules():
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    clf.eval()
    captured = {}
    clf.conv_reduce.register_forward_hook(lambda m, i, o: captured.update(f=i[0].detach().clone()))
    x = {"out": torch.randn(2, 48, 9, 12), "low": torch.randn(2, 256, 18, 24)}
    with torch.no_grad(), ref_import.cpu_only():
        logits, emb = clf(x, size=None)
    out["hfr_f"] = captured["f"].numpy()
    out["hfr_Wr"] = clf.conv_reduce.weight.detach().numpy()
    out["hfr_br"] = clf.conv_reduce.bias.detach().numpy()
    lin1, bn, lin2 = clf.wn_mlp[0], clf.wn_mlp[1], clf.wn_mlp[3]
    for name, t_ in (("W1", lin1.weight), ("b1", lin1.bias), ("bn_w", bn.weight), ("bn_b", bn.bias), ("bn_mean", bn.running_mean),
                     ("bn_var", bn.running_var), ("W2", lin2.weight), ("b2", lin2.bias), ("P", clf.conv_seg.P_MLR),
                     ("A", clf.conv_seg.A_MLR)):
        out["hfr_" + name] = t_.detach().numpy()
    out["hfr_bn_eps"] = np.float64(bn.eps)
    out["hfr_c"] = np.float64(clf.conv_seg.c)
    out["hfr_logits"] = logits.numpy()
    out["hfr_emb"] = emb.numpy()
    # the same classifier class in TRAINING mode (BatchNorm on batch statistics), one forward + backward on the CPU: what the
    # block conv_reduce + HFR + head (:526-554) receives (input of conv_reduce), returns (logits) and back-propagates
    # (gradient w.r.t. that input, gradients of every parameter of conv_reduce / wn_mlp / conv_seg, BatchNorm1d bookkeeping)
    torch.manual_seed(13)
    with ref_import.cpu_only():
        clf_t = side.classifier.DepthwiseSeparableASPP_Hyper(inplanes=48, dilation_series=[1, 2], padding_series=[1, 2],
                                                             num_classes=O, norm_layer=nn.BatchNorm2d, reduced_channels=C, hfr=True)
    with torch.no_grad():
        for mod in clf_t.modules():
            if isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d)):
                mod.running_mean.uniform_(-0.2, 0.2)
                mod.running_var.uniform_(0.5, 1.5)
                mod.weight.uniform_(0.5, 1.5)
                mod.bias.uniform_(-0.2, 0.2)
    clf_t.train()
    lin1, bn, lin2 = clf_t.wn_mlp[0], clf_t.wn_mlp[1], clf_t.wn_mlp[3]
    named = (("Wr", clf_t.conv_reduce.weight), ("br", clf_t.conv_reduce.bias), ("W1", lin1.weight), ("b1", lin1.bias),
             ("bn_w", bn.weight), ("bn_b", bn.bias), ("W2", lin2.weight), ("b2", lin2.bias), ("P", clf_t.conv_seg.P_MLR),
             ("A", clf_t.conv_seg.A_MLR))
    for name, t_ in named:
        out["hfrt_" + name] = t_.detach().numpy().copy()
    out["hfrt_bn_mean0"] = bn.running_mean.numpy().copy()
    out["hfrt_bn_var0"] = bn.running_var.numpy().copy()
    out["hfrt_bn_momentum"] = np.float64(bn.momentum)
    cap = {}
    clf_t.conv_reduce.register_forward_hook(lambda m, i, o: cap.update(f=i[0].detach().clone()))
    clf_t.conv_reduce.register_full_backward_hook(lambda m, gi, go: cap.update(df=gi[0].detach().clone()))
    xt = {"out": torch.randn(3, 48, 9, 12), "low": torch.randn(3, 256, 18, 24)}
    with ref_import.cpu_only():
        logits_t, _ = clf_t(xt, size=None)
    R = torch.randn(logits_t.shape, generator=torch.Generator().manual_seed(14))
    (logits_t * R).sum().backward()
    out["hfrt_f"] = cap["f"].numpy()
    out["hfrt_df"] = cap["df"].numpy()
    out["hfrt_R"] = R.numpy()
    out["hfrt_logits"] = logits_t.detach().numpy()
    for name, t_ in named:
        out["hfrt_d" + name] = t_.grad.detach().numpy()
    out["hfrt_bn_mean1"] = bn.running_mean.numpy().copy()
    out["hfrt_bn_var1"] = bn.running_var.numpy().copy()
    out["hfrt_bn_eps"] = np.float64(bn.eps)
    out["hfrt_c"] = np.float64(clf_t.conv_seg.c)
    # loss sequence
    neg_crit = side.NegativeLearningLoss(threshold=0.05)
    g = torch.Generator().manual_seed(12)
    for i, (N, h, w, H, W, wgt, lab_frac) in enumerate(((2, 10, 14, 37, 50, 1.0, 0.3), (1, 8, 8, 8, 8, 0.5, 1.0), (1, 6, 9, 24, 33, 1.0, 0.0))):
        lg = (torch.randn((N, O, h, w), generator=g) * 3.0).double().requires_grad_(True)
        lab = torch.randint(0, O, (N, H, W), generator=g)
        lab[torch.rand((N, H, W), generator=g) >= lab_frac] = 255
        tgt_out = F.interpolate(lg, size=(H, W), mode="bilinear", align_corners=True)      # classifier.py:556-557
        predict = torch.softmax(tgt_out, dim=1)                                            # train_learners.py:343
        loss = torch.zeros((), dtype=torch.float64)
        sup = torch.zeros((), dtype=torch.float64)
        if torch.sum(lab != 255) != 0:                                                     # :346
            sup = nn.CrossEntropyLoss(ignore_index=255)(tgt_out, lab)
            loss = loss + sup
        neg = neg_crit(predict) * wgt                                                      # :351-353
        loss = loss + neg
        loss.backward()
        tag = "loss%d_" % i
        out[tag + "logits"] = lg.detach().numpy()
        out[tag + "labels"] = lab.numpy().astype(np.uint8)
        out[tag + "size"] = np.array([H, W], dtype=np.int64)
        out[tag + "weight"] = np.float64(wgt)
        out[tag + "loss"] = loss.detach().numpy()
        out[tag + "sup"] = sup.detach().numpy()
        out[tag + "neg"] = neg.detach().numpy()
        out[tag + "grad"] = lg.grad.numpy()
    out["n_loss_cases"] = np.int64(3)
    return out


def main():
    ref = ref_import.load()
    if "--train-only" in sys.argv:
        np.savez_compressed(os.path.join(HERE, "train.npz"), **train_fixtures())
        print("train.npz", os.path.getsize(os.path.join(HERE, "train.npz")) // 1024, "KiB")
        return
    np.savez_compressed(os.path.join(HERE, "train.npz"), **train_fixtures())
    np.savez_compressed(os.path.join(HERE, "head.npz"), **head_fixtures(ref))
    np.savez_compressed(os.path.join(HERE, "score.npz"), **score_fixtures(ref))
    np.savez_compressed(os.path.join(HERE, "select.npz"), **select_fixtures(ref))
    for f in ("head.npz", "score.npz", "select.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
