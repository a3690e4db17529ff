"""GPU: channel reduction + HFR upstream of the head (halo_reduce_hfr_fwd through halo_b200.hfr.reduce_hfr) against the
reference block core/models/classifier.py:526-550 run by the oracle on the CPU in float64, and chained into the fused head."""
import copy

import pytest
import torch

import halo_b200
from halo_b200.hfr import reduce_hfr
from oracle import head as ohead
from oracle import hfr as ohfr
from tests.util import TOL, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("shape", [(2, 512, 64, 20, 40, True), (1, 512, 64, 17, 23, True), (2, 96, 32, 9, 12, True),
                                   (1, 512, 128, 16, 24, True), (2, 512, 64, 20, 40, False), (1, 70, 19, 5, 7, False),
                                   (1, 256, 256, 8, 8, False)])
def test_reduce_hfr_matches_reference_block(shape):
    N, Cin, C, H, W, hfr = shape
    conv_reduce, wn_mlp = ohfr.build_modules(Cin, C, hfr=hfr, seed=Cin + C)
    conv_reduce.eval()
    if wn_mlp is not None:
        wn_mlp.eval()
    f = torch.randn((N, Cin, H, W), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = ohfr.reduce_hfr(f.double(), copy.deepcopy(conv_reduce).double(), copy.deepcopy(wn_mlp).double() if hfr else None)
        out = reduce_hfr(f.to(DEV), conv_reduce, wn_mlp)
        again = reduce_hfr(f.to(DEV), conv_reduce, wn_mlp)
    assert out.dtype == torch.float32 and tuple(out.shape) == (N, C, H, W)
    assert rel_err(out, ref) <= TOL
    assert torch.equal(out, again)            # fixed-order reductions


def test_reduce_hfr_feeds_the_fused_head():
    N, Cin, C, O, H, W = 2, 512, 64, 19, 16, 32
    conv_reduce, wn_mlp = ohfr.build_modules(Cin, C, hfr=True, seed=3)
    conv_reduce.eval(); wn_mlp.eval()
    f = torch.randn((N, Cin, H, W), generator=torch.Generator().manual_seed(2))
    P, A = halo_b200.synth.head_params(O, C, seed=4, dtype=torch.float64)
    with torch.no_grad():
        z_ref = ohfr.reduce_hfr(f.double(), copy.deepcopy(conv_reduce).double(), copy.deepcopy(wn_mlp).double())
        logits_ref, _, rad_ref = ohead.head_forward(z_ref.float(), P, A, 1.0)
        z = reduce_hfr(f.to(DEV), conv_reduce, wn_mlp)
        res = halo_b200.head_forward(z, P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True)
    assert rel_err(res["logits"], logits_ref) <= 2 * TOL     # two fp32 stages chained against one fp64 chain
    assert rel_err(res["radius"], rad_ref) <= 2 * TOL
    with pytest.raises(RuntimeError, match="no CPU path"):
        reduce_hfr(f, conv_reduce, wn_mlp)


def _train_reference(f, conv_reduce, wn_mlp, dz, bn_train=True):
    """The reference block (oracle/hfr.py = classifier.py:526-550) in float64 with torch autograd: output, every gradient, and
    the BatchNorm running statistics after the step."""
    conv, mlp = copy.deepcopy(conv_reduce).double(), (copy.deepcopy(wn_mlp).double() if wn_mlp is not None else None)
    conv.train()
    if mlp is not None:
        mlp.train(bn_train)
    x = f.double().clone().requires_grad_(True)
    z = ohfr.reduce_hfr(x, conv, mlp)
    (z * dz.double()).sum().backward()
    grads = {"x": x.grad}
    for name, p in list(conv.named_parameters()) + ([("mlp." + k, v) for k, v in mlp.named_parameters()] if mlp is not None else []):
        grads[name] = p.grad
    return z.detach(), grads, mlp


@pytest.mark.parametrize("shape", [(2, 512, 64, 20, 40, True, True), (3, 96, 32, 9, 13, True, True), (1, 304, 64, 33, 65, True, True),
                                   (2, 512, 64, 20, 40, True, False), (2, 512, 64, 20, 40, False, True), (1, 70, 19, 5, 7, False, True),
                                   (2, 256, 48, 16, 16, True, True)])
def test_reduce_hfr_training_mode_forward_backward(shape):
    """Training step: BatchNorm1d on batch statistics (or left in eval mode: bn_train False), gradients to the features and to
    every parameter of conv_reduce / wn_mlp, running statistics updated like torch -- against float64 autograd through the
    reference block."""
    N, Cin, C, H, W, hfr, bn_train = shape
    conv_reduce, wn_mlp = ohfr.build_modules(Cin, C, hfr=hfr, seed=Cin + C + 1)
    f = torch.randn((N, Cin, H, W), generator=torch.Generator().manual_seed(5))
    dz = torch.randn((N, C, H, W), generator=torch.Generator().manual_seed(6))
    z_ref, g_ref, mlp_ref = _train_reference(f, conv_reduce, wn_mlp, dz, bn_train)

    conv, mlp = copy.deepcopy(conv_reduce).to(DEV), (copy.deepcopy(wn_mlp).to(DEV) if hfr else None)
    conv.train()
    if hfr:
        mlp.train(bn_train)
    x = f.to(DEV).requires_grad_(True)
    z = reduce_hfr(x, conv, mlp)
    assert z.requires_grad and tuple(z.shape) == (N, C, H, W)
    assert rel_err(z, z_ref) <= TOL
    (z * dz.to(DEV)).sum().backward()
    assert rel_err(x.grad, g_ref["x"]) <= 1e-4
    got = dict(list(conv.named_parameters()) + ([("mlp." + k, v) for k, v in mlp.named_parameters()] if hfr else []))
    # the bias in front of a batch-statistics BatchNorm has an exactly zero gradient (1e-16 in float64): measure every
    # parameter gradient against the larger of its own magnitude and 1e-2 of the largest parameter gradient (the zero comes
    # out of a cancelling fp32 sum over all pixels)
    floor = 1e-2 * max(float(v.abs().max()) for k, v in g_ref.items() if k != "x")
    for name, ref in g_ref.items():
        if name == "x":
            continue
        assert got[name].grad is not None, name
        err = float((got[name].grad.detach().cpu().double() - ref).abs().max()) / max(float(ref.abs().max()), floor)
        assert err <= 1e-4, (name, err)
    if hfr:
        bn, bn_ref = mlp[1], mlp_ref[1]
        assert rel_err(bn.running_mean, bn_ref.running_mean) <= TOL and rel_err(bn.running_var, bn_ref.running_var) <= TOL
        assert int(bn.num_batches_tracked) == int(bn_ref.num_batches_tracked)
    # fixed-order reductions: a second step from the same state returns the same bits
    conv2, mlp2 = copy.deepcopy(conv_reduce).to(DEV), (copy.deepcopy(wn_mlp).to(DEV) if hfr else None)
    conv2.train()
    if hfr:
        mlp2.train(bn_train)
    x2 = f.to(DEV).requires_grad_(True)
    z2 = reduce_hfr(x2, conv2, mlp2)
    (z2 * dz.to(DEV)).sum().backward()
    assert torch.equal(z, z2) and torch.equal(x.grad, x2.grad) and torch.equal(conv.weight.grad, conv2.weight.grad)


def test_reduce_hfr_training_mode_chains_into_the_fused_head_and_loss():
    """conv_reduce + HFR -> fused head -> fused loss, one autograd graph, against the same chain in float64 torch."""
    from halo_b200.losses import fused_seg_loss
    from oracle import loss as oloss

    N, Cin, C, O, H, W = 2, 128, 64, 19, 12, 20
    conv_reduce, wn_mlp = ohfr.build_modules(Cin, C, hfr=True, seed=11)
    P, A = halo_b200.synth.head_params(O, C, seed=4, dtype=torch.float64)
    f = torch.randn((N, Cin, H, W), generator=torch.Generator().manual_seed(7)) * 0.3
    labels = torch.randint(0, O, (N, 4 * H, 4 * W), generator=torch.Generator().manual_seed(8))
    labels[:, ::3] = 255
    # reference chain
    conv, mlp = copy.deepcopy(conv_reduce).double().train(), copy.deepcopy(wn_mlp).double().train()
    Pr, Ar = P.clone().requires_grad_(True), A.clone().requires_grad_(True)
    z = ohfr.reduce_hfr(f.double(), conv, mlp)
    logits = ohead.mlr_logits(ohead.expmap(z, 1.0, dim=1).double(), Pr, Ar, 1.0)
    up = torch.nn.functional.interpolate(logits, size=(4 * H, 4 * W), mode="bilinear", align_corners=True)
    loss_ref = torch.nn.functional.cross_entropy(up, labels.long(), ignore_index=255) + \
        oloss.negative_learning_loss(torch.softmax(up, dim=1), 0.05)
    loss_ref.backward()
    # fused chain
    convd, mlpd = copy.deepcopy(conv_reduce).to(DEV).train(), copy.deepcopy(wn_mlp).to(DEV).train()
    head = halo_b200.HyperMLR(C, O, c=1.0).to(DEV)
    with torch.no_grad():
        head.P_MLR.copy_(P); head.A_MLR.copy_(A)
    mapper = halo_b200.HyperMapper(c=1.0)
    zz = reduce_hfr(f.to(DEV), convd, mlpd)
    out = head(mapper.expmap(zz, dim=1))
    loss = fused_seg_loss(out, labels.to(DEV), (4 * H, 4 * W), 1.0, 0.05)[0]
    loss.backward()
    assert abs(float(loss.detach()) - float(loss_ref.detach())) <= 1e-5 * max(1.0, abs(float(loss_ref)))
    assert rel_err(convd.weight.grad, conv.weight.grad) <= 2e-4
    assert rel_err(mlpd[0].weight.grad, mlp[0].weight.grad) <= 2e-4
    assert rel_err(head.P_MLR.grad, Pr.grad) <= 2e-4


def test_reference_classifier_golden(golden):
    """conv_reduce + HFR + fused head on the GPU against what the reference's real classifier class
    (DepthwiseSeparableASPP_Hyper, core/models/classifier.py:526-558) returned for the same decoder features: its float64
    embedding and its logits (tests/golden/train.npz, frozen from the live reference)."""
    from tests.test_oracle_golden import _hfr_modules, t

    g = golden["train"]
    conv, mlp = _hfr_modules(g, torch.float32)
    c = float(g["hfr_c"])
    with torch.no_grad():
        z = reduce_hfr(t(g["hfr_f"]).to(DEV), conv, mlp)
        mapper = halo_b200.HyperMapper(c=c)
        mlr = halo_b200.HyperMLR(z.shape[1], 19, c=c).to(DEV)
        mlr.load_state_dict({"P_MLR": t(g["hfr_P"]), "A_MLR": t(g["hfr_A"])})
        emb = mapper.expmap(z, dim=1)                    # classifier.py:553
        out = mlr(emb.double()).float()                  # classifier.py:554
    assert rel_err(out, t(g["hfr_logits"])) <= 2 * TOL
    assert rel_err(emb.materialize(), t(g["hfr_emb"])) <= 2 * TOL


def test_reference_classifier_training_step_golden(golden):
    """One training step of conv_reduce + HFR + fused head on the GPU against the forward + backward of the reference's real
    classifier class in training mode (tests/golden/train.npz, frozen from the live reference): logits, the gradient w.r.t.
    the decoder features, every parameter gradient, BatchNorm1d running statistics."""
    from tests.test_oracle_golden import _hfr_train_modules, check_hfr_train_grads, t

    g = golden["train"]
    conv, mlp = _hfr_train_modules(g)
    conv, mlp = conv.to(DEV), mlp.to(DEV)
    c = float(g["hfrt_c"])
    mapper = halo_b200.HyperMapper(c=c)
    mlr = halo_b200.HyperMLR(conv.weight.shape[0], 19, c=c).to(DEV)
    mlr.load_state_dict({"P_MLR": t(g["hfrt_P"]), "A_MLR": t(g["hfrt_A"])})
    f = t(g["hfrt_f"]).to(DEV).requires_grad_(True)
    z = reduce_hfr(f, conv, mlp)                         # classifier.py:527-550, training mode
    out = mlr(mapper.expmap(z, dim=1).double()).float()  # :553-554
    assert rel_err(out, t(g["hfrt_logits"])) <= 2 * TOL
    (out * t(g["hfrt_R"]).to(DEV)).sum().backward()
    check_hfr_train_grads(g, conv, mlp, f.grad, mlr.P_MLR.grad, mlr.A_MLR.grad, 2e-4)
