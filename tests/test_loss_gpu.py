"""GPU: the fused training losses (halo_seg_loss through halo_b200.losses.fused_seg_loss) against the reference sequence
F.interpolate -> softmax -> CrossEntropyLoss(ignore 255) + NegativeLearningLoss -> autograd (core/train_learners.py:343-362,
core/loss/negative_learning_loss.py:6-16, core/models/classifier.py:556-557) run by the oracle in float64."""
import pytest
import torch

import halo_b200
from halo_b200.losses import fused_seg_loss
from oracle import loss as oloss
from tests.util import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(N, O, h, w, H, W, seed, labelled=0.3, scale=3.0):
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn((N, O, h, w), generator=g) * scale
    labels = torch.randint(0, O, (N, H, W), generator=g)
    labels[torch.rand((N, H, W), generator=g) > labelled] = 255
    return logits, labels


@pytest.mark.parametrize("shape", [(2, 19, 20, 40, 80, 160), (1, 19, 17, 23, 50, 71), (2, 16, 12, 12, 12, 12),
                                   (1, 5, 9, 7, 33, 20), (3, 19, 8, 16, 8, 61), (1, 19, 30, 30, 20, 25), (1, 32, 6, 6, 24, 24),
                                   (1, 19, 4, 5, 64, 70),      # 16x up-sampling: footprint too large for the tiled pass B -> untiled gather
                                   (2, 19, 40, 80, 160, 320)])  # several 8x8 tiles per image in both directions
@pytest.mark.parametrize("neg_weight", [1.0, 0.0])
def test_fused_loss_matches_reference_sequence(shape, neg_weight):
    N, O, h, w, H, W = shape
    logits, labels = _case(N, O, h, w, H, W, seed=sum(shape))
    ref_loss, ref_sup, ref_neg, ref_g = oloss.seg_loss(logits, labels, (H, W), neg_weight=neg_weight)
    x = logits.to(DEV).requires_grad_(True)
    loss, sup, neg = fused_seg_loss(x, labels.to(DEV), (H, W), neg_weight=neg_weight)
    loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= 1e-5 * max(1.0, abs(float(ref_loss)))
    assert abs(float(sup) - float(ref_sup)) <= 1e-5 * max(1.0, abs(float(ref_sup)))
    assert abs(float(neg) - float(ref_neg)) <= 1e-5 * max(1.0, abs(float(ref_neg)))
    # The negative-learning mask [p < threshold] is a step: a label-resolution pixel whose probability sits within fp32
    # rounding of the threshold may take the other side than the float64 reference (once per ~1e6 pixel-classes).  Its
    # whole contribution then differs, so the low-resolution pixels that read it are left out of the comparison.
    up = torch.nn.functional.interpolate(logits.double(), size=(H, W), mode="bilinear", align_corners=True)
    near = ((torch.softmax(up, dim=1) - 0.05).abs() < 2e-6).any(dim=1, keepdim=True).float()
    low = torch.nn.functional.adaptive_max_pool2d(near, (h, w))
    low = torch.nn.functional.max_pool2d(low, kernel_size=3, stride=1, padding=1)      # the stencil reaches one pixel further
    keep = (low == 0).expand_as(ref_g) if neg_weight > 0 else torch.ones_like(ref_g, dtype=torch.bool)
    assert keep.float().mean() > 0.9
    assert rel_err(x.grad.cpu() * keep, ref_g * keep) <= 1e-5


def test_fused_loss_edge_cases_and_determinism():
    N, O, h, w, H, W = 2, 19, 16, 32, 64, 128
    logits, labels = _case(N, O, h, w, H, W, seed=5)
    # nothing labelled: the supervised term is skipped (train_learners.py:345), the gradient is the negative term's alone
    none = torch.full_like(labels, 255)
    ref_loss, ref_sup, ref_neg, ref_g = oloss.seg_loss(logits, none, (H, W), neg_weight=0.5)
    x = logits.to(DEV).requires_grad_(True)
    loss, sup, neg = fused_seg_loss(x, none.to(DEV), (H, W), neg_weight=0.5)
    loss.backward()
    assert float(sup) == 0.0 and abs(float(loss) - float(ref_loss)) <= 1e-5 and rel_err(x.grad, ref_g) <= 1e-5
    # labels=None is the same thing
    x2 = logits.to(DEV).requires_grad_(True)
    loss2, _, _ = fused_seg_loss(x2, None, (H, W), neg_weight=0.5)
    loss2.backward()
    assert torch.equal(x2.grad, x.grad) and float(loss2) == float(loss)
    # an upstream scale flows through (loss * 3).backward()
    x3 = logits.to(DEV).requires_grad_(True)
    (fused_seg_loss(x3, labels.to(DEV), (H, W))[0] * 3.0).backward()
    x4 = logits.to(DEV).requires_grad_(True)
    fused_seg_loss(x4, labels.to(DEV), (H, W))[0].backward()
    assert torch.allclose(x3.grad, 3.0 * x4.grad, rtol=1e-6, atol=0)
    # bitwise reproducible (gather + fixed-order sums; torch's own bilinear backward scatters with atomics)
    x5 = logits.to(DEV).requires_grad_(True)
    l5 = fused_seg_loss(x5, labels.to(DEV), (H, W))[0]
    l5.backward()
    assert torch.equal(x5.grad, x4.grad)
    # forward only (no grad): no gradient buffer, same value
    with torch.no_grad():
        l6, _, _ = fused_seg_loss(logits.to(DEV), labels.to(DEV), (H, W))
    assert float(l6) == float(l5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        fused_seg_loss(logits, labels, (H, W))


def test_fused_loss_behind_the_fused_head():
    """The training step end to end on the drop-in modules: head (fused fwd) -> fused loss -> streaming backward, against the
    float64 oracle chain expmap -> HyperMLR -> interpolate -> CE + negative loss."""
    from halo_b200 import synth
    from oracle import head as ohead

    C, O, h, w, H, W = 64, 19, 16, 32, 64, 128
    P, A = synth.head_params(O, C, seed=3, dtype=torch.float64)
    u = torch.randn(2, C, h, w) * 0.1
    _, labels = _case(2, O, h, w, H, W, seed=9)
    u0, P0, A0 = u.clone().requires_grad_(True), P.clone().requires_grad_(True), A.clone().requires_grad_(True)
    lo = ohead.mlr_logits(ohead.expmap(u0, 1.0, dim=1), P0, A0, 1.0).float()
    out = torch.nn.functional.interpolate(lo.double(), size=(H, W), mode="bilinear", align_corners=True)
    ref = torch.nn.functional.cross_entropy(out, labels, ignore_index=255) + oloss.negative_learning_loss(torch.softmax(out, dim=1))
    ref.backward()
    mlr = halo_b200.HyperMLR(C, O, c=1.0).to(DEV)
    mlr.load_state_dict({"P_MLR": P, "A_MLR": A})
    u1 = u.to(DEV).requires_grad_(True)
    logits_lr = mlr(halo_b200.HyperMapper(1.0).expmap(u1, dim=1).double()).float()
    loss, _, _ = fused_seg_loss(logits_lr, labels.to(DEV), (H, W))
    loss.backward()
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    assert rel_err(u1.grad, u0.grad) <= 1e-4
    assert rel_err(mlr.P_MLR.grad, P0.grad) <= 1e-4 and rel_err(mlr.A_MLR.grad, A0.grad) <= 1e-4


def test_fused_loss_matches_golden(golden):
    """Against the learner's sequence run with the reference's own NegativeLearningLoss (tests/golden/train.npz)."""
    from tests.util import t

    g = golden["train"]
    for i in range(int(g["n_loss_cases"])):
        tag = "loss%d_" % i
        size = tuple(int(v) for v in g[tag + "size"])
        x = t(g[tag + "logits"]).float().to(DEV).requires_grad_(True)
        loss, sup, neg = fused_seg_loss(x, t(g[tag + "labels"]).to(DEV), size, neg_weight=float(g[tag + "weight"]))
        loss.backward()
        assert abs(float(loss) - float(g[tag + "loss"])) <= 1e-5 * max(1.0, abs(float(g[tag + "loss"])))
        assert abs(float(sup) - float(g[tag + "sup"])) <= 1e-5 * max(1.0, abs(float(g[tag + "sup"])))
        assert abs(float(neg) - float(g[tag + "neg"])) <= 1e-5 * max(1.0, abs(float(g[tag + "neg"])))
        assert rel_err(x.grad, t(g[tag + "grad"])) <= 1e-5
