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


def test_reduce_hfr_feeds_the_fused_head_and_refuses_training_mode():
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
    wn_mlp.train()
    with pytest.raises(NotImplementedError):
        reduce_hfr(f.to(DEV), conv_reduce, wn_mlp)
    with pytest.raises(RuntimeError, match="no CPU path"):
        wn_mlp.eval()
        reduce_hfr(f, conv_reduce, wn_mlp)


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
