"""GPU: FloatingRegionScore (halo_logits_stats + halo_score) against the golden vectors of every mode the
reference implements (core/active/floating_region.py:129-217)."""
import pytest
import torch

import halo_b200
from tests.golden.make_golden import SCORE_COMBOS
from tests.util import TOL, rel_err, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("i", range(len(SCORE_COMBOS)))
def test_score_modes_match_golden(golden, i):
    g = golden["score"]
    ctor, size, unc, pur, norm = SCORE_COMBOS[i]
    c = float(g["c"])
    logits, x, gt = t(g["logits"]).to(DEV), t(g["x"]).to(DEV), t(g["gt"]).to(DEV)
    frs = halo_b200.FloatingRegionScore(in_channels=19, size=size, purity_type=ctor, K=100, curvature=c)
    s, imp, un = frs(logits, decoder_out=x, unc_type=unc, pur_type=pur, normalize=norm, ground_truth=gt)
    refs = [t(g["s%d_%s" % (i, n)]) for n in ("score", "impurity", "uncertainty")]
    # "hyper" included: its K radius bins (round-half-even of an fp64 radius, floating_region.py:94-110) are formed in
    # fp64 from an fp64 norm (halo_radius_f64), so no pixel lands in a neighbouring bin
    for got, ref, name in zip((s, imp, un), refs, ("score", "impurity", "uncertainty")):
        assert rel_err(got, ref) <= TOL, name


def test_score_from_lazy_embedding(golden):
    """The fused path: decoder_out is the PoincareEmbedding handle, radius comes from the raw features."""
    g = golden["score"]
    c = float(g["c"])
    u = t(g["u"]).to(DEV)
    mapper = halo_b200.HyperMapper(c=c)
    mlr = halo_b200.HyperMLR(u.shape[1], 19, c=c).to(DEV)
    mlr.load_state_dict({"P_MLR": t(g["P"]), "A_MLR": t(g["A"])})
    with torch.no_grad():
        emb = mapper.expmap(u, dim=1)
        out = mlr(emb.double()).float()
    frs = halo_b200.FloatingRegionScore(in_channels=19, size=3, purity_type="radius", curvature=c)
    s, imp, un = frs(out, decoder_out=emb, unc_type="entropy", pur_type="radius", normalize=True)
    assert rel_err(s, t(g["s0_score"])) <= TOL
    assert rel_err(imp, t(g["s0_impurity"])) <= TOL
    assert rel_err(un, t(g["s0_uncertainty"])) <= TOL


def test_unknown_purity_raises():
    frs = halo_b200.FloatingRegionScore(in_channels=19, size=3, purity_type="radius", curvature=1.0)
    with pytest.raises(NotImplementedError):
        frs(torch.zeros(1, 19, 4, 4, device=DEV), unc_type="entropy", pur_type="bogus")
