"""Oracle for one image of the acquisition round (ORACLE -- test infrastructure, CPU only).

Restates the per-image body of RegionSelection, core/active/build.py:75-88 (config-derived
constants) and :137-160 (score -> mask already-labelled pixels -> region budget -> select),
without the model forward, the bilinear up-sampling (:122-135) and the file I/O (:162-166).
"""
import math

import torch

from . import head as _head
from . import score as _score
from . import select as _select


def region_budget(h, w, budget, n_rounds, radius_k):
    """build.py:75,78,148-150."""
    per_region = (2 * radius_k + 1) ** 2
    return math.ceil(h * w * (budget / n_rounds) / per_region)


def acquire_image(
    u, P, A, ground_truth, active, selected, active_mask, *, c=1.0, radius_k=1, mask_radius_k=5,
    budget=0.05, n_rounds=1, unc_type="entropy", pur_type="radius", normalize=True, K=100, fast_select=False,
):
    """u: (1,C,H,W) fp32 raw decoder features.  active/selected: (H,W) bool; active_mask, ground_truth: (H,W) int64.
    Returns dict(score, active, selected, active_mask, picks, logits, radius)."""
    num_classes = P.shape[0]
    logits, x, rad = _head.head_forward(u, P, A, c)
    h, w = logits.shape[-2:]
    score, imp, unc = _score.floating_region_score(
        logits, decoder_out=x, unc_type=unc_type, pur_type=pur_type, normalize=normalize,
        ground_truth=ground_truth, in_channels=num_classes, size=2 * radius_k + 1,
        ctor_purity_type=pur_type, K=K, c=c,
    )
    score_before = score.clone()
    score[active] = -float("inf")  # build.py:146
    n_regions = region_budget(h, w, budget, n_rounds, radius_k)
    if fast_select:
        s_np, a_np, sel_np, m_np = score.numpy(), active.numpy(), selected.numpy(), active_mask.numpy()
        _, _, _, _, picks = _select.select_numpy(s_np, n_regions, radius_k, mask_radius_k, a_np, sel_np, m_np,
                                                 ground_truth.numpy())
    else:
        _, _, _, _, picks = _select.select_sequential(score, n_regions, radius_k, mask_radius_k, active, selected,
                                                      active_mask, ground_truth)
    return dict(score=score_before, score_after=score, active=active, selected=selected, active_mask=active_mask,
                picks=picks, logits=logits, radius=rad[0], impurity=imp, uncertainty=unc, n_regions=n_regions)


def upsampled_score(logits_lr, x_lr, out_size, **score_kwargs):
    """build.py:122-144: bilinear (align_corners=True) up-sampling of the logits and of the (fp64) embedding to the
    label size, then FloatingRegionScore on the up-sampled tensors."""
    import torch.nn.functional as F

    out = F.interpolate(logits_lr, size=out_size, mode="bilinear", align_corners=True)
    emb = F.interpolate(x_lr, size=out_size, mode="bilinear", align_corners=True) if x_lr is not None else None
    return _score.floating_region_score(out, decoder_out=emb, **score_kwargs)

