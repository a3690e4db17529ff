"""Oracle for the budgeted greedy selection (ORACLE -- test infrastructure, CPU only).

Restates core/active/build.py:27-64 (select_pixels_to_label) twice:
  * `select_sequential`  -- the literal loop (nested torch.max, first index on ties), used as
    the checker and as the timed CPU baseline;
  * `select_numpy`       -- the same loop on numpy arrays with a column-max cache, ~50x faster,
    bit-identical (checked against `select_sequential` in tests) -- used for full-size images.
Tie-break (probed on the reference): torch.max(score, dim=0) keeps the FIRST row index per
column, torch.max(values, dim=0) the FIRST column, i.e. equal maxima resolve to smallest w,
then smallest h.  The loop stops at the first -inf maximum (:40-41).
"""
import numpy as np
import torch

NEG_INF = -float("inf")


def _window(center, radius):
    lo = center - radius
    return (lo if lo >= 0 else 0), center + radius + 1  # end is clipped by slicing (:45-53)


def select_sequential(score, active_regions, active_radius, mask_radius, active, selected, active_mask, ground_truth):
    """build.py:37-64.  Mutates all four tensors in place and returns them plus the pick list."""
    picks = []
    for _ in range(active_regions):
        col_best, col_arg = torch.max(score, dim=0)
        best, w_arg = torch.max(col_best, dim=0)
        if best == NEG_INF:
            break
        w = w_arg.item()
        h = col_arg[w].item()
        a_h0, a_h1 = _window(h, active_radius)
        a_w0, a_w1 = _window(w, active_radius)
        m_h0, m_h1 = _window(h, mask_radius)
        m_w0, m_w1 = _window(w, mask_radius)
        score[m_h0:m_h1, m_w0:m_w1] = NEG_INF
        active[m_h0:m_h1, m_w0:m_w1] = True
        selected[a_h0:a_h1, a_w0:a_w1] = True
        active_mask[a_h0:a_h1, a_w0:a_w1] = ground_truth[a_h0:a_h1, a_w0:a_w1]
        picks.append((h, w))
    return score, active, selected, active_mask, picks


def select_numpy(score, active_regions, active_radius, mask_radius, active, selected, active_mask, ground_truth):
    """Same semantics on numpy arrays (in place).  Keeps per-column maxima and refreshes only the
    columns a pick touched, so a 1280x640 image with 4 552 picks takes ~1 s instead of minutes."""
    H, W = score.shape
    picks = []
    if np.isnan(score).any():
        raise ValueError("select_numpy: NaN scores not supported by the fast oracle; use select_sequential")
    col_arg = score.argmax(axis=0)  # first max per column
    col_best = score[col_arg, np.arange(W)]
    for _ in range(active_regions):
        w = int(col_best.argmax())  # first max over columns
        if col_best[w] == NEG_INF:
            break
        h = int(col_arg[w])
        a_h0, a_h1 = _window(h, active_radius)
        a_w0, a_w1 = _window(w, active_radius)
        m_h0, m_h1 = _window(h, mask_radius)
        m_w0, m_w1 = _window(w, mask_radius)
        score[m_h0:m_h1, m_w0:m_w1] = NEG_INF
        active[m_h0:m_h1, m_w0:m_w1] = True
        selected[a_h0:a_h1, a_w0:a_w1] = True
        active_mask[a_h0:a_h1, a_w0:a_w1] = ground_truth[a_h0:a_h1, a_w0:a_w1]
        sub = score[:, m_w0:m_w1]
        arg = sub.argmax(axis=0)
        col_arg[m_w0:m_w1] = arg
        col_best[m_w0:m_w1] = sub[arg, np.arange(sub.shape[1])]
        picks.append((h, w))
    return score, active, selected, active_mask, picks
