"""Oracle for the round-delta exchange (ORACLE -- test infrastructure, CPU only).

Restates on CPU tensors what `halo_round_delta_pack` / `halo_round_delta_apply` do on the device: a selection round
labels the (2a+1)^2 window around every pick, clipped at the image border (core/active/build.py:45-48, 58-62:
`active_mask[h-a:h+a+1, w-a:w+a+1] = ground_truth[same]`), so the shards exchange the picks and the ground-truth labels
of their windows instead of the mask planes.
"""
import torch


def _windows(picks, n_picked, H, W, r):
    N, cap = picks.shape
    valid = (picks >= 0) & (torch.arange(cap)[None, :] < n_picked[:, None])
    p = picks.clamp_min(0).long()
    d = torch.arange(-r, r + 1)
    hh = (p // W)[..., None, None] + d[:, None]          # (N,cap,k,1)
    ww = (p % W)[..., None, None] + d[None, :]           # (N,cap,1,k)
    inside = (hh >= 0) & (hh < H) & (ww >= 0) & (ww < W) & valid[..., None, None]
    return hh, ww, inside


def pack_round_delta(picks, n_picked, gt, active_radius):
    """lab (N,cap,(2a+1)^2) uint8: gt of every pick's window, row-major; 255 outside the image / beyond n_picked."""
    N, cap = picks.shape
    H, W = gt.shape[-2:]
    r = int(active_radius)
    k = 2 * r + 1
    hh, ww, inside = _windows(picks, n_picked, H, W, r)
    flat = (hh.clamp(0, H - 1) * W + ww.clamp(0, W - 1)).reshape(N, cap * k * k)
    lab = torch.gather(gt.reshape(N, H * W), 1, flat)
    return torch.where(inside.reshape(N, cap * k * k), lab, torch.full_like(lab, 255)).reshape(N, cap, k * k)


def apply_round_delta(masks, row_image, picks, n_picked, lab, active_radius):
    """masks (n_images,H,W) uint8, in place: masks[row_image[j]][window of pick i] = lab[j][i] for labelled (!= 255) entries."""
    rows, cap = picks.shape
    H, W = masks.shape[-2:]
    r = int(active_radius)
    k = 2 * r + 1
    hh, ww, inside = _windows(picks, n_picked, H, W, r)
    labw = lab.reshape(rows, cap, k, k)
    inside = inside & (row_image[:, None, None, None] >= 0) & (labw != 255)
    flat = row_image.clamp_min(0).long()[:, None, None, None] * (H * W) + hh * W + ww
    masks.view(-1)[flat[inside]] = labw[inside]
    return masks
