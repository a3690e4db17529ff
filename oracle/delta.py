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


# ---- packed rows (halo_round_rows_pack / halo_round_rows_apply / halo_checksum64), restated with numpy -------------
def row_bytes(cap, active_radius):
    k2 = (2 * int(active_radius) + 1) ** 2
    return (4 + 4 * cap + cap * k2 + 15) // 16 * 16


def pack_rows(rows, picks, n_picked, gt, cap, active_radius):
    """rows (b,row_bytes) uint8, in place: [int32 count | int32 picks[cap] | uint8 lab[cap][(2a+1)^2]] per image;
    only the first `count` picks and their labels are written (the rest of the row is never read by apply)."""
    import numpy as np

    b = picks.shape[0]
    k2 = (2 * int(active_radius) + 1) ** 2
    lab = pack_round_delta(picks[:, :cap].contiguous(), n_picked, gt, active_radius).numpy()
    out = rows.numpy()
    for n in range(b):
        cnt = min(int(n_picked[n]), cap)
        out[n, 0:4] = np.frombuffer(np.int32(cnt).tobytes(), dtype=np.uint8)
        out[n, 4:4 + 4 * cnt] = np.frombuffer(picks[n, :cnt].numpy().astype(np.int32).tobytes(), dtype=np.uint8)
        out[n, 4 + 4 * cap:4 + 4 * cap + cnt * k2] = lab[n, :cnt].reshape(-1)
    return rows


def apply_rows(masks, row_image, rows, n_picked_out, cap, active_radius):
    """Replay packed rows onto masks (n_images,H,W) uint8 in place; counts go to n_picked_out[image]."""
    import numpy as np

    k2 = (2 * int(active_radius) + 1) ** 2
    raw = rows.numpy()
    n_rows = raw.shape[0]
    cnt = torch.from_numpy(raw[:, 0:4].copy().view(np.int32).reshape(n_rows).copy())
    picks = torch.from_numpy(raw[:, 4:4 + 4 * cap].copy().view(np.int32).reshape(n_rows, cap).copy())
    lab = torch.from_numpy(raw[:, 4 + 4 * cap:4 + 4 * cap + cap * k2].copy().reshape(n_rows, cap, k2))
    # entries past `count` hold stale bytes: mask them before the generic replay reads them
    live = torch.arange(cap)[None, :] < cnt[:, None]
    picks = torch.where(live, picks, torch.full_like(picks, -1))
    apply_round_delta(masks, row_image, picks, cnt, lab, active_radius)
    if n_picked_out is not None:
        keep = row_image >= 0
        n_picked_out[row_image[keep].long()] = cnt[keep]
    return masks


def checksum64(masks, n_picked):
    """sum over 8-byte little-endian words of word_i * (2 i + 1) mod 2^64 over masks then counts (zero-extended tails);
    returned in a 1-element int64 tensor like the device version."""
    import numpy as np

    def words(t):
        raw = t.contiguous().numpy().reshape(-1).view(np.uint8)
        pad = (-raw.size) % 8
        if pad:
            raw = np.concatenate([raw, np.zeros(pad, np.uint8)])
        return raw.view("<u8")

    total = np.uint64(0)
    off = 0
    with np.errstate(over="ignore"):
        for t in (masks, n_picked):
            w = words(t)
            idx = np.arange(off, off + w.size, dtype=np.uint64)
            total = total + (w * (np.uint64(2) * idx + np.uint64(1))).sum(dtype=np.uint64)
            off += (t.numel() * t.element_size() + 7) // 8
    return torch.from_numpy(np.array([total], dtype=np.uint64).view(np.int64).copy())
