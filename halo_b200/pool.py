"""Device-resident, image-sharded acquisition round (the hot path end to end).

`acquire_batch` runs K1 (fused head) -> K2 (region score) -> K3 (budgeted selection) on a batch of images
that is already in HBM, with no host synchronisation.  `acquire_pool` shards a pool of images over the ranks
of a torch.distributed job (one process per GPU; images are independent units -- score normalisation,
budget and suppression are all per image: floating_region.py:22-23, build.py:148-160) and all-gathers the
per-image pick counts and masks at the end; that all-gather is the only collective on the path.
The reference runs this loop on rank 0 alone, one image at a time (train_learners.py:308-322, build.py:92).
"""
import math
from dataclasses import dataclass

import torch

from . import _native as nat
from .active import select_planes
from .floating_region import modes_for, radius_f64, score_planes
from .hyperbolic import head_forward


@dataclass
class AcquisitionConfig:
    """The cfg keys the hot path reads (build.py:75-81,140; floating_region.py:68; defaults.py:66-79)."""
    num_classes: int = 19
    curvature: float = 1.0
    radius_k: int = 1            # ACTIVE.RADIUS_K  -> (2k+1)^2 region, active_radius
    mask_radius_k: int = 5       # ACTIVE.MASK_RADIUS_K
    budget: float = 0.05         # ACTIVE.BUDGET (total, over all rounds)
    n_rounds: int = 1            # len(ACTIVE.SELECT_ITER); per-round budget = budget / n_rounds (build.py:78)
    uncertainty: str = "entropy"
    purity: str = "radius"
    K: int = 100
    normalize: bool = True

    @classmethod
    def from_cfg(cls, cfg):
        return cls(num_classes=cfg.MODEL.NUM_CLASSES, curvature=cfg.MODEL.CURVATURE, radius_k=cfg.ACTIVE.RADIUS_K,
                   mask_radius_k=cfg.ACTIVE.MASK_RADIUS_K, budget=cfg.ACTIVE.BUDGET,
                   n_rounds=len(cfg.ACTIVE.SELECT_ITER), uncertainty=cfg.ACTIVE.UNCERTAINTY,
                   purity=cfg.ACTIVE.PURITY, K=cfg.ACTIVE.K, normalize=cfg.ACTIVE.NORMALIZE)

    def regions_per_image(self, h, w):
        per_region = (2 * self.radius_k + 1) ** 2
        return math.ceil(h * w * (self.budget / self.n_rounds) / per_region)  # build.py:148-150


def acquire_batch(feat, P, A, cfg, gt, active, selected, active_mask, *, want_score=False, want_picks=False, head_events=None):
    """One acquisition step on a resident batch.

    feat (B,C,H,W) fp32 raw decoder features; gt/active/selected/active_mask (B,H,W) uint8, the last three are
    updated IN PLACE exactly as select_pixels_to_label does.  Returns dict(n_picked (B,) int32 [, score, picks]).
    head_events = (start, end) CUDA events: recorded on the current stream around the fused-head launch (K1), for callers
    that time the dominant kernel inside their own timed region (bench.py)."""
    B, C, H, W = feat.shape
    unc_mode, pixunc_mode, pur_mode, label_mode, norm_mode = modes_for(cfg.uncertainty, cfg.purity)
    need_pixunc = unc_mode != nat.UNC_ZERO
    need_label = pur_mode == nat.PUR_LABEL_HIST
    need_radius = pur_mode == nat.PUR_NORM
    need_gt = pixunc_mode == "one_minus_pgt" or label_mode == "gt_filled"
    if head_events is not None:
        head_events[0].record()
    res = head_forward(feat, P, A, cfg.curvature, kind="tangent", want_logits=False, want_radius=need_radius,
                       want_pixunc=need_pixunc, want_label=need_label, want_stats=False,
                       gt=gt if need_gt else None, pixunc_mode=pixunc_mode, label_mode=label_mode, norm_mode=norm_mode)
    if head_events is not None:
        head_events[1].record()
    r64 = st64 = None
    if pur_mode == nat.PUR_RADIUS_BINS:
        # "hyper": the K radius bins follow the reference's fp64 arithmetic (floating_region.py:94-110), which costs a
        # second pass over the features (fp64 |u|^2); the other purities stay on the single fused pass
        r64, st64 = radius_f64(feat, cfg.curvature, "tangent")
    pixunc = res["pixunc"]
    if pixunc is None and res["radius"] is None and r64 is None:
        pixunc = torch.zeros((B, H, W), dtype=torch.float32, device=feat.device)
    k = 2 * cfg.radius_k + 1
    pk = 3 if cfg.purity == "hyper" else k
    n_bins = cfg.K if pur_mode == nat.PUR_RADIUS_BINS else cfg.num_classes
    score, _, _ = score_planes(pixunc, res["radius"], None, res["label"], active, unc_mode=unc_mode,
                               pur_mode=pur_mode, normalize=cfg.normalize, k=k, pk=pk, n_bins=n_bins,
                               want_impurity=False, want_maps=False, radius64=r64, radius_stats64=st64)
    out = {}
    if want_score:
        out["score"] = score.clone()
    n_regions = cfg.regions_per_image(H, W)
    n_picked, picks = select_planes(score, active, selected, active_mask, gt, n_regions, cfg.radius_k,
                                    cfg.mask_radius_k, want_picks=want_picks, keep_score=True)
    out["n_picked"] = n_picked
    out["n_regions"] = n_regions
    if want_picks:
        out["picks"] = picks
    return out


def shard_range(n_items, rank, world):
    """Contiguous block of image indices owned by `rank` (last blocks may be one shorter)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def acquire_pool(provider, n_images, P, A, cfg, *, batch_size=148, group=None, gather=True, device=None, masks=None,
                 exchange=None, keep_local=True, verify=False):
    """Run one acquisition round over a pool of `n_images` images sharded by image over the ranks of `group`
    (the reference's loop over the target pool, core/active/build.py:92, which it runs on rank 0 alone,
    core/train_learners.py:307-326).

    provider(lo, hi) -> dict(feat (b,C,H,W) fp32, gt, active, selected, active_mask (b,H,W) uint8), tensors on this
    rank's device for global image indices [lo, hi); the three state planes are updated in place.
    Every batch runs K1 -> K2 -> K3 and packs its picks into this shard's row buffer; there is NO collective and no host
    synchronisation until the shard is done.  Then ONE all-gather of the packed rows (on a side stream) and every rank
    replays all rows onto `masks`, its (n_images,H,W) uint8 replica of the pool's label masks (allocated, all 255, when not
    given; labels of earlier rounds stay).  Returns dict(rank, world, range, n_picked (n_images,) int32, active_mask
    (n_images,H,W) [, *_local planes when keep_local] [, checksum, replicas_agree when verify])."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    lo, hi = shard_range(n_images, rank, world)
    counts, lmasks, actives, selecteds = [], [], [], []
    ex = exchange
    for b0 in range(lo, hi, batch_size):
        b1 = min(b0 + batch_size, hi)
        item = provider(b0, b1)
        H, W = item["feat"].shape[-2:]
        if gather and ex is None:
            ex = RoundExchange(n_images, H, W, cfg.regions_per_image(H, W), cfg.radius_k, item["feat"].device, group=group)
        res = acquire_batch(item["feat"], P, A, cfg, item["gt"], item["active"], item["selected"], item["active_mask"],
                            want_picks=gather)
        if gather:
            ex.pack(b0 - lo, res["picks"], res["n_picked"], item["gt"])
        if keep_local:
            counts.append(res["n_picked"])
            lmasks.append(item["active_mask"])
            actives.append(item["active"])
            selecteds.append(item["selected"])
    out = {"rank": rank, "world": world, "range": (lo, hi)}
    if counts:
        out["n_picked_local"] = torch.cat(counts)
        out["active_mask_local"] = torch.cat(lmasks)
        out["active_local"] = torch.cat(actives)
        out["selected_local"] = torch.cat(selecteds)
    if gather:
        if ex is None:   # this rank owns no image (world > n_images) and was given no exchange: it still joins the gather
            raise ValueError("acquire_pool: a rank that owns no image needs `exchange=RoundExchange(...)` (it cannot "
                             "infer the image size)")
        if masks is None:
            masks = torch.full((n_images, ex.H, ex.W), 255, dtype=torch.uint8, device=ex.device)
        n_picked = torch.zeros((n_images,), dtype=torch.int32, device=ex.device)
        ex.exchange(masks, n_picked)
        ex.wait()
        out["n_picked"], out["active_mask"] = n_picked, masks
        if verify:
            out["checksum"], out["replicas_agree"] = ex.verify(masks, n_picked)
    return out


class RoundExchange:
    """The one exchange at the end of a sharded round: preallocated packed rows, ONE all-gather, replay.

    Per pool image one row [count | picks[cap] | window labels] (`halo_round_row_bytes`): 13 B per pick for 3x3 regions,
    59 KB per 1280x640 image at 4 552 picks instead of the 819 KB mask plane.  `cap` comes from the configuration
    (`AcquisitionConfig.regions_per_image`), so nothing is negotiated between the ranks and nothing synchronises with the
    host.  `pack` runs on the caller's stream right after a batch's selection; `exchange` runs the all-gather and the
    replay on a side stream behind an event, so it overlaps whatever the caller enqueues next; `wait` joins it.
    `pack_fn` / `apply_fn` default to the CUDA entry points; the world_size-2 gloo tests (no GPU) pass the oracle's
    restatement to exercise the sharding and the collective on CPU tensors."""

    def __init__(self, n_images, H, W, cap, active_radius, device, *, group=None, pack_fn=None, apply_fn=None, depth=2):
        import torch.distributed as dist

        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.n_images, self.H, self.W = int(n_images), int(H), int(W)
        self.cap, self.r = max(int(cap), 1), int(active_radius)
        self.device = torch.device(device)
        self.per = (self.n_images + self.world - 1) // self.world
        k2 = (2 * self.r + 1) ** 2
        self.row_bytes = (4 + 4 * self.cap + self.cap * k2 + 15) // 16 * 16
        self.pack_fn, self.apply_fn = pack_fn or pack_rows, apply_fn or apply_rows
        # `depth` send buffers used round-robin, so that the next round can be packed while the previous round's all-gather
        # is still reading its buffer (back-to-back rounds never wait for the slowest rank of the last one).  Zeroed once:
        # padding rows keep count = 0 for ever, real rows are rewritten by every round's pack.
        self.depth = max(1, int(depth))
        self._rows = [torch.zeros((self.per, self.row_bytes), dtype=torch.uint8, device=self.device) for _ in range(self.depth)]
        self._all = [(torch.empty((self.world * self.per, self.row_bytes), dtype=torch.uint8, device=self.device)
                      if self.distributed else self._rows[k]) for k in range(self.depth)]
        self.row_image = _row_maps(self.n_images, self.world, self.device)[0]
        self.cuda = self.device.type == "cuda"
        self.side = torch.cuda.Stream(device=self.device) if self.cuda else None
        self._done = [torch.cuda.Event() if self.cuda else None for _ in range(self.depth)]
        self._used = [False] * self.depth
        self._cur = 0           # buffer the current round packs into
        self._last = None       # buffer of the last exchange issued (what wait() joins)
        self._fresh = True      # nothing packed into the current buffer yet this round

    @property
    def rows_local(self):
        return self._rows[self._cur]

    def pack(self, local_row, picks, n_picked, gt):
        """Rows [local_row, local_row + b) of this shard <- the picks of a batch of b images (caller's stream)."""
        b = picks.shape[0]
        if self._fresh and self.cuda and self._used[self._cur]:
            # first pack of a round into a buffer an earlier exchange sent from: that exchange must have finished with it
            torch.cuda.current_stream(self.device).wait_event(self._done[self._cur])
        self._fresh = False
        self.pack_fn(self._rows[self._cur][local_row:local_row + b], picks, n_picked, gt, self.cap, self.r)

    def exchange(self, masks, n_picked_out):
        """All-gather the shard rows (one collective) and replay every rank's rows onto `masks` / `n_picked_out`."""
        import torch.distributed as dist

        k = self._cur
        rows, rows_all = self._rows[k], self._all[k]

        def run():
            if self.distributed:
                dist.all_gather_into_tensor(rows_all, rows, group=self.group)
            self.apply_fn(masks, self.row_image, rows_all, n_picked_out, self.cap, self.r)

        if self.cuda:
            self.side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side):
                run()
                self._done[k].record(self.side)
        else:
            run()
        self._used[k] = True
        self._last = k
        self._cur = (k + 1) % self.depth
        self._fresh = True

    def wait(self):
        """The caller's stream waits for the last exchange issued (masks / counts are then complete on it)."""
        if self.cuda and self._last is not None:
            torch.cuda.current_stream(self.device).wait_event(self._done[self._last])

    def bytes_per_rank(self):
        return self.per * self.row_bytes

    def verify(self, masks, n_picked, checksum_fn=None):
        """64-bit checksum of (masks, n_picked) on this rank and whether every rank holds the same one.
        Synchronises with the host: call it outside timed regions."""
        import torch.distributed as dist

        cs = (checksum_fn or checksum64)(masks, n_picked)
        agree = True
        if self.distributed:
            mine = cs.view(torch.int64).reshape(1)
            allc = torch.empty((self.world,), dtype=torch.int64, device=mine.device)
            dist.all_gather_into_tensor(allc, mine, group=self.group)
            agree = bool((allc == allc[0]).all().item())
        return int(cs.view(torch.int64).item()) & 0xFFFFFFFFFFFFFFFF, agree


def pack_rows(rows, picks, n_picked, gt, cap, active_radius):
    """`halo_round_rows_pack`: rows (b,row_bytes) uint8 view <- picks (b,stride) int32, n_picked (b,), gt (b,H,W)."""
    nat.require_cuda(picks, "picks")
    lib = nat.load()
    b, stride = picks.shape
    H, W = gt.shape[-2:]
    with torch.cuda.device(picks.device):
        rc = lib.halo_round_rows_pack(nat.ptr(picks), nat.ptr(n_picked), nat.ptr(gt), nat.ptr(rows), b, int(cap), stride, H, W,
                                      int(active_radius), nat.stream_of(picks))
    nat.check(rc, "halo_round_rows_pack")


def apply_rows(masks, row_image, rows, n_picked_out, cap, active_radius):
    """`halo_round_rows_apply`: replay packed rows onto masks (n_images,H,W) uint8 and scatter the counts."""
    nat.require_cuda(masks, "masks")
    lib = nat.load()
    H, W = masks.shape[-2:]
    with torch.cuda.device(masks.device):
        rc = lib.halo_round_rows_apply(nat.ptr(masks), nat.ptr(row_image), nat.ptr(rows), nat.ptr(n_picked_out), rows.shape[0],
                                       int(cap), H, W, int(active_radius), nat.stream_of(masks))
    nat.check(rc, "halo_round_rows_apply")


def checksum64(masks, n_picked):
    """Device uint64 (stored in an int64 tensor) checksum of the masks followed by the counts (`halo_checksum64`)."""
    nat.require_cuda(masks, "masks")
    lib = nat.load()
    out = torch.empty((1,), dtype=torch.int64, device=masks.device)
    m = masks.contiguous()
    c = n_picked.contiguous()
    with torch.cuda.device(masks.device):
        st = nat.stream_of(masks)
        rc = lib.halo_checksum64(nat.ptr(m), m.numel() * m.element_size(), 0, nat.ptr(out), 0, st)
        nat.check(rc, "halo_checksum64")
        rc = lib.halo_checksum64(nat.ptr(c), c.numel() * c.element_size(), (m.numel() * m.element_size() + 7) // 8, nat.ptr(out),
                                 1, st)
        nat.check(rc, "halo_checksum64")
    return out


def gather_round(n_picked_local, mask_local, n_images, *, group=None, device=None):
    """All-gather the per-shard pick counts and masks (the only collective on the path).

    Shards are padded to the largest shard so a single all_gather_into_tensor per tensor suffices; works on
    NCCL (CUDA tensors) and gloo (CPU tensors, used by the world_size-2 CPU tests)."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    if not distributed:
        return {"n_picked": n_picked_local, "active_mask": mask_local}
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    per = (n_images + world - 1) // world
    lo, hi = shard_range(n_images, rank, world)
    n_local = hi - lo
    if mask_local is not None:
        dev, hw = mask_local.device, tuple(mask_local.shape[1:])
    else:   # a rank that owns no image: NCCL needs CUDA tensors, gloo CPU ones
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
        dev, hw = torch.device(device), None
    # agree on the plane shape (a rank may own zero images when world > n_images)
    shape_t = torch.tensor(list(hw) if hw else [0, 0], dtype=torch.int64, device=dev)
    dist.all_reduce(shape_t, op=dist.ReduceOp.MAX, group=group)
    hw = (int(shape_t[0]), int(shape_t[1]))
    cnt_pad = torch.zeros((per,), dtype=torch.int32, device=dev)
    msk_pad = torch.full((per,) + hw, 255, dtype=torch.uint8, device=dev)
    if n_local:
        cnt_pad[:n_local] = n_picked_local
        msk_pad[:n_local] = mask_local
    cnt_all = torch.empty((world * per,), dtype=torch.int32, device=dev)
    msk_all = torch.empty((world * per,) + hw, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(cnt_all, cnt_pad, group=group)
    dist.all_gather_into_tensor(msk_all, msk_pad, group=group)
    keep = []
    for r in range(world):
        a, b = shard_range(n_images, r, world)
        keep.extend(range(r * per, r * per + (b - a)))
    keep_t = torch.tensor(keep, dtype=torch.long, device=dev)
    return {"n_picked": cnt_all.index_select(0, keep_t), "active_mask": msk_all.index_select(0, keep_t)}


# ---- compact exchange: round deltas instead of mask planes -----------------------------------------------------
def pack_round_delta(picks, n_picked, gt, active_radius):
    """Labels a round wrote around its picks: lab (N,cap,(2a+1)^2) uint8, 255 outside the image / beyond n_picked
    (build.py:58-62: active_mask[window] = ground_truth[window]) -- halo_round_delta_pack.  CUDA tensors only."""
    nat.require_cuda(picks, "picks")
    N, cap = picks.shape
    H, W = gt.shape[-2:]
    r = int(active_radius)
    k = 2 * r + 1
    lib = nat.load()
    lab = torch.empty((N, cap, k * k), dtype=torch.uint8, device=picks.device)
    with torch.cuda.device(picks.device):
        rc = lib.halo_round_delta_pack(nat.ptr(picks), nat.ptr(n_picked), nat.ptr(gt), nat.ptr(lab), N, cap, H, W, r,
                                       nat.stream_of(picks))
    nat.check(rc, "halo_round_delta_pack")
    return lab


def apply_round_delta(masks, row_image, picks, n_picked, lab, active_radius):
    """masks (n_images,H,W) uint8 replica, updated IN PLACE: row j of (picks, n_picked, lab) belongs to pool image
    row_image[j] (negative = padding row) -- halo_round_delta_apply.  CUDA tensors only."""
    nat.require_cuda(masks, "masks")
    rows, cap = picks.shape
    H, W = masks.shape[-2:]
    lib = nat.load()
    with torch.cuda.device(masks.device):
        rc = lib.halo_round_delta_apply(nat.ptr(masks), nat.ptr(row_image), nat.ptr(picks), nat.ptr(n_picked), nat.ptr(lab),
                                        rows, cap, H, W, int(active_radius), nat.stream_of(masks))
    nat.check(rc, "halo_round_delta_apply")
    return masks


def gather_round_delta(n_picked_local, picks_local, gt_local, masks, n_images, active_radius, *, group=None, cap=None,
                       pack=pack_round_delta, apply=apply_round_delta):
    """Compact form of `gather_round`: all-gather (pick counts, picks, window labels) -- 13 B per pick for 3x3 regions
    instead of H*W bytes per image -- and replay them onto `masks` (n_images,H,W) uint8, this rank's replica of the
    pool's label masks (the labels of earlier rounds stay).  Works without a process group (single shard) too.
    `pack` / `apply` are the two CUDA entry points above; the world_size-2 gloo test, which has no GPU, passes the
    oracle's restatement of them (oracle/delta.py) to exercise the sharding and the collective on CPU tensors.
    Superseded on the hot path by `RoundExchange` (one packed all-gather, preallocated, side stream); kept as the
    three-tensor form the round-1 tests pin.  Returns {"n_picked": (n_images,) int32, "active_mask": masks}."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    per = (n_images + world - 1) // world
    lo, hi = shard_range(n_images, rank, world)
    n_local = hi - lo
    dev = masks.device
    if cap is None:   # the pick capacity is the per-image budget, known from the configuration: pass it to skip this
        cap = picks_local.shape[1] if picks_local is not None else 0
        if distributed:   # (kept for callers that do not: a rank may own zero images, so the ranks must agree -- host sync)
            cap_t = torch.tensor([cap], dtype=torch.int64, device=dev)
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX, group=group)
            cap = int(cap_t[0])
    if cap == 0:
        return {"n_picked": torch.zeros((n_images,), dtype=torch.int32, device=dev), "active_mask": masks}
    k2 = (2 * int(active_radius) + 1) ** 2
    cnt_pad = torch.zeros((per,), dtype=torch.int32, device=dev)
    pk_pad = torch.full((per, cap), -1, dtype=torch.int32, device=dev)
    lab_pad = torch.full((per, cap, k2), 255, dtype=torch.uint8, device=dev)
    if n_local:
        cnt_pad[:n_local] = n_picked_local
        pk_pad[:n_local] = picks_local
        lab_pad[:n_local] = pack(picks_local, n_picked_local, gt_local, active_radius)
    if distributed:
        cnt_all = torch.empty((world * per,), dtype=torch.int32, device=dev)
        pk_all = torch.empty((world * per, cap), dtype=torch.int32, device=dev)
        lab_all = torch.empty((world * per, cap, k2), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(cnt_all, cnt_pad, group=group)
        dist.all_gather_into_tensor(pk_all, pk_pad, group=group)
        dist.all_gather_into_tensor(lab_all, lab_pad, group=group)
    else:
        cnt_all, pk_all, lab_all = cnt_pad, pk_pad, lab_pad
    row_image, keep_t = _row_maps(n_images, world, dev)
    apply(masks, row_image, pk_all, cnt_all, lab_all, active_radius)
    return {"n_picked": cnt_all.index_select(0, keep_t), "active_mask": masks}


_ROW_MAPS = {}


def _row_maps(n_images, world, dev):
    """Row j of the gathered (padded) buffers -> pool image index (-1 for padding rows), and the rows to keep; cached
    so that a steady-state round issues no host->device copy."""
    key = (n_images, world, str(dev))
    if key not in _ROW_MAPS:
        per = (n_images + world - 1) // world
        row_image = torch.full((world * per,), -1, dtype=torch.int32)
        keep = []
        for r in range(world):
            a, b = shard_range(n_images, r, world)
            row_image[r * per:r * per + (b - a)] = torch.arange(a, b, dtype=torch.int32)
            keep.extend(range(r * per, r * per + (b - a)))
        _ROW_MAPS[key] = (row_image.to(dev), torch.tensor(keep, dtype=torch.long, device=dev))
    return _ROW_MAPS[key]
