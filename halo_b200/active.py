"""Budgeted selection + acquisition driver -- B200 mirror of the reference's `core/active/build.py`.

`select_pixels_to_label` (:27-64) keeps the reference's signature and its in-place semantics on all four
tensors (the reference hands it a CUDA score, CPU bool `active`/`selected`, CUDA int64 `active_mask`/
`ground_truth`); the sequential arg-max loop is replaced by the exact parallel form in `halo_select_*`.
`RegionSelection` (:71-186) keeps its signature and side effects (eval()/train() toggles, uint8 mask PNG at
`path_to_mask`, `{"active","selected"}` indicator at `path_to_indicator`).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _native as nat
from .floating_region import FloatingRegionScore
from .hyperbolic import PoincareEmbedding
from .maskio import AsyncMaskWriter


def select_planes(score, active, selected, active_mask, gt, n_regions, active_radius, mask_radius, want_picks=False,
                  keep_score=False):
    """Batched in-place selection on device planes.  score (N,H,W) f32|f64; the rest (N,H,W) uint8.
    keep_score=True skips writing the -inf suppression windows back into `score` (the acquisition path never reads it).
    Returns (n_picked (N,) int32 device tensor, picks (N,n_regions) int32 | None)."""
    lib = nat.load()
    nat.require_cuda(score, "score")
    if score.dtype not in (torch.float32, torch.float64):
        raise TypeError("select_planes: score must be float32 or float64")
    for t, name in ((active, "active"), (selected, "selected"), (active_mask, "active_mask"), (gt, "ground_truth")):
        if t.dtype != torch.uint8 or not t.is_contiguous() or t.device != score.device or t.shape != score.shape:
            raise ValueError("select_planes: %s must be a contiguous uint8 tensor shaped like score on %s" % (name, score.device))
    if not score.is_contiguous():
        raise ValueError("select_planes: score must be contiguous")
    N, H, W = score.shape
    n_regions = int(n_regions)
    dev = score.device
    n_picked = torch.empty((N,), dtype=torch.int32, device=dev)
    picks = torch.empty((N, max(n_regions, 1)), dtype=torch.int32, device=dev) if want_picks else None
    ws = nat.workspace.get(dev, "select", lib.halo_select_workspace_bytes(N, H, W, n_regions))
    fn = lib.halo_select_f64 if score.dtype == torch.float64 else lib.halo_select_f32
    with torch.cuda.device(dev):
        rc = fn(nat.ptr(score), nat.ptr(active), nat.ptr(selected), nat.ptr(active_mask), nat.ptr(gt), n_regions,
                int(active_radius), int(mask_radius), nat.SELECT_KEEP_SCORE if keep_score else 0, nat.ptr(n_picked),
                nat.ptr(picks), N, H, W, nat.ptr(ws),
                ws.numel(), nat.stream_of(score))
    nat.check(rc, "halo_select")
    return n_picked, picks


def select_pixels_to_label(score, active_regions, active_radius, mask_radius, active, selected, active_mask,
                           ground_truth):
    """Reference `select_pixels_to_label` (build.py:27-64): mutates and returns (score, active, selected, active_mask).

    Tensors may live anywhere (the reference keeps active/selected on the CPU): they are staged to uint8
    device planes, selected on the GPU, and copied back IN PLACE."""
    nat.require_cuda(score, "score")
    dev = score.device
    H, W = score.shape
    s = score if (score.is_contiguous() and score.dtype in (torch.float32, torch.float64)) else score.float().contiguous()
    act = active.to(device=dev, dtype=torch.uint8).contiguous().clone().view(1, H, W)
    sel = selected.to(device=dev, dtype=torch.uint8).contiguous().clone().view(1, H, W)
    msk = active_mask.to(device=dev, dtype=torch.uint8).contiguous().clone().view(1, H, W)
    gt = ground_truth.to(device=dev, dtype=torch.uint8).contiguous().view(1, H, W)
    select_planes(s.view(1, H, W), act, sel, msk, gt, active_regions, active_radius, mask_radius)
    if s.data_ptr() != score.data_ptr():
        score.copy_(s)
    active.copy_(act[0].to(active.dtype))
    selected.copy_(sel[0].to(selected.dtype))
    active_mask.copy_(msk[0].to(active_mask.dtype))
    return score, active, selected, active_mask


def to_np_array(tensor):
    return np.array(tensor.cpu().numpy(), dtype=np.uint8)


def _to_u8_device(t, device):
    """Loader tensors arrive on the CPU as int64 / bool planes (cityscapes.py:274-286): narrow them to uint8 BEFORE the
    host->device copy (8x fewer PCIe bytes than the reference's `.cuda()` of the int64 plane)."""
    if t.is_cuda:
        return t.to(torch.uint8)
    return t.to(torch.uint8).to(device, non_blocking=True)


def RegionSelection(cfg, feature_extractor, classifier, tgt_epoch_loader, round_number):
    """Reference `RegionSelection` (build.py:71-186): one acquisition round over the target pool on this rank.

    Same loader item contract (core/datasets/cityscapes.py:274-286) and on-disk side effects.  The per-image body runs
    the CUDA path: classifier head (fused when the classifier was built with this package's HyperMapper/HyperMLR),
    FloatingRegionScore with `score[active] = -inf` (:146) fused into its last pass, and the budgeted selection --
    deferred and run for `ACTIVE.SELECT_BATCH` (optional key, default 8) images of equal size in one launch, because the
    selection kernel gives each image one SM (images are independent: build.py:137-160 touches nothing shared)."""
    feature_extractor.eval()
    classifier.eval()

    per_region_pixels = (2 * cfg.ACTIVE.RADIUS_K + 1) ** 2
    active_radius = cfg.ACTIVE.RADIUS_K
    mask_radius = cfg.ACTIVE.MASK_RADIUS_K
    active_budget = cfg.ACTIVE.BUDGET / len(cfg.ACTIVE.SELECT_ITER)
    uncertainty_type = cfg.ACTIVE.UNCERTAINTY
    purity_type = cfg.ACTIVE.PURITY
    K = cfg.ACTIVE.K

    floating_region_score = FloatingRegionScore(in_channels=cfg.MODEL.NUM_CLASSES, size=2 * active_radius + 1,
                                                purity_type=purity_type, K=K, curvature=cfg.MODEL.CURVATURE)
    needs_embedding = (uncertainty_type in ["certainty", "hyperbolic"]
                       or purity_type in ["hyper", "radius", "euc_norm"]
                       or (uncertainty_type == "none" and cfg.MODEL.HYPER))

    def optional(key, default):   # keys the reference's config does not define
        try:
            return int(getattr(cfg.ACTIVE, key))
        except (AttributeError, KeyError):
            return default

    # PNG encoding + torch.save leave the loop: same files, written by a thread pool from pinned
    # staging buffers; flushed before this function returns, as the caller expects (train_learners.py:318-322 reloads
    # the dataset right after)
    writer = AsyncMaskWriter(workers=optional("IO_WORKERS", 4))
    select_batch = max(1, optional("SELECT_BATCH", 8))
    pending = []   # images scored and waiting for the batched selection: (score, active, selected, mask, gt, paths)

    def flush_pending():
        if not pending:
            return
        size = tuple(pending[0][0].shape)
        score = torch.stack([q[0] for q in pending])
        act, sel, msk, gt = (torch.stack([q[k] for q in pending]) for k in (1, 2, 3, 4))
        active_regions = math.ceil(size[0] * size[1] * active_budget / per_region_pixels)    # build.py:148-150
        select_planes(score, act, sel, msk, gt, active_regions, active_radius, mask_radius, keep_score=True)
        for j, q in enumerate(pending):
            writer.write(msk[j], act[j], sel[j], q[5], q[6])                                  # build.py:162-166
        pending.clear()

    with torch.no_grad(), writer:
        idx = 0
        for tgt_data in tgt_epoch_loader:
            tgt_input = tgt_data["img"].cuda(non_blocking=True)
            path2mask, path2indicator = tgt_data["path_to_mask"], tgt_data["path_to_indicator"]
            origin_mask, origin_label = tgt_data["origin_mask"], tgt_data["origin_label"]
            origin_size = tgt_data["size"]
            active_indicator, selected_indicator = tgt_data["active"], tgt_data["selected"]
            if idx == 0:
                feature_extractor.to(tgt_input.device)
                classifier.to(tgt_input.device)
            dev = tgt_input.device
            tgt_size = tgt_input.shape[-2:]
            tgt_out, decoder_out = classifier(feature_extractor(tgt_input), size=tgt_size)

            for i in range(len(origin_mask)):
                size = (int(origin_size[i][0]), int(origin_size[i][1]))
                active_mask = _to_u8_device(origin_mask[i], dev).reshape(size)
                ground_truth = _to_u8_device(origin_label[i], dev).reshape(size)
                active = _to_u8_device(active_indicator[i], dev).reshape(size)
                selected = _to_u8_device(selected_indicator[i], dev).reshape(size)

                out_i = tgt_out[i:i + 1]
                emb = decoder_out[i:i + 1] if needs_embedding else None
                if tuple(out_i.shape[-2:]) == size and (emb is None or tuple(emb.shape[-2:]) == size):
                    # already at label size (reference: F.interpolate to the same size is the identity)
                    score, _, _ = floating_region_score(out_i, decoder_out=emb, normalize=cfg.ACTIVE.NORMALIZE,
                                                        unc_type=uncertainty_type, pur_type=purity_type,
                                                        ground_truth=ground_truth, active=active)
                else:
                    # reference build.py:122-135 up-samples logits and (fp64) embedding -- possibly from different
                    # resolutions (classifier.py:556-557) -- THEN scores; fused here, nothing is materialised
                    score, _, _ = floating_region_score.forward_upsampled(
                        out_i, emb, size, normalize=cfg.ACTIVE.NORMALIZE, unc_type=uncertainty_type,
                        pur_type=purity_type, ground_truth=ground_truth, active=active)
                if pending and tuple(pending[0][0].shape) != size:
                    flush_pending()
                pending.append((score, active, selected, active_mask, ground_truth, path2mask[i], path2indicator[i]))
                if len(pending) >= select_batch:
                    flush_pending()
            idx += 1
        flush_pending()

    feature_extractor.train()
    classifier.train()
