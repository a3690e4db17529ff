"""Floating-region acquisition score -- B200 mirror of the reference's `core/active/floating_region.py`.

`FloatingRegionScore` keeps the reference's constructor and forward signature (:26-29, :129-137) and its
quirks (entropy / log 19 whatever the class count; zero-padded box SUM; `count` = window population only for
the histogram purities; a module BUILT for "hyper" uses a 3x3 purity window whatever `size`, :54-55).
All arithmetic is `halo_logits_stats` + `halo_ball_norm` / the fused head's radius + `halo_score`.
Results are fp32 (the reference returns an fp64 score in "radius" mode because its radius is fp64).
"""
import torch
import torch.nn as nn

from . import _native as nat
from .hyperbolic import HyperMapper, PoincareEmbedding

_HIST_PURITIES = ("ripu", "oracle_ripu", "hyper")
_KNOWN_PURITIES = ("ripu", "oracle_ripu", "hyper", "none", "radius", "euc_norm")


def _reference_cfg(path, default):
    """Read a key of the reference's global yacs cfg when running inside its tree (floating_region.py:8,39,68)."""
    try:
        from core.configs import cfg  # type: ignore

        node = cfg
        for part in path.split("."):
            node = getattr(node, part)
        return node
    except Exception:
        return default


def radius_f64(feat, c, kind):
    """(N,H,W) fp64 radius plane + (N,2) fp64 {min,max}: the reference's own precision for the "hyper" bins
    (floating_region.py:94-110) -- `halo_radius_f64`.  kind: "tangent" (raw fp32 features) or "ball" (fp32/fp64 points)."""
    lib = nat.load()
    nat.require_cuda(feat, "decoder_out")
    if kind == "tangent":
        feat, fk = feat.detach().float().contiguous(), nat.FEAT_TANGENT_F32
    elif feat.dtype == torch.float64:
        feat, fk = feat.detach().contiguous(), nat.FEAT_BALL_F64
    else:
        feat, fk = feat.detach().float().contiguous(), nat.FEAT_BALL_F32
    N, C, H, W = feat.shape
    r64 = torch.empty((N, H, W), dtype=torch.float64, device=feat.device)
    st64 = torch.empty((N, 2), dtype=torch.float64, device=feat.device)
    with torch.cuda.device(feat.device):
        rc = lib.halo_radius_f64(nat.ptr(feat), fk, float(c), nat.ptr(r64), nat.ptr(st64), N, C, H, W, nat.stream_of(feat))
    nat.check(rc, "halo_radius_f64")
    return r64, st64


def score_planes(pixunc, radius, radius_stats, label, active, *, unc_mode, pur_mode, normalize, k, pk, n_bins,
                 want_impurity=True, want_maps=True, radius64=None, radius_stats64=None):
    """Batched `halo_score` on device planes (N,H,W).  Returns (score, impurity|None, uncertainty)."""
    lib = nat.load()
    ref = pixunc if pixunc is not None else (radius if radius is not None else radius64)
    if ref is None:
        raise ValueError("score_planes: need at least one input plane")
    N, H, W = ref.shape
    dev = ref.device
    score = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    unc = torch.empty((N, H, W), dtype=torch.float32, device=dev)
    need_imp = want_impurity or pur_mode in (nat.PUR_LABEL_HIST, nat.PUR_RADIUS_BINS)
    imp = torch.empty((N, H, W), dtype=torch.float32, device=dev) if need_imp else None
    ws = nat.workspace.get(dev, "score", lib.halo_score_workspace_bytes(N))
    with torch.cuda.device(dev):
        rc = lib.halo_score(nat.ptr(pixunc), nat.ptr(radius), nat.ptr(radius_stats), nat.ptr(radius64),
                            nat.ptr(radius_stats64), nat.ptr(label), nat.ptr(active),
                            unc_mode, pur_mode, (1 if want_maps else 2) if normalize else 0, k, pk, n_bins, nat.ptr(score), nat.ptr(imp),
                            nat.ptr(unc), N, H, W, nat.ptr(ws), ws.numel(), nat.stream_of(ref))
    nat.check(rc, "halo_score")
    return score, imp, unc


def _active_plane(active, N, H, W):
    if active is None:
        return None
    nat.require_cuda(active, "active")
    if active.dtype != torch.uint8:
        active = active.to(torch.uint8)
    return active.reshape(N, H, W).contiguous()


def modes_for(unc_type, pur_type):
    """Map the reference's strings to the kernel modes (floating_region.py:70-92,158-202)."""
    if pur_type not in _KNOWN_PURITIES:
        raise NotImplementedError("Error: purity type '{}' not implemented".format(pur_type))
    if unc_type == "pixel_entropy":
        unc_mode, pixunc_mode = nat.UNC_PIXEL, "entropy"
    elif unc_type == "entropy":
        unc_mode, pixunc_mode = nat.UNC_BOXSUM, "entropy"
    elif unc_type == "oracle_acc":
        unc_mode, pixunc_mode = nat.UNC_BOXSUM, "one_minus_pgt"
    else:  # "none" and unknown strings ("hyperbolic", "certainty"): zeros (box-summed zeros are zeros)
        unc_mode, pixunc_mode = nat.UNC_ZERO, "entropy"
    if pur_type == "ripu":
        pur_mode, label_mode = nat.PUR_LABEL_HIST, "argmax"
    elif pur_type == "oracle_ripu":
        pur_mode, label_mode = nat.PUR_LABEL_HIST, "gt_filled"
    elif pur_type == "hyper":
        pur_mode, label_mode = nat.PUR_RADIUS_BINS, "argmax"
    elif pur_type == "none":
        pur_mode, label_mode = nat.PUR_ZERO, "argmax"
    else:
        pur_mode, label_mode = nat.PUR_NORM, "argmax"
    norm_mode = "euclid" if pur_type == "euc_norm" else "radius"
    return unc_mode, pixunc_mode, pur_mode, label_mode, norm_mode


class FloatingRegionScore(nn.Module):
    def __init__(self, in_channels=19, padding_mode="zeros", size=33, purity_type=None, K=100, curvature=None):
        """purity window: size x size (3 x 3 when built for 'hyper'); entropy window: size x size."""
        super(FloatingRegionScore, self).__init__()
        self.in_channels = in_channels
        assert size % 2 == 1, "error size"
        if padding_mode != "zeros":
            raise NotImplementedError("FloatingRegionScore: only zero padding is implemented (the reference default)")
        if purity_type is None:
            purity_type = _reference_cfg("ACTIVE.PURITY", "hyper")
        self.size = size
        self.purity_size = size
        self.purity_type = purity_type
        if purity_type == "hyper":
            self.K, self.purity_size = K, 3  # reference :54-55
        if curvature is None:
            curvature = _reference_cfg("MODEL.CURVATURE", 1.0)
        self.mapper = HyperMapper(c=curvature)

    def _radius_plane(self, decoder_out, norm_mode, want_stats):
        """(N,H,W) radius / euclidean-norm plane (+ per-image min/max) of the embedding."""
        c = self.mapper.c
        if isinstance(decoder_out, PoincareEmbedding):
            if norm_mode == "radius":
                r = decoder_out.poincare_radius()
                return r, (decoder_out.radius_stats() if want_stats else None)
            return decoder_out.poincare_radius("euclid"), None
        lib = nat.load()
        x = nat.require_cuda(decoder_out, "decoder_out")
        if x.dtype not in (torch.float32, torch.float64):
            x = x.float()
        x = x.contiguous()
        N, C, H, W = x.shape
        out = torch.empty((N, H, W), dtype=torch.float32, device=x.device)
        stats = torch.empty((N, 4), dtype=torch.float32, device=x.device) if want_stats else None
        with torch.cuda.device(x.device):
            rc = lib.halo_ball_norm(nat.ptr(x), 1 if x.dtype == torch.float64 else 0, float(c),
                                    nat.NORM_EUCLID if norm_mode == "euclid" else nat.NORM_RADIUS, nat.ptr(out),
                                    nat.ptr(stats), N, C, H, W, nat.stream_of(x))
        nat.check(rc, "halo_ball_norm")
        return out, stats

    def forward(self, logit: torch.Tensor, decoder_out=None, unc_type: str = None, pur_type: str = None,
                normalize: bool = False, ground_truth=None, active=None):
        """Compute region score, impurity and uncertainty (reference :129-217).

        logit: (1,O,H,W) (a batch (N,O,H,W) is accepted as an extension and returns (N,H,W) maps);
        decoder_out: PoincareEmbedding or (N,C,H,W) tensor on the ball (needed by 'radius'/'hyper'/'euc_norm').
        active (extension): (N,H,W) / (H,W) uint8 device plane; score[active != 0] = -inf is fused into the last pass
        (what `RegionSelection` does next, core/active/build.py:146).
        Returns (score, region_impurity, prediction_uncertainty), each (H,W)."""
        lib = nat.load()
        nat.require_cuda(logit, "logit")
        unc_mode, pixunc_mode, pur_mode, label_mode, norm_mode = modes_for(unc_type, pur_type)
        if pur_type in ("ripu", "oracle_ripu") and self.purity_type == "hyper":
            raise RuntimeError("FloatingRegionScore built for 'hyper' purity cannot score '%s' "
                               "(the reference's purity_conv has K channels)" % pur_type)
        if pur_type == "hyper" and self.purity_type != "hyper":
            raise AttributeError("'FloatingRegionScore' object has no attribute 'K' (module was not built for 'hyper')")
        logit = logit.float().contiguous()
        N, O, H, W = logit.shape
        dev = logit.device
        gt8 = None
        if pixunc_mode == "one_minus_pgt" or label_mode == "gt_filled":
            if ground_truth is None:
                raise ValueError("ground_truth is required by unc_type/pur_type '%s'/'%s'" % (unc_type, pur_type))
            gt8 = nat.require_cuda(ground_truth, "ground_truth").to(torch.uint8).reshape(N, H, W).contiguous()
        need_pixunc = unc_mode != nat.UNC_ZERO
        need_label = pur_mode == nat.PUR_LABEL_HIST
        pixunc = torch.empty((N, H, W), dtype=torch.float32, device=dev) if need_pixunc else None
        label = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if need_label else None
        if need_pixunc or need_label:
            with torch.cuda.device(dev):
                rc = lib.halo_logits_stats(nat.ptr(logit), nat.ptr(gt8), nat.PIXUNC_ONE_MINUS_PGT if pixunc_mode == "one_minus_pgt" else nat.PIXUNC_ENTROPY,
                                           nat.LABEL_GT_FILLED if label_mode == "gt_filled" else nat.LABEL_ARGMAX,
                                           nat.ptr(pixunc), nat.ptr(label), N, O, H, W, nat.stream_of(logit))
            nat.check(rc, "halo_logits_stats")
        radius = stats = r64 = st64 = None
        if pur_mode in (nat.PUR_NORM, nat.PUR_RADIUS_BINS) and decoder_out is None:
            raise ValueError("decoder_out is required by pur_type '%s'" % pur_type)
        if pur_mode == nat.PUR_NORM:
            radius, stats = self._radius_plane(decoder_out, norm_mode, False)
        elif pur_mode == nat.PUR_RADIUS_BINS:   # bins follow the reference's fp64 arithmetic (:94-110)
            if isinstance(decoder_out, PoincareEmbedding):
                r64, st64 = radius_f64(decoder_out.u, self.mapper.c, "tangent")
            else:
                r64, st64 = radius_f64(decoder_out, self.mapper.c, "ball")
        n_bins = self.K if pur_mode == nat.PUR_RADIUS_BINS else self.in_channels
        if pixunc is None and radius is None and r64 is None:  # "none"/"none": the kernel still needs a shape carrier
            pixunc = torch.zeros((N, H, W), dtype=torch.float32, device=dev)
        score, imp, unc = score_planes(pixunc, radius, stats, label, _active_plane(active, N, H, W), unc_mode=unc_mode,
                                       pur_mode=pur_mode, normalize=normalize, k=self.size, pk=self.purity_size,
                                       n_bins=n_bins, want_impurity=True, radius64=r64, radius_stats64=st64)
        if N == 1:
            return score[0], imp[0], unc[0]
        return score, imp, unc

    def forward_upsampled(self, logit_lr, decoder_out_lr, size, unc_type=None, pur_type=None, normalize=False,
                          ground_truth=None, active=None):
        """Score at label resolution from LOW-resolution logits and embedding (extension).

        Equivalent to the reference sequence of `RegionSelection` (core/active/build.py:122-144):
            output = F.interpolate(logit_lr, size, mode="bilinear", align_corners=True)
            decoder_out = F.interpolate(decoder_out_lr, size, mode="bilinear", align_corners=True)
            floating_region_score(output, decoder_out=decoder_out, ...)
        but the two up-sampled tensors (O and C channels at label size, the latter float64 in the reference) are never
        materialised: `halo_upsample_score_inputs` interpolates per output pixel and writes the three planes K2 needs.
        decoder_out_lr: PoincareEmbedding (raw features, exp-map applied at the low-resolution pixels) or a tensor of
        ball points (fp32/fp64), or None."""
        lib = nat.load()
        nat.require_cuda(logit_lr, "logit")
        unc_mode, pixunc_mode, pur_mode, label_mode, norm_mode = modes_for(unc_type, pur_type)
        if pur_type in ("ripu", "oracle_ripu") and self.purity_type == "hyper":
            raise RuntimeError("FloatingRegionScore built for 'hyper' purity cannot score '%s'" % pur_type)
        logit_lr = logit_lr.float().contiguous()
        N, O, h, w = logit_lr.shape
        H, W = int(size[0]), int(size[1])
        dev = logit_lr.device
        gt8 = None
        if pixunc_mode == "one_minus_pgt" or label_mode == "gt_filled":
            if ground_truth is None:
                raise ValueError("ground_truth is required by unc_type/pur_type '%s'/'%s'" % (unc_type, pur_type))
            gt8 = nat.require_cuda(ground_truth, "ground_truth").to(torch.uint8).reshape(N, H, W).contiguous()
        need_pixunc = unc_mode != nat.UNC_ZERO
        need_label = pur_mode == nat.PUR_LABEL_HIST
        need_radius = pur_mode in (nat.PUR_NORM, nat.PUR_RADIUS_BINS)
        pixunc = torch.empty((N, H, W), dtype=torch.float32, device=dev) if need_pixunc else None
        label = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if need_label else None
        radius = torch.empty((N, H, W), dtype=torch.float32, device=dev) if need_radius else None
        stats = r64 = st64 = None
        if pur_mode == nat.PUR_RADIUS_BINS:   # bins follow the reference's fp64 arithmetic (:94-110)
            r64 = torch.empty((N, H, W), dtype=torch.float64, device=dev)
            st64 = torch.empty((N, 2), dtype=torch.float64, device=dev)
        emb, emb_kind, C = None, nat.FEAT_BALL_F32, 0
        if need_radius:
            if decoder_out_lr is None:
                raise ValueError("decoder_out is required by pur_type '%s'" % pur_type)
            if isinstance(decoder_out_lr, PoincareEmbedding):
                emb, emb_kind = decoder_out_lr.u.detach().float().contiguous(), nat.FEAT_TANGENT_F32
            else:
                emb = nat.require_cuda(decoder_out_lr, "decoder_out")
                if emb.dtype == torch.float64:
                    emb_kind = nat.FEAT_BALL_F64
                else:
                    emb, emb_kind = emb.float(), nat.FEAT_BALL_F32
                emb = emb.contiguous()
            if emb.shape[0] != N:
                raise ValueError("logits and embedding must share the batch size")
            C = emb.shape[1]
        eh, ew = (int(emb.shape[-2]), int(emb.shape[-1])) if emb is not None else (h, w)
        ws = nat.workspace.get(dev, "upsample", lib.halo_upsample_workspace_bytes(N, eh, ew))
        with torch.cuda.device(dev):
            rc = lib.halo_upsample_score_inputs(
                nat.ptr(logit_lr) if (need_pixunc or need_label) else None, nat.ptr(emb), emb_kind, float(self.mapper.c),
                nat.ptr(gt8), nat.PIXUNC_ONE_MINUS_PGT if pixunc_mode == "one_minus_pgt" else nat.PIXUNC_ENTROPY,
                nat.LABEL_GT_FILLED if label_mode == "gt_filled" else nat.LABEL_ARGMAX,
                nat.NORM_EUCLID if norm_mode == "euclid" else nat.NORM_RADIUS, nat.ptr(pixunc), nat.ptr(label),
                nat.ptr(radius), nat.ptr(stats), nat.ptr(r64), nat.ptr(st64), N, O, C, h, w, eh, ew, H, W, nat.ptr(ws),
                ws.numel(), nat.stream_of(logit_lr))
        nat.check(rc, "halo_upsample_score_inputs")
        n_bins = self.K if pur_mode == nat.PUR_RADIUS_BINS else self.in_channels
        if pixunc is None and radius is None:
            pixunc = torch.zeros((N, H, W), dtype=torch.float32, device=dev)
        score, imp, unc = score_planes(pixunc, radius, stats, label, _active_plane(active, N, H, W), unc_mode=unc_mode,
                                       pur_mode=pur_mode, normalize=normalize, k=self.size, pk=self.purity_size,
                                       n_bins=n_bins, want_impurity=True, radius64=r64, radius_stats64=st64)
        if N == 1:
            return score[0], imp[0], unc[0]
        return score, imp, unc

