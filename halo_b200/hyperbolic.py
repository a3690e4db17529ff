"""Poincare-ball classifier head -- B200 mirror of the reference's `core/utils/hyperbolic.py`.

Same names and call surface as the reference (HyperMapper :16-97, HyperMLR :100-188) so that
`core/models/classifier.py:361-362,372-373,481-482,553-554` run unmodified on top of it:

    decoder_out = self.mapper.expmap(decoder_out, dim=1)          # -> PoincareEmbedding (lazy)
    out = self.conv_seg(decoder_out.double()).float()             # -> ONE fused CUDA pass over the raw features

`HyperMapper.expmap` on an (N,C,H,W) fp32 CUDA tensor returns a lazy `PoincareEmbedding` that remembers the raw
features; `HyperMLR.forward` recognises it and runs expmap0 + project + MLR logits + radius in a single kernel
(`halo_head_fwd`).  Any other use of the handle (F.interpolate, .cpu(), indexing, TEST.SAVE_EMBED ...) materialises
the embedding through `halo_expmap0_project`.  Dtype contract: the reference computes the head in float64 and
returns fp64 embeddings; here logits / radius are fp32 (within 1e-5 of the fp64 oracle, tests/test_head_gpu.py)
and a materialised embedding is fp64 like the reference's.  There is no CPU path.
"""
import math

import torch
import torch.nn as nn
from torch.nn.init import kaiming_uniform_
from torch.nn.parameter import Parameter

from . import _native as nat

PROJ_EPS = 1e-3  # reference hyperbolic.py:13 (the MLR's own projection radius)

_PIXUNC = {"entropy": nat.PIXUNC_ENTROPY, "one_minus_pgt": nat.PIXUNC_ONE_MINUS_PGT}
_LABEL = {"argmax": nat.LABEL_ARGMAX, "gt_filled": nat.LABEL_GT_FILLED}
_NORM = {"radius": nat.NORM_RADIUS, "euclid": nat.NORM_EUCLID}


def _as_u8(t, name):
    if t is None:
        return None
    nat.require_cuda(t, name)
    if t.dtype != torch.uint8:
        t = t.to(torch.uint8)
    return t.contiguous()


def head_forward(feat, P, A, c=1.0, *, kind="tangent", want_logits=True, want_radius=False, want_pixunc=False,
                 want_label=False, want_stats=False, want_saved=False, gt=None, pixunc_mode="entropy",
                 label_mode="argmax", norm_mode="radius", tensor_cores=True):
    """One fused pass of the head over `feat` (N,C,H,W).  Returns a dict with the requested planes.

    kind: "tangent" (raw fp32 features, expmap fused) or "ball" (points already on the ball, fp32/fp64).
    tensor_cores=False keeps the contraction on the fp32 CUDA cores (the tcgen05 3xTF32 path is the default
    whenever the shape allows it: raw features, C % 32 == 0, C <= 256, H*W % 4 == 0).
    Wraps `halo_head_fwd` (include/halo_b200.h); replaces hyperbolic.py:28-39,74-83,120-188 and the
    softmax-entropy / argmax prologue of floating_region.py:151-166.
    want_saved: also return `saved`, the per-pixel contractions a training forward keeps for `head_backward` (None when
    the shape cannot save: halo_head_saved_rows == 0)."""
    lib = nat.load()
    nat.require_cuda(feat, "feat")
    if feat.dim() != 4:
        raise ValueError("head_forward: feat must be (N,C,H,W), got %s" % (tuple(feat.shape),))
    if kind == "tangent":
        feat = feat.contiguous() if feat.dtype == torch.float32 else feat.float().contiguous()
        fk = nat.FEAT_TANGENT_F32 | (0 if tensor_cores else nat.FEAT_FLAG_NO_TENSOR_CORE)
    elif kind == "ball":
        if feat.dtype == torch.float64:
            fk = nat.FEAT_BALL_F64
        else:
            feat = feat.float()
            fk = nat.FEAT_BALL_F32
        feat = feat.contiguous()
    else:
        raise ValueError("head_forward: kind must be 'tangent' or 'ball'")
    N, C, H, W = feat.shape
    O = P.shape[0]
    if tuple(P.shape) != (O, C) or tuple(A.shape) != (O, C):
        raise ValueError("head_forward: P/A must be (O,C)=(%d,%d), got %s %s" % (O, C, tuple(P.shape), tuple(A.shape)))
    dev = feat.device
    Pf = P.detach().to(device=dev, dtype=torch.float32).contiguous()
    Af = A.detach().to(device=dev, dtype=torch.float32).contiguous()
    out = {}
    logits = torch.empty((N, O, H, W), dtype=torch.float32, device=dev) if want_logits else None
    radius = torch.empty((N, H, W), dtype=torch.float32, device=dev) if want_radius else None
    pixunc = torch.empty((N, H, W), dtype=torch.float32, device=dev) if want_pixunc else None
    label = torch.empty((N, H, W), dtype=torch.uint8, device=dev) if want_label else None
    stats = torch.empty((N, 4), dtype=torch.float32, device=dev) if want_stats else None
    saved = None
    if want_saved and kind == "tangent" and tensor_cores:
        rows = lib.halo_head_saved_rows(C, O, H, W)
        if rows > 0 and feat.data_ptr() % 16 == 0:
            saved = torch.empty((N, rows, H, W), dtype=torch.float32, device=dev)
    gt8 = _as_u8(gt, "gt")
    need = lib.halo_head_workspace_bytes(O, C)
    ws = nat.workspace.get(dev, "head", need)
    with torch.cuda.device(dev):
        rc = lib.halo_head_fwd(nat.ptr(feat), fk, nat.ptr(Pf), nat.ptr(Af), float(c), nat.ptr(logits), nat.ptr(radius),
                               nat.ptr(pixunc), nat.ptr(label), nat.ptr(stats), nat.ptr(saved), nat.ptr(gt8), _PIXUNC[pixunc_mode],
                               _LABEL[label_mode], _NORM[norm_mode], N, C, O, H, W, nat.ptr(ws), ws.numel(),
                               nat.stream_of(feat))
    nat.check(rc, "halo_head_fwd")
    out.update(logits=logits, radius=radius, pixunc=pixunc, label=label, stats=stats, saved=saved)
    return out


def head_backward(feat, P, A, c, dlogits, saved=None):
    """Fused backward of expmap + MLR (autograd of train_learners.py:362): returns (dfeat, dP, dA), all fp32.
    saved: the planes `head_forward(..., want_saved=True)` returned for the same inputs (the backward then reads the
    features once); None = recompute them."""
    lib = nat.load()
    nat.require_cuda(feat, "feat")
    feat = feat.float().contiguous()
    dlogits = dlogits.float().contiguous()
    N, C, H, W = feat.shape
    O = P.shape[0]
    dev = feat.device
    Pf = P.detach().to(device=dev, dtype=torch.float32).contiguous()
    Af = A.detach().to(device=dev, dtype=torch.float32).contiguous()
    dfeat = torch.empty_like(feat)
    dP = torch.empty((O, C), dtype=torch.float32, device=dev)
    dA = torch.empty((O, C), dtype=torch.float32, device=dev)
    need = lib.halo_head_bwd_workspace_bytes(N, C, O, H, W)
    ws = nat.workspace.get(dev, "head_bwd", need)
    with torch.cuda.device(dev):
        rc = lib.halo_head_bwd(nat.ptr(feat), nat.ptr(Pf), nat.ptr(Af), float(c), nat.ptr(dlogits), nat.ptr(saved), nat.ptr(dfeat),
                               nat.ptr(dP), nat.ptr(dA), N, C, O, H, W, nat.ptr(ws), ws.numel(), nat.stream_of(feat))
    nat.check(rc, "halo_head_bwd")
    return dfeat, dP, dA


class _FusedHead(torch.autograd.Function):
    """logits = HyperMLR(expmap(u)) with the fused CUDA forward / backward."""

    @staticmethod
    def forward(ctx, u, P, A, c, handle):
        res = head_forward(u, P, A, c, kind="tangent", want_logits=True, want_radius=True, want_stats=True, want_saved=True)
        if handle is not None:
            handle._radius = res["radius"]
            handle._radius_stats = res["stats"]
        ctx.save_for_backward(u, P, A)
        ctx.saved_planes = res["saved"]     # S_k, T_k, |u|^2 per pixel (or None): the backward reads u once
        ctx.c = c
        return res["logits"]

    @staticmethod
    def backward(ctx, dlogits):
        u, P, A = ctx.saved_tensors
        du, dP, dA = head_backward(u, P, A, ctx.c, dlogits, saved=ctx.saved_planes)
        ctx.saved_planes = None
        return du.to(u.dtype), dP.to(P.dtype), dA.to(A.dtype), None, None


class PoincareEmbedding(object):
    """Lazy result of HyperMapper.expmap on raw (N,C,H,W) features.

    Behaves like the fp64 tensor the reference returns (hyperbolic.py:37-39) for every consumer on the
    acquisition / inference path, but only materialises it when something other than HyperMLR /
    FloatingRegionScore / poincare_distance_origin touches it."""

    def __init__(self, u, c):
        self.u = u
        self.c = float(c)
        self._radius = None
        self._radius_stats = None
        self._x = {}

    # -- tensor-like facade ------------------------------------------------------------------------
    @property
    def shape(self):
        return self.u.shape

    def size(self, *a):
        return self.u.size(*a)

    def dim(self):
        return self.u.dim()

    @property
    def device(self):
        return self.u.device

    @property
    def dtype(self):
        return torch.float64

    @property
    def requires_grad(self):
        return self.u.requires_grad

    def double(self):
        return self  # classifier.py:554 `conv_seg(decoder_out.double())` keeps the fused path

    def float(self):
        return self.materialize(torch.float32)

    def detach(self):
        e = PoincareEmbedding(self.u.detach(), self.c)
        e._radius, e._radius_stats = self._radius, self._radius_stats
        return e

    def __len__(self):
        return self.u.shape[0]

    def __getitem__(self, idx):
        # batch slicing (build.py:132 `decoder_out[i:i+1, :, :, :]`) keeps the handle lazy
        first = idx[0] if isinstance(idx, tuple) else idx
        rest = idx[1:] if isinstance(idx, tuple) else ()
        if isinstance(first, slice) and all(isinstance(r, slice) and r == slice(None) for r in rest):
            e = PoincareEmbedding(self.u[first], self.c)
            if self._radius is not None:
                e._radius = self._radius[first]
                e._radius_stats = self._radius_stats[first] if self._radius_stats is not None else None
            return e
        return self.materialize()[idx]

    def materialize(self, dtype=torch.float64):
        """The embedding as a real tensor (halo_expmap0_project).  Differentiable only through torch ops
        applied afterwards, not back to the features."""
        if dtype not in self._x:
            lib = nat.load()
            u = nat.require_cuda(self.u.detach(), "features").float().contiguous()
            N, C, H, W = u.shape
            x = torch.empty((N, C, H, W), dtype=dtype, device=u.device)
            with torch.cuda.device(u.device):
                rc = lib.halo_expmap0_project(nat.ptr(u), nat.ptr(x), 1 if dtype == torch.float64 else 0, self.c,
                                              N, C, H, W, nat.stream_of(u))
            nat.check(rc, "halo_expmap0_project")
            self._x[dtype] = x
        return self._x[dtype]

    def poincare_radius(self, norm_mode="radius"):
        """(N,H,W) fp32 distance to the origin, straight from the raw features (no embedding needed)."""
        if norm_mode == "radius" and self._radius is not None:
            return self._radius
        return _norm_from_tangent(self.u.detach(), self.c, norm_mode)[0]

    def radius_stats(self):
        if self._radius_stats is None:
            self._radius, self._radius_stats = _norm_from_tangent(self.u.detach(), self.c, "radius")
        return self._radius_stats

    def norm(self, *a, **k):
        return self.materialize().norm(*a, **k)

    def cpu(self):
        return self.materialize().cpu()

    def __getattr__(self, name):  # anything else: behave like the materialised fp64 tensor
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}

        def conv(o):
            if isinstance(o, PoincareEmbedding):
                return o.materialize()
            if isinstance(o, (list, tuple)):
                return type(o)(conv(v) for v in o)
            return o

        return func(*conv(args), **{k: conv(v) for k, v in kwargs.items()})


def _logmap0(x, c):
    """u with expmap0(u) == x for (N,C,H,W) points on the ball, in torch ops (differentiable w.r.t. x):
    u = artanh(sqrt(c)|x|) x / (sqrt(c)|x|)   (geoopt logmap0, k<0 branch; hyperbolic.py:60 is its only caller in the
    reference).  Computed in float64 and returned in float32 (what the fused head reads)."""
    xd = x.double()
    s = math.sqrt(float(c))
    n = xd.norm(dim=1, keepdim=True).clamp_min(1e-15)
    t = (s * n).clamp(max=1.0 - 1e-7)
    return (torch.atanh(t) / (s * n) * xd).float()


def _norm_from_tangent(u, c, norm_mode):
    """radius / |x| plane + per-image min/max from RAW features: r = min(2|u|, 2 artanh(1-1e-5)/sqrt(c))."""
    lib = nat.load()
    u = nat.require_cuda(u, "features").float().contiguous()
    N, C, H, W = u.shape
    # closed form on the tangent side: |x| = tanh(sqrt(c)|u|)/sqrt(c) clipped at (1-1e-5)/sqrt(c)
    n = torch.empty((N, H, W), dtype=torch.float32, device=u.device)
    stats = torch.empty((N, 4), dtype=torch.float32, device=u.device)
    # reuse the head kernel with a single dummy class: the contraction cost is negligible at O=1
    z = torch.zeros((1, C), dtype=torch.float32, device=u.device)
    ws = nat.workspace.get(u.device, "head", lib.halo_head_workspace_bytes(1, C))
    with torch.cuda.device(u.device):
        rc = lib.halo_head_fwd(nat.ptr(u), nat.FEAT_TANGENT_F32, nat.ptr(z), nat.ptr(z), float(c), None, nat.ptr(n), None,
                               None, nat.ptr(stats), None, None, 0, 0, _NORM[norm_mode], N, C, 1, H, W, nat.ptr(ws),
                               ws.numel(), nat.stream_of(u))
    nat.check(rc, "halo_head_fwd")
    return n, stats


def _ball_norm(x, c, norm_mode, dim):
    """poincare_distance_origin / |x| for a materialised tensor of any shape, reducing over `dim`."""
    lib = nat.load()
    nat.require_cuda(x, "x")
    if x.dtype not in (torch.float32, torch.float64):
        x = x.float()
    xm = x.movedim(dim, 0)                    # (C, ...)
    rest = xm.shape[1:]
    L = int(math.prod(rest)) if len(rest) else 1
    x4 = xm.reshape(1, xm.shape[0], 1, L).contiguous()
    out = torch.empty((1, 1, L), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.halo_ball_norm(nat.ptr(x4), 1 if x4.dtype == torch.float64 else 0, float(c), _NORM[norm_mode],
                                nat.ptr(out), None, 1, x4.shape[1], 1, L, nat.stream_of(x4))
    nat.check(rc, "halo_ball_norm")
    return out.reshape(rest)


class HyperMapper(object):
    """A class to map between euclidean and hyperbolic space and compute distances (reference :16-97)."""

    def __init__(self, c=1.) -> None:
        self.c = c
        self.K = torch.tensor(-self.c, dtype=float)  # reference :26 (negative curvature for geoopt)

    def expmap(self, x, dim=-1):
        """Exponential map at the origin followed by the projection onto the ball (reference :28-39).

        (N,C,H,W) fp32 CUDA features with dim=1 -> lazy PoincareEmbedding (fused later); any other
        shape / dim -> eager fp64 tensor like the reference."""
        if isinstance(x, PoincareEmbedding):
            raise TypeError("expmap: input is already a PoincareEmbedding")
        nat.require_cuda(x, "x")
        d = dim if dim >= 0 else x.dim() + dim
        if x.dim() == 4 and d == 1:
            return PoincareEmbedding(x if x.dtype == torch.float32 else x.float(), self.c)
        if torch.is_grad_enabled() and x.requires_grad:
            raise NotImplementedError(
                "HyperMapper.expmap: the eager path (anything but (N,C,H,W) features with dim=1) is forward-only; "
                "detach the input or use the (N,C,H,W) layout, whose fused head is differentiable")
        lib = nat.load()
        xm = x.detach().float().movedim(d, 0)
        rest = xm.shape[1:]
        L = int(math.prod(rest)) if len(rest) else 1
        u4 = xm.reshape(1, xm.shape[0], 1, L).contiguous()
        out = torch.empty(u4.shape, dtype=torch.float64, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.halo_expmap0_project(nat.ptr(u4), nat.ptr(out), 1, float(self.c), 1, u4.shape[1], 1, L,
                                          nat.stream_of(u4))
        nat.check(rc, "halo_expmap0_project")
        return out.reshape((xm.shape[0],) + tuple(rest)).movedim(0, d)

    def poincare_distance_origin(self, x, dim=-1):
        """Poincare distance to the origin (reference :74-83).  fp32 result."""
        if isinstance(x, PoincareEmbedding):
            d = dim if dim >= 0 else 4 + dim
            if d != 1:
                return _ball_norm(x.materialize(), self.c, "radius", d)
            return x.poincare_radius()
        d = dim if dim >= 0 else x.dim() + dim
        return _ball_norm(x, self.c, "radius", d)


class HyperMLR(nn.Module):
    """Multinomial logistic regression in hyperbolic space (reference :100-188)."""

    def __init__(self, out_channels, num_classes, c=1.):
        super().__init__()
        self.c = c
        self.K = torch.tensor(c, dtype=float)  # reference :113 (+c here, -c in the mapper)
        self.num_classes = num_classes
        # Same names / shapes as the reference (state-dict compatible; fp64 checkpoints are cast on load).
        self.P_MLR = Parameter(torch.empty((num_classes, out_channels), dtype=torch.float32))
        self.A_MLR = Parameter(torch.empty((num_classes, out_channels), dtype=torch.float32))
        kaiming_uniform_(self.P_MLR, a=math.sqrt(5))
        kaiming_uniform_(self.A_MLR, a=math.sqrt(5))

    def _hyper_logits(self, inputs):
        """(B,C,H,W) -> (B,O,H,W) logits.  `inputs` is a PoincareEmbedding (fused, differentiable) or a
        tensor of points already on the ball (reference semantics, forward only)."""
        if isinstance(inputs, PoincareEmbedding):
            u = inputs.u
            needs_grad = torch.is_grad_enabled() and (u.requires_grad or self.P_MLR.requires_grad or self.A_MLR.requires_grad)
            if needs_grad:
                return _FusedHead.apply(u, self.P_MLR, self.A_MLR, float(self.c), inputs)
            res = head_forward(u, self.P_MLR, self.A_MLR, self.c, kind="tangent", want_logits=True,
                               want_radius=True, want_stats=True)
            inputs._radius, inputs._radius_stats = res["radius"], res["stats"]
            return res["logits"]
        if torch.is_grad_enabled() and (inputs.requires_grad or self.P_MLR.requires_grad or self.A_MLR.requires_grad):
            # Differentiable route for points ALREADY on the ball (the reference's HyperMLR is differentiable w.r.t. x, P
            # and A on this path too): pull the points back to the tangent space with torch ops (logmap0, differentiable,
            # fp64 like the reference's embedding) and run the fused forward / backward on the result -- for points inside
            # the projection radius expmap0(logmap0(x)) == x, so logits and all three gradients are the reference's.
            nat.require_cuda(inputs, "inputs")
            return _FusedHead.apply(_logmap0(inputs, self.c), self.P_MLR, self.A_MLR, float(self.c), None)
        return head_forward(inputs, self.P_MLR, self.A_MLR, self.c, kind="ball", want_logits=True)["logits"]

    def forward(self, x):
        return self._hyper_logits(x)
