"""Restatement of the three geoopt functions the hot path calls (ORACLE -- test infrastructure).

Third-party dependency: ``geoopt`` (``requirements.txt:15`` of the reference, version unpinned,
not vendored, not installable here).  Call sites in the reference:
``core/utils/hyperbolic.py:37`` (expmap0), ``:38`` (project), ``:83`` (dist0),
plus cold ``:49,60,72``.  Follows geoopt ``manifolds/stereographic/math.py`` for negative
curvature ``k`` (Poincare ball): tan_k = tanh, artan_k = artanh, with geoopt's guards.
"""
import torch

_MIN_NORM = 1e-15


def _sabs(k, eps=1e-15):
    return k.abs() + eps


def _tanh(x):
    return torch.tanh(x.clamp(-15.0, 15.0))


def _artanh(x):
    x = x.clamp(-1.0 + 1e-7, 1.0 - 1e-7)
    # geoopt's Artanh.forward: (log(1+z) - log(1-z)) / 2 on the clamped input
    return 0.5 * (torch.log(1.0 + x) - torch.log(1.0 - x))


def _k(k, like):
    if not torch.is_tensor(k):
        k = torch.tensor(float(k), dtype=like.dtype)
    if bool((k >= 0).any()):
        raise NotImplementedError("oracle restates only the k<0 (Poincare ball) branch")
    return k


def tan_k(x, k):
    rk = _sabs(k).sqrt()
    return _tanh(x * rk) / rk


def artan_k(x, k):
    rk = _sabs(k).sqrt()
    return _artanh(x * rk) / rk


def expmap0(u, *, k, dim=-1):
    k = _k(k, u)
    n = u.norm(dim=dim, p=2, keepdim=True).clamp_min(_MIN_NORM)
    return tan_k(n, k) * (u / n)


def project(x, *, k, dim=-1, eps=-1.0):
    k = _k(k, x)
    if eps < 0:
        eps = 4e-3 if x.dtype == torch.float32 else 1e-5
    maxnorm = (1.0 - eps) / _sabs(k).sqrt()
    n = x.norm(dim=dim, p=2, keepdim=True).clamp_min(_MIN_NORM)
    return torch.where(n > maxnorm, x / n * maxnorm, x)


def logmap0(y, *, k, dim=-1):
    k = _k(k, y)
    n = y.norm(dim=dim, p=2, keepdim=True).clamp_min(_MIN_NORM)
    return (y / n) * artan_k(n, k)


def dist0(x, *, k, dim=-1, keepdim=False):
    k = _k(k, x)
    return 2.0 * artan_k(x.norm(dim=dim, p=2, keepdim=keepdim), k)


def dist(x, y, *, k, keepdim=False, dim=-1):
    """Only needed so the reference module's cold paths import; Mobius-add distance."""
    k = _k(k, x)
    mx = -x
    x2 = (mx * mx).sum(dim=dim, keepdim=True)
    y2 = (y * y).sum(dim=dim, keepdim=True)
    xy = (mx * y).sum(dim=dim, keepdim=True)
    num = (1 - 2 * k * xy - k * y2) * mx + (1 + k * x2) * y
    den = (1 - 2 * k * xy + k ** 2 * x2 * y2).clamp_min(_MIN_NORM)
    return 2.0 * artan_k((num / den).norm(dim=dim, p=2, keepdim=keepdim), k)
