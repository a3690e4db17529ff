"""Oracle for the Poincare-ball classifier head (ORACLE -- test infrastructure, CPU only).

Restates, in float64 exactly as the reference runs it:
  * HyperMapper.expmap               core/utils/hyperbolic.py:28-39   (expmap0 then project, eps=1e-5 in fp64)
  * HyperMapper.poincare_distance_origin  core/utils/hyperbolic.py:74-83
  * HyperMLR._hyper_logits / forward  core/utils/hyperbolic.py:120-188
  * parameter init of HyperMLR        core/utils/hyperbolic.py:103-118
  * the autograd backward the learners take through both (core/train_learners.py:362)
"""
import math

import torch

from . import geoopt_math as gm

PROJ_EPS_MLR = 1e-3  # core/utils/hyperbolic.py:13 (used only inside the MLR, :162)


def init_mlr_params(num_classes, channels, seed=0, dtype=torch.float64):
    """kaiming_uniform_(a=sqrt(5)) on (O, C) -> U(-1/sqrt(C), 1/sqrt(C)); hyperbolic.py:115-118."""
    g = torch.Generator().manual_seed(seed)
    bound = 1.0 / math.sqrt(channels)
    P = (torch.rand((num_classes, channels), generator=g, dtype=torch.float64) * 2 - 1) * bound
    A = (torch.rand((num_classes, channels), generator=g, dtype=torch.float64) * 2 - 1) * bound
    return P.to(dtype), A.to(dtype)


def expmap(u, c=1.0, dim=1):
    """hyperbolic.py:37-38: x = project(expmap0(u.double(), k=-c), k=-c)."""
    k = torch.tensor(-float(c), dtype=torch.float64)  # hyperbolic.py:26
    x = gm.expmap0(u.double(), k=k, dim=dim)
    return gm.project(x, k=k, dim=dim)


def radius(x, c=1.0, dim=1):
    """hyperbolic.py:83: distance to the origin of the ball (x already on the ball)."""
    k = torch.tensor(-float(c), dtype=torch.float64)
    return gm.dist0(x, k=k, dim=dim)


def mlr_logits(x, P, A, c=1.0):
    """hyperbolic.py:120-184.  x: (B,C,H,W) on the ball; P, A: (O,C).  Returns (B,O,H,W) in x.dtype.

    Variable names follow the reference's algebra: Mobius addition (-p) (+) x is written
    alpha*(-p) + beta*x, its squared norm is formed in closed form, clipped to the MLR's own
    projection radius (1-1e-3)/sqrt(c), and the signed distance to the hyperplane goes through asinh.
    """
    K = torch.tensor(float(c), dtype=torch.float64)  # hyperbolic.py:113 (+c here, -c in the mapper)
    tiny = torch.tensor(1e-12, dtype=x.dtype)
    q = -P  # (O,C)
    xx = x.norm(dim=1) ** 2  # :136  (B,H,W)
    pp = (q.norm(dim=1) ** 2)[None, :, None, None]  # :137,143
    px = torch.einsum("bchw,oc->bohw", x, q)  # :141-142 (1x1 conv with -P)
    xx_ = xx[:, None]
    sqsq = K * xx_ * K * pp  # :146
    num_a = 1 + 2 * K * px + K * xx_  # :150
    num_b = 1 - K * pp  # :151
    den = torch.maximum(1 + 2 * K * px + sqsq, tiny)  # :152-153
    alpha = num_a / den
    beta = num_b / den
    mob = alpha ** 2 * pp + beta ** 2 * xx_ + 2 * alpha * beta * px  # :159
    maxnorm = (1.0 - PROJ_EPS_MLR) / torch.sqrt(K)  # :162
    root = torch.sqrt(mob)
    shrink = torch.where(root > maxnorm, maxnorm / torch.maximum(root, tiny), torch.ones_like(mob))  # :163-166
    mob_clipped = torch.where(root < maxnorm, mob, torch.ones_like(mob) * maxnorm ** 2)  # :167-170
    a_norm = A.norm(dim=1)  # :172
    a_hat = torch.nn.functional.normalize(A, dim=1)  # :173 (eps 1e-12)
    xa = beta * torch.einsum("bchw,oc->bohw", x, a_hat)  # :175
    pa = alpha * (q * a_hat).sum(dim=1)[None, :, None, None]  # :176
    dot = (xa + pa) * shrink  # :177-178
    lam = 2.0 / torch.maximum(1 - K * mob_clipped, tiny)  # :179
    arg = torch.sqrt(K) * dot * lam  # :180
    return 2.0 / torch.sqrt(K) * a_norm[None, :, None, None] * torch.asinh(arg)  # :181-183


def head_forward(u, P, A, c=1.0):
    """The reference call-site sequence core/models/classifier.py:553-554 (and :372-373):
    decoder_out = mapper.expmap(u, dim=1); out = conv_seg(decoder_out.double()).float().
    Returns (logits fp32, x fp64, radius fp64)."""
    x = expmap(u, c, dim=1)
    logits = mlr_logits(x.double(), P.double(), A.double(), c).float()
    return logits, x, radius(x, c, dim=1)


def head_grads(u, P, A, dlogits, c=1.0):
    """Autograd through expmap + MLR exactly as manual_backward does (train_learners.py:362).
    Returns (du fp32-shaped-as-u, dP fp64, dA fp64)."""
    u = u.detach().clone().requires_grad_(True)
    P = P.detach().double().clone().requires_grad_(True)
    A = A.detach().double().clone().requires_grad_(True)
    x = expmap(u, c, dim=1)
    out = mlr_logits(x.double(), P, A, c).float()
    du, dP, dA = torch.autograd.grad(out, (u, P, A), grad_outputs=dlogits.float())
    return du, dP, dA


def nonsmooth_pixels(u, P, c=1.0, rel=1e-5):
    """(B,H,W) bool: pixels where some class sits within `rel` (relative, in 1 - c*m) of the MLR's projection
    switch root == maxnorm (hyperbolic.py:163-170).  The two `where`s make the logit continuous but its derivative
    discontinuous there, so an fp32 recompute may legitimately take the other branch than fp64 autograd
    (SURVEY a-14 "non-smooth points"); gradient parity tests leave these pixels out of the du comparison."""
    K = float(c)
    x = expmap(u, c, dim=1).double()
    q = -P.double()
    xx = (x * x).sum(dim=1, keepdim=True)
    pp = (q * q).sum(dim=1)[None, :, None, None]
    px = torch.einsum("bchw,oc->bohw", x, q)
    den = torch.clamp(1 + 2 * K * px + K * K * xx * pp, min=1e-12)
    one_minus_cm = (1 - K * pp) * (1 - K * xx) / den  # the Mobius-norm identity, DESIGN.md section 4
    thresh = 1 - (1.0 - PROJ_EPS_MLR) ** 2
    return ((one_minus_cm / thresh - 1).abs() < rel).any(dim=1)
