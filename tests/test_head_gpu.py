"""GPU: the fused head (halo_head_fwd / halo_head_bwd through the Python mirror of core/utils/hyperbolic.py)
against the golden vectors frozen from the reference and against the fp64 oracle on seeded sweeps."""
import pytest
import torch

import halo_b200
from halo_b200 import synth
from oracle import head as ohead
from tests.util import TOL, rel_err, t

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_golden_logits_radius_fused(golden):
    g = golden["head"]
    for k in range(int(g["n_cases"])):
        tag = "h%d_" % k
        c = float(g[tag + "c"])
        u = t(g[tag + "u"]).to(DEV)
        res = halo_b200.head_forward(u, t(g[tag + "P"]).to(DEV), t(g[tag + "A"]).to(DEV), c, want_logits=True,
                                     want_radius=True, want_stats=True)
        assert rel_err(res["logits"], t(g[tag + "logits"])) <= TOL, k
        assert rel_err(res["radius"], t(g[tag + "radius"])) <= TOL, k
        ref_r = t(g[tag + "radius"]).float()
        assert torch.allclose(res["stats"][:, 0].cpu(), ref_r.amin(dim=(1, 2)), rtol=1e-5)
        assert torch.allclose(res["stats"][:, 1].cpu(), ref_r.amax(dim=(1, 2)), rtol=1e-5)


def test_golden_logits_from_ball_points(golden):
    """HyperMLR fed the reference's fp64 embedding (the reference call: conv_seg(decoder_out.double()))."""
    g = golden["head"]
    for k in range(int(g["n_cases"])):
        tag = "h%d_" % k
        c = float(g[tag + "c"])
        x = t(g[tag + "x"]).to(DEV)
        res = halo_b200.head_forward(x, t(g[tag + "P"]).to(DEV), t(g[tag + "A"]).to(DEV), c, kind="ball",
                                     want_logits=True, want_radius=True)
        assert rel_err(res["logits"], t(g[tag + "logits"])) <= TOL, k
        assert rel_err(res["radius"], t(g[tag + "radius"])) <= TOL, k


@pytest.mark.parametrize("tensor_cores", [True, False], ids=["tcgen05", "cuda_core"])
@pytest.mark.parametrize("c", [1.0, 0.5])
@pytest.mark.parametrize("sigma", [0.01, 0.1, 0.3, 1.0])
def test_sweep_vs_oracle_c256(c, sigma, tensor_cores):
    """Interior (sigma 0.01, 0.1), MLR-projection branch (0.3) and expmap clip (1.0) at the BASELINE channel count,
    on both contraction paths (tcgen05 3xTF32 and fp32 CUDA cores)."""
    C, O, H, W = 256, 19, 24, 40
    P, A = synth.head_params(O, C, seed=3, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(2)])
    logits, x, rad = ohead.head_forward(u, P, A, c)
    res = halo_b200.head_forward(u.to(DEV), P.to(DEV), A.to(DEV), c, want_logits=True, want_radius=True,
                                 want_pixunc=True, want_label=True, tensor_cores=tensor_cores)
    assert rel_err(res["logits"], logits) <= TOL
    assert rel_err(res["radius"], rad) <= TOL
    p = torch.softmax(logits, dim=1)
    ent = torch.sum(-p * torch.log(p + 1e-6), dim=1) / torch.log(torch.tensor(19.0))
    assert rel_err(res["pixunc"], ent) <= TOL
    agree = (res["label"].cpu().long() == p.argmax(dim=1)).float().mean().item()
    assert agree >= 0.999


@pytest.mark.parametrize("shape", [(16, 64, 9, 17), (19, 62, 7, 11), (3, 30, 5, 5), (32, 128, 8, 16), (21, 100, 6, 9)])
def test_odd_shapes_and_class_counts(shape):
    O, C, H, W = shape
    P, A = synth.head_params(O, C, seed=5, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=0.25) for i in range(3)])
    logits, x, rad = ohead.head_forward(u, P, A, 1.0)
    res = halo_b200.head_forward(u.to(DEV), P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True)
    assert rel_err(res["logits"], logits) <= TOL
    assert rel_err(res["radius"], rad) <= TOL


@pytest.mark.parametrize("shape", [(19, 32, 8, 16, 1), (16, 64, 24, 40, 3), (3, 64, 10, 10, 1), (32, 128, 9, 12, 2),
                                   (19, 256, 130, 126, 2), (19, 64, 160, 320, 1), (8, 96, 31, 36, 2)])
def test_tensor_core_path_shapes(shape):
    """tcgen05 path: class counts that pad differently (N = 16/32/48/64 columns), several channel counts (1-8
    pipeline stages), tiles that straddle the end of an image (H*W % 128 != 0) and many tiles per CTA."""
    O, C, H, W, N = shape
    P, A = synth.head_params(O, C, seed=9, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=0.2) for i in range(N)])
    gt = torch.stack([synth.image_labels(i, O, H, W) for i in range(N)])
    logits, x, rad = ohead.head_forward(u, P, A, 1.0)
    tc = halo_b200.head_forward(u.to(DEV), P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True, want_pixunc=True,
                                want_label=True, want_stats=True, gt=gt.to(DEV), pixunc_mode="one_minus_pgt",
                                label_mode="gt_filled")
    cc = halo_b200.head_forward(u.to(DEV), P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True, want_pixunc=True,
                                want_label=True, want_stats=True, gt=gt.to(DEV), pixunc_mode="one_minus_pgt",
                                label_mode="gt_filled", tensor_cores=False)
    assert rel_err(tc["logits"], logits) <= TOL
    assert rel_err(tc["radius"], rad) <= TOL
    assert rel_err(tc["radius"], cc["radius"]) <= 2e-6             # |u|^2: two interleaved fp32 partial sums (FFMA2) vs one chain
    assert torch.allclose(tc["stats"][:, :2], cc["stats"][:, :2])
    assert rel_err(tc["pixunc"], cc["pixunc"]) <= 1e-4
    assert (tc["label"] == cc["label"]).float().mean().item() >= 0.999


def test_points_outside_and_zero_features():
    C, O, H, W = 64, 19, 8, 8
    P, A = synth.head_params(O, C, seed=7, dtype=torch.float64)
    P[2] *= 9.0  # |p| > 1/sqrt(c): B_k < 0 branch
    u = torch.randn(1, C, H, W) * 0.3
    u[0, :, 0, 0] = 0.0  # a pixel at the origin
    u[0, :, 1, 1] *= 100.0  # far outside: clipped by project
    logits, x, rad = ohead.head_forward(u, P, A, 1.0)
    res = halo_b200.head_forward(u.to(DEV), P.to(DEV), A.to(DEV), 1.0, want_logits=True, want_radius=True)
    assert rel_err(res["logits"], logits) <= TOL
    assert rel_err(res["radius"], rad) <= TOL


def test_module_dropin_surface(golden):
    g = golden["head"]
    tag = "h1_"
    c = float(g[tag + "c"])
    u = t(g[tag + "u"]).to(DEV)
    mapper = halo_b200.HyperMapper(c=c)
    mlr = halo_b200.HyperMLR(u.shape[1], 19, c=c).to(DEV)
    mlr.load_state_dict({"P_MLR": t(g[tag + "P"]), "A_MLR": t(g[tag + "A"])})
    with torch.no_grad():
        emb = mapper.expmap(u, dim=1)                    # classifier.py:553
        out = mlr(emb.double()).float()                  # classifier.py:554
    assert isinstance(emb, halo_b200.PoincareEmbedding)
    assert rel_err(out, t(g[tag + "logits"])) <= TOL
    assert rel_err(mapper.poincare_distance_origin(emb, dim=1), t(g[tag + "radius"])) <= TOL
    x = emb.materialize()
    assert x.dtype == torch.float64 and rel_err(x, t(g[tag + "x"])) <= 1e-6
    # the handle behaves like a tensor for torch functions (build.py:133 F.interpolate(decoder_out, ...))
    up = torch.nn.functional.interpolate(emb, size=(24, 40), mode="bilinear", align_corners=True)
    ref_up = torch.nn.functional.interpolate(t(g[tag + "x"]), size=(24, 40), mode="bilinear", align_corners=True)
    assert rel_err(up, ref_up) <= 1e-6
    assert emb[0:1].shape[0] == 1 and isinstance(emb[0:1], halo_b200.PoincareEmbedding)
    # eager paths on other layouts
    v = torch.randn(5, 7, 33, device=DEV)
    xe = mapper.expmap(v, dim=-1)
    assert rel_err(xe, ohead.expmap(v.cpu(), c, dim=-1)) <= 1e-6
    assert rel_err(mapper.poincare_distance_origin(xe, dim=-1), ohead.radius(ohead.expmap(v.cpu(), c, dim=-1), c, dim=-1)) <= TOL


def test_backward_matches_golden(golden):
    g = golden["head"]
    for k in range(8):
        tag = "h%d_" % k
        c = float(g[tag + "c"])
        du, dP, dA = halo_b200.head_backward(t(g[tag + "u"]).to(DEV), t(g[tag + "P"]).to(DEV), t(g[tag + "A"]).to(DEV),
                                             c, t(g[tag + "dlogits"]).to(DEV))
        # tolerance: 1e-4 of the largest gradient entry (fp32 recompute + analytic derivative vs fp64 autograd)
        assert rel_err(du, t(g[tag + "du"])) <= 1e-4, (k, "du")
        assert rel_err(dP, t(g[tag + "dP"])) <= 1e-4, (k, "dP")
        assert rel_err(dA, t(g[tag + "dA"])) <= 1e-4, (k, "dA")


@pytest.mark.parametrize("shape", [(19, 64, 24, 40, 2, 0.1), (19, 256, 16, 24, 1, 0.1), (16, 128, 20, 20, 2, 0.3),
                                   (19, 256, 130, 126, 1, 0.1), (5, 32, 9, 12, 3, 1.0), (24, 96, 16, 16, 1, 0.2),
                                   (19, 160, 20, 24, 2, 0.1), (8, 224, 12, 20, 1, 0.2)])
def test_backward_tensor_core_pixel_pass(shape, monkeypatch):
    """K4a on the tensor cores (two chained tcgen05 GEMMs) against fp64
    autograd through the oracle and against the fp32 CUDA-core pixel pass."""
    O, C, H, W, N, sigma = shape
    P, A = synth.head_params(O, C, seed=13, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    g = torch.Generator().manual_seed(3)
    dl = torch.randn((N, O, H, W), generator=g) * 1e-3
    du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
    args = (u.to(DEV), P.to(DEV), A.to(DEV), 1.0, dl.to(DEV))
    monkeypatch.delenv("HALO_BWD_CUDA_CORE", raising=False)
    monkeypatch.delenv("HALO_BWD_DW_CUDA_CORE", raising=False)
    monkeypatch.setenv("HALO_BWD_TWO_KERNEL", "1")           # this test pins the round-1 two-kernel tensor-core path
    du, dP, dA = halo_b200.head_backward(*args)              # tcgen05 pixel pass + tcgen05 weight gradient (C = 128, 256)
    monkeypatch.setenv("HALO_BWD_DW_CUDA_CORE", "1")
    _, dP_mix, dA_mix = halo_b200.head_backward(*args)       # tcgen05 pixel pass + fp32 CUDA-core weight gradient
    assert rel_err(dP_mix, dP_ref) <= 1e-4 and rel_err(dA_mix, dA_ref) <= 1e-4
    assert rel_err(dP, dP_mix) <= 2e-5 and rel_err(dA, dA_mix) <= 2e-5
    monkeypatch.setenv("HALO_BWD_CUDA_CORE", "1")
    du_cc, dP_cc, dA_cc = halo_b200.head_backward(*args)
    # pixels sitting on the MLR projection switch (derivative discontinuity) may take either branch in fp32:
    # (16,128,20,20) holds one (image 0, pixel 29, class 9: root - maxnorm = 1.6e-10).  They stay out of the du check.
    smooth = (~ohead.nonsmooth_pixels(u, P, 1.0, rel=1e-5))[:, None].to(du_ref.dtype)
    assert smooth.mean() > 0.99
    du, du_cc, du_ref = du.cpu() * smooth, du_cc.cpu() * smooth, du_ref * smooth
    for got, cc, ref, name in ((du, du_cc, du_ref, "du"), (dP, dP_cc, dP_ref, "dP"), (dA, dA_cc, dA_ref, "dA")):
        assert rel_err(got, ref) <= 1e-4, name
        assert rel_err(cc, ref) <= 1e-4, name
        assert rel_err(got, cc.cpu()) <= 1e-4, name


def test_autograd_through_modules():
    C, O, H, W = 64, 19, 12, 16
    c = 1.0
    P, A = synth.head_params(O, C, seed=11, dtype=torch.float64)
    u = torch.randn(2, C, H, W) * 0.1
    target = torch.randint(0, O, (2, H, W))
    # oracle: CE loss through the fp64 head
    u0 = u.clone().requires_grad_(True)
    P0, A0 = P.clone().requires_grad_(True), A.clone().requires_grad_(True)
    lo = ohead.mlr_logits(ohead.expmap(u0, c, dim=1), P0, A0, c).float()
    torch.nn.functional.cross_entropy(lo, target).backward()
    mapper = halo_b200.HyperMapper(c=c)
    mlr = halo_b200.HyperMLR(C, O, c=c).to(DEV)
    mlr.load_state_dict({"P_MLR": P, "A_MLR": A})
    u1 = u.to(DEV).requires_grad_(True)
    out = mlr(mapper.expmap(u1, dim=1).double()).float()
    torch.nn.functional.cross_entropy(out, target.to(DEV)).backward()
    assert rel_err(u1.grad, u0.grad) <= 1e-4
    assert rel_err(mlr.P_MLR.grad, P0.grad) <= 1e-4
    assert rel_err(mlr.A_MLR.grad, A0.grad) <= 1e-4


def test_backward_is_deterministic():
    C, O, H, W = 32, 19, 40, 64
    P, A = synth.head_params(O, C, seed=1)
    u = (torch.randn(2, C, H, W) * 0.2).to(DEV)
    dl = (torch.randn(2, O, H, W) * 1e-3).to(DEV)
    a = halo_b200.head_backward(u, P.to(DEV), A.to(DEV), 1.0, dl)
    b = halo_b200.head_backward(u, P.to(DEV), A.to(DEV), 1.0, dl)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


@pytest.mark.parametrize("shape", [(19, 256, 640, 1280, 2), (16, 128, 512, 1024, 3)])
def test_backward_full_size_tensor_core_vs_cuda_core(shape, monkeypatch):
    """BASELINE configs[4] geometry: many tiles per CTA, several accumulator drains of the weight-gradient kernel, the
    double-buffered TMEM regions wrapping hundreds of times.  The fp64 oracle takes minutes at this size, so the two
    independent GPU paths (tcgen05 vs fp32 CUDA cores, each pinned to the oracle at small sizes) are compared."""
    O, C, H, W, N = shape
    P, A = synth.head_params(O, C, seed=2, device=DEV)
    u = torch.stack([synth.image_features(i, C, H, W, device=DEV) for i in range(N)])
    dl = torch.randn((N, O, H, W), device=DEV, generator=torch.Generator(device=DEV).manual_seed(5)) * 1e-3
    monkeypatch.delenv("HALO_BWD_CUDA_CORE", raising=False)
    monkeypatch.delenv("HALO_BWD_DW_CUDA_CORE", raising=False)
    monkeypatch.delenv("HALO_BWD_TWO_KERNEL", raising=False)
    from halo_b200 import _native as nat
    saved = halo_b200.head_forward(u, P, A, 1.0, want_logits=False, want_saved=True)["saved"]
    assert saved is not None
    du, dP, dA = halo_b200.head_backward(u, P, A, 1.0, dl, saved=saved)        # streaming kernel, features read once
    assert nat.last_path() == ("bwd:stream_tcgen05",)
    du2, dP2, dA2 = halo_b200.head_backward(u, P, A, 1.0, dl, saved=saved)
    assert torch.equal(du, du2) and torch.equal(dP, dP2) and torch.equal(dA, dA2)   # fixed-order reductions
    du3, dP3, dA3 = halo_b200.head_backward(u, P, A, 1.0, dl)                  # same kernel after recomputing the contractions
    assert nat.last_path() == ("bwd:stream_tcgen05", "bwd:recompute")
    assert torch.equal(du, du3) and torch.equal(dP, dP3) and torch.equal(dA, dA3)
    monkeypatch.setenv("HALO_BWD_TWO_KERNEL", "1")
    du_tk, dP_tk, dA_tk = halo_b200.head_backward(u, P, A, 1.0, dl)            # round-1 two-kernel tensor-core path
    assert nat.last_path()[0] == "bwd_pix:tcgen05"
    monkeypatch.setenv("HALO_BWD_CUDA_CORE", "1")
    du_cc, dP_cc, dA_cc = halo_b200.head_backward(u, P, A, 1.0, dl)
    for a, b in ((du, du_cc), (dP, dP_cc), (dA, dA_cc), (du, du_tk), (dP, dP_tk), (dA, dA_tk)):
        assert rel_err(a, b) <= 1e-4


@pytest.mark.parametrize("shape", [(19, 256, 2, 2, 1), (19, 128, 2, 4, 3), (19, 256, 4, 4, 37), (3, 64, 2, 2, 5),
                                   (19, 256, 1, 132, 2), (16, 128, 2, 2, 300)])
def test_tiny_planes_forward_and_backward(shape):
    """Planes much smaller than a TMA box (128-pixel tiles, 32-pixel chunks of the weight-gradient kernel), a single live
    pixel row per tile, hundreds of images: the tensor-core forward and both tensor-core backward kernels against fp64."""
    O, C, H, W, N = shape
    P, A = synth.head_params(O, C, seed=1, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=0.1) for i in range(N)])
    dl = torch.randn((N, O, H, W), generator=torch.Generator().manual_seed(1)) * 1e-3
    du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
    logits_ref, _, _ = ohead.head_forward(u, P, A, 1.0)
    args = (u.to(DEV), P.to(DEV), A.to(DEV), 1.0)
    assert rel_err(halo_b200.head_forward(*args, want_logits=True)["logits"], logits_ref) <= TOL
    du, dP, dA = halo_b200.head_backward(*args, dl.to(DEV))
    assert rel_err(du, du_ref) <= 1e-4 and rel_err(dP, dP_ref) <= 1e-4 and rel_err(dA, dA_ref) <= 1e-4


def test_autograd_from_points_already_on_the_ball():
    """HyperMLR fed a plain tensor of ball points (not the lazy handle) stays differentiable w.r.t. x, P and A like the
    reference's module (hyperbolic.py:120-188): frozen-backbone / precomputed-embedding callers get gradients, not a
    silent None."""
    C, O, H, W = 64, 19, 10, 12
    P, A = synth.head_params(O, C, seed=21, dtype=torch.float64)
    x = ohead.expmap(torch.randn(2, C, H, W) * 0.1, 1.0, dim=1)
    target = torch.randint(0, O, (2, H, W))
    x0, P0, A0 = x.clone().requires_grad_(True), P.clone().requires_grad_(True), A.clone().requires_grad_(True)
    torch.nn.functional.cross_entropy(ohead.mlr_logits(x0, P0, A0, 1.0).float(), target).backward()
    mlr = halo_b200.HyperMLR(C, O, c=1.0).to(DEV)
    mlr.load_state_dict({"P_MLR": P, "A_MLR": A})
    for x_needs_grad in (True, False):          # detached embeddings must still give parameter gradients
        mlr.zero_grad()
        x1 = x.to(DEV).requires_grad_(x_needs_grad)
        out = mlr(x1)
        assert out.grad_fn is not None
        torch.nn.functional.cross_entropy(out.float(), target.to(DEV)).backward()
        assert rel_err(mlr.P_MLR.grad, P0.grad) <= 1e-4 and rel_err(mlr.A_MLR.grad, A0.grad) <= 1e-4
        if x_needs_grad:
            assert rel_err(x1.grad, x0.grad) <= 1e-4
    with torch.no_grad():                        # forward-only: the ball kernel, same logits
        assert rel_err(mlr(x.to(DEV)), ohead.mlr_logits(x, P, A, 1.0)) <= TOL
    with pytest.raises(NotImplementedError):     # the eager (non NCHW) expmap is forward-only and says so
        halo_b200.HyperMapper(1.0).expmap(torch.randn(5, 33, device=DEV, requires_grad=True), dim=-1)


def test_last_path_reports_the_kernel_variant():
    """halo_last_path(): shapes that fall off the tensor-core envelope are visible to the caller."""
    from halo_b200 import _native as nat

    P, A = synth.head_params(19, 64, seed=1, device=DEV)
    u = torch.randn(1, 64, 16, 16, device=DEV) * 0.1
    halo_b200.head_forward(u, P, A, 1.0)
    assert nat.last_path() == ("fwd:tcgen05",)
    halo_b200.head_forward(u, P, A, 1.0, tensor_cores=False)
    assert nat.last_path() == ("fwd:cuda_core",)
    P2, A2 = synth.head_params(19, 48, seed=1, device=DEV)          # C % 32 != 0: CUDA cores
    halo_b200.head_forward(torch.randn(1, 48, 16, 16, device=DEV) * 0.1, P2, A2, 1.0)
    assert nat.last_path() == ("fwd:cuda_core",)
    dl = torch.randn(1, 19, 16, 16, device=DEV) * 1e-3
    halo_b200.head_backward(u, P, A, 1.0, dl)                        # C = 64, the shipped HALO channel count: streaming kernel
    assert nat.last_path() == ("bwd:stream_tcgen05", "bwd:recompute")
    P3, A3 = synth.head_params(19, 96, seed=1, device=DEV)          # C = 96: two-kernel tensor-core pixel pass + CUDA-core dW
    halo_b200.head_backward(torch.randn(1, 96, 16, 16, device=DEV) * 0.1, P3, A3, 1.0, dl)
    assert nat.last_path() == ("bwd_pix:tcgen05", "bwd_dw:cuda_core")
    halo_b200.head_backward(torch.randn(1, 48, 16, 16, device=DEV) * 0.1, P2, A2, 1.0, dl)
    assert nat.last_path() == ("bwd_pix:cuda_core", "bwd_dw:cuda_core")


@pytest.mark.parametrize("shape", [(19, 64, 24, 40, 2, 0.1), (19, 256, 16, 24, 1, 0.1), (16, 128, 20, 20, 2, 0.3),
                                   (19, 256, 130, 126, 1, 0.1), (5, 64, 9, 12, 3, 1.0), (24, 128, 16, 16, 1, 0.2),
                                   (19, 64, 160, 320, 1, 0.1), (8, 256, 12, 20, 1, 0.2), (19, 128, 2, 2, 5, 0.1)])
def test_backward_streaming_kernel(shape, monkeypatch):
    """K4s (head_bwd_stream_tc.cu): du and dW from ONE pass over the features, driven by the contractions the forward saved,
    against fp64 autograd through the oracle -- every supported channel count (64 = the shipped HALO configs,
    core/configs/defaults.py:14; 128; 256), class paddings 8 / 16 / 20 / 24, ragged tiles, tiny planes."""
    from halo_b200 import _native as nat

    for k in ("HALO_BWD_CUDA_CORE", "HALO_BWD_DW_CUDA_CORE", "HALO_BWD_TWO_KERNEL"):
        monkeypatch.delenv(k, raising=False)
    O, C, H, W, N, sigma = shape
    P, A = synth.head_params(O, C, seed=13, dtype=torch.float64)
    u = torch.stack([synth.image_features(i, C, H, W, sigma=sigma) for i in range(N)])
    dl = torch.randn((N, O, H, W), generator=torch.Generator().manual_seed(3)) * 1e-3
    du_ref, dP_ref, dA_ref = ohead.head_grads(u, P, A, dl, 1.0)
    ud, Pd, Ad = u.to(DEV), P.to(DEV), A.to(DEV)
    fwd = halo_b200.head_forward(ud, Pd, Ad, 1.0, want_logits=True, want_saved=True)
    assert fwd["saved"] is not None and tuple(fwd["saved"].shape) == (N, 2 * ((O + 3) // 4 * 4) + 1, H, W)
    logits_ref, _, _ = ohead.head_forward(u, P, A, 1.0)
    assert rel_err(fwd["logits"], logits_ref) <= TOL                      # saving does not disturb the forward
    du, dP, dA = halo_b200.head_backward(ud, Pd, Ad, 1.0, dl.to(DEV), saved=fwd["saved"])
    assert nat.last_path() == ("bwd:stream_tcgen05",)
    smooth = (~ohead.nonsmooth_pixels(u, P, 1.0, rel=1e-5))[:, None].to(du_ref.dtype)
    assert smooth.mean() > 0.99
    assert rel_err(du.cpu() * smooth, du_ref * smooth) <= 1e-4
    assert rel_err(dP, dP_ref) <= 1e-4 and rel_err(dA, dA_ref) <= 1e-4
    du2, dP2, dA2 = halo_b200.head_backward(ud, Pd, Ad, 1.0, dl.to(DEV))  # recompute instead of saved planes: same bits
    assert torch.equal(du, du2) and torch.equal(dP, dP2) and torch.equal(dA, dA2)


@pytest.mark.parametrize("case", [(256, 3, 333, 500, 0), (256, 3, 320, 500, 4), (128, 3, 333, 500, 0), (64, 4, 328, 500, 12)],
                         ids=["c256-ragged", "c256-misaligned", "c128-ragged", "c64-ragged-misaligned"])
def test_backward_streaming_kernel_is_reproducible_over_many_launches(case, monkeypatch):
    """The per-CTA partials are added in a fixed order, so EVERY launch must return the same bits.  This is the regression test
    of a round-2 race (a ring stage released before its rows had been consumed: once per ~1 000 launches a few channels of
    one CTA's dP / dA partial were built from the next load's rows; profiles/r2_k4.md): 600 launches on shapes whose
    ragged last tiles and 16-byte-misaligned feature rows exposed it most often, each compared bit for bit with the first."""
    for k in ("HALO_BWD_CUDA_CORE", "HALO_BWD_DW_CUDA_CORE", "HALO_BWD_TWO_KERNEL"):
        monkeypatch.delenv(k, raising=False)
    C, N, H, W, off = case
    O = 19
    P, A = synth.head_params(O, C, seed=0, device=DEV)
    store = torch.empty(N * C * H * W + off, device=DEV)
    feat = store[off:].view(N, C, H, W)
    for i in range(N):
        feat[i] = synth.image_features(i, C, H, W, device=DEV)
    dl = torch.randn((N, O, H, W), device=DEV, generator=torch.Generator(device=DEV).manual_seed(1)) * 1e-3
    fwd = halo_b200.head_forward(feat, P, A, 1.0, want_logits=True, want_saved=True)
    first = [t.clone() for t in halo_b200.head_backward(feat, P, A, 1.0, dl, saved=fwd["saved"])]
    bad = torch.zeros((), dtype=torch.int32, device=DEV)
    for _ in range(600):
        out = halo_b200.head_backward(feat, P, A, 1.0, dl, saved=fwd["saved"])
        for a, b in zip(out, first):
            bad += (a != b).any()
    assert int(bad) == 0
