"""Channel reduction + hyperbolic feature re-weighting in front of the head -- SURVEY section 8f row 3.

`reduce_hfr(feats, conv_reduce, wn_mlp)` computes what the reference's classifier computes between its decoder and
`mapper.expmap` (core/models/classifier.py:526-550; v2 head :187-214) with the classifier in evaluation mode -- how the
acquisition round runs it (core/active/build.py:72-73):

    decoder_out = self.conv_reduce(decoder_out)
    if self.wn_mlp is not None: ... F.normalize over the pixels of each channel * clamp(mean(wn_mlp(pixels)), 1e-5)

from the modules' own parameters (`conv_reduce`: 1x1 nn.Conv2d; `wn_mlp`: nn.Sequential(Linear, BatchNorm1d, ReLU, Linear) or
None), in three CUDA kernels (`halo_reduce_hfr_fwd`).  Forward only: in training mode (batch statistics, autograd through
the normalisation) the classifier keeps its torch modules, and this function raises.  There is no CPU path.
"""
import torch

from . import _native as nat


def reduce_hfr(feats, conv_reduce, wn_mlp=None, return_scale=False):
    """feats (N,Cin,H,W) CUDA float tensor -> (N,C,H,W) fp32 features for HyperMapper.expmap / the fused head."""
    lib = nat.load()
    nat.require_cuda(feats, "feats")
    params = list(conv_reduce.parameters()) + (list(wn_mlp.parameters()) if wn_mlp is not None else [])
    bn_training = wn_mlp is not None and wn_mlp[1].training
    if bn_training or (torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in params))):
        if return_scale:
            raise ValueError("reduce_hfr: return_scale is an evaluation-mode diagnostic")
        return _reduce_hfr_train(feats, conv_reduce, wn_mlp)
    w = conv_reduce.weight
    if w.dim() != 4 or w.shape[2] != 1 or w.shape[3] != 1:
        raise ValueError("reduce_hfr: conv_reduce must be a 1x1 convolution, got weight %s" % (tuple(w.shape),))
    x = feats.detach().float().contiguous()
    N, Cin, H, W = x.shape
    C = w.shape[0]
    if w.shape[1] != Cin:
        raise ValueError("reduce_hfr: conv_reduce expects %d input channels, features have %d" % (w.shape[1], Cin))
    dev = x.device

    def f32(t):
        return None if t is None else t.detach().to(device=dev, dtype=torch.float32).contiguous()

    Wr, br = f32(w.reshape(C, Cin)), f32(conv_reduce.bias)
    W1 = b1 = g = b = m = v = W2 = b2 = None
    eps = 1e-5
    if wn_mlp is not None:
        lin1, bn, lin2 = wn_mlp[0], wn_mlp[1], wn_mlp[3]
        W1, b1, W2, b2 = f32(lin1.weight), f32(lin1.bias), f32(lin2.weight), f32(lin2.bias)
        ones = torch.ones((C,), dtype=torch.float32, device=dev)
        g = f32(bn.weight) if bn.weight is not None else ones
        b = f32(bn.bias) if bn.bias is not None else torch.zeros_like(ones)
        m, v, eps = f32(bn.running_mean), f32(bn.running_var), float(bn.eps)
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
    scale = torch.empty((N, C), dtype=torch.float32, device=dev) if (return_scale and wn_mlp is not None) else None
    ws = nat.workspace.get(dev, "reduce_hfr", lib.halo_reduce_hfr_workspace_bytes(N, C, H, W))
    with torch.cuda.device(dev):
        rc = lib.halo_reduce_hfr_fwd(nat.ptr(x), nat.ptr(Wr), nat.ptr(br), nat.ptr(W1), nat.ptr(b1), nat.ptr(g), nat.ptr(b),
                                     nat.ptr(m), nat.ptr(v), eps, nat.ptr(W2), nat.ptr(b2), nat.ptr(out), nat.ptr(scale), N, Cin,
                                     C, H, W, nat.ptr(ws), ws.numel(), nat.stream_of(x))
    nat.check(rc, "halo_reduce_hfr_fwd")
    return (out, scale) if return_scale else out


class _ReduceHFRTrain(torch.autograd.Function):
    """conv_reduce + HFR with autograd: forward `halo_reduce_hfr_train_fwd`, backward `halo_reduce_hfr_train_bwd`."""

    @staticmethod
    def forward(ctx, x, Wr, br, W1, b1, g, b, W2, b2, fixed_stats, eps, dims):
        lib = nat.load()
        N, Cin, C, H, W = dims
        dev = x.device
        hfr = W1 is not None
        y = torch.empty((N, C, H, W), dtype=torch.float32, device=dev)
        z = torch.empty_like(y) if hfr else None
        a = torch.empty_like(y) if hfr else None
        stats = torch.empty((2, C), dtype=torch.float32, device=dev) if hfr else None
        small = torch.empty((N, 3, C), dtype=torch.float32, device=dev) if hfr else None
        ws = nat.workspace.get(dev, "reduce_hfr_train", lib.halo_reduce_hfr_train_workspace_bytes(N, Cin, C, H, W))
        with torch.cuda.device(dev):
            rc = lib.halo_reduce_hfr_train_fwd(nat.ptr(x), nat.ptr(Wr), nat.ptr(br), nat.ptr(W1), nat.ptr(b1), nat.ptr(g), nat.ptr(b),
                                               eps, nat.ptr(W2), nat.ptr(b2), nat.ptr(fixed_stats), nat.ptr(y), nat.ptr(a), nat.ptr(z),
                                               nat.ptr(stats), nat.ptr(small), N, Cin, C, H, W, nat.ptr(ws), ws.numel(),
                                               nat.stream_of(x))
        nat.check(rc, "halo_reduce_hfr_train_fwd")
        ctx.save_for_backward(x, Wr, W1, g, b, W2, y, a, stats, small)
        ctx.meta = (eps, dims, fixed_stats is None, br is not None)
        ctx.mark_non_differentiable(*([stats] if hfr else []))
        return (z, stats) if hfr else (y, None)

    @staticmethod
    def backward(ctx, dz, _dstats):
        lib = nat.load()
        x, Wr, W1, g, b, W2, y, a, stats, small = ctx.saved_tensors
        eps, (N, Cin, C, H, W), batch, has_br = ctx.meta
        dev = x.device
        hfr = W1 is not None
        dz = dz.contiguous().float()
        f32 = dict(dtype=torch.float32, device=dev)
        dx = torch.empty((N, Cin, H, W), **f32) if ctx.needs_input_grad[0] else None
        dWr, dbr = torch.empty((C, Cin), **f32), (torch.empty((C,), **f32) if has_br else None)
        dW1 = db1 = dg = db = dW2 = db2 = None
        if hfr:
            dW1, dW2 = torch.empty((C, C), **f32), torch.empty((C, C), **f32)
            db1, dg, db, db2 = (torch.empty((C,), **f32) for _ in range(4))
        ws = nat.workspace.get(dev, "reduce_hfr_train", lib.halo_reduce_hfr_train_workspace_bytes(N, Cin, C, H, W))
        with torch.cuda.device(dev):
            rc = lib.halo_reduce_hfr_train_bwd(nat.ptr(x), nat.ptr(Wr), nat.ptr(W1), nat.ptr(g), nat.ptr(b), eps,
                                               nat.ptr(W2), nat.ptr(y), nat.ptr(a), nat.ptr(stats), nat.ptr(small), nat.ptr(dz), nat.ptr(dx),
                                               nat.ptr(dWr), nat.ptr(dbr), nat.ptr(dW1), nat.ptr(db1), nat.ptr(dg), nat.ptr(db),
                                               nat.ptr(dW2), nat.ptr(db2), 1 if batch else 0, N, Cin, C, H, W, nat.ptr(ws),
                                               ws.numel(), nat.stream_of(x))
        nat.check(rc, "halo_reduce_hfr_train_bwd")
        return dx, dWr, dbr, dW1, db1, dg, db, dW2, db2, None, None, None


def _reduce_hfr_train(feats, conv_reduce, wn_mlp):
    w = conv_reduce.weight
    if w.dim() != 4 or w.shape[2] != 1 or w.shape[3] != 1:
        raise ValueError("reduce_hfr: conv_reduce must be a 1x1 convolution, got weight %s" % (tuple(w.shape),))
    x = feats.float().contiguous()
    N, Cin, H, W = x.shape
    C = w.shape[0]
    if w.shape[1] != Cin:
        raise ValueError("reduce_hfr: conv_reduce expects %d input channels, features have %d" % (w.shape[1], Cin))
    dev = x.device

    def f32(t):
        return None if t is None else t.to(device=dev, dtype=torch.float32).contiguous()

    Wr, br = f32(w.reshape(C, Cin)), f32(conv_reduce.bias)
    if wn_mlp is None:
        out, _ = _ReduceHFRTrain.apply(x, Wr, br, None, None, None, None, None, None, None, 0.0, (N, Cin, C, H, W))
        return out
    lin1, bn, lin2 = wn_mlp[0], wn_mlp[1], wn_mlp[3]
    ones = torch.ones((C,), dtype=torch.float32, device=dev)
    g = f32(bn.weight) if bn.weight is not None else ones
    b = f32(bn.bias) if bn.bias is not None else torch.zeros_like(ones)
    use_batch = bn.training or bn.running_mean is None
    fixed = None if use_batch else torch.stack([f32(bn.running_mean.detach()), f32(bn.running_var.detach())])
    out, stats = _ReduceHFRTrain.apply(x, Wr, br, f32(lin1.weight), f32(lin1.bias), g, b, f32(lin2.weight), f32(lin2.bias), fixed,
                                       float(bn.eps), (N, Cin, C, H, W))
    if bn.training and bn.track_running_stats and bn.running_mean is not None:
        # torch.nn.BatchNorm1d's bookkeeping: exponential (or cumulative when momentum is None) average of the batch mean
        # and of the UNBIASED batch variance
        with torch.no_grad():
            m = N * H * W
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            rm, rv = bn.running_mean, bn.running_var
            rm.mul_(1.0 - mom).add_(stats[0].to(device=rm.device, dtype=rm.dtype), alpha=mom)
            rv.mul_(1.0 - mom).add_(stats[1].to(device=rv.device, dtype=rv.dtype) * (m / max(m - 1, 1)), alpha=mom)
    return out
