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
    if torch.is_grad_enabled() and (feats.requires_grad or any(p.requires_grad for p in conv_reduce.parameters())) and conv_reduce.training:
        raise NotImplementedError("reduce_hfr is the evaluation-mode forward (acquisition / inference); in training mode "
                                  "keep the classifier's own conv_reduce / wn_mlp modules (autograd, BatchNorm batch statistics)")
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
        if bn.training:
            raise NotImplementedError("reduce_hfr: BatchNorm1d of wn_mlp is in training mode (batch statistics); call "
                                      "classifier.eval() first, as RegionSelection does (build.py:72-73)")
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
