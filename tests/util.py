import numpy as np
import torch

TOL = 1e-5  # north_star: logits, radius and scores within 1e-5 relative (of the max magnitude) in fp32


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_err(got, ref):
    """max|got-ref| / max|ref| over finite entries; NaN/inf patterns must match exactly."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(got), fin), "non-finite pattern differs"
    if fin.sum() == 0:
        return 0.0
    scale = ref[fin].abs().max().item()
    return (got[fin] - ref[fin]).abs().max().item() / max(scale, 1e-30)
