"""Oracle for the channel reduction + hyperbolic feature re-weighting upstream of the head (ORACLE -- test infrastructure,
CPU only).  Restates core/models/classifier.py:526-550 (DepthwiseSeparableASPP_Hyper.forward; the v2 head repeats the block
at :187-214) with the same torch calls, on modules built like :478-494 (conv_reduce, wn_mlp)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def build_modules(cin, c, hfr=True, seed=0):
    """conv_reduce (classifier.py:478-480) and wn_mlp (:486-492) with default torch initialisation, deterministic."""
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    conv_reduce = nn.Conv2d(cin, c, kernel_size=1)
    wn_mlp = None
    if hfr:
        wn_mlp = nn.Sequential(nn.Linear(c, c), nn.BatchNorm1d(c), nn.ReLU(), nn.Linear(c, c))
        with torch.no_grad():   # running statistics as after some training, not the identity
            wn_mlp[1].running_mean.uniform_(-0.2, 0.2)
            wn_mlp[1].running_var.uniform_(0.5, 1.5)
            wn_mlp[1].weight.uniform_(0.5, 1.5)
            wn_mlp[1].bias.uniform_(-0.2, 0.2)
    torch.random.set_rng_state(g)
    return conv_reduce, wn_mlp


def reduce_hfr(decoder_out, conv_reduce, wn_mlp):
    """classifier.py:527-550, verbatim call sequence."""
    decoder_out = conv_reduce(decoder_out)
    if wn_mlp is not None:
        temp_out = decoder_out.permute(0, 2, 3, 1).contiguous().view(-1, decoder_out.size(1))
        norm_weights = wn_mlp(temp_out)
        norm_weights = norm_weights.view(-1, decoder_out.size(2) * decoder_out.size(3), decoder_out.size(1))
        norm_weights = torch.mean(norm_weights, dim=1, keepdim=False)
        norm_weights = norm_weights.view(-1, decoder_out.size(1), 1, 1)
        norm_weights = torch.clamp(norm_weights, min=1e-5)
        temp_out = decoder_out.reshape(-1, decoder_out.size(1), decoder_out.size(2) * decoder_out.size(3))
        temp_out = F.normalize(temp_out, dim=-1)
        temp_out = temp_out.reshape(-1, decoder_out.size(1), decoder_out.size(2), decoder_out.size(3))
        decoder_out = temp_out * norm_weights
    return decoder_out
