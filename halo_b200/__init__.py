"""halo_b200 -- B200-native (sm_100a) implementation of HALO's per-pixel hyperbolic hot path.

Mirrors the reference's Python call surface for that path and nothing else:

    core/utils/hyperbolic.py       -> halo_b200.hyperbolic       (HyperMapper, HyperMLR)
    core/active/floating_region.py -> halo_b200.floating_region  (FloatingRegionScore)
    core/active/build.py           -> halo_b200.active           (select_pixels_to_label, RegionSelection)
    mask / indicator files         -> halo_b200.maskio           (build.py:162-166, cityscapes.py:234,245-251)
    CE + negative-learning loss    -> halo_b200.losses           (train_learners.py:343-356, loss/negative_learning_loss.py)
    conv_reduce + HFR (eval/train) -> halo_b200.hfr              (models/classifier.py:526-550)

All arithmetic runs in hand-written CUDA behind the C ABI of include/halo_b200.h
(halo_b200/libhalo_sm100.so); there is no CPU path and no fallback.
"""
from . import _native  # noqa: F401
from .hyperbolic import HyperMapper, HyperMLR, PoincareEmbedding, head_forward, head_backward  # noqa: F401
from .floating_region import FloatingRegionScore  # noqa: F401
from .active import RegionSelection, select_pixels_to_label, select_planes  # noqa: F401
from .pool import AcquisitionConfig, acquire_batch, acquire_pool  # noqa: F401
from .maskio import AsyncMaskWriter, read_indicator, read_mask  # noqa: F401
from .losses import fused_seg_loss  # noqa: F401
from .hfr import reduce_hfr  # noqa: F401
from . import synth  # noqa: F401
from .dropin import install  # noqa: F401

__version__ = "0.1.0"
