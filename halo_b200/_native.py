"""ctypes binding of libhalo_sm100.so (the C ABI declared in include/halo_b200.h).

There is deliberately NO fallback: if the shared library is missing it is built with nvcc; if that
fails, or a CUDA device is absent when a kernel is requested, the call raises.  Tensors cross the
boundary as raw device pointers (`tensor.data_ptr()`), the stream as `torch.cuda.current_stream()`.
"""
import ctypes
import os
import threading

import torch

from . import _build

_lock = threading.Lock()
_lib = None

# enums of include/halo_b200.h
FEAT_TANGENT_F32, FEAT_BALL_F32, FEAT_BALL_F64 = 0, 1, 2
FEAT_FLAG_NO_TENSOR_CORE = 0x100
PIXUNC_ENTROPY, PIXUNC_ONE_MINUS_PGT = 0, 1
LABEL_ARGMAX, LABEL_GT_FILLED = 0, 1
NORM_RADIUS, NORM_EUCLID = 0, 1
UNC_BOXSUM, UNC_PIXEL, UNC_ZERO = 0, 1, 2
PUR_NORM, PUR_LABEL_HIST, PUR_RADIUS_BINS, PUR_ZERO = 0, 1, 2, 3
SELECT_KEEP_SCORE = 0x1

ERR_BAD_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = -1, -2, -3, -4
PATH_FWD_TC, PATH_FWD_CUDA_CORE, PATH_BWD_PIX_TC, PATH_BWD_PIX_CUDA_CORE, PATH_BWD_DW_TC, PATH_BWD_DW_CUDA_CORE = 1, 2, 4, 8, 16, 32
PATH_BWD_STREAM_TC, PATH_BWD_RECOMPUTE = 64, 128
ABI_VERSION = 2

_vp, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

_SIGNATURES = {
    "halo_abi_version": (_i, []),
    "halo_last_error": (ctypes.c_char_p, []),
    "halo_last_path": (_i, []),
    "halo_source_hash": (ctypes.c_char_p, []),
    "halo_head_workspace_bytes": (_sz, [_i, _i]),
    "halo_head_saved_rows": (_i, [_i, _i, _i, _i]),
    "halo_head_fwd": (_i, [_vp, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "halo_head_bwd_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "halo_head_bwd": (_i, [_vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "halo_expmap0_project": (_i, [_vp, _vp, _i, _f, _i, _i, _i, _i, _vp]),
    "halo_ball_norm": (_i, [_vp, _i, _f, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "halo_radius_f64": (_i, [_vp, _i, _f, _vp, _vp, _i, _i, _i, _i, _vp]),
    "halo_logits_stats": (_i, [_vp, _vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "halo_upsample_workspace_bytes": (_sz, [_i, _i, _i]),
    "halo_upsample_score_inputs": (_i, [_vp, _vp, _i, _f, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "halo_score_workspace_bytes": (_sz, [_i]),
    "halo_score": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "halo_select_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "halo_select_f32": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "halo_select_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "halo_round_delta_pack": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "halo_round_delta_apply": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "halo_round_row_bytes": (_sz, [_i, _i]),
    "halo_round_rows_pack": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "halo_round_rows_apply": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "halo_checksum64": (_i, [_vp, _sz, ctypes.c_ulonglong, _vp, _i, _vp]),
    "halo_reduce_hfr_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "halo_reduce_hfr_train_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "halo_reduce_hfr_train_fwd": (_i, [_vp] * 7 + [_f] + [_vp] * 8 + [_i] * 5 + [_vp, _sz, _vp]),
    "halo_reduce_hfr_train_bwd": (_i, [_vp] * 5 + [_f] + [_vp] * 15 + [_i] * 6 + [_vp, _sz, _vp]),
    "halo_reduce_hfr_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
    "halo_seg_loss_workspace_bytes": (_sz, []),
    "halo_seg_loss": (_i, [_vp, _vp, _f, _f, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _sz, _vp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _build.LIB


def load(build_if_missing=True):
    """Load (building first if needed) libhalo_sm100.so and type its entry points."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = os.environ.get("HALO_B200_LIB")  # override: A/B experiments with differently built libraries
        if path is None:
            path = _build.LIB
            if _build.stale():      # missing, or built from other sources than the ones in the tree (content hash)
                if not build_if_missing:
                    raise RuntimeError("libhalo_sm100.so is missing or stale: run `python -m halo_b200._build`")
                _build.build()
        elif not os.path.exists(path):
            raise RuntimeError("HALO_B200_LIB=%s does not exist" % path)
        lib = ctypes.CDLL(path)
        override = os.environ.get("HALO_B200_LIB") is not None
        for name, (res, args) in _SIGNATURES.items():
            try:
                fn = getattr(lib, name)  # AttributeError if the symbol is not exported: fail loudly
            except AttributeError:
                if override:             # A/B runs against a library built from an older commit: tolerate missing entry points
                    continue
                raise
            fn.restype = res
            fn.argtypes = args
        if lib.halo_abi_version() != ABI_VERSION:
            raise RuntimeError("libhalo_sm100.so ABI version mismatch")
        _lib = lib
    return _lib


def last_error():
    return load().halo_last_error().decode("utf-8", "replace")


_PATH_NAMES = ((PATH_FWD_TC, "fwd:tcgen05"), (PATH_FWD_CUDA_CORE, "fwd:cuda_core"), (PATH_BWD_PIX_TC, "bwd_pix:tcgen05"),
               (PATH_BWD_PIX_CUDA_CORE, "bwd_pix:cuda_core"), (PATH_BWD_DW_TC, "bwd_dw:tcgen05"),
               (PATH_BWD_DW_CUDA_CORE, "bwd_dw:cuda_core"), (PATH_BWD_STREAM_TC, "bwd:stream_tcgen05"),
               (PATH_BWD_RECOMPUTE, "bwd:recompute"))


def last_path():
    """Kernel variants the last head forward / backward call of this thread launched, e.g. ("fwd:tcgen05",)."""
    bits = load().halo_last_path()
    return tuple(name for bit, name in _PATH_NAMES if bits & bit)


def check(rc, what):
    if rc == 0:
        return
    msg = "%s failed (%d): %s" % (what, rc, last_error())
    if rc == ERR_BAD_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def require_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(
            "halo_b200: %s must be a CUDA tensor -- this package has no CPU path (got device %s)" % (name, t.device))
    return t


class Workspace:
    """Grow-only uint8 device scratch, one per (device, stream, purpose); owned by torch's allocator.

    Keyed by the CURRENT stream: every entry point writes its scratch (packed class parameters, extrema, partials) on
    the caller's stream, so two streams (or threads driving different streams) must never share a buffer; and a buffer is
    only ever used on the stream it was allocated on, which is what makes handing it back to torch's caching allocator on
    re-growth safe without record_stream."""

    def __init__(self):
        self._bufs = {}
        self._lock = threading.Lock()

    def get(self, device, key, nbytes):
        k = (device.index, torch.cuda.current_stream(device).cuda_stream, key)
        with self._lock:
            buf = self._bufs.get(k)
            if buf is None or buf.numel() < nbytes:
                with torch.cuda.device(device):
                    buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
                self._bufs[k] = buf
        return buf


workspace = Workspace()
