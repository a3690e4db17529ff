"""CPU: the C-ABI library builds/loads, exports every symbol include/halo_b200.h declares, and rejects bad
arguments through the documented error channel -- no kernel launches, no GPU needed."""
import ctypes
import os
import re

import pytest
import torch

from halo_b200 import _native as nat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "halo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(halo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = nat.load()
    syms = header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libhalo_sm100.so does not export %s" % s
    assert set(syms) == set(nat.EXPORTED_SYMBOLS), "python binding and header disagree"
    assert lib.halo_abi_version() == nat.ABI_VERSION == 2


def test_library_carries_the_hash_of_its_sources():
    """A stale library (sources edited, .so not rebuilt) is detected by content, not by mtime."""
    from halo_b200 import _build

    lib = nat.load()
    assert lib.halo_source_hash().decode() == "HALO_SRC_SHA256=" + _build.source_hash()
    assert _build.built_hash() == _build.source_hash() and not _build.stale()
    assert lib.halo_last_path() == 0 or lib.halo_last_path() > 0   # exported, callable without a GPU


def test_workspace_queries_are_pure():
    lib = nat.load()
    std = (256 * 40 + 80) * 4  # CUDA-core pack: Wt[C][2*OP] + cls[4][OP]
    tc = (2 * 48 * 256 + 80) * 4  # tensor-core pack: hi/lo planes [2][C/4][NP][4] + cls
    assert lib.halo_head_workspace_bytes(19, 256) == (std + 255) // 256 * 256 + tc
    assert lib.halo_head_workspace_bytes(0, 256) == 0
    assert lib.halo_score_workspace_bytes(3) == 48
    assert lib.halo_select_workspace_bytes(2, 640, 1280, 4552) >= 2 * 4552 * 4 + 2 * 640 * 1280 // 8
    assert lib.halo_head_bwd_workspace_bytes(1, 64, 19, 16, 16) > 0


def test_bad_arguments_return_status_not_crash():
    lib = nat.load()
    rc = lib.halo_head_fwd(None, 0, None, None, 1.0, None, None, None, None, None, None, None, 0, 0, 0, 1, 8, 19, 4, 4, None, 0, None)
    assert rc == nat.ERR_BAD_ARG and "NULL" in nat.last_error()
    one = ctypes.c_void_p(16)  # never dereferenced: argument checks come first
    rc = lib.halo_head_fwd(one, 0, one, one, -1.0, None, None, None, None, None, None, None, 0, 0, 0, 1, 8, 19, 4, 4, one, 1 << 20, None)
    assert rc == nat.ERR_BAD_ARG and "curvature" in nat.last_error()
    rc = lib.halo_head_fwd(one, 0, one, one, 1.0, None, None, None, None, None, None, None, 0, 0, 0, 1, 8, 40, 4, 4, one, 1 << 20, None)
    assert rc == nat.ERR_UNSUPPORTED
    rc = lib.halo_head_fwd(one, 0, one, one, 1.0, None, None, None, None, None, None, None, 0, 0, 0, 1, 8, 19, 4, 4, one, 8, None)
    assert rc == nat.ERR_WORKSPACE
    assert lib.halo_head_saved_rows(256, 19, 640, 1280) == 41 and lib.halo_head_saved_rows(64, 16, 8, 8) == 33
    assert lib.halo_head_saved_rows(48, 19, 8, 8) == 0 and lib.halo_head_saved_rows(256, 28, 8, 8) == 0
    rc = lib.halo_head_fwd(one, 0, one, one, 1.0, None, None, None, None, None, one, None, 0, 0, 0, 1, 48, 19, 4, 4, one, 1 << 20, None)
    assert rc == nat.ERR_UNSUPPORTED and "saved" in nat.last_error()   # a shape that cannot save says so
    rc = lib.halo_score(one, one, None, None, None, None, None, 0, 0, 1, 4, 3, 19, one, None, one, 1, 8, 8, one, 64, None)
    assert rc == nat.ERR_BAD_ARG and "odd" in nat.last_error()
    rc = lib.halo_round_delta_pack(None, one, one, one, 1, 4, 8, 8, 1, None)
    assert rc == nat.ERR_BAD_ARG and "halo_round_delta_pack" in nat.last_error()
    rc = lib.halo_round_delta_apply(one, one, one, one, one, 1, 0, 8, 8, 1, None)
    assert rc == nat.ERR_BAD_ARG and "halo_round_delta_apply" in nat.last_error()
    rc = lib.halo_upsample_score_inputs(None, one, 0, 1.0, None, 0, 0, 0, None, None, one, None, None, None, 1, 0, 8, 0, 0, 4, 4, 8, 8, None, 0, None)
    assert rc == nat.ERR_WORKSPACE   # an embedding always needs the Gram workspace
    rc = lib.halo_score(one, None, None, one, None, None, None, 0, 2, 1, 3, 3, 100, one, one, one, 1, 8, 8, one, 64, None)
    assert rc == nat.ERR_BAD_ARG and "radius" in nat.last_error()   # fp64 radius plane without its extrema
    rc = lib.halo_radius_f64(one, 7, 1.0, one, one, 1, 8, 4, 4, None)
    assert rc == nat.ERR_BAD_ARG and "feat_kind" in nat.last_error()
    rc = lib.halo_select_f32(one, one, one, one, one, -1, 1, 5, 0, one, None, 1, 8, 8, one, 1 << 20, None)
    assert rc == nat.ERR_BAD_ARG
    with pytest.raises(ValueError):
        nat.check(nat.ERR_BAD_ARG, "x")
    with pytest.raises(NotImplementedError):
        nat.check(nat.ERR_UNSUPPORTED, "x")


def test_no_cpu_path():
    """The product must fail loudly on CPU tensors instead of falling back."""
    import halo_b200

    u = torch.randn(1, 8, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        halo_b200.HyperMapper(1.0).expmap(u, dim=1)
    with pytest.raises(RuntimeError, match="no CPU path"):
        halo_b200.head_forward(u, torch.randn(19, 8), torch.randn(19, 8), 1.0)
    with pytest.raises(RuntimeError, match="no CPU path"):
        halo_b200.select_pixels_to_label(torch.rand(4, 4), 1, 1, 1, torch.zeros(4, 4, dtype=torch.bool),
                                         torch.zeros(4, 4, dtype=torch.bool), torch.zeros(4, 4, dtype=torch.long),
                                         torch.zeros(4, 4, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "halo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


def test_header_is_plain_c():
    """The boundary is a C ABI: include/halo_b200.h compiles as C99 and as C++17 on its own (no torch / CUDA types)."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "halo_b200.h")
    for cc, args in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++17", "-x", "c++"])):
        exe = shutil.which(cc)
        if exe is None:
            pytest.skip("%s not installed" % cc)
        r = subprocess.run([exe, "-fsyntax-only", "-Wall", "-Wextra", "-Werror"] + args + [hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    assert "at::Tensor" not in text and "#include <torch" not in text   # plain pointers and sizes only
