"""Build libhalo_sm100.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhalo_sm100.so")
SOURCES = ["abi.cu", "head_fwd.cu", "head_fwd_tc.cu", "head_misc.cu", "head_bwd.cu", "head_bwd_tc.cu", "head_bwd_dw_tc.cu", "head_bwd_stream_tc.cu", "score.cu", "upsample.cu", "select.cu", "delta.cu", "seg_loss.cu", "reduce_hfr.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; libhalo_sm100.so cannot be built")
    return exe


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _deps():
    deps = sources() + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cuh")]
    return deps + [os.path.join(HERE, "..", "include", "halo_b200.h")]


def source_hash():
    """Content hash of everything the library is built from (mtimes do not survive the copy to the GPU box)."""
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in _deps():
        with open(d, "rb") as f:
            h.update(os.path.basename(d).encode() + b"\0" + f.read())
    return h.hexdigest()


HASH_MARK = b"HALO_SRC_SHA256="


def built_hash(path=None):
    """The source hash compiled into the library (abi.cu embeds it as a string: it travels with the .so)."""
    path = path or LIB
    if not os.path.exists(path):
        return None
    with open(path, "rb") as f:
        blob = f.read()
    i = blob.find(HASH_MARK)
    return blob[i + len(HASH_MARK):i + len(HASH_MARK) + 64].decode("ascii", "replace") if i >= 0 else None


def stale():
    """True when the library is missing or was built from different sources than the ones in the tree."""
    return built_hash() != source_hash()


def build(force=False, verbose=False, extra_flags=None):
    """Compile every .cu to an object (in parallel) and link the shared library.  Returns the path."""
    if not force and not stale():
        return LIB
    import fcntl

    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    with open(os.path.join(HERE, "build", ".lock"), "w") as lock:   # ranks of one job may all find the library stale
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not stale():
            return LIB
        return _build_locked(force, verbose, extra_flags)


def _build_locked(force, verbose, extra_flags):
    nvcc = _nvcc()
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    extra = list(extra_flags or os.environ.get("HALO_EXTRA_NVCC", "").split())
    stamp = os.path.join(objdir, "flags.txt")
    flags_changed = (not os.path.exists(stamp)) or open(stamp).read() != " ".join(NVCC_FLAGS + extra)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + \
        [os.path.join(HERE, "..", "include", "halo_b200.h")]
    newest_header = max(os.path.getmtime(h) for h in headers)
    up_to_date = []
    sha = source_hash()
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        is_abi = os.path.basename(src) == "abi.cu"   # carries the source hash: recompiled on every build (a second)
        if (not force and not flags_changed and not is_abi and os.path.exists(obj)
                and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header)):
            up_to_date.append(obj)  # incremental: only recompile what changed
            continue
        cmd = [nvcc] + NVCC_FLAGS + extra + (['-DHALO_SOURCE_HASH="%s"' % sha] if is_abi else []) + ["-c", src, "-o", obj]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, logs = list(up_to_date), []
    for src, obj, p in procs:
        out, _ = p.communicate()
        logs.append(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        objs.append(obj)
    with open(stamp, "w") as f:
        f.write(" ".join(NVCC_FLAGS + extra))
    with open(os.path.join(objdir, "ptxas.log"), "a" if up_to_date else "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
