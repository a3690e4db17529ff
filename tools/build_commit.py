"""Build libhalo_sm100.so as of an older commit into halo_b200/variants/<name>.so (bisecting with HALO_B200_LIB=...).
Usage: python tools/build_commit.py <commit> <name> [-DFLAG ...]"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from halo_b200 import _build  # noqa: E402

commit, name, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
out = os.path.join(ROOT, "halo_b200", "variants")
src = os.path.join(out, "src_" + name)
os.makedirs(src, exist_ok=True)
tar = subprocess.run(["git", "-C", ROOT, "archive", commit, "halo_b200/csrc", "include"], stdout=subprocess.PIPE, check=True).stdout
subprocess.run(["tar", "-x", "-C", src], input=tar, check=True)
csrc = os.path.join(src, "halo_b200", "csrc")
cus = sorted(f for f in os.listdir(csrc) if f.endswith(".cu"))


def cc(f):
    o = os.path.join(src, f[:-3] + ".o")
    flags = [x for x in _build.NVCC_FLAGS if x not in ("-Xptxas", "-v")]
    r = subprocess.run([_build._nvcc()] + flags + extra + ['-DHALO_SOURCE_HASH="variant"', "-c", os.path.join(csrc, f), "-o", o],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        raise RuntimeError(r.stdout)
    return o


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, cus))
lib = os.path.join(out, name + ".so")
subprocess.run([_build._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs, check=True)
for o in objs:
    os.remove(o)
print(lib)
