"""Build A/B variants of libhalo_sm100.so (extra -D flags) into halo_b200/variants/<name>.so; select one at run time with
HALO_B200_LIB=<path>.  Usage: python tools/build_variants.py name1:-DFLAG1,-DFLAG2 name2: ..."""
import os
import shutil
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from halo_b200 import _build  # noqa: E402

out = os.path.join(os.path.dirname(_build.LIB), "variants")
os.makedirs(out, exist_ok=True)
for spec in sys.argv[1:]:
    name, _, flags = spec.partition(":")
    _build.build(force=True, extra_flags=[f for f in flags.split(",") if f])
    shutil.copy(_build.LIB, os.path.join(out, name + ".so"))
    print(name, flags)
_build.build(force=True)   # leave the default library in place
