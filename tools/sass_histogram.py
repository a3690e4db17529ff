"""profiles/sass_<kernel file>.txt: instruction-mnemonic histogram of the tensor-core kernels' SASS (evidence that the hot
paths are tcgen05 / TMEM / TMA code).   python tools/sass_histogram.py   (after a build; reads halo_b200/build/*.o)"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "FADD2", "FMUL2", "FFMA2", "MUFU", "SHFL",
        "LDS", "STS", "LDG", "STG", "LDL", "STL", "FFMA", "FMUL", "FADD", "IMAD", "BAR", "ELECT", "USETMAXREG"]
for name in ("head_fwd_tc", "head_bwd_stream_tc", "head_bwd_tc", "head_bwd_dw_tc"):
    obj = os.path.join(ROOT, "halo_b200", "build", name + ".o")
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    kernels = len(re.findall(r"^\s*Function : ", sass, flags=re.M))
    ops = collections.Counter()
    for m in re.finditer(r"^\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", sass, flags=re.M):
        ops[m.group(1)] += 1
    with open(os.path.join(ROOT, "profiles", "sass_%s.txt" % name), "w") as f:
        f.write("# cuobjdump -sass halo_b200/build/%s.o : %d kernels (template instantiations), %d instructions\n" % (name, kernels, sum(ops.values())))
        f.write("# mnemonic count (Blackwell evidence: UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA load, SYNCS = mbarrier,\n"
                "# FADD2/FMUL2/FFMA2 = packed fp32x2, USETMAXREG = setmaxnreg)\n")
        for k in KEEP:
            if ops.get(k):
                f.write("%-12s %d\n" % (k, ops[k]))
        f.write("# top 25 overall\n")
        for k, v in ops.most_common(25):
            f.write("%-12s %d\n" % (k, v))
    print(name, kernels, sum(ops.values()), {k: ops[k] for k in KEEP[:8] if ops.get(k)})
