"""Per-source-line stall samples / executed instructions of an ncu capture taken with -lineinfo and --import-source on.

    python tools/src_hot.py <file.ncu-rep> [top_n] [kernel-regex]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = "?"
agg = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ci_s, ci_e = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue   # keep the per-line summary rows (Address == "-")
    try:
        s, e = int(r[ci_s] or 0), int(r[ci_e] or 0)
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    st = agg.setdefault(key, [0, 0, r[1].strip(), {}])
    st[0] += s
    st[1] += e
    for i, name in stall_cols:
        try:
            st[3][name] = st[3].get(name, 0) + int(r[i] or 0)
        except ValueError:
            pass
tot_s = sum(v[0] for v in agg.values()) or 1
tot_e = sum(v[1] for v in agg.values()) or 1
print("total samples %d, warp instructions %d" % (tot_s, tot_e))
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    why = sorted(((c, n) for n, c in v[3].items()), reverse=True)[:2]
    print("%-24s:%4d %5.1f%% smp %5.1f%% exe  %-74s %s" % (f[:24], ln, 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_e, v[2][:74],
                                                      " ".join("%s=%d" % (n, c) for c, n in why if c)))
