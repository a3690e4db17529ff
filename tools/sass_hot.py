"""Hot spots of an `ncu --set full --import-source on` capture: top SASS instructions by stall samples and by executed count.

    python tools/sass_hot.py <file.ncu-rep> [top_n] [kernel-regex]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if len(sys.argv) > 3:
    cmd += ["-k", "regex:" + sys.argv[3]]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for n, r in enumerate(rows[hdr_i + 1:]):
    if len(r) < len(hdr) or r[0] == "Address":
        continue
    try:
        data.append((n, r[ci["Source"]], int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), r))
    except ValueError:
        pass
tot_s = sum(d[2] for d in data) or 1
tot_i = sum(d[3] for d in data) or 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples %d, total warp instructions %d, %d SASS lines" % (tot_s, tot_i, len(data)))
print("--- top by stall samples")
for n, src, s, e, r in sorted(data, key=lambda d: -d[2])[:top]:
    why = sorted(((int(r[ci[h]] or 0), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5d %5.1f%% smp %5.1f%% exe  %-70s %s" % (n, 100.0 * s / tot_s, 100.0 * e / tot_i, src[:70], " ".join("%s=%d" % (h, c) for c, h in why if c)))
print("--- top by executed instructions")
for n, src, s, e, r in sorted(data, key=lambda d: -d[3])[:top]:
    print("%5d %5.1f%% exe %5.1f%% smp  %s" % (n, 100.0 * e / tot_i, 100.0 * s / tot_s, src[:90]))
