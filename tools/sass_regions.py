"""Stall samples of an ncu capture bucketed by SASS address window, with the notable opcodes of each window (role finder).

    python tools/sass_regions.py <file.ncu-rep> [window] [tiles-for-normalisation]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
W = int(sys.argv[2]) if len(sys.argv) > 2 else 150
tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 51200.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    try:
        data.append((r[ci["Source"]], int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0)))
    except ValueError:
        pass
tot = sum(d[1] for d in data) or 1
KEYS = ("UTMALDG", "UTCHMMA", "LDTM", "STTM", "SHFL", "MUFU", "STG", "LDG", "LDS.128", "STS", "BAR.SYNC", "LDL", "STL")
for i in range(0, len(data), W):
    chunk = data[i:i + W]
    s = sum(d[1] for d in chunk)
    e = sum(d[2] for d in chunk)
    ops = {}
    for src, _, _ in chunk:
        for key in KEYS:
            if key in src:
                ops[key] = ops.get(key, 0) + 1
    wait = sum(d[1] for d in chunk if "NANOSLEEP" in d[0] or "SYNCS" in d[0] or ("BRA" in d[0] and d[1] > 0.002 * tot))
    print("%5d %5.1f%% smp (%4.1f%% in waits) %6.2fK exe/tile  %s" % (i, 100.0 * s / tot, 100.0 * wait / tot, e / tiles / 1e3,
                                                                      " ".join("%s:%d" % kv for kv in sorted(ops.items()))))
