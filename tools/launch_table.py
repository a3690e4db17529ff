"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # launches to skip per kernel (warm-up)
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        us = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        agg.setdefault(row["Kernel Name"].split("(")[0][:60], []).append(us)
    tot = sum(sum(v[skip:]) for v in agg.values())
    print("| kernel | launches | avg µs | share |\n|---|---|---|---|")
    for k, v in agg.items():
        w = v[skip:] or v
        print("| `%s` | %d | %.1f | %.1f %% |" % (k, len(w), sum(w) / len(w), 100 * sum(w) / tot))


if __name__ == "__main__":
    main()
