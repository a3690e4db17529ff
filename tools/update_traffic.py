"""Write profiles/k1_traffic.json from an `ncu --set full` capture of K1 (head_fwd_tc_kernel): DRAM bytes per launch and per
image, tagged with the sha256 of csrc/head_fwd_tc.cu so that bench.py drops the figure as soon as the kernel changes.

    python tools/update_traffic.py gpurun_out/prof_k1.ncu-rep <images_per_launch> [note]
"""
import csv
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, images = sys.argv[1], int(sys.argv[2])
    note = sys.argv[3] if len(sys.argv) > 3 else os.path.basename(rep)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i, rd_i, wr_i = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")

    def to_bytes(v, u):
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u]
        return float(v.replace(",", "")) * scale

    k1 = [r for r in rows[2:] if "head_fwd_tc_kernel" in r[name_i]]
    if not k1:
        sys.exit("no head_fwd_tc_kernel launch in %s" % rep)
    rd = sum(to_bytes(r[rd_i], units[rd_i]) for r in k1) / len(k1)
    wr = sum(to_bytes(r[wr_i], units[wr_i]) for r in k1) / len(k1)
    with open(os.path.join(ROOT, "halo_b200", "csrc", "head_fwd_tc.cu"), "rb") as f:
        sha = hashlib.sha256(f.read()).hexdigest()
    doc = {"kernel": k1[0][name_i][:60], "source": "ncu --set full --clock-control none, %s" % note, "images_per_launch": images,
           "launches_averaged": len(k1), "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_image": (rd + wr) / images, "kernel_source_sha256": sha}
    with open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w") as f:
        json.dump(doc, f, indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
