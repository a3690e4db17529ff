"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of numbers the roofline needs."""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit",
    "sm__inst_executed_pipe_fma", "sm__pipe_fma_cycles_active", "sm__pipe_fmaheavy", "sm__inst_executed_pipe_lsu",
    "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_xu", "sm__pipe_tensor", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "smsp__issue_active.avg.pct",
    "smsp__average_warp", "smsp__warps_issue_stalled", "local_load", "local_store", "lts__t_sector_hit_rate", "sm__cycles_elapsed.max",
    "l1tex__t_bytes_pipe_lsu_mem_local", "smsp__inst_executed_op_local", "lts__t_bytes.sum", "l1tex__throughput.avg.pct",
    "lts__throughput.avg.pct", "smsp__cycles_active.avg", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
    "smsp__pcsamp_warps_issue_stalled", "sm__warps_active", "achieved_occupancy", "launch__waves_per_multiprocessor",
]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")][:100], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            if any(k in h for k in KEYS + extra):
                print("  %-95s %-14s %s" % (h, u, v))


if __name__ == "__main__":
    main()
