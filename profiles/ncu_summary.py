#!/usr/bin/env python
"""Print the handful of ncu raw metrics we track for a kernel (reads an .ncu-rep here, no GPU).
Usage: python profiles/ncu_summary.py <report.ncu-rep> [kernel-index]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    d = dict(zip(hdr, zip(units, vals)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1][:100])
    for k in KEYS:
        if k in d:
            print("%-82s %-10s %s" % (k, d[k][0], d[k][1]))
    st = [(float(v[1]), k[len(STALL):].replace("_per_issue_active.ratio", "")) for k, v in d.items()
          if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v[1]]
    print("stall cycles per issued instruction:", ", ".join("%s %.2f" % (n, x) for x, n in sorted(st, reverse=True)[:9]))


if __name__ == "__main__":
    main()
