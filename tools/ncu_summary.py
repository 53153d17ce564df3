#!/usr/bin/env python3
"""Key metrics of every kernel in an .ncu-rep (raw page): python tools/ncu_summary.py file.ncu-rep"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread ", "launch__occupancy_limit", "smsp__issue_active.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum "]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
names = [r[hdr.index("Kernel Name")][:60] for r in rows[2:]]
print("kernels:", names)
for i, h in enumerate(hdr):
    if any((h + " ").startswith(w) or w in h + " " for w in WANT):
        vals = [r[i] for r in rows[2:]]
        if any(v not in ("0", "", "0.000000") for v in vals):
            print("%-90s %-10s %s" % (h, rows[1][i], vals))
