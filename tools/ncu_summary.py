#!/usr/bin/env python3
"""tools/ncu_summary.py REPORT.ncu-rep [...] > summary.csv - the metrics profiles/README.md quotes, one row per
profiled launch, from `ncu --set full` reports (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

KEEP = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "gpu__time_duration.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]
w = csv.writer(sys.stdout)
first = True
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    stall = [k for k in h if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
             and "not_issued" not in k]
    cols = [k for k in KEEP if k in h] + stall
    if first:
        w.writerow(["report"] + [c.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for c in cols])
        w.writerow([""] + [units[h.index(c)] for c in cols])
        first = False
    for r in rows[2:]:
        d = dict(zip(h, r))
        w.writerow([rep.split("/")[-1]] + [d[c] for c in cols])
