#!/bin/sh
# ncu counters of the demod kernels of configs 3 / 4 (NT3 speech, NT3 FACCH, NT9, RACH): instructions per burst
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,launch__grid_size,launch__registers_per_thread --clock-control none --kernel-name-base demangled -k regex:"demod_fast_kernel<\(int\)[457]|demod_fast_kernel<[457],|demod_kernel" -c 10 --csv --log-file gpurun_out/c3_demod.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/c3_demod.csv")) if len(r) > 10]
h = rows[0]
per = {}
for r in rows[1:]:
    d = dict(zip(h, r))
    per.setdefault((d["ID"], d["Kernel Name"][:70]), {})[d["Metric Name"]] = d["Metric Value"]
for (i, k), m in per.items():
    print(i, k, m)
PY
