#!/bin/sh
# channeliser kernels: time split and counters
mkdir -p gpurun_out
python tools/bench_chan.py
python tools/bench_chan.py --chans 320
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:"pfb_kernel|resamp_kernel" -s 4 -c 2 --csv --log-file gpurun_out/s_chan.csv python tools/bench_chan.py --reps 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/s_chan.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d["Kernel Name"][:40], d["Metric Name"], d["Metric Value"])
PY
