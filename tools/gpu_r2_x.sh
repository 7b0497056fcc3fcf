#!/bin/sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sdr_gpu.py tests/test_chan_gpu.py -q -m gpu 2>&1 | tail -4
sh tools/gpu_r2_w.sh
python tools/bench_chan.py
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"pfb_|resamp_kernel" -s 4 -c 2 --csv --log-file gpurun_out/x_chan.csv python tools/bench_chan.py --reps 2 > /dev/null 2>&1
grep "gpu__time\|inst_executed\|issue_active" gpurun_out/x_chan.csv | awk -F'","' '{print substr($5,1,36), $13, $15}'
