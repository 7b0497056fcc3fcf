#!/bin/sh
# channeliser, second generation: tests, timing, counters, wideband leg of bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chan_gpu.py -q > gpurun_out/t_pytest.log 2>&1; tail -15 gpurun_out/t_pytest.log
python tools/bench_chan.py
python tools/bench_chan.py --chans 256
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_bytes.sum --clock-control none -k regex:"pfb_|resamp_kernel" -s 4 -c 2 --csv --log-file gpurun_out/t_chan.csv python tools/bench_chan.py --reps 2 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/t_chan.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d["Kernel Name"][:40], d["Metric Name"], d["Metric Value"])
PY
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
tail -5 gpurun_out/t_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/t_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2))
print(json.dumps(d["wideband"], indent=1))
PY
