#!/bin/sh
# FCCH in the frequency domain: tests, A/B of the search kernels in the bench (headline + config 4), counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sdr_gpu.py tests/test_chain_gpu.py tests/test_pool_gpu.py tests/test_rxsched_gpu.py tests/test_rxcall_gpu.py -q -m gpu > gpurun_out/v_pytest.log 2>&1; tail -12 gpurun_out/v_pytest.log
for mode in fft direct; do
  if [ $mode = direct ]; then export GMR1B200_FCCH_FFT=0; else unset GMR1B200_FCCH_FFT; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/v_bench_$mode.json 2> gpurun_out/v_bench_$mode.err
  tail -2 gpurun_out/v_bench_$mode.err
  python - $mode <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/v_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
c4 = d["configs"]["4"]
print(sys.argv[1], "value", round(d["value"] / 1e6, 1), "fcch ms", round(d["fcch"]["ms_per_step"], 4), "found", d["fcch"]["found_frac"],
      "| config4", round(c4["bursts_per_s"] / 1e6, 1), "fcch grid ms", c4["ms"]["fcch_5_shift_search_and_fine"], "share", round(c4["fcch_share_of_chain"], 3),
      "parity", c4["parity_vs_cpu_reference"]["identical"])
PY
done
unset GMR1B200_FCCH_FFT
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:"fcch_" -s 6 -c 4 --csv --log-file gpurun_out/v_fcch.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --no-configs --min-seconds 0 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/v_fcch.csv")) if len(r) > 10]
h = rows[0]
for r in rows[1:]:
    d = dict(zip(h, r))
    print(d["ID"], d["Kernel Name"][:44], d["Metric Name"], d["Metric Value"])
PY
