#!/bin/sh
# FCCH search kernels A/B: headline FCCH leg and config 4's grid, fft vs direct
mkdir -p gpurun_out
for mode in fft direct; do
  if [ $mode = direct ]; then export GMR1B200_FCCH_FFT=0; else unset GMR1B200_FCCH_FFT; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/w_bench_$mode.json 2> gpurun_out/w_bench_$mode.err
  tail -2 gpurun_out/w_bench_$mode.err
  python - $mode <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/w_bench_{sys.argv[1]}.json").read().strip().splitlines()[-1])
c4 = d["configs"]["4"]
print(sys.argv[1], "value", round(d["value"] / 1e6, 1), "fcch ms", round(d["fcch"]["ms_per_step"], 4), "found", d["fcch"]["found_frac"],
      "| config4", round(c4["bursts_per_s"] / 1e6, 1), "fcch grid ms", c4["ms"]["fcch_5_shift_search_and_fine"], "share", round(c4["fcch_share_of_chain"], 3))
PY
done
