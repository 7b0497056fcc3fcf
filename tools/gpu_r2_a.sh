#!/bin/sh
# round 2, GPU call A: parity of the per-format demod kernels + A/B against the generic kernel + instruction counts
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
timeout 900 python -m pytest tests/test_demod_gpu.py -x -q -m gpu > gpurun_out/a_pytest_demod.log 2>&1; echo "pytest demod rc=$?" 
tail -5 gpurun_out/a_pytest_demod.log
for pass in 1 2; do
  for v in fast generic; do
    f=""; [ $v = generic ] && f="--demod-generic"
    timeout 600 python bench.py --steps 30 --warmup 3 --no-cpu-baseline $f > gpurun_out/a_bench_${v}_$pass.json 2> gpurun_out/a_bench_${v}_$pass.err
    python - $v gpurun_out/a_bench_${v}_$pass.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1], round(d["value"] / 1e6, 1), "Mb/s step", round(d["ms_per_step"], 4), "demod ms", round(r["ms_per_launch"], 4),
          "frac", round(r["frac"], 4), "vit", round(d["viterbi"]["ms_per_launch"], 4), "fcch", round(d["fcch"]["ms_per_step"], 4),
          d["e2e"]["same_results_as_device_path"], d["crc_ok_frac"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
done
sh tools/ncu_demod.sh a_ncu
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/a_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -5 gpurun_out/a_pytest_all.log
