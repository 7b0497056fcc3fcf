#!/bin/sh
# FCCH second-generation search: parity, A/B against the first-generation kernel (GMR1B200_FCCH_OLD=1), ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sdr_gpu.py tests/test_fullsize_gpu.py tests/test_rxsched_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in new old; do
  if [ $v = old ]; then export GMR1B200_FCCH_OLD=1; else unset GMR1B200_FCCH_OLD; fi
  python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-sweep --min-seconds 0 > gpurun_out/m_bench_$v.json 2> gpurun_out/m_bench_$v.err
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/m_bench_{v}.json").read().strip().splitlines()[-1])
    print(v, "value", round(d["value"] / 1e6, 1), "fcch", d["fcch"]["ms_per_step"], d["fcch"]["toa_identical_to_reference"] if "toa_identical_to_reference" in d["fcch"] else None,
          "cfg4", d["configs"]["4"]["ms"], d["configs"]["4"]["bursts_per_s"] / 1e6, d["configs"]["4"].get("parity_vs_cpu_reference", {}).get("identical"))
except Exception as e:
    print(v, "FAILED", e)
PY
  tail -2 gpurun_out/m_bench_$v.err
done
unset GMR1B200_FCCH_OLD
ncu --set full --import-source on --clock-control none -k regex:fcch_grid -s 3 -c 1 -f -o gpurun_out/m_fcch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 --streams 1 > /dev/null 2>&1
ncu -i gpurun_out/m_fcch.ncu-rep --page source --csv > gpurun_out/m_fcch_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/m_fcch.ncu-rep > gpurun_out/m_fcch_summary.csv
rm -f gpurun_out/m_fcch.ncu-rep
cat gpurun_out/m_fcch_summary.csv | python -c "
import sys, csv
r = list(csv.reader(sys.stdin))
for a, b in zip(r[0], r[2]): print(a, b)
"
