#!/bin/sh
mkdir -p gpurun_out
python tools/kernel_counters.py --out gpurun_out/k_kernel_counters.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/k_kernel_counters.json'))
for k,v in d['kernels'].items(): print(k, {a:(round(b,1) if isinstance(b,float) else b) for a,b in v.items() if a!='kernel'})
PY
ncu --set full --import-source on --clock-control none -k regex:fcch_rough -s 3 -c 1 -f -o gpurun_out/k_fcch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 --streams 1 > /dev/null 2>&1
ncu -i gpurun_out/k_fcch.ncu-rep --page source --csv > gpurun_out/k_fcch_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/k_fcch.ncu-rep > gpurun_out/k_fcch_summary.csv
rm -f gpurun_out/k_fcch.ncu-rep
