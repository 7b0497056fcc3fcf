#!/bin/sh
# source-level profile of the TCH3 decode kernel (two codewords per thread) of config 3
mkdir -p gpurun_out
python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"decode_tpc_kernel<\(int\)[0-9]+, \(bool\)1>" -s 2 -c 1 -f -o gpurun_out/t3src python tools/bench_configs.py > /dev/null 2>&1
ncu -i gpurun_out/t3src.ncu-rep --page source --csv > gpurun_out/t3src_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/t3src.ncu-rep > gpurun_out/t3src_summary.csv
rm -f gpurun_out/t3src.ncu-rep
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/t3src_bench.json 2>gpurun_out/t3src_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/t3src_bench.json'))
print(d['value'], d['viterbi']['ms_per_launch'], d['configs']['3']['ms'], d['configs']['4']['ms'])
P
