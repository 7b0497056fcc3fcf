#!/bin/sh
# ncu --set full of the packed BCCH decode kernel (source-level stalls)
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:decode_tpc -s 4 -c 1 -f -o gpurun_out/o_dec python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 --streams 1 > /dev/null 2>&1
ncu -i gpurun_out/o_dec.ncu-rep --page source --csv > gpurun_out/o_dec_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/o_dec.ncu-rep > gpurun_out/o_dec_summary.csv
rm -f gpurun_out/o_dec.ncu-rep
cat gpurun_out/o_dec_summary.csv | python -c "
import sys, csv
r = list(csv.reader(sys.stdin))
for a, b in zip(r[0], r[2]): print(a, b)
"
