#!/bin/sh
# source-level profiles of the decode kernels of configs 3 / 4 (second launch of each): TCH3 pairs (8,1), TCH9-9k6 (6,0),
# FACCH9 (3,0), RACH (7,0), and the RACH / NT9 demod kernels
mkdir -p gpurun_out
cap() {
	timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$2" -s 1 -c 1 -f -o gpurun_out/s4_$1 python tools/bench_configs.py --reps 1 > /dev/null 2>&1
	ncu -i gpurun_out/s4_$1.ncu-rep --page source --csv > gpurun_out/s4_$1_source.csv 2>/dev/null
}
cap tch3 "decode_tpc_kernel<\(int\)8, \(bool\)1>"
cap tch9 "decode_tpc_kernel<\(int\)6, \(bool\)0>"
cap facch9 "decode_tpc_kernel<\(int\)3, \(bool\)0>"
cap rachdec "decode_tpc_kernel<\(int\)7, \(bool\)0>"
cap rachdem "demod_fast_kernel<\(int\)8, "
cap nt9dem "demod_fast_kernel<\(int\)7, "
python tools/ncu_summary.py gpurun_out/s4_*.ncu-rep > gpurun_out/s4_summary.csv
rm -f gpurun_out/s4_*.ncu-rep
ls -la gpurun_out/s4_*
