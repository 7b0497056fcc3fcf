#!/bin/sh
mkdir -p gpurun_out
V=osmo_gmr_b200/build/variants
sh tools/ab_bench.sh d_ab $V/libch1.so $V/libch2.so $V/libch3.so $V/libch3c9.so $V/libch3c10.so
sh tools/ncu_demod.sh d_ncu $V/libch3c9.so $V/libch3c10.so
sh tools/ncu_full.sh d_full
ncu -i gpurun_out/d_full_main.ncu-rep --page source --csv > gpurun_out/d_full_main_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/d_full_main.ncu-rep > gpurun_out/d_full_summary.csv
rm -f gpurun_out/d_full_main.ncu-rep
