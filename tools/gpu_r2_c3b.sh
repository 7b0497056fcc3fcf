#!/bin/sh
# source-level profile of the NT3-FACCH and NT3-speech demod kernels (config 3)
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"demod_fast_kernel<\(int\)5|demod_fast_kernel<5," -c 1 -f -o gpurun_out/c3_facch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"demod_fast_kernel<\(int\)4|demod_fast_kernel<4," -c 1 -f -o gpurun_out/c3_speech python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > /dev/null 2>&1
ncu -i gpurun_out/c3_facch.ncu-rep --page source --csv > gpurun_out/c3_facch_source.csv 2>/dev/null
ncu -i gpurun_out/c3_speech.ncu-rep --page source --csv > gpurun_out/c3_speech_source.csv 2>/dev/null
rm -f gpurun_out/c3_facch.ncu-rep gpurun_out/c3_speech.ncu-rep
ls -la gpurun_out/c3_*source.csv
