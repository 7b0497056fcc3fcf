#!/bin/sh
# source-level profile of the BCCH decode kernel and the BCCH demod kernel of a bench step
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"decode_tpc_kernel" -s 3 -c 1 -f -o gpurun_out/dsrc_dec python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --no-wideband --min-seconds 0 --streams 1 > /dev/null 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"demod_fast_kernel<\(int\)0|demod_fast_kernel<0," -s 3 -c 1 -f -o gpurun_out/dsrc_dem python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --no-wideband --min-seconds 0 --streams 1 > /dev/null 2>&1
ncu -i gpurun_out/dsrc_dec.ncu-rep --page source --csv > gpurun_out/dsrc_dec_source.csv 2>/dev/null
ncu -i gpurun_out/dsrc_dem.ncu-rep --page source --csv > gpurun_out/dsrc_dem_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/dsrc_dec.ncu-rep gpurun_out/dsrc_dem.ncu-rep > gpurun_out/dsrc_summary.csv
rm -f gpurun_out/dsrc_dec.ncu-rep gpurun_out/dsrc_dem.ncu-rep
ls -la gpurun_out/dsrc_*
