#!/bin/sh
# the 5-shift FCCH grid kernel's row of the ncu summary (round_artifacts.sh's regex predated the second template argument)
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"fcch_fft_kernel<\(bool\)1|fcch_fft_kernel<true" -c 1 -f -o gpurun_out/r2e_full_grid \
	python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/r2e_full_grid.ncu-rep > gpurun_out/r2e_grid_summary.csv
rm -f gpurun_out/r2e_full_grid.ncu-rep
