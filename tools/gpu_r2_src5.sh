#!/bin/sh
# source-level profiles: FCCH search (single shift, 5-shift grid), FCCH fine
mkdir -p gpurun_out
cap() {
	timeout 600 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$2" -s $3 -c 1 -f -o gpurun_out/s5_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 --streams 1 > /dev/null 2>&1
	ncu -i gpurun_out/s5_$1.ncu-rep --page source --csv > gpurun_out/s5_$1_source.csv 2>/dev/null
}
cap fft1 "fcch_fft_kernel<\(bool\)0>|fcch_fft_kernel<false>" 2
cap fft5 "fcch_fft_kernel<\(bool\)1>|fcch_fft_kernel<true>" 0
cap fine "fcch_fine_kernel" 2
python tools/ncu_summary.py gpurun_out/s5_*.ncu-rep > gpurun_out/s5_summary.csv
rm -f gpurun_out/s5_*.ncu-rep
ls -la gpurun_out/s5_*
