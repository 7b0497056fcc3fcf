#!/bin/sh
# ncu --set full of the FCCH search kernels (single shift and 5-shift grid) and the resampler; summaries to gpurun_out
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"fcch_fft_kernel" -s 2 -c 1 -o gpurun_out/y_fcch1 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --no-configs --min-seconds 0 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fcch_fft_kernel<1>|fcch_fft_kernelILb1" -c 1 -o gpurun_out/y_fcch5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/y_fcch5.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"resamp_kernel" -s 2 -c 1 -o gpurun_out/y_resamp -f python tools/bench_chan.py --reps 2 > /dev/null 2>&1
ls -la gpurun_out/y_*.ncu-rep
python tools/ncu_summary.py gpurun_out/y_fcch1.ncu-rep gpurun_out/y_fcch5.ncu-rep gpurun_out/y_resamp.ncu-rep > gpurun_out/y_summary.csv
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/y_summary.csv")))
h = rows[0]
for r in rows[2:]:
    print("----", r[1][:50])
    for k, v in zip(h[2:], r[2:]):
        try:
            fv = float(v)
        except ValueError:
            continue
        if k.startswith("stall_") and fv < 0.3:
            continue
        print("  ", k, round(fv, 3))
PY
