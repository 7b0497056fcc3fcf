#!/bin/sh
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_chan_gpu.py -q -m gpu 2>&1 | tail -3
python tools/bench_chan.py
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"pfb_|resamp_kernel" -s 4 -c 2 --csv --log-file gpurun_out/z_chan.csv python tools/bench_chan.py --reps 2 > /dev/null 2>&1
grep "gpu__time\|inst_executed\|issue_active" gpurun_out/z_chan.csv | awk -F'","' '{print substr($5,1,36), $13, $15}'
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"fcch_fft_kernel<\(bool\)1>|fcch_fft_kernel<true>" -c 1 -o gpurun_out/y_fcch5 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/y_fcch5.log 2>&1
ls -la gpurun_out/y_fcch5.ncu-rep && python tools/ncu_summary.py gpurun_out/y_fcch5.ncu-rep > gpurun_out/z_summary.csv
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/z_summary.csv")))
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
