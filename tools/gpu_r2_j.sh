#!/bin/sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rxsched_gpu.py tests/test_rxcall_gpu.py tests/test_dropin_gpu.py -x -q -m gpu 2>&1 | tail -8
for n in 1024 4096 16384; do python tools/bench_rxloop.py --channels $n; done > gpurun_out/j_rxloop.jsonl 2> gpurun_out/j_rxloop.err
python tools/bench_rxloop.py --channels 1024 --lockstep >> gpurun_out/j_rxloop.jsonl 2>> gpurun_out/j_rxloop.err
cat gpurun_out/j_rxloop.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['schedule'], d['channels'], round(d['ms'],3), 'ms', round(d['bursts_per_s']/1e6,2), 'Mb/s', d['kernel_launches'], d['crc_ok_frac'], d['bursts'])
"
tail -3 gpurun_out/j_rxloop.err
