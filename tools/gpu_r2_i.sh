#!/bin/sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rxcall_gpu.py tests/test_rxsched_gpu.py -x -q -m gpu 2>&1 | tail -5
for n in 1024 4096 16384; do python tools/bench_rxloop.py --channels $n; done > gpurun_out/i_rxloop_graph.jsonl 2> gpurun_out/i_rxloop.err
for n in 1024 4096; do GMR1B200_RX_NOGRAPH=1 python tools/bench_rxloop.py --channels $n; done > gpurun_out/i_rxloop_nograph.jsonl 2>> gpurun_out/i_rxloop.err
cat gpurun_out/i_rxloop_graph.jsonl gpurun_out/i_rxloop_nograph.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['channels'], round(d['ms'],3), 'ms', round(d['bursts_per_s']/1e6,2), 'Mb/s', d['kernel_launches'], d['crc_ok_frac'])
"
tail -3 gpurun_out/i_rxloop.err
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
