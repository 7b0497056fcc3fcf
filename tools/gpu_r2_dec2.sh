#!/bin/sh
# decode tests (incl. ciphered channels and the frame loops that use them), memcheck on the ciphered ones, bench
python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py tests/test_chain_gpu.py tests/test_rxcall_gpu.py tests/test_rxsched_gpu.py -m gpu -x -q 2>&1 | tail -2
compute-sanitizer --tool memcheck python -m pytest tests/test_decode_gpu.py -m gpu -x -q -k "tch3 or facch3 or facch9" 2>&1 | tail -3
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-sweep --no-wideband --min-seconds 0 > gpurun_out/dec2_bench.json 2>gpurun_out/dec2_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/dec2_bench.json'))
print(d['value'], d['viterbi']['ms_per_launch'], d['configs']['3']['ms'], d['configs']['4']['ms'])
P
