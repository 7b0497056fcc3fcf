#!/bin/sh
# A/B on one box: branch metrics of the one-codeword-per-thread Viterbi through the byte tables (main) vs arithmetic
# (build/variants/libnolut.so = -DTPC_LUT=0)
python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py tests/test_chain_gpu.py -m gpu -x -q 2>&1 | tail -3
sh tools/ab_bench.sh lut osmo_gmr_b200/build/variants/libnolut.so
for v in "" osmo_gmr_b200/build/variants/libnolut.so; do
  lib=""; [ -n "$v" ] && lib=$PWD/$v
  GMR1B200_LIB=$lib python tools/bench_configs.py 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin: print(json.loads(l)['ms'])"
done
