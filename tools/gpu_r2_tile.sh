#!/bin/sh
# A/B: 64-codeword tiles for FACCH9 / TCH9 (main) vs 128 (build/variants/libt32r64.so)
python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py tests/test_chain_gpu.py tests/test_rxcall_gpu.py -m gpu -x -q 2>&1 | tail -2
for pass in 1 2; do
for v in "" osmo_gmr_b200/build/variants/libt32r64.so; do
  lib=""; [ -n "$v" ] && lib=$PWD/$v
  echo "lib=$v"; GMR1B200_LIB=$lib python tools/bench_configs.py 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin: print(json.loads(l)['ms'])"
done; done
