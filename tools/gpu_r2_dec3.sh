#!/bin/sh
python -m pytest tests/test_decode_gpu.py tests/test_fullsize_gpu.py tests/test_chain_gpu.py tests/test_rxcall_gpu.py -m gpu -x -q 2>&1 | tail -2
compute-sanitizer --tool memcheck python -m pytest tests/test_decode_gpu.py -m gpu -x -q -k "tch3 or facch3 or facch9" 2>&1 | tail -3
for v in "" osmo_gmr_b200/build/variants/libt9e16.so; do
  lib=""; [ -n "$v" ] && lib=$PWD/$v
  echo "lib=$v"; GMR1B200_LIB=$lib python tools/bench_configs.py 2>&1 | tail -2 | python -c "
import sys,json
for l in sys.stdin: print(json.loads(l)['ms'])"
done
