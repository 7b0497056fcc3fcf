#!/bin/sh
# A/B: BCCH / CCCH decode tiles of 128 (main), 64, 32 codewords
python -m pytest tests/test_decode_gpu.py -m gpu -x -q 2>&1 | tail -1
sh tools/ab_bench.sh tile osmo_gmr_b200/build/variants/libx64.so osmo_gmr_b200/build/variants/libx32.so
