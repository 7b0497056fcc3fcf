#!/bin/sh
# channeliser: GPU tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_chan_gpu.py -x -q -m gpu 2>&1 | tail -40
