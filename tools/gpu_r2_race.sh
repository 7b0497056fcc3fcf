#!/bin/sh
# racecheck (shared-memory hazards) over the kernels changed last: two-block FCCH search, cipher pass / tiles of the
# decode kernel, RACH rows of the per-format demod kernel
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_sdr_gpu.py -m gpu -x -q -k "rough_grid and fft" 2>&1 | tail -4
timeout 200 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_decode_gpu.py -m gpu -x -q -k "unaligned or facch9 or tch3" 2>&1 | tail -4
timeout 120 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_demod_gpu.py -m gpu -x -q -k "rach" 2>&1 | tail -4
