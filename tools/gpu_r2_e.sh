#!/bin/sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demod_gpu.py -x -q -m gpu > gpurun_out/e_pytest_demod.log 2>&1; echo "pytest demod rc=$?"
tail -3 gpurun_out/e_pytest_demod.log
V=osmo_gmr_b200/build/variants
sh tools/ab_bench.sh e_ab $V/libpipe0.so $V/libpf4.so $V/libc9.so $V/libpf4c9.so
sh tools/ncu_demod.sh e_ncu $V/libpf4.so
