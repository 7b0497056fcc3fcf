#!/bin/sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demod_gpu.py -x -q -m gpu > gpurun_out/c_pytest_demod.log 2>&1; echo "pytest demod rc=$?"
tail -3 gpurun_out/c_pytest_demod.log
V=osmo_gmr_b200/build/variants
sh tools/ab_bench.sh c_ab $V/libpf3.so $V/libpf0.so $V/libwalk0.so $V/libsb8.so $V/libpf3sb8.so
sh tools/ncu_demod.sh c_ncu $V/libpf3.so
