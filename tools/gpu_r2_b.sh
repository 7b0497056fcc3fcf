#!/bin/sh
# round 2, GPU call B: demod parity, A/B of per-format kernel variants, one full ncu capture of the BCCH launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_demod_gpu.py -x -q -m gpu > gpurun_out/b_pytest_demod.log 2>&1; echo "pytest demod rc=$?"
tail -3 gpurun_out/b_pytest_demod.log
V=osmo_gmr_b200/build/variants
sh tools/ab_bench.sh b_ab $V/libpf0.so $V/libpf2.so $V/libpf2all.so $V/libctas9.so $V/libctas10.so $V/libsb16.so $V/libsb4.so
sh tools/ncu_demod.sh b_ncu $V/libctas10.so
sh tools/ncu_full.sh b_full
ncu -i gpurun_out/b_full_main.ncu-rep --page source --csv > gpurun_out/b_full_main_source.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/b_full_main.ncu-rep > gpurun_out/b_full_summary.csv
