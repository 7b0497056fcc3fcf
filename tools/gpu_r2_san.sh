#!/bin/sh
# compute-sanitizer over the kernels new in this round: channeliser (both bank kernels, resampler, pieces), FCCH in the
# frequency domain, vocoder stream
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_chan_gpu.py tests/test_sdr_gpu.py -q -m gpu -x -k "not 2048 and not 1024 and not pieces" > gpurun_out/san_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/san_$tool.log | tail -4
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_rxcall_gpu.py tests/test_pool_gpu.py -q -m gpu -x > gpurun_out/san_memcheck2.log 2>&1
echo "memcheck rxcall/pool rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_memcheck2.log | tail -3
