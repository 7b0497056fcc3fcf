#!/bin/sh
# compute-sanitizer memcheck over the GPU suite (full-size and 1024+-channel cases left out: minutes each under the tool)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -q -m gpu --deselect tests/test_fullsize_gpu.py -k "not 2048 and not 1024 and not pieces" > gpurun_out/san_all.log 2>&1
echo "memcheck all rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/san_all.log | tail -3
grep -A1 "Invalid\|Program hit" gpurun_out/san_all.log | grep " at \|hit" | sort | uniq -c | head
