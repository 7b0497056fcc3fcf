#!/bin/sh
# whole GPU suite on the working tree + config-3 timing of the TCH3 decode (traceback with batched decision loads)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/p_pytest.log 2>&1; tail -4 gpurun_out/p_pytest.log
python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-sweep --min-seconds 0 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/p_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 1), "ms/step", round(d["ms_per_step"], 4), "roofline", round(d["roofline"]["frac"], 3), d["fcch"])
for c in ("3", "4"):
    print("  cfg", c, round(d["configs"][c]["bursts_per_s"] / 1e6, 1), d["configs"][c]["ms"])
PY
tail -2 gpurun_out/p_bench.err
