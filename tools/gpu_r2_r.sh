#!/bin/sh
# wideband leg of bench.py
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -5 gpurun_out/r_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r_bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2))
print(json.dumps(d["wideband"], indent=1))
PY
