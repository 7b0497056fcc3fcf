#!/bin/sh
mkdir -p gpurun_out
T0=$(date +%s); python bench.py --steps 20 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$? seconds=$(( $(date +%s) - T0 ))"; tail -5 gpurun_out/g_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/g_bench.json').read().strip().splitlines()[-1])
    print("value",d["value"]/1e6,"sust",d["sustained"],"e2e",{k:v for k,v in d["e2e"].items() if k!="how" and k!="ceiling_how"})
    print("roof",d["roofline"]["frac"],d["roofline"]["per_format"],"cpu",d["cpu_baseline"])
    for k,v in (d["configs"] or {}).items(): print("cfg",k,v["bursts_per_s"]/1e6,v["roofline"]["frac"],v["roofline"]["per_format"],v["ms"],v.get("parity_vs_cpu_reference"),v.get("cpu_baseline"))
    print("sweep",d["sweep"]["chunk_decodes_correctly"]); [print(p) for p in d["sweep"]["points"]]
except Exception as e: print("PARSE FAILED",e)
PY
