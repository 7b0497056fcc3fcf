#!/bin/sh
# the driver's N-rank launch with DEFAULT flags (all legs), timed; then the reference arm the same way
N=${1:-8}
mkdir -p gpurun_out
T0=$(date +%s)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err
echo "torchrun N=$N default rc=$? seconds=$(( $(date +%s) - T0 ))"
grep -v "^W1\|OMP_NUM\|^\*\*\*" gpurun_out/n${N}_bench.err | tail -5
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/n{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
w = d["wideband"]; s = w.get("shared_capture")
print("N", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2), "frac_of_ceiling", round(d["e2e"]["frac_of_ceiling"], 3), "clocks", d["clocks"])
print("wideband weak e2e", round(w["e2e"]["value"] / 1e6, 2), "shared e2e", round(s["e2e"]["value"] / 1e6, 2) if s else None)
print("sweep", [(p["arfcns"], round(p["resident_bursts_per_s"]/1e6,1), round(p["streamed_bursts_per_s"]/1e6,2)) for p in d["sweep"]["points"]])
PY
T0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 3 --warmup 3 > gpurun_out/n${N}_ref.json 2> gpurun_out/n${N}_ref.err
echo "reference arm N=$N rc=$? seconds=$(( $(date +%s) - T0 ))"; tail -c 200 gpurun_out/n${N}_ref.json
