#!/bin/sh
# the driver's N = 2 launch with DEFAULT flags (all legs), timed
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
echo "torchrun N=2 default rc=$? seconds=$(( $(date +%s) - T0 ))"
tail -3 gpurun_out/n2_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/n2_bench.json").read().strip().splitlines()[-1])
w = d["wideband"]
print("N=2 value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2), "frac_of_ceiling", round(d["e2e"]["frac_of_ceiling"], 3),
      "| wideband e2e", round(w["e2e"]["value"] / 1e6, 2), "resident", round(w["device_resident"]["bursts_per_s"] / 1e6, 1), w["crc_ok_frac"], w["n_gpus"])
print("sweep", [(p["arfcns"], round(p["resident_bursts_per_s"]/1e6,1), round(p["streamed_bursts_per_s"]/1e6,2)) for p in d["sweep"]["points"]])
PY
T0=$(date +%s)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/n2_ref.json 2> gpurun_out/n2_ref.err
echo "reference arm N=2 rc=$? seconds=$(( $(date +%s) - T0 ))"; tail -c 300 gpurun_out/n2_ref.json
