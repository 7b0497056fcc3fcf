#!/bin/sh
# shared-capture wideband leg (one recording, NCCL all-gather over NVLink) on N ranks, lean bench
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/sh${N}_bench.json 2> gpurun_out/sh${N}_bench.err
echo "torchrun N=$N rc=$? seconds=$(( $(date +%s) - T0 ))"
grep -v "^W1\|OMP_NUM\|^\*\*\*" gpurun_out/sh${N}_bench.err | tail -12
python - $N <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/sh{sys.argv[1]}_bench.json").read().strip().splitlines()[-1])
w = d["wideband"]; s = w["shared_capture"]
print("N", d["n_gpus"], "value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2))
print("wideband weak   e2e", round(w["e2e"]["value"] / 1e6, 2), "ms", round(w["e2e"]["ms_per_step"], 2), "resident", round(w["device_resident"]["bursts_per_s"] / 1e6, 1))
print("wideband shared e2e", round(s["e2e"]["value"] / 1e6, 2), "ms", round(s["e2e"]["ms_per_step"], 2), "resident", round(s["device_resident"]["bursts_per_s"] / 1e6, 1),
      s["device_resident"]["ms"], "crc", s["crc_ok_frac"], "wrong", s["crc_ok_but_payload_wrong"], "fcch", s["fcch_found_frac"])
PY
