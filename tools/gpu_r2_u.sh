#!/bin/sh
# full GPU test suite + default bench + 2-rank lean bench with the wideband leg (needs --gpus 2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/u_pytest.log 2>&1; tail -8 gpurun_out/u_pytest.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/u_bench_n2.json 2> gpurun_out/u_bench_n2.err
tail -3 gpurun_out/u_bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/u_bench_n2.json").read().strip().splitlines()[-1])
w = d["wideband"]
print("N=2 value", round(d["value"] / 1e6, 1), "e2e", round(d["e2e"]["value"] / 1e6, 2), "wideband e2e", round(w["e2e"]["value"] / 1e6, 2),
      "resident", round(w["device_resident"]["bursts_per_s"] / 1e6, 1), w["crc_ok_frac"])
PY
fi
