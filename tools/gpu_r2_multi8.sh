#!/bin/sh
# 8-GPU check, bounded: lean torchrun bench (e2e + H2D ceiling), single-process device pool, pool test
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
T0=$(date +%s)
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --min-seconds 0 > gpurun_out/m${N}_bench.json 2> gpurun_out/m${N}_bench.err
echo "torchrun N=$N rc=$? seconds=$(( $(date +%s) - T0 ))"
T0=$(date +%s)
timeout 240 python bench.py --gpus $N --single-process --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-sweep --no-wideband --min-seconds 0 > gpurun_out/m${N}_pool.json 2> gpurun_out/m${N}_pool.err
echo "single-process N=$N rc=$? seconds=$(( $(date +%s) - T0 ))"
timeout 200 python -m pytest tests/test_pool_gpu.py -x -q -m gpu 2>&1 | tail -3
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("bench", "pool"):
    try:
        d = json.loads(open(f"gpurun_out/m{n}_{tag}.json").read().strip().splitlines()[-1])
        e = d["e2e"]
        print(tag, "value", round(d["value"] / 1e6, 1), "e2e", round(e["value"] / 1e6, 2), "h2d/gpu", round(e["h2d_gbs_per_gpu"], 1),
              "ceiling/gpu", round(e["h2d_ceiling_gbs_per_gpu"], 1), "frac", round(e["frac_of_ceiling"], 3), "pool", e["single_process_pool"])
    except Exception as ex:
        print(tag, "FAILED", ex)
PY
tail -3 gpurun_out/m${N}_bench.err gpurun_out/m${N}_pool.err
