#!/bin/sh
# bitsliced A5: tests, config 3 with both kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_rxcall_gpu.py tests/test_fullsize_gpu.py -q -m gpu -k "a5 or rxcall or calls or config3 or tch9" 2>&1 | tail -4
python - <<'PY'
import ctypes, json, subprocess, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import osmo_gmr_b200
L = osmo_gmr_b200.lib(); L.init(0)
n, nbits = 393216, 208
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(0, 256, (n, 8), dtype=torch.uint8, device="cuda", generator=g)
fn = torch.randint(0, 1 << 19, (n,), dtype=torch.int32, device="cuda", generator=g)
dl = torch.zeros((n, nbits), dtype=torch.uint8, device="cuda")
for mode, name in ((0, "one unit per thread"), (1, "bitsliced")):
    L.c.gmr1b200_set_a5_bitslice(mode)
    for _ in range(3):
        L.call("gmr1b200_a5_batch", None, 1, keys, fn, nbits, nbits, dl, None, n, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        L.call("gmr1b200_a5_batch", None, 1, keys, fn, nbits, nbits, dl, None, n, None)
    e1.record(); torch.cuda.synchronize()
    print(name, round(e0.elapsed_time(e1) / 10, 4), "ms per", n, "streams of", nbits, "bits")
L.c.gmr1b200_set_a5_bitslice(-1)
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep --no-wideband --min-seconds 0 > gpurun_out/a5_bench.json 2> gpurun_out/a5_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/a5_bench.json").read().strip().splitlines()[-1])
c3 = d["configs"]["3"]
print("config3", round(c3["bursts_per_s"] / 1e6, 1), c3["ms"], c3["parity_vs_cpu_reference"]["identical"])
PY
