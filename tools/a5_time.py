#!/usr/bin/env python3
"""tools/a5_time.py [--n 393216] [--mode 1] - gmr1b200_a5_batch alone on device-resident keys (run under ncu for counters)"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import osmo_gmr_b200
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=393216)
ap.add_argument("--mode", type=int, default=1)
ap.add_argument("--nbits", type=int, default=208)
a = ap.parse_args()
L = osmo_gmr_b200.lib(); L.init(0)
g = torch.Generator(device="cuda").manual_seed(1)
keys = torch.randint(0, 256, (a.n, 8), dtype=torch.uint8, device="cuda", generator=g)
fn = torch.randint(0, 1 << 19, (a.n,), dtype=torch.int32, device="cuda", generator=g)
dl = torch.zeros((a.n, a.nbits), dtype=torch.uint8, device="cuda")
L.c.gmr1b200_set_a5_bitslice(a.mode)
for _ in range(3):
    L.call("gmr1b200_a5_batch", None, 1, keys, fn, a.nbits, a.nbits, dl, None, a.n, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    L.call("gmr1b200_a5_batch", None, 1, keys, fn, a.nbits, a.nbits, dl, None, a.n, None)
e1.record(); torch.cuda.synchronize()
print("mode", a.mode, round(e0.elapsed_time(e1) / 10, 4), "ms per", a.n, "streams of", a.nbits, "bits")
