#!/usr/bin/env python3
"""tools/bench_chan.py [--chans 1024] [--seconds 1.0] - the channeliser alone on a random int16 recording
(device-resident): CUDA-event time per call; run under ncu for the per-kernel split."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import osmo_gmr_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--chans", type=int, default=1024)
ap.add_argument("--seconds", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--fmt", type=int, default=1)
a = ap.parse_args()
L = osmo_gmr_b200.lib()
L.init(0)
h = ctypes.c_void_p()
assert L.c.gmr1b200_chan_create(a.chans, 4, ctypes.byref(h)) == 0
n_wide = int(a.seconds * a.chans * 31250)
g = torch.Generator(device="cuda").manual_seed(1)
if a.fmt == 1:
    wide = torch.randint(-8000, 8000, (n_wide, 2), dtype=torch.int16, device="cuda", generator=g)
else:
    wide = torch.randn((n_wide, 2), dtype=torch.float32, device="cuda", generator=g)
n_out = int(L.c.gmr1b200_chan_out_len(h, n_wide))
out = torch.empty((a.chans, n_out, 2), dtype=torch.float32, device="cuda")
for _ in range(2):
    L.call("gmr1b200_channelize", h.value, wide, a.fmt, n_wide, None, a.chans, out, n_out, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    L.call("gmr1b200_channelize", h.value, wide, a.fmt, n_wide, None, a.chans, out, n_out, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
steps = n_wide // (a.chans // 2)
print(json.dumps({"chans": a.chans, "seconds_of_signal": a.seconds, "fmt": a.fmt, "ms": ms, "realtime_factor": a.seconds / (ms * 1e-3),
                  "input_msps": n_wide / ms / 1e3, "n_out": n_out,
                  "algorithmic_gb": (n_wide * (4 if a.fmt else 8) + 2 * steps * a.chans * 8 + a.chans * n_out * 8) / 1e9}))
