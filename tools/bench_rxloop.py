#!/usr/bin/env python
"""Secondary measurement: the receiver frame loop gmr1b200_rx_bcch_batch over many
channels, device-resident recordings.  Each channel is its own copy in HBM of one of a few synthetic 2.2 s
recordings (FCCH + BCCH + CCCH frames, tests/recording.py), so the windows the loop cuts are real DRAM reads.
Prints one JSON line: channels, frames walked, bursts decoded, ms, bursts/s, launches."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=4096)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--lockstep", action="store_true", help="frame-by-frame schedule instead of the BCCH-paced one")
    args = ap.parse_args()
    import torch
    import osmo_gmr_b200
    import recording             # the recording generator of the tests (numpy); nothing under oracle/ is used here
    L = osmo_gmr_b200.lib()
    L.call("gmr1b200_set_rx_lockstep", 1 if args.lockstep else 0)

    def enc(chan, nbits):        # the library's own channel encoders (gmr1b200_xcch_encode_batch, host code)
        def f(l2):
            out = np.zeros(nbits, np.uint8)
            L.call("gmr1b200_xcch_encode_batch", chan, out, np.ascontiguousarray(l2, np.uint8), 1)
            return out
        return f
    SPS = 4
    base = []
    for seed, (snr, cfo) in enumerate([(15.0, 300.0), (10.0, -200.0), (12.0, 90.0), (20.0, 600.0)]):
        x, _ = recording.make(enc(0, 424), enc(1, 432), esn0_db=snr, cfo_hz=cfo, seed=seed + 1)
        base.append(x)
    rl = len(base[0])
    n = args.channels
    dev = torch.device("cuda", 0)
    iq = torch.empty((n, rl, 2), dtype=torch.float32, device=dev)
    for k, x in enumerate(base):
        iq[k::len(base)] = torch.from_numpy(x.view(np.float32).reshape(rl, 2)).to(dev)
    rec_ofs = (torch.arange(n, dtype=torch.int64, device=dev) * rl)
    rec_len = torch.full((n,), rl, dtype=torch.int32, device=dev)
    # FCCH acquisition of every channel (search window at START_DISCARD = 8000, gmr1_rx.c:56,612)
    W = (330 * 23400 * SPS) // 1000
    rough = torch.empty(n, dtype=torch.int32, device=dev)
    align = torch.empty(n, dtype=torch.int32, device=dev)
    ferr = torch.empty(n, dtype=torch.float32, device=dev)
    L.call("gmr1b200_fcch_acquire_batch", 0, iq, n * rl, rec_ofs + 8000, 0, W, SPS, rough, align, ferr, n, None)
    align0 = (align + 8000).contiguous()
    F = 64
    out = dict(kind=torch.empty((n, F), dtype=torch.int32, device=dev), fn=torch.empty((n, F), dtype=torch.int32, device=dev),
               crc=torch.empty((n, F), dtype=torch.int32, device=dev), conv=torch.empty((n, F), dtype=torch.int32, device=dev),
               l2=torch.empty((n, F, 24), dtype=torch.uint8, device=dev), nfr=torch.empty(n, dtype=torch.int32, device=dev))
    times = []
    l0 = L.c.gmr1b200_kernel_launches()
    for rep in range(args.reps + 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        L.call("gmr1b200_rx_bcch_batch", iq, n * rl, rec_ofs, rec_len, align0, ferr, SPS, n, F,
               out["kind"], out["fn"], out["crc"], out["conv"], out["l2"], out["nfr"], None, None, None)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            times.append(e0.elapsed_time(e1))
    launches = (L.c.gmr1b200_kernel_launches() - l0) // (args.reps + 1)
    kind, crc = out["kind"].cpu().numpy(), out["crc"].cpu().numpy()
    bursts = int((kind > 0).sum())
    ms = float(np.mean(times))
    print(json.dumps({"what": "gmr1b200_rx_bcch_batch: BCCH/CCCH frame loop with tracking feedback, device-resident",
                      "schedule": "lockstep" if args.lockstep else "paced", "channels": n, "frames_per_channel": int(out["nfr"].cpu().numpy().max()), "bursts": bursts,
                      "crc_ok_frac": float((crc[kind > 0] == 0).mean()), "ms": ms, "bursts_per_s": bursts / (ms * 1e-3),
                      "iq_bytes": int(iq.numel() * 4), "kernel_launches": int(launches)}))


if __name__ == "__main__":
    main()
