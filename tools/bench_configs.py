#!/usr/bin/env python3
"""tools/bench_configs.py - device-resident throughput of BASELINE.json configs 3 and 4 on one GPU (secondary
numbers: bench.py measures the headline config 2; parity of these workloads at size is tests/test_fullsize_gpu.py).

  config 3: NT3 traffic across 4096 ARFCNs x 128 frames: 75 % speech bursts (pi/4-CQPSK demod, A5/1 masks for half
            of them made on the device, TCH3 K7 tail-biting decode), 25 % FACCH3 in groups of 4 (pi/4-CBPSK demod,
            K5 r1/4 decode + CRC16)
  config 4: 8192 ARFCNs x 64 bursts: 60 % NT9 (half FACCH9, half TCH9-9k6 with the depth-3 inter-burst interleaver),
            40 % RACH, plus per ARFCN one 330 ms FCCH search repeated over a grid of 5 frequency shifts and one fine
            estimate

Payload bits are random (the kernels' work does not depend on them); IQ is made by the GPU synthesiser.  One JSON
line per config: bursts/s over the whole chain, per-kernel times (CUDA events, serial), demod GB/s of algorithmic
bytes against MEASURED_PEAKS.json."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import osmo_gmr_b200  # noqa: E402

SPS = 4
BT = {"nt3_speech": 4, "nt3_facch": 5, "nt9": 7, "rach": 8}
DEV = torch.device("cuda", 0)


def peak_gbs():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6552.0


def synth(L, name, n, win, seed, chunk=65536):
    eb = L.c.gmr1b200_burst_ebits(BT[name])
    wl = L.c.gmr1b200_burst_len(BT[name]) * SPS + win
    iq = torch.empty((n, wl, 2), dtype=torch.float32, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(seed)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        hard = torch.randint(0, 2, (m, eb), dtype=torch.uint8, device=DEV, generator=g)
        toa = torch.rand(m, device=DEV, generator=g) * (win - 3) + 1.5
        cfo = (torch.rand(m, device=DEV, generator=g) - 0.5) * 0.01
        ph = torch.rand(m, device=DEV, generator=g) * 6.28
        L.call("gmr1b200_synth_bursts", BT[name], hard, eb, None, SPS, wl, toa, 0.0, cfo, 0.0, ph, 0.0, None, 15.0,
               None, 1.0, seed + s, iq[s:s + m], m * wl, None, wl, m, None)
    torch.cuda.synchronize()
    return iq, wl, eb


class Timer:
    def __init__(self):
        self.t = {}

    def run(self, name, fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        self.t[name] = a.elapsed_time(b) / reps


def config3(L, arfcns, frames, reps):
    n = arfcns * frames
    n_sp, n_fa = (n * 3 // 4), (n // 4) // 4 * 4
    iq_sp, wl, eb_sp = synth(L, "nt3_speech", n_sp, 6, 31)
    iq_fa, _, eb_fa = synth(L, "nt3_facch", n_fa, 6, 32)
    e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=DEV)
    ebits_sp, ebits_fa = e(n_sp, eb_sp, dt=torch.int8), e(n_fa, eb_fa, dt=torch.int8)
    keys = torch.randint(0, 256, (n_sp, 8), dtype=torch.uint8, device=DEV)
    fn = torch.randint(0, 1 << 19, (n_sp,), dtype=torch.int32, device=DEV)
    alg = (torch.arange(n_sp, device=DEV) % 2).to(torch.int32)
    ciph = e(n_sp, 208)
    f0, f1, bs = e(n_sp, 10), e(n_sp, 10), e(n_sp, 4)
    l2, bs3, crc = e(n_fa // 4, 10), e(n_fa // 4, 32), e(n_fa // 4, dt=torch.int32)
    steps = {
        "a5_masks": lambda: L.call("gmr1b200_a5_batch", alg, 0, keys, fn, 208, 208, ciph, None, n_sp, None),
        "demod_nt3_speech": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt3_speech"], iq_sp, n_sp * wl, None, wl, wl,
                                           SPS, None, 0.0, ebits_sp, eb_sp, None, None, None, None, n_sp, None),
        "decode_tch3": lambda: L.call("gmr1b200_tch3_decode_batch", f0, f1, bs, ebits_sp, ciph, 0, None, None, n_sp, None),
        "demod_nt3_facch": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt3_facch"], iq_fa, n_fa * wl, None, wl, wl,
                                          SPS, None, 0.0, ebits_fa, eb_fa, None, None, None, None, n_fa, None),
        "decode_facch3": lambda: L.call("gmr1b200_facch3_decode_batch", l2, bs3, ebits_fa, None, None, crc, n_fa // 4, None),
    }
    tm = Timer()
    for k, f in steps.items():
        tm.run(k, f, reps)
    tm.run("whole_chain", lambda: [f() for f in steps.values()], reps)
    byt = n_sp * (8 * wl + eb_sp + 16) + n_fa * (8 * wl + eb_fa + 16)
    dem_ms = tm.t["demod_nt3_speech"] + tm.t["demod_nt3_facch"]
    return {"config": "3: NT3 TCH3 speech (half A5/1) + FACCH3, %d ARFCNs x %d frames" % (arfcns, frames), "bursts": n_sp + n_fa,
            "iq_bytes": (n_sp + n_fa) * wl * 8, "ms": {k: round(v, 4) for k, v in tm.t.items()},
            "bursts_per_s": (n_sp + n_fa) / tm.t["whole_chain"] * 1e3,
            "demod_gbs": byt / dem_ms / 1e6, "demod_frac_of_hbm_peak": byt / dem_ms / 1e6 / peak_gbs(),
            "tch3_acs_per_s": n_sp * 12288 / tm.t["decode_tch3"] * 1e3}


def config4(L, arfcns, per, reps):
    n = arfcns * per
    n9 = n * 6 // 10 // 6 * 6
    n_f9 = n9 // 2
    n_t9 = n9 - n_f9
    n_ra = n - n9
    iq9, wl9, eb9 = synth(L, "nt9", n9, 6, 41)
    iqr, wlr, ebr = synth(L, "rach", n_ra, 6, 42)
    e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=DEV)
    ebits9, ebitsr = e(n9, eb9, dt=torch.int8), e(n_ra, ebr, dt=torch.int8)
    # TCH9: runs of 3 consecutive bursts per channel: prev1 / prev2 index the one / two bursts before
    idx = torch.arange(n_t9, device=DEV, dtype=torch.int32)
    prev1 = torch.where(idx % 3 >= 1, idx - 1, torch.full_like(idx, -1))
    prev2 = torch.where(idx % 3 >= 2, idx - 2, torch.full_like(idx, -1))
    l2f, crcf = e(n_f9, 38), e(n_f9, dt=torch.int32)
    l2t = e(n_t9, 60)
    rach, crcr = e(n_ra, 18), e(n_ra, dt=torch.int32)
    W = (330 * 23400 * SPS) // 1000
    fw = torch.randn((arfcns, W, 2), dtype=torch.float32, device=DEV)
    toa, peak = e(5, arfcns, dt=torch.int32), e(5, arfcns, dt=torch.float32)
    ftoa, ferr = e(arfcns, dt=torch.int32), e(arfcns, dt=torch.float32)
    grid = [-0.54, -0.27, 0.0, 0.27, 0.54]           # +-2, +-1, 0 kHz as rad/symbol at 23.4 ksym/s

    def fcch():
        for k, fs in enumerate(grid):
            L.call("gmr1b200_fcch_rough_batch", 0, fw, arfcns * W, None, W, W, SPS, None, fs, toa[k], peak[k], arfcns, None)
        L.call("gmr1b200_fcch_fine_batch", 0, fw, arfcns * W, None, W, SPS, None, 0.0, ftoa, ferr, arfcns, None)

    steps = {
        "demod_nt9": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt9"], iq9, n9 * wl9, None, wl9, wl9, SPS, None, 0.0,
                                    ebits9, eb9, None, None, None, None, n9, None),
        "decode_facch9": lambda: L.call("gmr1b200_facch9_decode_batch", l2f, None, None, ebits9[:n_f9], None, None, crcf, n_f9, None),
        "decode_tch9_9k6": lambda: L.call("gmr1b200_tch9_decode_batch", l2t, None, None, ebits9[n_f9:], 2, None, prev1, prev2,
                                          None, n_t9, None),
        "demod_rach": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["rach"], iqr, n_ra * wlr, None, wlr, wlr, SPS, None, 0.0,
                                     ebitsr, ebr, None, None, None, None, n_ra, None),
        "decode_rach": lambda: L.call("gmr1b200_rach_decode_batch", rach, ebitsr, None, 0, None, None, crcr, n_ra, None),
        "fcch_5_shift_search_and_fine": fcch,
    }
    tm = Timer()
    for k, f in steps.items():
        tm.run(k, f, reps)
    tm.run("whole_chain", lambda: [f() for f in steps.values()], reps)
    byt = n9 * (8 * wl9 + eb9 + 16) + n_ra * (8 * wlr + ebr + 16)
    dem_ms = tm.t["demod_nt9"] + tm.t["demod_rach"]
    return {"config": "4: NT9 FACCH9 + TCH9-9k6, RACH, 5-shift FCCH search, %d ARFCNs x %d bursts" % (arfcns, per), "bursts": n,
            "iq_bytes": n9 * wl9 * 8 + n_ra * wlr * 8 + arfcns * W * 8, "ms": {k: round(v, 4) for k, v in tm.t.items()},
            "bursts_per_s": n / tm.t["whole_chain"] * 1e3, "fcch_searches_per_s": arfcns * 5 / tm.t["fcch_5_shift_search_and_fine"] * 1e3,
            "demod_gbs": byt / dem_ms / 1e6, "demod_frac_of_hbm_peak": byt / dem_ms / 1e6 / peak_gbs()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the ARFCN counts (smoke runs)")
    args = ap.parse_args()
    L = osmo_gmr_b200.lib()
    L.init(0)
    print(json.dumps(config3(L, int(4096 * args.scale), 128, args.reps)), flush=True)
    torch.cuda.empty_cache()
    print(json.dumps(config4(L, int(8192 * args.scale), 64, args.reps)), flush=True)


if __name__ == "__main__":
    main()
