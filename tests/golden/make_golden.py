#!/usr/bin/env python
"""Generate tests/golden/*.npz from the reference's own C sources (oracle/_ref = /root/reference
compiled against oracle/shim).  Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

The fixtures are small seeded input/output pairs for every stage boundary of the hot path; they pin
the oracle port (and, on the GPU box, the CUDA path) to what the reference computed here.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib  # noqa: E402
import sigen  # noqa: E402
import vectors  # noqa: E402


def main():
    ref = os.path.join(oracle_lib.ROOT, "oracle", "_ref", "libgmr1_ref.so")
    assert os.path.exists(ref), "build oracle/_ref first (make -C oracle ref)"
    o = oracle_lib.Oracle(ref, "reference")
    rng = np.random.default_rng(20261017)
    g = {}

    # ---- stage 3: soft bits -> L2
    for name, nbits in (("bcch", 424), ("ccch", 432), ("xch_dc12", 432)):
        e = vectors.simple(o, rng, name, nbits, 24)
        out = [o.simple_decode(name, e[i]) for i in range(len(e))]
        g[f"{name}_e"] = e
        g[f"{name}_l2"] = np.stack([x[0] for x in out])
        g[f"{name}_crc"] = np.array([x[1] for x in out], np.int32)
        g[f"{name}_conv"] = np.array([x[2] for x in out], np.int32)
    e, ciph = vectors.facch3(o, rng, 16, True)
    out = [o.facch3_decode(e[i], ciph[i]) for i in range(16)]
    g.update(facch3_e=e, facch3_ciph=ciph, facch3_l2=np.stack([x[0] for x in out]),
             facch3_s=np.stack([x[1] for x in out]), facch3_crc=np.array([x[2] for x in out], np.int32),
             facch3_conv=np.array([x[3] for x in out], np.int32))
    e, ciph = vectors.facch9(o, rng, 12, True)
    out = [o.facch9_decode(e[i], ciph[i]) for i in range(12)]
    g.update(facch9_e=e, facch9_ciph=ciph, facch9_l2=np.stack([x[0] for x in out]),
             facch9_sacch=np.stack([x[1] for x in out]), facch9_status=np.stack([x[2] for x in out]),
             facch9_crc=np.array([x[3] for x in out], np.int32), facch9_conv=np.array([x[4] for x in out], np.int32))
    for mode in (0, 1, 2):
        e, ciph, p1, p2 = vectors.tch9(o, rng, mode, 2, 5, True)
        out = []
        for c in range(2):
            il = o.interleaver()
            for b in range(5):
                out.append(o.tch9_decode(e[c * 5 + b], mode, ciph[c * 5 + b], il))
        g.update({f"tch9_{mode}_e": e, f"tch9_{mode}_ciph": ciph, f"tch9_{mode}_prev1": p1, f"tch9_{mode}_prev2": p2,
                  f"tch9_{mode}_l2": np.stack([x[0] for x in out]),
                  f"tch9_{mode}_conv": np.array([x[3] for x in out], np.int32)})
    e, masks = vectors.rach(o, rng, 16)
    out = [o.rach_decode(e[i], masks[i]) for i in range(16)]
    g.update(rach_e=e, rach_mask=masks, rach_l2=np.stack([x[0] for x in out]),
             rach_crc=np.array([x[1] for x in out], np.int32), rach_conv=np.array([x[2] for x in out], np.int32),
             rach_crc2=np.array([x[3] for x in out], np.int32))
    e, ciph = vectors.tch3(rng, 16, True)
    out = [o.tch3_decode(e[i], ciph[i], i % 2) for i in range(16)]
    g.update(tch3_e=e, tch3_ciph=ciph, tch3_f0=np.stack([x[0] for x in out]), tch3_f1=np.stack([x[1] for x in out]),
             tch3_s=np.stack([x[2] for x in out]), tch3_c0=np.array([x[3] for x in out], np.int32),
             tch3_c1=np.array([x[4] for x in out], np.int32))
    g["a5_key"] = rng.integers(0, 256, 8, dtype=np.uint8)
    g["a5_dl"] = o.a5(1, g["a5_key"], 0x2a5c7, 208)

    # ---- stage 2: IQ -> soft bits (kept small: 3 bursts per format)
    for name, win in (("bcch", 80), ("dc6", 40), ("nt3_speech", 6), ("nt3_facch", 6), ("nt9", 6), ("rach", 6)):
        hard = rng.integers(0, 2, (3, sigen.burst_ebits(name)), dtype=np.uint8)
        x = sigen.modulate(name, hard, 4, win, rng.uniform(2, win - 2, 3), rng.uniform(-0.013, 0.013, 3),
                           rng.uniform(0, 6, 3), np.array([8.0, 15.0, 30.0]), rng)
        out = [o.demod(name, x[i], 4, 0.0) for i in range(3)]
        g.update({f"demod_{name}_iq": x, f"demod_{name}_ebits": np.stack([r[1] for r in out]),
                  f"demod_{name}_sync": np.array([r[2] for r in out], np.int32),
                  f"demod_{name}_toa": np.array([r[3] for r in out], np.float32),
                  f"demod_{name}_ferr": np.array([r[4] for r in out], np.float32)})

    # ---- stage 1: FCCH fine / snr on 468-sample bursts (rough windows are 247 KB each: one only)
    xs = np.stack([sigen.fcch_window(468 + 64, 4, 32 + k, 0.05 * (k - 2), 12.0, rng)[32:32 + 468] for k in range(5)])
    out = [o.fcch_fine(xs[i], 4, 0.0) for i in range(5)]
    g.update(fcch_fine_iq=xs, fcch_fine_toa=np.array([r[1] for r in out], np.int32),
             fcch_fine_ferr=np.array([r[2] for r in out], np.float32),
             fcch_snr=np.array([o.fcch_snr(xs[i], 4, 0.0)[1] for i in range(5)], np.float32))
    xr = sigen.fcch_window(30888, 4, 17321, 0.04, 8.0, rng)
    g.update(fcch_rough_iq=xr.astype(np.complex64), fcch_rough_toa=np.array([o.fcch_rough(xr, 4, 0.0)[1]], np.int32))

    np.savez_compressed(os.path.join(HERE, "golden.npz"), **g)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
