"""Synthetic single-ARFCN recording for BASELINE.json config 1 (TEST INFRASTRUCTURE).

Frame f (24 slots x 39 symbols = 3744 samples at sps 4): f % 8 == 0 carries the FCCH dual chirp,
f % 8 == 2 a BCCH burst, every other frame a DC6/CCCH burst, all starting at the frame boundary -
the layout gmr1_rx's process_bcch() walks with fn = 0 on the FCCH frame and sa_bcch_stn = 0
(src/gmr1_rx.c:853-895).  One global carrier offset, fractional timing offset and AWGN.
"""
import numpy as np

import sigen

SPS = 4
FRAME = 24 * 39 * SPS


def make(enc_bcch, enc_ccch, seconds=2.2, esn0_db=15.0, cfo_hz=300.0, frac=0.37, start=9000, seed=1, tdma_si1=False,
         imm_ass=None):
    """enc_bcch / enc_ccch: l2[24] -> hard bits.  imm_ass: {frame: (tn, p)} CCCH frames that carry an IMMEDIATE
    ASSIGNMENT (ccch_is_imm_ass / ccch_imm_ass_parse, gmr1_rx.c:236-245).
    Returns (complex64 samples, list of (frame, kind, l2))."""
    rng = np.random.default_rng(seed)
    n = int(seconds * 23400 * SPS)
    x = np.zeros(n, np.complex128)
    truth = []
    f = 0
    pos = start
    while pos + 6 * 39 * SPS + 64 < n:
        if f % 8 == 0:
            c = sigen.fcch_chirp(SPS)
            # fractional delay of the chirp: evaluate it on the shifted grid
            t = (np.arange(len(c)) - frac) / SPS - 117 / 2.0
            x[pos:pos + len(c)] += np.sqrt(2.0) * np.cos(0.32 * 2 * np.pi / 117 * t * t)
            truth.append((f, "fcch", None))
        else:
            kind = "bcch" if f % 8 == 2 else "dc6"
            l2 = rng.integers(0, 256, 24, dtype=np.uint8)
            if kind == "bcch" and tdma_si1:
                # SI1 with a segment 2Abis (bcch_tdma_align, gmr1_rx.c:194-236): SA_BCCH_STN 0 (the burst sits on
                # slot 0), random SA_SIRFN_DELAY / superframe / multiframe -> the frame number jumps, its phase
                # within the 8-frame SI cycle stays
                l2[0] = 0x08 | (l2[0] & 0x07)
                l2[9] = 0x80 | (l2[9] & 0x03)
                l2[10] = (l2[10] & 0x80) | (int(rng.integers(0, 16)) << 3)
                l2[11] &= 0x3f
            elif kind == "bcch":
                l2[0] = (l2[0] & 0x07) | 0x10            # not an SI1 header: bcch_tdma_align() is a no-op
            else:
                l2[1] = 0x01                             # never an IMM.ASS (gmr1_rx.c:236-239)
                if imm_ass and f in imm_ass:
                    tn, p = imm_ass[f]
                    l2[1], l2[2] = 0x06, 0x3f
                    l2[8] = ((p & 0x3f) << 2) | ((tn >> 3) & 0x03)
                    l2[9] = ((tn & 0x07) << 5) | (l2[9] & 0x1f)
            hard = (enc_bcch if kind == "bcch" else enc_ccch)(l2)
            w = sigen.modulate(kind, hard[None, :], SPS, 32, 16.0 + frac, 0.0, 0.0, 200.0, rng)[0]
            x[pos - 16:pos - 16 + len(w)] += w
            truth.append((f, kind, l2))
        f += 1
        pos += FRAME
    cfo = 2 * np.pi * cfo_hz / 23400.0                   # rad/symbol
    x *= np.exp(1j * (cfo * np.arange(n) / SPS + 0.7))
    sig = 10.0 ** (-esn0_db / 20.0) / np.sqrt(2.0)
    x += sig * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64), truth
