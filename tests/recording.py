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


def make_call(enc_bcch, enc_ccch, enc_speech, enc_facch3, plan, tn=7, p=3, ass_frame=3, kc=None, a5=None,
              seconds=2.6, esn0_db=22.0, cfo_hz=100.0, frac=0.37, start=9000, seed=1,
              csd=None):
    """A BCCH recording whose CCCH burst in frame `ass_frame` is an IMMEDIATE ASSIGNMENT to timeslot `tn` with DKAB
    position `p`, and the traffic-channel recording that goes with it (same clock: same length, carrier offset,
    timing) - what `gmr1_rx sps bcch.cfile tch.cfile [key]` takes (src/gmr1_rx.c:538-600, rx_tch3).

    plan: one character per frame from ass_frame on: 's' NT3 speech burst, 'f' NT3 FACCH3 burst (a codeword is four
    consecutive 'f' frames starting where fn & 3 == 0), 'd' DKAB, '-' nothing.  Every FACCH3 burst carries sync
    sequence 1: the reference's sync search never clears its accumulator between candidate sequences
    (pi4cxpsk.c:207,232), so it always answers the last candidate and cannot demodulate a burst sent with sequence 0.
    enc_speech(frame0[10], frame1[10], bits_s[4], ciph[208] or None) -> 212 hard bits;
    enc_facch3(l2[10], bits_s[32], ciph[384] or None) -> 416 hard bits.  a5(kc, fn, nbits) -> cipher bits when the
    FACCH3 codewords and the speech bursts after the first FACCH3 are to be ciphered (kc given).
    csd = (codeword, tn9, plan9, enc_tch9): the FACCH3 codeword number `codeword` (0 = first) carries an ASSIGNMENT
    COMMAND 1 to timeslot tn9 (facch3_is_ass_cmd_1 / facch3_ass_cmd_1_parse, gmr1_rx.c:247-257), and a third recording
    carries what plan9 says per frame from the frame after that codeword on: 't' = NT9 burst with a TCH9 block
    (enc_tch9(fn) -> 662 hard bits, ciphered and interleaved by the caller's closure), '-' nothing (rx_tch9 :281-355).
    Returns (bcch samples, tch samples, truth) with truth = list of (frame, kind, payload); with csd, a fourth
    element: the csd samples."""
    bcch, truth = make(enc_bcch, enc_ccch, seconds=seconds, esn0_db=esn0_db, cfo_hz=cfo_hz, frac=frac, start=start,
                       seed=seed, imm_ass={ass_frame: (tn, p)})
    rng = np.random.default_rng(seed + 1000)
    n = len(bcch)
    x = np.zeros(n, np.complex128)
    ofs = SPS * tn * 39
    sync_id, group, ciphered = 1, None, False
    n_group, csd_from = 0, None
    for k, what in enumerate(plan):
        f = ass_frame + k
        pos = start + f * FRAME + ofs
        if pos + 117 * SPS + 64 > n:
            break
        if what == "s":
            f0, f1 = rng.integers(0, 256, 10, dtype=np.uint8), rng.integers(0, 256, 10, dtype=np.uint8)
            f0[6:] &= 0; f1[6:] &= 0                       # keep to the 48 protected bits: bytes 6..9 are class 2
            bs = rng.integers(0, 2, 4, dtype=np.uint8)
            c = a5(kc, f, 208) if (kc is not None and ciphered) else None
            hard = enc_speech(f0, f1, bs, c)
            w = sigen.modulate("nt3_speech", hard[None, :], SPS, 32, 16.0 + frac, 0.0, 0.0, 200.0, rng)[0]
            truth.append((f, "tch3", (f0, f1)))
        elif what == "f":
            bi = f & 3
            if bi == 0 or group is None:
                l2 = rng.integers(0, 256, 10, dtype=np.uint8)
                l2[3] = 0x01                               # never an ASSIGNMENT COMMAND 1 (gmr1_rx.c:247-251)
                if csd and n_group == csd[0]:
                    l2[3], l2[4] = 0x06, 0x2e
                    l2[5] = (l2[5] & 0xfc) | ((csd[1] >> 3) & 0x03)
                    l2[6] = ((csd[1] & 0x07) << 5) | (l2[6] & 0x1f)
                    csd_from = f - bi + 4
                n_group += 1
                bs = rng.integers(0, 2, 32, dtype=np.uint8)
                c = None
                if kc is not None:
                    c = np.concatenate([a5(kc, f - bi + i, 96) for i in range(4)])
                group = (enc_facch3(l2, bs, c), l2)
            hard = group[0][104 * bi:104 * bi + 104]
            w = sigen.modulate("nt3_facch", hard[None, :], SPS, 32, 16.0 + frac, 0.0, 0.0, 200.0, rng, sync_id=sync_id)[0]
            if bi == 3:
                truth.append((f, "facch3", group[1]))
                group = None
                ciphered = kc is not None
        elif what == "d":
            w = sigen.modulate_symbols(sigen.dkab_symbols(1, p, rng), SPS, 32, 16.0 + frac, 0.0, 0.0, 200.0, rng)[0]
            truth.append((f, "dkab", None))
        else:
            continue
        x[pos - 16:pos - 16 + len(w)] += w
    cfo = 2 * np.pi * cfo_hz / 23400.0
    rot = np.exp(1j * (cfo * np.arange(n) / SPS + 0.7))
    sig = 10.0 ** (-esn0_db / 20.0) / np.sqrt(2.0)
    x = x * rot + sig * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    if not csd:
        return bcch, x.astype(np.complex64), truth
    y = np.zeros(n, np.complex128)
    for k, what in enumerate(csd[2]):
        f = csd_from + k
        pos = start + f * FRAME + SPS * csd[1] * 39
        if what != "t" or pos + 351 * SPS + 64 > n:
            continue
        w = sigen.modulate("nt9", csd[3](f)[None, :], SPS, 32, 16.0 + frac, 0.0, 0.0, 200.0, rng, sync_id=1)[0]
        y[pos - 16:pos - 16 + len(w)] += w
        truth.append((f, "tch9", None))
    y = y * rot + sig * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return bcch, x.astype(np.complex64), truth, y.astype(np.complex64)
