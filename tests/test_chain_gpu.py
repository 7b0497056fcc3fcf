"""GPU parity of whole chains (IQ -> soft bits -> L2) for the traffic-channel configurations of
BASELINE.json (configs 3 and 4 at test size): NT3 speech -> TCH3 (A5/1 ciphered), NT3 FACCH x4 ->
FACCH3, NT9 -> FACCH9 / TCH9-9k6 (3-deep inter-burst interleaver), RACH, and the +-frequency-offset
FCCH search, each against the oracle chain on identical synthetic IQ.  Bars: decoded bytes / CRC
identical wherever the oracle's CRC passes (all, for CRC-less channels at >= 10 dB), soft bits +-1.
"""
import numpy as np
import pytest

import sigen

pytestmark = pytest.mark.gpu
SPS = 4


def _iq(x):
    return np.ascontiguousarray(x).view(np.float32)


def _demod(L, name, x, sps=SPS):
    n, wl = x.shape
    neb = sigen.burst_ebits(name)
    eb = np.zeros((n, neb), np.int8)
    sid = np.zeros(n, np.int32)
    toa = np.zeros(n, np.float32)
    L.call("gmr1b200_pi4cxpsk_demod_batch", sigen.BT_ID[name], _iq(x), n * wl, None, wl, wl, sps, None, 0.0,
           eb, neb, sid, toa, None, None, n, None)
    return eb, sid, toa


def _mod(name, hard, win, rng, snr_lo=10.0, sync_id=0, sps=SPS, cfo=0.012):
    """cfo: single-chunk formats (NT3, DC2) have no burst-level frequency estimate in the reference
    (pi4cxpsk.c:399-403) and rely on the BCCH-tracked frequency: keep their residual small"""
    n = hard.shape[0]
    snr = np.where(np.arange(n) % 2, 30.0, snr_lo)
    return sigen.modulate(name, hard, sps, win, rng.uniform(1.5, win - 1.5, n), rng.uniform(-cfo, cfo, n),
                          rng.uniform(0, 6.28, n), snr, rng, sync_id=sync_id)


def test_tch3_speech_ciphered(gpu_lib, oracle):
    rng = np.random.default_rng(41)
    n = 96
    key = rng.integers(0, 256, 8, dtype=np.uint8)
    f0 = rng.integers(0, 256, (n, 10), dtype=np.uint8)
    f1 = rng.integers(0, 256, (n, 10), dtype=np.uint8)
    ciph = np.stack([oracle.a5(1 if i % 2 else 0, key, 1000 + i, 208) for i in range(n)])   # A5/1 and A5/0 halves
    # the same masks from the device-side generator; the decoder below consumes these
    alg = (np.arange(n) % 2).astype(np.int32)
    ciph_g = np.zeros((n, 208), np.uint8)
    gpu_lib.call("gmr1b200_a5_batch", alg, 0, np.tile(key, (n, 1)), (1000 + np.arange(n)).astype(np.uint32), 208, 208,
                 ciph_g, None, n, None)
    assert (ciph_g == ciph).all()
    hard = np.zeros((n, 212), np.uint8)
    for i in range(n):
        gpu_lib.call("gmr1b200_tch3_encode", hard[i], f0[i], f1[i], rng.integers(0, 2, 4, dtype=np.uint8), ciph[i], 0)
    x = _mod("nt3_speech", hard, 6, rng, cfo=0.001)
    eb, _, _ = _demod(gpu_lib, "nt3_speech", x)
    g0 = np.zeros((n, 10), np.uint8)
    g1 = np.zeros((n, 10), np.uint8)
    gpu_lib.call("gmr1b200_tch3_decode_batch", g0, g1, None, eb, ciph_g, 0, None, None, n, None)
    for i in range(n):
        _, eb_o, _, _, _ = oracle.demod("nt3_speech", x[i], SPS, 0.0)
        o0, o1, _, _, _ = oracle.tch3_decode(eb_o, ciph[i], 0)
        assert (g0[i] == o0).all() and (g1[i] == o1).all(), i
    hi = np.arange(n) % 2 == 1
    assert (g0[hi, :6] == f0[hi, :6]).all()          # the 48 protected bits arrive at 30 dB


def test_facch3_four_bursts(gpu_lib, oracle):
    rng = np.random.default_rng(42)
    n = 40
    l2 = rng.integers(0, 256, (n, 10), dtype=np.uint8)
    l2[:, 9] &= 0x0F
    e_all = np.zeros((n, 416), np.int8)
    e_ora = np.zeros((n, 416), np.int8)
    for i in range(n):
        hard = oracle.facch3_encode(l2[i], rng.integers(0, 2, 32, dtype=np.uint8))
        x = _mod("nt3_facch", hard.reshape(4, 104), 6, rng, sync_id=1, cfo=0.001)
        e_all[i] = _demod(gpu_lib, "nt3_facch", x)[0].reshape(-1)
        e_ora[i] = np.concatenate([oracle.demod("nt3_facch", x[b], SPS, 0.0)[1] for b in range(4)])
    out = np.zeros((n, 10), np.uint8)
    crc = np.zeros(n, np.int32)
    gpu_lib.call("gmr1b200_facch3_decode_batch", out, None, e_all, None, None, crc, n, None)
    for i in range(n):
        l2_o, _, crc_o, _ = oracle.facch3_decode(e_ora[i])
        assert crc[i] == crc_o and (out[i] == l2_o).all(), i
    assert (crc == 0).mean() > 0.9 and (out[crc == 0] == l2[crc == 0]).all()


def test_facch9_and_tch9_over_nt9(gpu_lib, oracle, port_noquirk):
    """NT9 carries FACCH9 (sync 0) or TCH9 (sync 1).  With the reference's accumulator quirk sync 1
    always wins and FACCH9 never decodes (checked: GPU == reference).  With the opt-in reset both
    are told apart; the checker for that mode is the oracle port with the same switch."""
    rng = np.random.default_rng(43)
    nch, nb = 6, 6
    n = nch * nb
    l2 = rng.integers(0, 256, (n, 38), dtype=np.uint8)
    l2[:, 37] &= 0x0F
    hard = np.stack([oracle.facch9_encode(l2[i], rng.integers(0, 2, 10, dtype=np.uint8), rng.integers(0, 2, 4, dtype=np.uint8))
                     for i in range(n)])
    x = _mod("nt9", hard, 6, rng, sync_id=0)
    # reference behaviour: identical to the reference, which mis-identifies the sequence
    eb, sid, _ = _demod(gpu_lib, "nt9", x)
    for i in range(n):
        _, eb_o, sid_o, _, _ = oracle.demod("nt9", x[i], SPS, 0.0)
        assert sid[i] == sid_o == 1 and np.abs(eb[i].astype(int) - eb_o.astype(int)).max() <= 2
    # opt-in reset: FACCH9 is found and decodes
    prev = gpu_lib.call("gmr1b200_set_sync_accumulator_reset", 1)
    try:
        eb, sid, _ = _demod(gpu_lib, "nt9", x)
        out = np.zeros((n, 38), np.uint8)
        crc = np.zeros(n, np.int32)
        gpu_lib.call("gmr1b200_facch9_decode_batch", out, None, None, eb, None, None, crc, n, None)
        for i in range(n):
            _, eb_o, sid_o, _, _ = port_noquirk.demod("nt9", x[i], SPS, 0.0)
            l2_o, _, _, crc_o, _ = port_noquirk.facch9_decode(eb_o)
            assert sid[i] == sid_o and crc[i] == crc_o and (out[i] == l2_o).all(), i
        assert (sid == 0).all() and (crc == 0).mean() > 0.9 and (out[crc == 0] == l2[crc == 0]).all()
    finally:
        gpu_lib.call("gmr1b200_set_sync_accumulator_reset", prev)
    # TCH9 9k6 (sync 1, found either way): channel-major streams of nb consecutive bursts
    pay = rng.integers(0, 256, (n, 60), dtype=np.uint8)
    hard = np.zeros((n, 662), np.uint8)
    p1 = np.full(n, -1, np.int32)
    p2 = np.full(n, -1, np.int32)
    for c in range(nch):
        il = oracle.interleaver()
        for b in range(nb):
            i = c * nb + b
            hard[i] = oracle.tch9_encode(pay[i], 2, rng.integers(0, 2, 10, dtype=np.uint8), rng.integers(0, 2, 4, dtype=np.uint8), None, il)
            p1[i] = i - 1 if b >= 1 else -1
            p2[i] = i - 2 if b >= 2 else -1
    x = _mod("nt9", hard, 6, rng, snr_lo=15.0, sync_id=1)
    eb, _, _ = _demod(gpu_lib, "nt9", x)
    out = np.zeros((n, 60), np.uint8)
    gpu_lib.call("gmr1b200_tch9_decode_batch", out, None, None, eb, 2, None, p1, p2, None, n, None)
    for c in range(nch):
        il = oracle.interleaver()
        for b in range(nb):
            i = c * nb + b
            l2_o, _, _, _ = oracle.tch9_decode(oracle.demod("nt9", x[i], SPS, 0.0)[1], 2, None, il)
            assert (out[i] == l2_o).all(), i
            if b >= 2:
                assert (out[i] == pay[i - 2]).all()      # the interleaver delays the payload by two bursts


def test_rach(gpu_lib, oracle):
    rng = np.random.default_rng(44)
    n = 48
    pay = rng.integers(0, 256, (n, 18), dtype=np.uint8)
    pay[:, 17] &= 0x07
    hard = np.stack([oracle.rach_encode(pay[i], 0x3C) for i in range(n)])
    x = _mod("rach", hard, 6, rng)
    eb, _, _ = _demod(gpu_lib, "rach", x)
    out = np.zeros((n, 18), np.uint8)
    crc = np.zeros(n, np.int32)
    gpu_lib.call("gmr1b200_rach_decode_batch", out, eb, None, 0x3C, None, None, crc, n, None)
    for i in range(n):
        r_o, crc_o, _, _ = oracle.rach_decode(oracle.demod("rach", x[i], SPS, 0.0)[1], 0x3C)
        assert crc[i] == crc_o and (out[i] == r_o).all(), i
    assert (crc == 0).mean() > 0.9 and (out[crc == 0] == pay[crc == 0]).all()


@pytest.mark.parametrize("sps,win", [(8, 160), (2, 40), (3, 60), (1, 20)])
def test_other_oversampling(gpu_lib, oracle, sps, win):
    """sps = 8: run-time-sps kernel; sps < 4: the reference's sinc-interpolating symbol alignment
    (src/sdr/pi4cxpsk.c:298-343)"""
    rng = np.random.default_rng(45 + sps)
    n = 24
    hard = rng.integers(0, 2, (n, 424), dtype=np.uint8)
    x = _mod("bcch", hard, win, rng, sps=sps)
    eb, sid, toa = _demod(gpu_lib, "bcch", x, sps=sps)
    exact = 0
    for i in range(n):
        rc, eb_o, sid_o, toa_o, _ = oracle.demod("bcch", x[i], sps, 0.0)
        assert rc == 0 and sid[i] == sid_o and abs(toa[i] - toa_o) <= 0.02, (i, toa[i], toa_o)
        dlt = np.abs(eb[i].astype(int) - eb_o.astype(int))
        assert dlt.max() <= 2, (i, dlt.max())
        exact += int((dlt == 0).sum())
    assert exact >= 0.99 * n * 424
    hi = np.arange(n) % 2 == 1                         # 30 dB bursts demodulate (1 sample/symbol cannot
    ok = ((eb[hi] < 0) == (hard[hi] > 0)).mean()          # resolve the timing offset: a few errors remain)
    assert ok > (0.999 if sps >= 2 else 0.98)


def test_fcch_search_over_frequency_grid(gpu_lib, oracle):
    """config 4: the +-frequency-offset FCCH search = the same window under several pre-rotations"""
    rng = np.random.default_rng(46)
    L = 30888
    true_cfo = 2 * np.pi * 1800.0 / 23400.0            # +1.8 kHz
    x = sigen.fcch_window(L, SPS, 11111, true_cfo, 12.0, rng)
    grid_hz = np.array([-2000.0, -1000.0, 0.0, 1000.0, 2000.0])
    fsh = (-2 * np.pi * grid_hz / 23400.0).astype(np.float32)
    ofs = np.zeros(5, np.int64)
    toa = np.zeros(5, np.int32)
    peak = np.zeros(5, np.float32)
    gpu_lib.call("gmr1b200_fcch_rough_batch", 0, _iq(x), L, ofs, 0, L, SPS, fsh, 0.0, toa, peak, 5, None)
    for k in range(5):
        assert abs(int(toa[k]) - oracle.fcch_rough(x, SPS, fsh[k])[1]) <= 1
    best = int(np.argmax(peak))
    assert grid_hz[best] == 2000.0 and abs(int(toa[best]) - 11111) <= 2 * SPS
    # fine estimate on the winning hypothesis recovers the residual -200 Hz
    w = x[toa[best]:toa[best] + 468]
    t = np.zeros(1, np.int32)
    fe = np.zeros(1, np.float32)
    gpu_lib.call("gmr1b200_fcch_fine_batch", 0, _iq(w), 468, None, 468, SPS, None, float(fsh[best]), t, fe, 1, None)
    assert abs(fe[0] * 23400.0 / (2 * np.pi) - (-200.0)) < 25.0


@pytest.mark.parametrize("chan,name,neb", [(0, "bcch", 424), (1, "dc6", 432), (2, "dc12", 432)])
def test_fused_burst_to_l2(gpu_lib, oracle, chan, name, neb):
    """gmr1b200_rx_xcch_batch == demod_batch followed by the channel's decode_batch, and == the oracle's L2 / CRC"""
    rng = np.random.default_rng(60 + chan)
    n, win = 70, 40
    l2 = rng.integers(0, 256, (n, 24), dtype=np.uint8)
    enc = {0: "bcch", 1: "ccch", 2: "xch_dc12"}[chan]
    hard = np.stack([oracle.encode(enc, neb, l2[i]) for i in range(n)])
    x = _mod(name, hard, win, rng, snr_lo=8.0)
    wl = x.shape[1]
    out = np.zeros((n, 24), np.uint8)
    crc = np.zeros(n, np.int32)
    conv = np.zeros(n, np.int32)
    toa = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_rx_xcch_batch", chan, np.ascontiguousarray(x).view(np.float32), n * wl, None, wl, wl, SPS,
                 None, 0.0, out, crc, conv, toa, None, n, None)
    eb, _, toa2 = _demod(gpu_lib, name, x)
    out2 = np.zeros((n, 24), np.uint8)
    crc2 = np.zeros(n, np.int32)
    conv2 = np.zeros(n, np.int32)
    dec = {0: "gmr1b200_bcch_decode_batch", 1: "gmr1b200_ccch_decode_batch", 2: "gmr1b200_xch_dc12_decode_batch"}[chan]
    gpu_lib.call(dec, out2, eb, conv2, crc2, n, None)
    assert (out == out2).all() and (crc == crc2).all() and (conv == conv2).all() and (toa == toa2).all()
    for i in range(0, n, 5):
        _, eb_o, _, _, _ = oracle.demod(name, x[i], SPS, 0.0)
        l2_o, crc_o, _ = oracle.simple_decode(enc, eb_o)
        assert crc[i] == crc_o and (out[i] == l2_o).all(), i
    assert (crc == 0).mean() > 0.9 and (out[crc == 0] == l2[crc == 0]).all()
