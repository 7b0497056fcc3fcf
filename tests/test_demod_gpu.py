"""GPU parity, stage 2: batched pi/4-CxPSK demod / detect through the C ABI vs the oracle
(reference gmr1_pi4cxpsk_demod / _detect, src/sdr/pi4cxpsk.c:520,617) on identical synthetic IQ.

Float contract: sync_id identical; TOA within 0.01 sample; freq_err within 2e-5
rad/symbol; soft bits equal within +-1 LSB with >= 99.5 % exactly equal, never off by more than
2; and the decoded L2 + CRC of BCCH / CCCH bursts identical after stage 3.
"""
import numpy as np
import pytest

import sigen

pytestmark = pytest.mark.gpu

SNRS = np.array([6.0, 10.0, 15.0, 30.0])


def gen(name, n, win, rng, sync_ids=1, hard=None):
    neb = sigen.burst_ebits(name)
    if hard is None:
        hard = rng.integers(0, 2, (n, neb), dtype=np.uint8)
    toa = rng.uniform(2, max(win - 2, 2.5), n)
    cfo = rng.uniform(-0.0134, 0.0134, n)          # +-50 Hz residual, inside the estimator's range
    ph = rng.uniform(0, 2 * np.pi, n)
    snr = SNRS[np.arange(n) % 4]
    sid = np.arange(n) % sync_ids
    x = np.zeros((n, sigen.burst_len(name) * 4 + win), np.complex64)
    for s in range(sync_ids):
        m = sid == s
        x[m] = sigen.modulate(name, hard[m], 4, win, toa[m], cfo[m], ph[m], snr[m], rng, sync_id=s)
    return x, sid, hard


def gpu_demod(L, name, x, freq_shift=None, device=False):
    n, wl = x.shape
    neb = sigen.burst_ebits(name)
    eb = np.full((n, neb), 99, np.int8)
    sid = np.full(n, -9, np.int32)
    toa = np.zeros(n, np.float32)
    fe = np.zeros(n, np.float32)
    pw = np.zeros(n, np.float32)
    iq = np.ascontiguousarray(x).view(np.float32)
    if device:
        import torch
        t = [torch.from_numpy(a).cuda() for a in (iq, eb, sid, toa, fe, pw)]
        fs = None if freq_shift is None else torch.from_numpy(freq_shift).cuda()
        L.call("gmr1b200_pi4cxpsk_demod_batch", sigen.BT_ID[name], t[0], n * wl, None, wl, wl, 4, fs, 0.0,
               t[1], neb, t[2], t[3], t[4], t[5], n, None)
        torch.cuda.synchronize()
        eb, sid, toa, fe, pw = [a.cpu().numpy() for a in t[1:]]
    else:
        L.call("gmr1b200_pi4cxpsk_demod_batch", sigen.BT_ID[name], iq, n * wl, None, wl, wl, 4, freq_shift, 0.0,
               eb, neb, sid, toa, fe, pw, n, None)
    return eb, sid, toa, fe, pw


def compare(name, oracle, x, got, freq_shift=None):
    eb, sid, toa, fe, pw = got
    n = x.shape[0]
    exact = total = 0
    for i in range(n):
        rc, eb_o, sid_o, toa_o, fe_o = oracle.demod(name, x[i], 4, 0.0 if freq_shift is None else freq_shift[i])
        assert rc == 0
        assert sid[i] == sid_o, f"{name}[{i}] sync_id {sid[i]} != {sid_o}"
        assert abs(toa[i] - toa_o) <= 0.01, f"{name}[{i}] toa {toa[i]} vs {toa_o}"
        assert abs(fe[i] - fe_o) <= 2e-5, f"{name}[{i}] freq_err {fe[i]} vs {fe_o}"
        d = np.abs(eb[i].astype(int) - eb_o.astype(int))
        assert d.max() <= 2, f"{name}[{i}] ebits differ by {d.max()}"
        exact += int((d == 0).sum())
        total += d.size
    assert exact / total >= 0.995, f"{name}: only {exact / total:.4%} of soft bits identical"
    return exact / total


@pytest.mark.parametrize("name,win,sync_ids", [
    ("bcch", 80, 1), ("dc6", 40, 1), ("nt3_speech", 6, 1), ("nt3_facch", 6, 2), ("nt9", 6, 2),
    ("rach", 6, 1), ("sdcch", 40, 4), ("dc2", 24, 1), ("nt6", 6, 2), ("dc12", 40, 1),
    # search windows of 1, 2 and 3 rows of 32 offsets and beyond (the kernel is specialised on the row count,
    # more than 96 offsets take the generic loop)
    ("bcch", 30, 1), ("bcch", 62, 1), ("bcch", 95, 1), ("bcch", 130, 1), ("nt9", 100, 2)])
@pytest.mark.parametrize("generic", [0, 1])
def test_demod_parity(gpu_lib, oracle, name, win, sync_ids, generic):
    """generic = 0: the standard formats at their standard search widths run on the per-format kernels
    (csrc/demod_fast.cu), everything else on the generic kernel; generic = 1: the generic kernel for all."""
    rng = np.random.default_rng(100 + sigen.BT_ID[name])
    n = 70
    x, sid_true, _ = gen(name, n, win, rng, sync_ids)
    prev = gpu_lib.call("gmr1b200_set_demod_generic", generic)
    try:
        got = gpu_demod(gpu_lib, name, x)
    finally:
        gpu_lib.call("gmr1b200_set_demod_generic", prev)
    compare(name, oracle, x, got)
    if sync_ids == 1:
        assert (got[1] == 0).all()
    # With several candidate sequences the reference scores candidate i on the SUM of the
    # correlations of candidates 0..i (accumulator cleared once, pi4cxpsk.c:207,232) and so
    # favours the last one; the GPU path reproduces that (checked against the oracle above),
    # hence no comparison with the transmitted sync id here.


def test_sync_power_fast_vs_generic(gpu_lib):
    """sync power (an output the reference computes only inside its detect function, pi4cxpsk.c:256-262): the
    per-format kernels run the search on unscaled samples and scale the result, the generic kernel scales every
    chunk correlation as the reference does - same number"""
    rng = np.random.default_rng(3)
    for name, win, sync_ids in (("bcch", 80, 1), ("nt3_facch", 6, 2), ("nt9", 6, 2)):
        x, _, _ = gen(name, 40, win, rng, sync_ids)
        fast = gpu_demod(gpu_lib, name, x)
        prev = gpu_lib.call("gmr1b200_set_demod_generic", 1)
        try:
            generic = gpu_demod(gpu_lib, name, x)
        finally:
            gpu_lib.call("gmr1b200_set_demod_generic", prev)
        assert (fast[1] == generic[1]).all() and np.abs(fast[2] - generic[2]).max() <= 0.004
        assert (fast[4] > 0).all() and np.abs(fast[4] / generic[4] - 1).max() < 1e-3, name
        d = np.abs(fast[0].astype(int) - generic[0].astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.995


@pytest.mark.parametrize("name,win", [("bcch", 80), ("rach", 6)])
def test_demod_device_pointers_and_freq_shift(gpu_lib, oracle, name, win):
    """a frequency shift per window (the per-format kernels rebuild their rotated taps when it changes; RACH has its
    own tap layout: one row of lanes per chunk)"""
    rng = np.random.default_rng(7)
    x, _, _ = gen(name, 40, win, rng)
    fsh = rng.uniform(-0.01, 0.01, 40).astype(np.float32)
    got = gpu_demod(gpu_lib, name, x, freq_shift=fsh, device=True)
    compare(name, oracle, x, got, freq_shift=fsh)


@pytest.mark.parametrize("name,win", [("bcch", 80), ("rach", 6)])
@pytest.mark.parametrize("generic", [0, 1])
def test_demod_hot_path_layouts(gpu_lib, oracle, generic, name, win):
    prev = gpu_lib.call("gmr1b200_set_demod_generic", generic)
    try:
        _hot_path_layouts(gpu_lib, oracle, name, win)
    finally:
        gpu_lib.call("gmr1b200_set_demod_generic", prev)


def _hot_path_layouts(gpu_lib, oracle, name="bcch", win=80):
    """The same bursts through the kernel's usual path (no sync power requested, 16-byte aligned windows, even
    soft-bit rows) and through its out-of-line variants (odd soft-bit row stride -> byte stores, windows at an odd
    sample offset -> unaligned statistics loop + separate region copy, sync power requested) give the same soft bits."""
    rng = np.random.default_rng(77)
    n = 48
    x, _, _ = gen(name, n, win, rng)
    neb = sigen.burst_ebits(name)
    ref = gpu_demod(gpu_lib, name, x)                      # sync power requested (cold statistics variant)
    compare(name, oracle, x, ref)
    wl = x.shape[1]
    iq = np.ascontiguousarray(x).view(np.float32)

    def run(iq_buf, iq_len, ofs, stride, eb_stride):
        eb = np.full((n, eb_stride), 99, np.int8)
        sid = np.full(n, -9, np.int32)
        toa = np.zeros(n, np.float32)
        fe = np.zeros(n, np.float32)
        gpu_lib.call("gmr1b200_pi4cxpsk_demod_batch", sigen.BT_ID[name], iq_buf, iq_len, ofs, stride, wl, 4, None, 0.0,
                     eb, eb_stride, sid, toa, fe, None, n, None)
        return eb, sid, toa, fe

    for eb_stride in (neb, neb + 1, neb + 7):
        eb, sid, toa, fe = run(iq, n * wl, None, wl, eb_stride)
        assert (eb[:, :neb] == ref[0]).all() and (sid == ref[1]).all()
        assert np.abs(toa - ref[2]).max() < 1e-6 and np.abs(fe - ref[3]).max() < 1e-7
    # windows at odd sample offsets (8-byte but not 16-byte aligned)
    pad = np.zeros((n, 2 * (wl + 1)), np.float32)
    pad[:, 2:] = iq.reshape(n, 2 * wl)
    ofs = (np.arange(n, dtype=np.int64) * (wl + 1) + 1)
    eb, sid, toa, fe = run(pad, n * (wl + 1), ofs, 0, neb)
    d = np.abs(eb.astype(int) - ref[0].astype(int))
    assert (sid == ref[1]).all() and np.abs(toa - ref[2]).max() <= 0.004 and d.max() <= 1 and (d == 0).mean() > 0.999


@pytest.mark.parametrize("name,chan,win", [("bcch", "bcch", 80), ("dc6", "ccch", 40)])
def test_demod_then_decode_bit_exact(gpu_lib, oracle, name, chan, win):
    """IQ -> ebits -> L2 on the GPU vs the oracle chain: L2 bytes and CRC identical on the SNR grid"""
    rng = np.random.default_rng(11)
    n = 128
    neb = sigen.burst_ebits(name)
    l2 = rng.integers(0, 256, (n, 24), dtype=np.uint8)
    hard = np.stack([oracle.encode(chan, neb, l2[i]) for i in range(n)])
    x, _, _ = gen(name, n, win, rng, hard=hard)
    eb = gpu_demod(gpu_lib, name, x)[0]
    out = np.zeros((n, 24), np.uint8)
    crc = np.zeros(n, np.int32)
    gpu_lib.call(f"gmr1b200_{chan}_decode_batch", out, eb, None, crc, n, None)
    n_ok = 0
    for i in range(n):
        _, eb_o, _, _, _ = oracle.demod(name, x[i], 4, 0.0)
        l2_o, crc_o, _ = oracle.simple_decode(chan, eb_o)
        assert crc[i] == crc_o and (out[i] == l2_o).all(), f"{name}[{i}]"
        n_ok += int(crc_o == 0 and (l2_o == l2[i]).all())
    assert n_ok >= 0.9 * n


def test_detect_parity(gpu_lib, oracle):
    """gmr1_pi4cxpsk_detect as gmr1_rx uses it: NT3 FACCH vs speech (src/gmr1_rx.c:534-588)"""
    rng = np.random.default_rng(5)
    n = 64
    xs, kinds = [], []
    for i in range(n):
        name = "nt3_facch" if i % 2 else "nt3_speech"
        x, _, _ = gen(name, 1, 6, rng, sync_ids=1)
        xs.append(x[0])
        kinds.append(i % 2)
    x = np.stack(xs)
    types = np.array([sigen.BT_ID["nt3_facch"], sigen.BT_ID["nt3_speech"]], np.int32)
    bt = np.full(n, -9, np.int32)
    sid = np.full(n, -9, np.int32)
    toa = np.zeros(n, np.float32)
    wl = x.shape[1]
    gpu_lib.call("gmr1b200_pi4cxpsk_detect_batch", types, 2, None, 3.0, np.ascontiguousarray(x).view(np.float32),
                 n * wl, None, wl, wl, 4, None, 0.0, bt, sid, toa, n, None)
    for i in range(n):
        rc, bt_o, sid_o, toa_o = oracle.detect(["nt3_facch", "nt3_speech"], 3.0, x[i], 4, 0.0)
        assert rc == 0 and bt[i] == bt_o and sid[i] == sid_o and abs(toa[i] - toa_o) <= 0.01, i
