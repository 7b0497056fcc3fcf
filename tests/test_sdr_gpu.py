"""GPU parity, stage 1 + small SDR entry points: FCCH rough / fine / snr, DKAB, modulation order
vs the oracle (reference src/sdr/fcch.c, dkab.c, pi4cxpsk.c:693) on identical synthetic IQ."""
import numpy as np
import pytest

import sigen

pytestmark = pytest.mark.gpu
SPS = 4


def _iq(x):
    return np.ascontiguousarray(x).view(np.float32)


def test_fcch_rough(gpu_lib, oracle):
    rng = np.random.default_rng(31)
    n, L = 12, 30888                      # 330 ms at sps 4 (src/gmr1_rx.c:612)
    pos = rng.integers(600, L - 1200, n)
    x = np.stack([sigen.fcch_window(L, SPS, int(pos[i]), rng.uniform(-0.08, 0.08), [3.0, 10.0, 20.0][i % 3], rng)
                  for i in range(n)])
    fsh = np.where(np.arange(n) % 2, rng.uniform(-0.05, 0.05, n), 0.0).astype(np.float32)
    toa = np.full(n, -1, np.int32)
    peak = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_fcch_rough_batch", 0, _iq(x), n * L, None, L, L, SPS, fsh, 0.0, toa, peak, n, None)
    same = 0
    for i in range(n):
        rc, t = oracle.fcch_rough(x[i], SPS, fsh[i])
        assert rc == 0 and abs(int(toa[i]) - t) <= 1, (i, toa[i], t)
        assert abs(t - pos[i]) <= 2 * SPS                  # the oracle itself finds the burst
        same += int(toa[i] == t)
    assert same >= n - 1


@pytest.fixture(params=["fft", "fft-one-block", "direct"])
def fcch_kernel(request, gpu_lib):
    """the coarse search through the frequency-domain kernel (csrc/fcch_fft.cu, the default: two 4096-point blocks for
    the standard window), through its one-block form (8192 points) and through the direct-correlation kernel
    (csrc/fcch_grid.cu)"""
    prev = gpu_lib.c.gmr1b200_set_fcch_fft({"fft": 1, "fft-one-block": 2, "direct": 0}[request.param])
    yield request.param
    gpu_lib.c.gmr1b200_set_fcch_fft(prev)


@pytest.mark.parametrize("grid,L", [([0.0, 0.27, -0.27, 0.54, -0.54], 30888 - 4 * 37), ([0.1], 30888 - 4 * 37),
                                    ([-0.2, 0.0], 30888 - 4 * 37), ([0.05, 0.1, 0.15, -0.15], 30888 - 4 * 37),
                                    # window lengths either side of the two-block form of the frequency-domain kernel
                                    # (4097 .. 7808 decimated samples): 7808 (its last), 7809 and 8190 (one block of 8192),
                                    # 4098 (its first, block 1 almost empty), 4000 (one block)
                                    ([0.0, 0.2], 4 * 7808), ([0.0, -0.2], 4 * 7809 + 3), ([0.1], 4 * 8190),
                                    ([0.0, 0.27, -0.27], 4 * 4098 + 1), ([0.27], 4 * 4000)])
def test_fcch_rough_grid(gpu_lib, oracle, grid, L, fcch_kernel):
    """gmr1b200_fcch_rough_grid_batch: every shift of the grid answers what gmr1_fcch_rough (src/sdr/fcch.c:211)
    answers for that freq_shift on the same window - paired (+-f), unpaired and zero shifts, ragged window
    lengths (the last round of outputs partly empty), host and strided windows; all kernels."""
    rng = np.random.default_rng(131 + len(grid))
    n = 10
    stride = L + 24
    pos = rng.integers(600, L - 1200, n)
    cfo = rng.choice(np.array(grid, np.float64), n) * -1.0 + rng.uniform(-0.05, 0.05, n)
    buf = np.zeros((n, stride), np.complex64)
    for i in range(n):
        buf[i, :L] = sigen.fcch_window(L, SPS, int(pos[i]), cfo[i], [3.0, 10.0, 20.0][i % 3], rng)
    g = np.array(grid, np.float32)
    toa = np.full((len(grid), n), -1, np.int32)
    peak = np.zeros((len(grid), n), np.float32)
    gpu_lib.call("gmr1b200_fcch_rough_grid_batch", 0, _iq(buf), n * stride, None, stride, L, SPS, g, len(grid), toa,
                 peak, n, None)
    same = 0
    for k, fs in enumerate(grid):
        for i in range(n):
            rc, t = oracle.fcch_rough(buf[i, :L], SPS, float(g[k]))
            assert rc == 0 and abs(int(toa[k, i]) - t) <= 1, (k, i, toa[k, i], t)
            same += int(toa[k, i] == t)
    assert same >= len(grid) * n - 2
    assert (peak > 0).all()
    # single-shift entry on the same windows: same answers as the grid's column
    t1 = np.full(n, -1, np.int32)
    p1 = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_fcch_rough_batch", 0, _iq(buf), n * stride, None, stride, L, SPS, None, float(g[0]), t1, p1, n,
                 None)
    assert (t1 == toa[0]).all() and np.allclose(p1, peak[0], rtol=1e-4)
    # the other kernel on the same windows: same positions (rounding ties aside), same window energies
    other = gpu_lib.c.gmr1b200_set_fcch_fft(0 if fcch_kernel.startswith("fft") else 1)
    try:
        toa2 = np.full((len(grid), n), -1, np.int32)
        peak2 = np.zeros((len(grid), n), np.float32)
        gpu_lib.call("gmr1b200_fcch_rough_grid_batch", 0, _iq(buf), n * stride, None, stride, L, SPS, g, len(grid), toa2,
                     peak2, n, None)
    finally:
        gpu_lib.c.gmr1b200_set_fcch_fft(other)
    assert np.abs(toa2 - toa).max() <= 1 and (toa2 == toa).mean() > 0.95 and np.allclose(peak2, peak, rtol=2e-4)


def test_fcch_fine_and_snr(gpu_lib, oracle):
    rng = np.random.default_rng(32)
    n, W = 96, 117 * SPS
    xs, fsh = [], []
    for i in range(n):
        off = int(rng.integers(-12, 13))                   # residual timing error after rough
        full = sigen.fcch_window(W + 64, SPS, 32 + off, rng.uniform(-0.15, 0.15), [6.0, 12.0, 25.0][i % 3], rng)
        xs.append(full[32:32 + W])
        fsh.append(rng.uniform(-0.03, 0.03) if i % 2 else 0.0)
    x = np.stack(xs)
    fsh = np.array(fsh, np.float32)
    toa = np.full(n, -999, np.int32)
    fe = np.zeros(n, np.float32)
    snr = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_fcch_fine_batch", 0, _iq(x), n * W, None, W, SPS, fsh, 0.0, toa, fe, n, None)
    gpu_lib.call("gmr1b200_fcch_snr_batch", 0, _iq(x), n * W, None, W, SPS, fsh, 0.0, snr, n, None)
    same = 0
    for i in range(n):
        rc, t, f = oracle.fcch_fine(x[i], SPS, fsh[i])
        assert rc == 0 and abs(int(toa[i]) - t) <= 1 and abs(fe[i] - f) <= 1e-5, (i, toa[i], t, fe[i], f)
        same += int(toa[i] == t)
        rc, s = oracle.fcch_snr(x[i], SPS, fsh[i])
        assert rc == 0 and abs(snr[i] - s) <= 2e-3 * abs(s), (i, snr[i], s)
    assert same >= 0.97 * n


@pytest.mark.parametrize("sps,ftype", [(4, 0), (2, 0), (1, 0), (3, 0), (4, 1)])
def test_fcch_acquire(gpu_lib, oracle, sps, ftype):
    """fcch_single_init (src/gmr1_rx.c:606-639) in one call: rough over the search window, fine on the burst
    found; every other sps the reference accepts and the 12-slot FCCH3 format too."""
    rng = np.random.default_rng(33 + sps + 10 * ftype)
    blen = [117, 468, 468][ftype]
    n, L = 10, (330 * 23400 * sps) // 1000 + (0 if ftype == 0 else 468 * sps)
    pos = rng.integers(150 * sps, L - (blen + 150) * sps, n)
    cfo = rng.uniform(-0.2, 0.2, n) * (1.0 if ftype == 0 else 0.25)
    x = np.stack([sigen.fcch_window(L, sps, int(pos[i]), cfo[i], [4.0, 10.0, 20.0][i % 3], rng,
                                    [0.32, 0.32, 0.16][ftype], blen) for i in range(n)])
    rough = np.full(n, -1, np.int32)
    align = np.full(n, -1, np.int32)
    fe = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_fcch_acquire_batch", ftype, _iq(x), n * L, None, L, L, sps, rough, align, fe, n, None)
    same = 0
    for i in range(n):
        rc, t = oracle.fcch_rough(x[i], sps, 0.0, ftype)
        assert rc == 0 and abs(int(rough[i]) - t) <= 1, (i, rough[i], t)
        a = int(rough[i])                                   # fine stage judged on the GPU's own rough TOA
        rc, ft, f = oracle.fcch_fine(x[i][a:a + blen * sps], sps, 0.0, ftype)
        assert rc == 0 and abs(int(align[i]) - (a + ft)) <= 1 and abs(fe[i] - f) <= 1e-5, (i, align[i], a + ft, fe[i], f)
        same += int(rough[i] == t and align[i] == a + ft)
        assert abs(int(align[i]) - pos[i]) <= 3 * sps and abs(fe[i] - cfo[i]) < 0.02
    assert same >= n - 1


def test_dkab(gpu_lib, oracle):
    rng = np.random.default_rng(33)
    n, win = 64, 6                          # gmr1_rx maps DKABs with the NT3 window (src/gmr1_rx.c:549)
    p = rng.integers(0, 40, n).astype(np.int32)
    L = 117 * SPS + win
    x = np.zeros((n, L), np.complex64)
    for i in range(n):
        if i % 4 == 3:                       # not a DKAB: a full NT3 burst
            hard = rng.integers(0, 2, (1, 212), dtype=np.uint8)
            x[i] = sigen.modulate("nt3_speech", hard, SPS, win, rng.uniform(1, 5), 0.0, rng.uniform(0, 6), 15.0, rng)[0]
        else:
            x[i] = sigen.modulate_symbols(sigen.dkab_symbols(1, int(p[i]), rng), SPS, win, rng.uniform(1, 5), 0.0,
                                          rng.uniform(0, 6), [10.0, 20.0, 30.0][i % 3], rng)[0]
    eb = np.zeros((n, 8), np.int8)
    toa = np.zeros(n, np.float32)
    rv = np.full(n, -9, np.int32)
    gpu_lib.call("gmr1b200_dkab_demod_batch", _iq(x), n * L, None, L, L, SPS, None, 0.0, p, 0, eb, toa, rv, n, None)
    found = 0
    for i in range(n):
        rc, eb_o, toa_o = oracle.dkab_demod(x[i], SPS, 0.0, int(p[i]))
        assert rv[i] == rc, (i, rv[i], rc)
        assert abs(toa[i] - toa_o) <= 2e-3 * max(1.0, abs(toa_o)), (i, toa[i], toa_o)
        if rc == 0:
            found += 1
            assert np.abs(eb[i].astype(int) - eb_o.astype(int)).max() <= 1, (i, eb[i], eb_o)
    assert found >= n // 2


def test_mod_order(gpu_lib, oracle):
    rng = np.random.default_rng(34)
    n, win = 40, 6
    L = 117 * SPS + win
    x = np.zeros((n, L), np.complex64)
    for i in range(n):
        name = "nt3_facch" if i % 2 else "nt3_speech"
        hard = rng.integers(0, 2, (1, sigen.burst_ebits(name)), dtype=np.uint8)
        x[i] = sigen.modulate(name, hard, SPS, win, 3.0, rng.uniform(-0.01, 0.01), rng.uniform(0, 6), 20.0, rng)[0]
    order = np.zeros(n, np.int32)
    gpu_lib.call("gmr1b200_pi4cxpsk_mod_order_batch", _iq(x), n * L, None, L, L, SPS, None, 0.0, order, n, None)
    for i in range(n):
        assert order[i] == oracle.mod_order(x[i], SPS, 0.0), i
    # the heuristic itself is allowed an occasional miss (parity with the oracle is the check above)
    assert (order[1::2] == 2).mean() >= 0.9 and (order[0::2] == 4).mean() >= 0.9


def test_fcch_rough_multi(gpu_lib, oracle):
    """gmr1b200_fcch_rough_multi vs gmr1_fcch_rough_multi (src/sdr/fcch.c:341-483): 650 ms windows with the FCCHs of
    one, two and three cells (each twice, 7488 symbols apart), found with the same count and the same TOAs."""
    rng = np.random.default_rng(21)
    W = (650 * 23400 * SPS) // 1000
    per = 7488 * SPS
    chirp = sigen.fcch_chirp(SPS, 0.32, 117)
    n_found = 0
    for case, cells in enumerate([[(5000, 1.0)], [(5000, 1.0), (17000, 0.7)], [(1234, 0.8), (9000, 1.0), (23456, 0.6)],
                                  [(29000, 1.0), (12000, 0.9)]]):
        sig = 10.0 ** (-10.0 / 20.0) / np.sqrt(2.0)
        x = sig * (rng.standard_normal(W) + 1j * rng.standard_normal(W))
        for pos, amp in cells:
            for k in range(2):
                p0 = pos + k * per
                cfo = 0.02 * (case - 1)
                x[p0:p0 + len(chirp)] += amp * chirp * np.exp(1j * (cfo * np.arange(len(chirp)) / SPS + rng.uniform(0, 6.28)))
        x = x.astype(np.complex64)
        rc_o, toa_o = oracle.fcch_rough_multi(x, SPS, 0.0)
        toa = np.zeros(16, np.int32)
        rc = gpu_lib.call("gmr1b200_fcch_rough_multi", 0, np.ascontiguousarray(x).view(np.float32), W, SPS, 0.0, toa, 16, None)
        assert rc == rc_o, (case, rc, rc_o)
        assert rc_o == len(cells), (case, rc_o, toa_o)                       # the generator is understood by the reference
        for a, b in zip(sorted(toa[:rc].tolist()), sorted(toa_o)):
            assert abs(a - b) <= 1, (case, toa[:rc], toa_o)
        for (pos, _), t in zip(sorted(cells), sorted(toa_o)):
            assert abs(t - pos) <= 2 * SPS, (case, toa_o)
        n_found += rc
    assert n_found == 8
    # noise only: whatever the reference answers (nothing found, or -EINVAL when the two cycles do not line up)
    x = (rng.standard_normal(W) + 1j * rng.standard_normal(W)).astype(np.complex64)
    rc_o, _ = oracle.fcch_rough_multi(x, SPS, 0.0)
    toa = np.zeros(16, np.int32)
    try:
        rc = gpu_lib.call("gmr1b200_fcch_rough_multi", 0, np.ascontiguousarray(x).view(np.float32), W, SPS, 0.0, toa, 16, None)
    except Exception as e:                                                   # the wrapper raises on negative codes
        rc = -22 if "rc=-22" in str(e) else None
    assert rc == rc_o or (rc_o > 0 and rc > 0), (rc, rc_o)

