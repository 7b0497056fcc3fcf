"""Wideband channeliser, CPU side: the oracle restatement of the GNU Radio blocks that
utils/gmr1_rx_sdr.py wires (oracle/chan_port.py) against first principles, and the product's host-side plan
(filter designs, phase walk, output length - no GPU needed) against the oracle.

GNU Radio is not in the reference tree nor in this image: PARITY UNPINNED for this row (see oracle/chan_port.py)."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import chan_port as cp  # noqa: E402
import osmo_gmr_b200  # noqa: E402


class Info(ctypes.Structure):              # include/gmr1_b200.h: struct gmr1b200_chan_info
    _fields_ = [("n_chans", ctypes.c_int32), ("sps", ctypes.c_int32), ("n_taps", ctypes.c_int32),
                ("taps_per_branch", ctypes.c_int32), ("n_taps_resamp", ctypes.c_int32), ("fft_stages", ctypes.c_int32),
                ("samp_rate", ctypes.c_double), ("mid_rate", ctypes.c_double), ("resamp", ctypes.c_double),
                ("delay_out", ctypes.c_double)]


def plan_of(L, n_chans, sps=4):
    h = ctypes.c_void_p()
    rc = L.c.gmr1b200_chan_create(n_chans, sps, ctypes.byref(h))
    return rc, h


def test_filter_designs_follow_firdes():
    # firdes.low_pass(1, 16 x 31 250, 15 625, 7 812.5): 53 dB Hamming => int(53 fs / (22 tw)) taps made odd, unit DC gain
    pl = cp.Plan(16)
    assert len(pl.taps) == 155 and pl.taps_per_branch == 10
    assert abs(float(pl.taps.astype(np.float64).sum()) - 1.0) < 1e-6 and np.allclose(pl.taps, pl.taps[::-1])
    assert len(cp.Plan(1024).taps) == 9867                                   # int(53 * 32e6 / (22 * 7812.5)) = 9867 (odd)
    # firdes.root_raised_cosine(32, 32 x 62 500, 23 400, 0.35, int(11 x 32 x 62 500 / 23 400)): 940 -> 941 taps, DC gain 32
    assert len(pl.taps_resamp) == 941 and abs(float(pl.taps_resamp.astype(np.float64).sum()) - 32.0) < 1e-4
    assert np.allclose(pl.taps_resamp, pl.taps_resamp[::-1], atol=1e-7) and pl.taps_resamp.argmax() == 470
    # RRC * RRC is Nyquist: the autocorrelation of the prototype vanishes at multiples of the symbol period
    spb = 32 * 62500 / 23400.0
    ac = np.correlate(pl.taps_resamp.astype(np.float64), pl.taps_resamp.astype(np.float64), "full")
    mid = len(ac) // 2
    for k in (1, 2, 3):
        lo, hi = int(np.floor(k * spb)), int(np.ceil(k * spb))
        v = ac[mid + lo] + (ac[mid + hi] - ac[mid + lo]) * (k * spb - lo)
        assert abs(v) < 0.02 * ac[mid]
    assert abs(pl.resamp - 1.4976) < 1e-12


def test_bank_polyphase_form_equals_its_defining_sum_and_separates_carriers():
    rng = np.random.default_rng(3)
    for n_chans in (8, 12, 16):
        pl = cp.Plan(n_chans)
        x = (rng.standard_normal(n_chans * 30) + 1j * rng.standard_normal(n_chans * 30)).astype(np.complex64)
        a = cp.pfb_channelize_direct(x, pl.taps, n_chans, list(range(n_chans)))
        b = cp.pfb_channelize(x, pl.taps, n_chans)
        assert a.shape == b.shape == (n_chans, 60) and np.abs(a - b).max() < 1e-6
    # a carrier 2 kHz above the centre of channel 3 (and one in channel 13 = -3): shows up there, at +2 kHz, at DC gain 1
    pl = cp.Plan(16)
    n = np.arange(16 * 300)
    fs = 16 * cp.CHAN_WIDTH
    for k in (3, 13):
        fc = (k if k < 8 else k - 16) * cp.CHAN_WIDTH + 2000.0
        y = cp.pfb_channelize(np.exp(2j * np.pi * fc * n / fs).astype(np.complex64), pl.taps, 16)[:, 40:]
        pw = (np.abs(y) ** 2).mean(axis=1)
        assert abs(pw[k] - 1.0) < 0.02 and np.delete(pw, k).max() < 1e-5
        f = np.angle(y[k, 1:] * np.conj(y[k, :-1])).mean() * 62500.0 / (2 * np.pi)
        assert abs(f - 2000.0) < 1.0


def test_resampler_walk_and_alignment():
    pl = cp.Plan(16)
    ii, jj, aa = cp.arb_resampler_schedule(pl.resamp, 5000, len(pl.taps_resamp))
    assert jj[0] == 470 % 32 and ii[0] == 0 and aa[0] == 0.0              # start phase (ntaps / 2) % 32
    assert abs(len(ii) / 5000.0 - pl.resamp) < 1e-3                         # 1.4976 outputs per input
    pos = ii + (jj + aa) / 32.0                                             # position of every output on the input axis
    assert np.abs(np.diff(pos) - 1.0 / pl.resamp).max() < 1e-4             # evenly spaced (float32 accumulator drift only)
    # a smooth pulse sent down channel 5 comes out of bank + resampler delay_out samples later, at unit gain
    n = np.arange(500)
    n0 = 200.3
    s = np.exp(-0.5 * ((n - n0) / 6.0) ** 2).astype(np.complex64)
    y = cp.channelize(cp.synth_wideband(s[None, :], [5], 16), pl, [5])[0]
    m = np.abs(y)
    k = int(m.argmax())
    peak = k + 0.5 * (m[k - 1] - m[k + 1]) / (m[k - 1] - 2 * m[k] + m[k + 1])
    assert abs(peak - (n0 + pl.delay_out)) < 0.02 and abs(m[k] - 1.0) < 0.02


def test_product_plan_matches_the_oracle():
    """chan_plan.cu (host code of the product) designs the same filters, walks the same phases and reports the same
    output length and delay as the independent numpy restatement."""
    L = osmo_gmr_b200.Lib()
    for n_chans, sps in ((16, 4), (12, 4), (46, 4), (1024, 4), (64, 2)):
        rc, h = plan_of(L, n_chans, sps)
        assert rc == 0
        pl = cp.Plan(n_chans, sps)
        info = Info()
        assert L.c.gmr1b200_chan_info(h, ctypes.byref(info)) == 0
        assert (info.n_chans, info.sps, info.n_taps, info.taps_per_branch, info.n_taps_resamp) == \
            (n_chans, sps, len(pl.taps), pl.taps_per_branch, len(pl.taps_resamp))
        assert info.samp_rate == pl.samp_rate and info.mid_rate == 62500.0 and abs(info.resamp - pl.resamp) < 1e-12
        assert abs(info.delay_out - pl.delay_out) < 1e-9
        taps = np.zeros(info.n_taps, np.float32)
        rrc = np.zeros(info.n_taps_resamp, np.float32)
        assert L.c.gmr1b200_chan_taps(h, taps.ctypes.data_as(ctypes.c_void_p), len(taps),
                                      rrc.ctypes.data_as(ctypes.c_void_p), len(rrc)) == 0
        assert np.abs(taps - pl.taps).max() <= 1e-9 + 2e-7 * np.abs(pl.taps).max()
        assert np.abs(rrc - pl.taps_resamp).max() <= 2e-6
        for n_wide in (0, n_chans // 2 - 1, n_chans * 37, n_chans * 1000 + 3):
            steps = n_wide // (n_chans // 2)
            want = len(cp.arb_resampler_schedule(pl.resamp, steps, len(pl.taps_resamp))[0]) if steps else 0
            # the walk emits every output of the last consumed step; the product keeps those whose step is in range
            ii = cp.arb_resampler_schedule(pl.resamp, steps, len(pl.taps_resamp))[0] if steps else np.zeros(0, np.int32)
            want = int((ii < steps).sum())
            assert L.c.gmr1b200_chan_out_len(h, n_wide) == want
        L.c.gmr1b200_chan_destroy(h)
    # bank sizes the FFT cannot take, odd sizes, bad sps
    for bad in ((15, 4), (0, 4), (2 * 37, 4), (8192, 4), (16, 0)):
        rc, h = plan_of(L, *bad)
        assert rc < 0 and b"chan_create" in L.c.gmr1b200_last_error()


def test_register_butterfly_fft_on_the_cpu():
    """chan_fft.cuh (the radix-16 register butterflies and Stockham index arithmetic of pfb_fast_kernel and of the
    FCCH search) compiled for the host by tests/emu/chan_emu.cpp and run with the kernels' barrier structure: equals
    numpy's reverse FFT for every size the kernels use (stage plans 16, 16x2 ... 16x16x16x2)."""
    import ctypes
    import subprocess
    emu_dir = os.path.join(ROOT, "tests", "emu")
    so, src = os.path.join(emu_dir, "libchan_emu.so"), os.path.join(emu_dir, "chan_emu.cpp")
    hdr = os.path.join(ROOT, "osmo_gmr_b200", "csrc", "chan_fft.cuh")
    hdr2 = os.path.join(ROOT, "osmo_gmr_b200", "csrc", "a5_bitslice.cuh")
    if not os.path.exists(so) or max(os.path.getmtime(src), os.path.getmtime(hdr), os.path.getmtime(hdr2)) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I" + os.path.dirname(hdr), "-o", so, src])
    emu = ctypes.CDLL(so)
    rng = np.random.default_rng(3)
    for log2n in range(4, 14):
        n = 1 << log2n
        x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
        y = np.zeros(n, np.complex64)
        assert emu.chan_emu_fft(log2n, x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p)) == 0
        want = np.fft.ifft(x.astype(np.complex128)) * n
        assert np.abs(y - want).max() < 1e-6 * np.abs(want).max()
    assert emu.chan_emu_fft(3, None, None) == -1
