"""numpy pi/4-CxPSK burst generator for tests (TEST INFRASTRUCTURE).

The reference has no general modulator (its gmr1_pi4cxpsk_mod is 1 sample/symbol, no pulse,
src/sdr/pi4cxpsk.c:741-799), so test signals are synthesised here: symbol mapping per the burst
descriptors (src/sdr/nb.c), pi/4 (pi/2) continuous rotation, raised-cosine pulse (RRC 0.35 at
the transmitter x RRC 0.35 matched filter, which is what gmr1_rx expects its input to have been
through, utils/gmr1_rx_sdr.py:523-529), fractional timing offset, carrier frequency offset,
phase and AWGN.  It is validated by the oracle demodulating and decoding what it produces.
"""
import numpy as np

# burst descriptors (facts of ETSI TS 101 376-5-2 section 7.4, as in the reference's nb.c):
# name: (rotation divisor, bits/symbol, symbols, ebits, [sync sequences: [(pos, "symbols")]], [(data pos, len)])
BURSTS = {
    "bcch": (4, 2, 234, 424, [[(28, "02200020222"), (119, "220"), (197, "220")]],
             [(2, 26), (39, 80), (122, 75), (200, 31)]),
    "dc2": (4, 2, 78, 132, [[(28, "0123030")]], [(2, 26), (35, 40)]),
    "dc6": (4, 2, 234, 432, [[(28, "0002202"), (119, "030"), (197, "311")]],
            [(2, 26), (35, 84), (122, 75), (200, 31)]),
    "dc12": (2, 1, 468, 432, [[(10, "0010001111"), (228, "00100011101"), (447, "0010001111")]],
             [(2, 8), (20, 208), (239, 208), (457, 8)]),
    "nt3_speech": (4, 2, 117, 212, [[(28, "033123")]], [(2, 26), (34, 80)]),
    "nt3_facch": (4, 1, 117, 104, [[(28, "10101010")], [(28, "11001001")]], [(2, 26), (36, 78)]),
    "nt6": (4, 2, 234, 434, [[(28, "022323"), (119, "010"), (197, "230")],
                             [(28, "000220"), (119, "130"), (197, "213")]],
            [(2, 26), (34, 85), (122, 75), (200, 31)]),
    "nt9": (4, 2, 351, 662, [[(28, "022323"), (119, "122"), (197, "010"), (275, "230")],
                             [(28, "000220"), (119, "020"), (197, "130"), (275, "213")]],
            [(2, 26), (34, 85), (122, 75), (200, 75), (278, 70)]),
    "rach": (4, 2, 351, 494, [[(78, "02200020222220220"), (127, "2" * 32), (191, "2" * 32),
                               (255, "02200020222220220"), (347, "0")]],
             [(2, 76), (95, 32), (159, 32), (223, 32), (272, 75)]),
    "sdcch": (4, 1, 234, 208, [[(28, "0101010"), (115, "1010101"), (197, "0101011")],
                               [(28, "0011001"), (115, "1001100"), (197, "1100111")],
                               [(28, "0000111"), (115, "1000011"), (197, "1100001")],
                               [(28, "0110100"), (115, "1011010"), (197, "0101101")]],
              [(2, 26), (35, 80), (122, 75), (204, 27)]),
}
BT_ID = {n: i for i, n in enumerate(["bcch", "dc2", "dc6", "dc12", "nt3_speech", "nt3_facch", "nt6", "nt9",
                                     "rach", "sdcch"])}


def burst_len(name):
    return BURSTS[name][2]


def burst_ebits(name):
    return BURSTS[name][3]


def symbols(name, ebits, sync_id=0):
    """hard bits [n, ebits] -> complex symbols [n, len] incl. continuous rotation"""
    rdiv, nbits, ln, neb, syncs, data = BURSTS[name]
    ebits = np.atleast_2d(ebits)
    n = ebits.shape[0]
    ph = np.full((n, ln), -1, np.int64)           # phase index (x pi/2), -1 = guard (no energy)
    for pos, seq in syncs[sync_id]:
        v = np.array([int(c) for c in seq])
        ph[:, pos:pos + len(seq)] = (2 * v if nbits == 1 else v)[None, :]
    k = 0
    for pos, ln_c in data:
        b = ebits[:, k:k + ln_c * nbits].reshape(n, ln_c, nbits).astype(np.int64)
        if nbits == 2:
            sym = np.array([0, 1, 3, 2])[(b[:, :, 0] << 1) | b[:, :, 1]]    # Gray: 00 01 11 10
        else:
            sym = 2 * b[:, :, 0]
        ph[:, pos:pos + ln_c] = sym
        k += ln_c * nbits
    s = np.where(ph >= 0, np.exp(1j * (np.pi / 2) * np.maximum(ph, 0)), 0)
    s = s * np.exp(1j * (np.pi / rdiv) * np.arange(ln))[None, :]
    return s


def raised_cosine(t, alpha=0.35):
    t = np.asarray(t, np.float64)
    den = 1.0 - (2.0 * alpha * t) ** 2
    sing = np.abs(den) < 1e-9
    out = np.sinc(t) * np.cos(np.pi * alpha * t) / np.where(sing, 1.0, den)
    return np.where(sing, (np.pi / 4) * np.sinc(1.0 / (2 * alpha)), out)


def modulate(name, ebits, sps, win, toa, cfo, phase, esn0_db, rng, sync_id=0, amp=1.0):
    """-> complex64 windows [n, len*sps + win].
    toa: position (samples, fractional) of symbol 0 within the window;  cfo: rad/symbol;
    phase: rad;  esn0_db: per burst (scalar or [n])."""
    return modulate_symbols(symbols(name, ebits, sync_id), sps, win, toa, cfo, phase, esn0_db, rng, amp)


def modulate_symbols(s, sps, win, toa, cfo, phase, esn0_db, rng, amp=1.0):
    """pulse-shape arbitrary complex symbol rows [n, len] (0 = silent symbol)"""
    s = np.atleast_2d(s)
    n, ln = s.shape
    L = ln * sps + win
    toa = np.broadcast_to(np.asarray(toa, np.float64), (n,))[:, None]
    cfo = np.broadcast_to(np.asarray(cfo, np.float64), (n,))[:, None]
    phase = np.broadcast_to(np.asarray(phase, np.float64), (n,))[:, None]
    t = (np.arange(L)[None, :] - toa) / sps                 # time in symbols
    k0 = np.floor(t).astype(np.int64)
    x = np.zeros((n, L), np.complex128)
    rows = np.arange(n)[:, None]
    for j in range(-5, 7):
        k = k0 + j
        ok = (k >= 0) & (k < ln)
        x += np.where(ok, s[rows, np.clip(k, 0, ln - 1)], 0) * raised_cosine(t - k)
    x *= amp * np.exp(1j * (cfo * t + phase))
    sig = np.broadcast_to(amp * 10.0 ** (-np.asarray(esn0_db, np.float64) / 20.0) / np.sqrt(2.0), (n,))[:, None]
    x += sig * (rng.standard_normal((n, L)) + 1j * rng.standard_normal((n, L)))
    return x.astype(np.complex64)


def dkab_symbols(n, p, rng):
    """DKAB: 117-symbol slot, silent except two 5-symbol keep-alive pulses at symbols 2+p and
    2+p+59 (reference src/sdr/dkab.c:73-75), random BPSK with the continuous pi/4 rotation"""
    s = np.zeros((n, 117), np.complex128)
    for o in (2 + p, 2 + p + 59):
        s[:, o:o + 5] = 1.0 - 2.0 * rng.integers(0, 2, (n, 5))
    return s * np.exp(1j * (np.pi / 4) * np.arange(117))[None, :]


def fcch_chirp(sps, freq=0.32, ln=117):
    """real dual chirp sqrt(2) cos(freq*2pi/len * (t - len/2)^2), t in symbols (fcch.c:167-193)"""
    t = np.arange(ln * sps) / sps - ln / 2.0
    return np.sqrt(2.0) * np.cos(freq * 2 * np.pi / ln * t * t)


def fcch_window(L, sps, pos, cfo, esn0_db, rng, freq=0.32, ln=117, amp=1.0):
    """L-sample window of noise with one FCCH burst starting at integer sample `pos`, carrier
    offset cfo (rad/symbol)"""
    sig = amp * 10.0 ** (-esn0_db / 20.0) / np.sqrt(2.0)
    x = sig * (rng.standard_normal(L) + 1j * rng.standard_normal(L))
    c = fcch_chirp(sps, freq, ln)
    n = np.arange(len(c))
    x[pos:pos + len(c)] += amp * c * np.exp(1j * (cfo * n / sps + rng.uniform(0, 2 * np.pi)))
    return x.astype(np.complex64)
