"""oracle/chan_port.py - CPU oracle of the wideband channeliser (TEST INFRASTRUCTURE).

Nothing in the product imports this file; only tests/ and bench.py's checker legs do.

The reference's channeliser is a GNU Radio flowgraph (utils/gmr1_rx_sdr.py, "PFB Channelizer mode",
:391-604): PFBBase wires  rotator_cc -> pfb.arb_resampler_ccf (only when the sample rate is off the
channel grid) -> pfb.channelizer_ccf(n_chans, taps, oversample 2)  (:452-478), and one PFBOutputBranch
per ARFCN wires  delay -> pfb.arb_resampler_ccf(resamp, RRC 0.35 taps, flt_size 32) -> file_sink
(:587-599).  The script holds only the wiring and the filter specifications; the arithmetic is in
GNU Radio (gr-filter: firdes, pfb_channelizer_ccf, pfb_arb_resampler), a THIRD-PARTY dependency that
is not vendored under /root/reference, is not pinned by the reference (the script says only
`from gnuradio import ...`, Python 2 era => GNU Radio 3.7.x) and is not installed in this image.

PARITY UNPINNED: there is no reference output to compare with.  What follows restates the published
algorithms of those GNU Radio 3.7 blocks:
  * firdes::low_pass / root_raised_cosine / window (gr-filter/lib/firdes.cc),
  * pfb_channelizer_ccf: its defining sum (a critically aligned, 2x oversampled polyphase analysis
    bank; channel k centred at +k * fs / N, reverse FFT) evaluated DIRECTLY, sample by sample, so the
    GPU's polyphase + FFT factorisation is checked against an independent form,
  * pfb_arb_resampler kernel (gr-filter/lib/pfb_arb_resampler.cc: 32 phases, derivative filters,
    float accumulator `d_acc`, start phase (ntaps / 2) % 32).
Output sample alignment (a constant delay) follows the causal convention y[n] = sum_t h[t] x[n - t]
with zeros in front of the recording, which is what a GNU Radio block with history does at stream
start.  Parity tests anchor on what the reference's flowgraph is FOR: the channelised stream of every
ARFCN, fed to the reference's own C receive path (oracle/_ref), must decode to the L2 that was sent.
"""
import math

import numpy as np

CHAN_WIDTH = 31250.0          # GMR-1 carrier spacing, utils/gmr1_rx_sdr.py:51 (gmr1_dl_freq steps of 31.25 kHz)
SYM_RATE = 23400.0            # utils/gmr1_rx_sdr.py (sym_rate), 23.4 ksym/s
FLT_SIZE = 32                 # pfb.arb_resampler_ccf(..., flt_size = 32), gmr1_rx_sdr.py:594


# ---- gr::filter::firdes ---------------------------------------------------------------------------------------------
def _window_hamming(ntaps):
    m = ntaps - 1
    n = np.arange(ntaps, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * math.pi * n / m)).astype(np.float32)


def firdes_low_pass(gain, fs, fc, tw):
    """firdes::low_pass with the default Hamming window (gmr1_rx_sdr.py:433-438 passes no window).
    ntaps = (int)(53 * fs / (22 * tw)) made odd; taps sin(n w)/(n pi) * window, scaled to DC gain."""
    ntaps = int(53.0 * fs / (22.0 * tw))
    if (ntaps & 1) == 0:
        ntaps += 1
    w = _window_hamming(ntaps).astype(np.float64)
    m = (ntaps - 1) // 2
    fwt0 = 2.0 * math.pi * fc / fs
    taps = np.zeros(ntaps, np.float32)
    for n in range(-m, m + 1):
        if n == 0:
            taps[n + m] = np.float32(fwt0 / math.pi * w[n + m])
        else:
            taps[n + m] = np.float32(math.sin(n * fwt0) / (n * math.pi) * w[n + m])
    fmax = float(taps[m])
    for n in range(1, m + 1):
        fmax += 2.0 * float(taps[n + m])
    g = gain / fmax
    return (taps.astype(np.float64) * g).astype(np.float32)


def firdes_root_raised_cosine(gain, fs, sym_rate, alpha, ntaps):
    """firdes::root_raised_cosine (gmr1_rx_sdr.py:523-529: gain 32, fs 32 * chan_rate * 2, 0.35,
    11 symbols * 32 phases)."""
    ntaps |= 1
    spb = fs / sym_rate
    taps = np.zeros(ntaps, np.float64)
    scale = 0.0
    for i in range(ntaps):
        xindx = i - ntaps // 2
        x1 = math.pi * xindx / spb
        x2 = 4.0 * alpha * xindx / spb
        x3 = x2 * x2 - 1.0
        if abs(x3) >= 0.000001:
            if i != ntaps // 2:
                num = math.cos((1 + alpha) * x1) + math.sin((1 - alpha) * x1) / (4 * alpha * xindx / spb)
            else:
                num = math.cos((1 + alpha) * x1) + (1 - alpha) * math.pi / (4 * alpha)
            den = x3 * math.pi
        else:
            if alpha == 1:
                taps[i] = -1
                scale += taps[i]
                continue
            x3 = (1 - alpha) * x1
            x2 = (1 + alpha) * x1
            num = (math.sin(x2) * (1 + alpha) * math.pi
                   - math.cos(x3) * ((1 - alpha) * math.pi * spb) / (4 * alpha * xindx)
                   + math.sin(x3) * spb * spb / (4 * alpha * xindx * xindx))
            den = -32 * math.pi * alpha * alpha * xindx / spb
        taps[i] = 4 * alpha * num / den
        scale += taps[i]
    return (taps * gain / scale).astype(np.float32)


# ---- plan: what PFBBase / PFBOutputParameters compute (gmr1_rx_sdr.py:393-447, 501-529) -------------------------------
class Plan:
    """samp_rate must sit on the channel grid (n_chans * 31.25 kHz, n_chans even): the pre-resampler of
    gmr1_rx_sdr.py:455-462 is then absent (`self.resamp == 1`)."""

    def __init__(self, n_chans, sps=4):
        assert n_chans >= 2 and n_chans % 2 == 0
        self.n_chans = n_chans
        self.sps = sps
        self.samp_rate = n_chans * CHAN_WIDTH
        # :433-438 "Use a looser filter to reduce CPU"
        self.taps = firdes_low_pass(1.0, self.samp_rate, CHAN_WIDTH * 0.50, CHAN_WIDTH * 0.25)
        self.taps_per_branch = -(-len(self.taps) // n_chans)
        chan_rate = CHAN_WIDTH                      # width 1 (:513-517)
        oversample = 2                              # PFBOutputParameters.OVERSAMPLE
        self.mid_rate = chan_rate * oversample
        self.resamp = (SYM_RATE * sps) / self.mid_rate
        self.taps_resamp = firdes_root_raised_cosine(32.0, 32.0 * self.mid_rate, SYM_RATE, 0.35,
                                                     int(11.0 * 32 * self.mid_rate / SYM_RATE))
        # group delay of the chain in output samples (gmr1_rx_sdr.py min_delay(), :546-549, is the same idea)
        # (the resampler starts its walk at phase (ntaps / 2) % 32, i.e. that fraction of a bank step EARLY)
        self.delay_out = (((len(self.taps) - 1) / 2.0) / self.samp_rate
                          + ((len(self.taps_resamp) - 1) / 2.0 - (len(self.taps_resamp) // 2) % FLT_SIZE)
                          / (FLT_SIZE * self.mid_rate)) * SYM_RATE * sps


# ---- pfb.channelizer_ccf(n_chans, taps, oversample_rate = 2) ------------------------------------------------------------
def pfb_channelize_direct(x, taps, n_chans, chans, n_steps=None):
    """Defining sum of the 2x oversampled analysis bank, one output at a time:
         y_k[m] = sum_n h[n] x[m D - n] exp(-j 2 pi k (m D - n) / N),   D = N / 2,
    i.e. channel k (centre +k fs / N) mixed to DC, low-passed by h and kept every D samples.
    float64 accumulation; returns complex64 [len(chans), n_steps]."""
    N, D = n_chans, n_chans // 2
    x = np.asarray(x, np.complex128)
    h = np.asarray(taps, np.float64)
    M = len(x) // D if n_steps is None else n_steps
    out = np.zeros((len(chans), M), np.complex128)
    nn = np.arange(len(h))
    for m in range(M):
        idx = m * D - nn
        ok = (idx >= 0) & (idx < len(x))
        xs = np.where(ok, x[np.clip(idx, 0, len(x) - 1)], 0.0)
        for ci, k in enumerate(chans):
            ph = np.exp(-2j * math.pi * ((k * idx) % N) / N)
            out[ci, m] = np.sum(h * xs * ph)
    return out.astype(np.complex64)


def pfb_channelize(x, taps, n_chans, n_steps=None):
    """The same bank in its polyphase form (what pfb_channelizer_ccf computes: N branch filters of
    ceil(ntaps / N) taps, reverse FFT of size N per output step), all channels; [N, n_steps] complex64.
    Used for sizes where the direct sum is too slow; tests check the two agree."""
    N, D = n_chans, n_chans // 2
    x = np.asarray(x, np.complex64)
    P = -(-len(taps) // N)
    h = np.zeros(P * N, np.float64)
    h[:len(taps)] = taps
    M = len(x) // D if n_steps is None else n_steps
    pad = P * N + N
    xp = np.concatenate([np.zeros(pad, np.complex128), x.astype(np.complex128), np.zeros(N, np.complex128)])
    out = np.zeros((N, M), np.complex64)
    p = np.arange(N)
    sign = np.where(p % 2 == 1, -1.0, 1.0)
    for m in range(M):
        u = np.zeros(N, np.complex128)
        for q in range(P):
            u += h[p + q * N] * xp[pad + m * D - p - q * N]
        v = np.fft.ifft(u) * N                     # reverse (unnormalised) FFT: sum_p u_p e^{+j 2 pi k p / N}
        if m & 1:
            v = v * sign                           # e^{-j 2 pi k m D / N} = (-1)^(k m)
        out[:, m] = v
    return out


# ---- pfb.arb_resampler_ccf(rate, taps, flt_size = 32) -----------------------------------------------------------------
def arb_resampler_filters(taps, flt_size=FLT_SIZE):
    """Polyphase split of the prototype and of its first difference (pfb_arb_resampler::set_taps /
    create_diff_taps: diff[i] = taps[i + 1] - taps[i], last 0).  -> (filt [32, tpf], dfilt [32, tpf])"""
    taps = np.asarray(taps, np.float32)
    diff = np.zeros_like(taps)
    diff[:-1] = taps[1:] - taps[:-1]
    tpf = -(-len(taps) // flt_size)
    f = np.zeros(tpf * flt_size, np.float32)
    d = np.zeros(tpf * flt_size, np.float32)
    f[:len(taps)] = taps
    d[:len(taps)] = diff
    return f.reshape(tpf, flt_size).T.copy(), d.reshape(tpf, flt_size).T.copy()


def arb_resampler_schedule(rate, n_in, ntaps, flt_size=FLT_SIZE):
    """The phase walk of pfb_arb_resampler::filter, which is the same for every channel: for output n the input
    index i_in[n], the filter j[n] and the interpolation weight acc[n] (float32 arithmetic as in the block:
    d_acc += d_flt_rate; j += d_dec_rate + floor(d_acc); d_acc = fmodf(d_acc, 1)).  Starts at filter
    (ntaps / 2) % flt_size.  Stops when the input is used up."""
    dec_rate = int(math.floor(flt_size / rate))
    flt_rate = np.float32(flt_size / rate - dec_rate)
    acc = np.float32(0.0)
    j = (ntaps // 2) % flt_size
    i_in = 0
    ii, jj, aa = [], [], []
    while i_in < n_in:
        while j < flt_size:
            ii.append(i_in)
            jj.append(j)
            aa.append(acc)
            acc = np.float32(acc + flt_rate)
            j += dec_rate + int(math.floor(acc))
            acc = np.float32(math.fmod(acc, 1.0))
        i_in += j // flt_size
        j = j % flt_size
    return np.array(ii, np.int32), np.array(jj, np.int32), np.array(aa, np.float32)


def arb_resample(x, rate, taps, n_out=None, flt_size=FLT_SIZE):
    """pfb_arb_resampler_ccf on one stream: out[n] = f_j(x, i) + acc * df_j(x, i) with
    f_j(x, i) = sum_t filt[j][t] x[i - t] (zeros in front of the stream).  complex64."""
    filt, dfilt = arb_resampler_filters(taps, flt_size)
    tpf = filt.shape[1]
    ii, jj, aa = arb_resampler_schedule(rate, len(x), len(taps), flt_size)
    if n_out is not None:
        ii, jj, aa = ii[:n_out], jj[:n_out], aa[:n_out]
    xp = np.concatenate([np.zeros(tpf, np.complex64), np.asarray(x, np.complex64)])
    out = np.zeros(len(ii), np.complex64)
    t = np.arange(tpf)
    for n in range(len(ii)):
        seg = xp[tpf + ii[n] - t]
        o0 = np.sum(filt[jj[n]].astype(np.float64) * seg)
        o1 = np.sum(dfilt[jj[n]].astype(np.float64) * seg)
        out[n] = o0 + o1 * float(aa[n])
    return out


def channelize(x, plan, chans, n_out=None, direct=False):
    """Wideband recording -> one sps-oversampled stream per wanted channel, [len(chans), n_out] complex64:
    what gmr1_rx_sdr.py writes into the per-ARFCN files (without its start-up `delay` alignment block)."""
    if direct:
        mid = pfb_channelize_direct(x, plan.taps, plan.n_chans, chans)
    else:
        mid = pfb_channelize(x, plan.taps, plan.n_chans)[list(chans)]
    outs = [arb_resample(mid[i], plan.resamp, plan.taps_resamp, n_out) for i in range(len(chans))]
    n = min(len(o) for o in outs)
    return np.stack([o[:n] for o in outs])


# ---- a wideband test signal from per-channel streams (the inverse direction; test-side construction only) --------------
def synth_wideband(streams, chans, n_chans, sps=4, n_wide=None):
    """Per-channel complex streams at sps * 23.4 kS/s -> one wideband recording at n_chans * 31.25 kS/s: cubic
    (4-point Lagrange) interpolation to the wideband rate, mixed up to +k * 31.25 kHz and summed.  The GPU
    generator (gmr1b200_synth_wideband) does the same arithmetic."""
    streams = np.asarray(streams, np.complex64)
    num, den = 1872 * sps // 4, 625 * n_chans            # fs_ch / fs_wide = (23400 sps) / (31250 N), reduced by 12.5 x 4
    n_ch = streams.shape[1]
    if n_wide is None:
        n_wide = ((n_ch - 3) * den) // num
    t = np.arange(n_wide, dtype=np.int64)
    pos = t * num
    i = pos // den
    f = ((pos % den).astype(np.float64) / den).astype(np.float32)
    w0 = -f * (f - 1) * (f - 2) / 6
    w1 = (f + 1) * (f - 1) * (f - 2) / 2
    w2 = -(f + 1) * f * (f - 2) / 2
    w3 = (f + 1) * f * (f - 1) / 6
    out = np.zeros(n_wide, np.complex64)
    sp = np.concatenate([np.zeros((len(chans), 1), np.complex64), streams, np.zeros((len(chans), 3), np.complex64)], axis=1)
    for ci, k in enumerate(chans):
        s = sp[ci]
        v = w0 * s[i] + w1 * s[i + 1] + w2 * s[i + 2] + w3 * s[i + 3]
        ph = np.exp(2j * math.pi * ((k * t) % n_chans) / n_chans).astype(np.complex64)
        out += (v * ph).astype(np.complex64)
    return out
