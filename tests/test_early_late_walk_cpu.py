"""The demod kernel runs the early/late peak search of osmo_cxvec_peak_energy_find(PEAK_EARLY_LATE) (libosmo-dsp as
restated in oracle/shim/shim_dsp.c:251-281; called from src/sdr/pi4cxpsk.c:240) three bisection steps at a time:
after the first step (which sits on an integer position) steps 2-4 can only visit early + m/8, m in {0, +-1, +-2, +-3},
steps 5-7 the same on a grid of 1/64, steps 8-9 on 1/512; each round evaluates its grid in parallel and then walks
the decisions (csrc/demod_kernels.cu, peak_early_late).  This test restates both procedures over the same comparison
function and checks that the walk visits exactly the positions the sequential search visits: identical result for
every input, including the ties that end the search early.  (The CUDA code itself is covered by
tests/test_demod_gpu.py; this is the combinatorial argument, kept runnable.)"""
import numpy as np


def interp(acc, pos):
    """osmo_cxvec_interpolate_point: 10 taps either side, samples outside the vector count as 0"""
    b = int(np.floor(pos)) - 10
    i = np.arange(max(b, 0), min(b + 21, len(acc)))
    x = np.pi * (i - pos)
    w = np.where(np.abs(x) < 0.01, 1.0, np.sin(x) / np.where(x == 0, 1.0, x))
    return float((acc[i] * w).sum())


def compare(acc, early, quant):
    """-1: early gate stronger (move left), +1: late gate stronger (move right), 0: tie (search ends)"""
    e, l = interp(acc, early) ** 2, interp(acc, early + 2.0) ** 2
    if quant:                      # coarse values make exact ties frequent
        e, l = round(e, quant), round(l, quant)
    return (e < l) - (e > l)


def sequential(acc, mwi, quant):
    early, incr = float(mwi - 1), 0.5
    visited = []
    while incr > 1.0 / 1024.0:
        visited.append(early)
        c = compare(acc, early, quant)
        if c == 0:
            break
        early += c * incr
        incr /= 2.0
    return early, visited


def radix8(acc, mwi, quant):
    early = float(mwi - 1)
    visited = [early]
    c = compare(acc, early, quant)
    if c == 0:
        return early, visited
    early += 0.5 * c
    h = 0.125
    for rnd in range(3):
        grid = {m: compare(acc, early + m * h, quant) for m in range(-3, 4)}      # what the 8 quads evaluate
        m, half, live = 0, 0.0, True
        visited.append(early)
        if grid[0] == 0:
            live = False
        else:
            m = 2 * grid[0]
            visited.append(early + m * h)
            if grid[m] == 0:
                live = False
            else:
                m += grid[m]
                if rnd < 2:
                    visited.append(early + m * h)
                    if grid[m] == 0:
                        live = False
                    else:
                        half = 0.5 * grid[m]
        early += (m + half) * h
        h *= 0.125
        if not live:
            break
    return early, visited


def test_walk_equals_sequential_search():
    rng = np.random.default_rng(11)
    n_tie = 0
    for trial in range(3000):
        w = int(rng.integers(7, 90))
        peak = rng.uniform(1.0, w - 2.0)
        k = np.arange(w)
        width = rng.uniform(0.6, 3.0)
        acc = np.sinc((k - peak) / width) ** 2 * rng.uniform(5, 50) + np.abs(rng.normal(0, rng.uniform(0.0, 2.0), w))
        if trial % 5 == 0:
            acc = np.round(acc)                      # plateaus: ties in the first steps
        quant = 2 if trial % 3 == 0 else 0
        mwi = int(np.argmax(acc))
        a, va = sequential(acc, mwi, quant)
        b, vb = radix8(acc, mwi, quant)
        assert a == b and va == vb, (trial, a, b)
        n_tie += len(va) < 9
    assert n_tie > 20                                  # the early exits were exercised
