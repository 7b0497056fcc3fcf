"""The demod kernel takes the soft bits of a data symbol from a table over floor(256 * sv) mod 256 * 2^nbits
(csrc/demod_kernels.cu, soft_word / soft_lut) instead of evaluating the reference's rule per symbol
(_gmr1_pi4cxpsk_soft_bits, src/sdr/pi4cxpsk.c:468-503).  That is exact because the rule is piecewise constant in the
symbol value sv with every breakpoint on a multiple of 1/256 and periodic with 2^nbits.  This test states the rule in
numpy (float32, as the C code computes it) and checks the claim: dense random symbol values give the same soft bits
through the table as through the rule, and values next to a breakpoint differ by at most one LSB.
(The CUDA code itself is covered by tests/test_demod_gpu.py; this is the arithmetic argument, kept runnable.)"""
import numpy as np
import pytest

GRAY = {1: np.array([[0], [1]], np.uint8),
        2: np.array([[0, 0], [0, 1], [1, 1], [1, 0]], np.uint8)}     # gmr1_pi2cbpsk / gmr1_pi4cqpsk symbol bits


def rule(sv, nbits):
    """pi4cxpsk.c:485-498 for an array of symbol values -> soft bits [n][nbits]"""
    sv = sv.astype(np.float32)
    mask = (1 << nbits) - 1
    svr = np.where(sv >= 0, np.floor(sv + np.float32(0.5)), np.ceil(sv - np.float32(0.5))).astype(np.float32)   # roundf
    sp = svr.astype(np.int64) & mask
    ss = np.where(svr > sv, sp - 1, sp + 1) & mask
    x = (np.float32(2.0) * np.abs(svr - sv)) * np.float32(64.0)
    d = np.floor(x + np.float32(0.5)).astype(np.int64)                                                           # roundf, x >= 0
    vp, vs = GRAY[nbits][sp], GRAY[nbits][ss]
    v = 127 - np.where(vp ^ vs, d[:, None], d[:, None] >> 1)
    return np.where(vp, -v, v).astype(np.int8)


def table(nbits):
    cells = 256 << nbits
    return rule((np.arange(cells, dtype=np.float32) + np.float32(0.5)) / np.float32(256.0), nbits)


@pytest.mark.parametrize("nbits", [1, 2])
def test_rule_is_constant_on_cells_of_1_256(nbits):
    rng = np.random.default_rng(nbits)
    tab = table(nbits)
    cells = 256 << nbits
    sv = rng.uniform(-3 * (1 << nbits), 3 * (1 << nbits), 2_000_000).astype(np.float32)
    s256 = sv.astype(np.float64) * 256.0
    inside = np.abs(s256 - np.rint(s256)) > 1e-3          # not on a breakpoint (float32 symbol values 1e-3 / 256 away)
    cell = np.floor(s256).astype(np.int64) % cells
    want = rule(sv, nbits)
    got = tab[cell]
    assert inside.mean() > 0.99
    assert (got[inside] == want[inside]).all()
    # next to a breakpoint the two may fall on different sides: never more than one LSB apart, same hard decision
    # unless the breakpoint is the symbol boundary itself
    diff = np.abs(got.astype(int) - want.astype(int))
    near = ~inside
    same_symbol = np.abs((s256[near] / 256.0) % 1.0 - 0.5) > 1e-3
    assert (diff[near][same_symbol] <= 1).all()


@pytest.mark.parametrize("nbits", [1, 2])
def test_every_breakpoint_is_a_multiple_of_1_256(nbits):
    """walk sv in steps of 1/4096 across one period and collect where the soft bits change"""
    per = 1 << nbits
    sv = (np.arange(per * 4096 + 1, dtype=np.float64) / 4096.0).astype(np.float32)
    out = rule(sv, nbits)
    changes = np.nonzero((out[1:] != out[:-1]).any(axis=1))[0] + 1          # first index of the new value
    # a change between samples i-1 and i means a breakpoint in ((i-1)/4096, i/4096]: it must contain a multiple of 1/256
    assert len(changes) > 100
    assert ((changes % 16 == 0) | ((changes - 1) % 16 == 0)).all()
