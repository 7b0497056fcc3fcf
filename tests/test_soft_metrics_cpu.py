"""The arithmetic argument behind soft_metrics_rel (osmo_gmr_b200/csrc/viterbi_tpc.cuh), runnable on the CPU.

osmo_conv_decode charges a soft bit x against a transmitted 0 / 1 with m0 = ((x - 127)^2) >> 9 and
m1 = ((x + 127)^2) >> 9, and nothing for an erased bit (x == 0) (oracle/shim).  The decode kernel
works with (m0, m1 - m0) and forms the second square from the first: (x + 127)^2 = (x - 127)^2 + 508 x.  This test
walks all 256 int8 inputs: the difference is exact, it is 0 for x == 0 without a special case, and the int8 negation
the gather program applies (-128 stays -128) swaps m0 and m1 for every input but -128.  The kernel code itself is
compared with the oracle bit for bit in test_decode_emu.py (CPU build) and test_decode_gpu.py."""
import numpy as np


def _ref(x):
    if x == 0:
        return 0, 0
    return ((x - 127) ** 2) >> 9, ((x + 127) ** 2) >> 9


def _rel(x):
    a = (x - 127) ** 2
    r0 = a >> 9
    assert a + 508 * x >= 0                      # the shift in the kernel is on a non-negative int
    d = ((a + 508 * x) >> 9) - r0
    return (r0 if x else 0), d


def test_relative_metric_identity_for_every_int8():
    for x in range(-128, 128):
        m0, m1 = _ref(x)
        r0, d = _rel(x)
        assert r0 == m0 and d == m1 - m0, x
        assert 0 <= m0 <= 127 and 0 <= m1 <= 127 and abs(d) <= 127


def test_int8_negation_swaps_the_metrics_except_for_minus_128():
    for x in range(-128, 128):
        nx = int(np.int8(np.int16(-x).astype(np.int8)))        # sbit_neg: (int8)(-x)
        if x == -128:
            assert nx == -128 and _ref(nx) == _ref(x)
        else:
            assert _ref(nx) == _ref(x)[::-1], x
