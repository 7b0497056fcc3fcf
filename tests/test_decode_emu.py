"""CPU-only: the product's decode code (compiled for the host by tests/emu) vs the oracle."""
import pytest

import decode_parity as dp
from decode_backends import CH, EmuBackend


@pytest.fixture(scope="module", params=["one-per-thread", "one-per-thread-metric-table", "two-per-thread-16-bit"])
def be(emu, request):
    return EmuBackend(emu, p16=request.param == "two-per-thread-16-bit", lut=request.param == "one-per-thread-metric-table")


def test_bcch(be, oracle):
    dp.check_simple(be, oracle, "bcch", CH["BCCH"], 424, 96, 11)


def test_ccch(be, oracle):
    dp.check_simple(be, oracle, "ccch", CH["CCCH"], 432, 96, 12)


@pytest.mark.parametrize("use_ciph", [False, True])
def test_facch3(be, oracle, use_ciph):
    dp.check_facch3(be, oracle, 64, 13, use_ciph)


@pytest.mark.parametrize("use_ciph", [False, True])
def test_facch9(be, oracle, use_ciph):
    dp.check_facch9(be, oracle, 48, 14, use_ciph)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("use_ciph", [False, True])
def test_tch9(be, oracle, mode, use_ciph):
    dp.check_tch9(be, oracle, mode, 4, 7, 15 + mode, use_ciph)


def test_rach(be, oracle):
    dp.check_rach(be, oracle, 64, 16)


@pytest.mark.parametrize("m", [0, 1])
@pytest.mark.parametrize("use_ciph", [False, True])
def test_tch3(be, oracle, use_ciph, m):
    dp.check_tch3(be, oracle, 64, 17, use_ciph, m)


@pytest.mark.parametrize("ch,n_in", [("BCCH", 424), ("CCCH", 432), ("FACCH3", 416), ("FACCH9", 662), ("RACH", 494),
                                     ("TCH3", 212)])
@pytest.mark.parametrize("kind", ["uniform", "extremes", "near-zero"])
def test_p16_equals_one_per_thread_on_hostile_input(emu, ch, n_in, kind):
    """The 16-bit packed form against the 32-bit form on inputs no demodulator produces: full-range noise, only
    -128 / 127 / 0 (largest metric growth and many ties), and values around 0 (smallest decision margins); odd batch
    (the last thread has no second codeword).  L2, CRC and the Viterbi metric must be identical."""
    import numpy as np
    rng = np.random.default_rng(sum(map(ord, ch + kind)))
    n = 37
    if kind == "uniform":
        e = rng.integers(-128, 128, (n, n_in)).astype(np.int8)
    elif kind == "extremes":
        e = rng.choice(np.array([-128, -127, 127, 0], np.int8), (n, n_in))
    else:
        e = rng.integers(-2, 3, (n, n_in)).astype(np.int8)
    a = EmuBackend(emu).decode(CH[ch], e)
    b = EmuBackend(emu, p16=True).decode(CH[ch], e)
    c = EmuBackend(emu, lut=True).decode(CH[ch], e)
    for k in a:
        assert (a[k] == b[k]).all(), (ch, kind, k)
        assert (a[k] == c[k]).all(), (ch, kind, k, "metric table")
