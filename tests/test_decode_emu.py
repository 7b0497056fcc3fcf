"""CPU-only: the product's decode code (compiled for the host by tests/emu) vs the oracle."""
import pytest

import decode_parity as dp
from decode_backends import CH, EmuBackend


@pytest.fixture(scope="module")
def be(emu):
    return EmuBackend(emu)


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
