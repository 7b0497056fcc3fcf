"""GPU parity: CUDA channel-decode kernels through the C ABI vs the oracle, bit-exact.
Sizes straddle the 128-codeword CTA tile so both the TMA bulk path (full tiles) and the
cooperative tail path are exercised; host-pointer and device-pointer entry are both covered."""
import pytest

import decode_parity as dp
from decode_backends import CH, GpuBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=[False, True], ids=["hostptr", "devptr"])
def be(request, gpu_lib):
    return GpuBackend(gpu_lib, device=request.param)


def test_bcch(be, oracle):
    dp.check_simple(be, oracle, "bcch", CH["BCCH"], 424, 300, 21)


def test_ccch(be, oracle):
    dp.check_simple(be, oracle, "ccch", CH["CCCH"], 432, 257, 22)


@pytest.mark.parametrize("use_ciph", [False, True])
def test_facch3(be, oracle, use_ciph):
    dp.check_facch3(be, oracle, 200, 23, use_ciph)


@pytest.mark.parametrize("use_ciph", [False, True])
def test_facch9(be, oracle, use_ciph):
    dp.check_facch9(be, oracle, 150, 24, use_ciph)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("use_ciph", [False, True])
def test_tch9(be, oracle, mode, use_ciph):
    dp.check_tch9(be, oracle, mode, 20, 7, 25 + mode, use_ciph)


def test_rach(be, oracle):
    dp.check_rach(be, oracle, 200, 26)


@pytest.mark.parametrize("m", [0, 1])
@pytest.mark.parametrize("use_ciph", [False, True])
def test_tch3(be, oracle, use_ciph, m):
    dp.check_tch3(be, oracle, 200, 27, use_ciph, m)


def test_dc12(be, oracle):
    dp.check_dc12(be, oracle, 70, 28)


def test_empty_and_single(gpu_lib, oracle):
    import numpy as np
    e = np.zeros((0, 424), np.int8)
    l2 = np.zeros((0, 24), np.uint8)
    gpu_lib.call("gmr1b200_bcch_decode_batch", l2, e, None, None, 0, None)   # n = 0 is a no-op
    be = GpuBackend(gpu_lib)
    dp.check_simple(be, oracle, "bcch", CH["BCCH"], 424, 9, 29)               # < one tile
