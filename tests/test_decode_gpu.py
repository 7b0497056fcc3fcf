"""GPU parity: CUDA channel-decode kernels through the C ABI vs the oracle, bit-exact.
Sizes straddle the 128-codeword CTA tile so both the TMA bulk path (full tiles) and the
cooperative tail path are exercised; host-pointer and device-pointer entry are both covered."""
import numpy as np
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


@pytest.mark.parametrize("off", [1, 2, 8])
def test_ciphered_channels_unaligned_inputs(gpu_lib, oracle, off):
    """soft bits and cipher bytes that do not start on a 16-byte boundary: the tile comes in by the plain copy instead
    of the bulk copy, the cipher pass takes byte loads where its words are not aligned; full tiles + a ragged one"""
    be = GpuBackend(gpu_lib, device=True, misalign=off)
    dp.check_tch3(be, oracle, 128 * 2 + 37, 31, True, 0)
    dp.check_facch3(be, oracle, 128 + 5, 32, True)
    dp.check_facch9(be, oracle, 32 * 3 + 7, 33, True)
    dp.check_simple(be, oracle, "bcch", CH["BCCH"], 424, 128 + 9, 34)


def test_dc12(be, oracle):
    dp.check_dc12(be, oracle, 70, 28)


def test_empty_and_single(gpu_lib, oracle):
    import numpy as np
    e = np.zeros((0, 424), np.int8)
    l2 = np.zeros((0, 24), np.uint8)
    gpu_lib.call("gmr1b200_bcch_decode_batch", l2, e, None, None, 0, None)   # n = 0 is a no-op
    be = GpuBackend(gpu_lib)
    dp.check_simple(be, oracle, "bcch", CH["BCCH"], 424, 9, 29)               # < one tile


@pytest.fixture(params=["thread", "bitsliced"])
def a5_kernel(request, gpu_lib):
    """the cipher streams from the one-unit-per-thread kernel and from the bitsliced one (32 units per thread,
    csrc/a5_bitslice.cuh; the default from 163 840 units up)"""
    prev = gpu_lib.c.gmr1b200_set_a5_bitslice(1 if request.param == "bitsliced" else 0)
    yield request.param
    gpu_lib.c.gmr1b200_set_a5_bitslice(prev)


@pytest.mark.parametrize("nbits,stride", [(208, 208), (658, 658), (658, 660), (96, 100), (5, 7)])
def test_a5_batch(gpu_lib, oracle, nbits, stride, a5_kernel):
    """gmr1b200_a5_batch vs gmr1_a5 (src/l1/a5.c:57): downlink and uplink streams, A5/0 and A5/1 units mixed,
    word-aligned and odd row strides, host and device pointers; batch sizes that end inside a warp's 1024 units"""
    rng = np.random.default_rng(nbits * 7 + stride)
    n = 300 if nbits != 208 else 1024 + 37
    keys = rng.integers(0, 256, (n, 8), dtype=np.uint8)
    keys[0] = 0
    keys[1] = 255
    fn = rng.integers(0, 1 << 19, n).astype(np.uint32)
    fn[:3] = [0, (1 << 19) - 1, 0x5a5a5]
    alg = rng.integers(0, 2, n).astype(np.int32)
    alg[:3] = 1
    dl = np.full((n, stride), 9, np.uint8)
    ul = np.full((n, stride), 9, np.uint8)
    gpu_lib.call("gmr1b200_a5_batch", alg, 0, keys, fn, nbits, stride, dl, ul, n, None)
    for i in range(n):
        d, u = oracle.a5(int(alg[i]), keys[i], int(fn[i]), nbits, both=True)
        assert (dl[i, :nbits] == d).all() and (ul[i, :nbits] == u).all(), i
    # downlink only, one algorithm for all, device memory
    import torch
    dk, dfn = torch.from_numpy(keys).cuda(), torch.from_numpy(fn.view(np.int32)).cuda()
    ddl = torch.zeros((n, stride), dtype=torch.uint8, device="cuda")
    gpu_lib.call("gmr1b200_a5_batch", None, 1, dk, dfn, nbits, stride, ddl, None, n, None)
    torch.cuda.synchronize()
    got = ddl.cpu().numpy()
    for i in range(0, n, 17):
        assert (got[i, :nbits] == oracle.a5(1, keys[i], int(fn[i]), nbits)).all(), i


def test_a5_large_batch_both_kernels_agree(gpu_lib, oracle):
    """40 000 units: both kernels (and the size-based default) give the same streams, spot-checked against gmr1_a5"""
    import torch
    rng = np.random.default_rng(77)
    n, nbits = 40000, 208
    keys = torch.from_numpy(rng.integers(0, 256, (n, 8), dtype=np.uint8)).cuda()
    fn = torch.from_numpy(rng.integers(0, 1 << 19, n).astype(np.int32)).cuda()
    out = []
    for mode in (0, 1, -1):
        prev = gpu_lib.c.gmr1b200_set_a5_bitslice(mode)
        dl = torch.full((n, nbits), 7, dtype=torch.uint8, device="cuda")
        ul = torch.full((n, nbits), 7, dtype=torch.uint8, device="cuda")
        gpu_lib.call("gmr1b200_a5_batch", None, 1, keys, fn, nbits, nbits, dl, ul, n, None)
        torch.cuda.synchronize()
        gpu_lib.c.gmr1b200_set_a5_bitslice(prev)
        out.append((dl.cpu().numpy(), ul.cpu().numpy()))
    assert all((out[0][0] == o[0]).all() and (out[0][1] == o[1]).all() for o in out[1:])
    k, f = keys.cpu().numpy(), fn.cpu().numpy()
    for i in (0, 1, 1023, 1024, 33333, n - 1):
        d, u = oracle.a5(1, k[i], int(f[i]), nbits, both=True)
        assert (out[1][0][i] == d).all() and (out[1][1][i] == u).all(), i
