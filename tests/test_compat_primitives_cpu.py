"""The host-side primitives libgmr1_b200.so exports under the reference's names (csrc/compat.c, a5.cpp, encode.cpp):
gmr1_scramble_sbit / _ubit (src/l1/scramb.c:64,82), gmr1_(de)interleave_intra (interleave.c:49,74), the stateful
inter-burst (de)interleaver (interleave.c:95-185), gmr1_a5 / gmr1_a5_1 (a5.c:57,226) - against the reference build on
random inputs, byte for byte.  Pure host code: no GPU involved."""
import ctypes
import os

import numpy as np
import pytest

import osmo_gmr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libgmr1_ref.so")
P = ctypes.c_void_p


class IL(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("K", ctypes.c_int), ("n", ctypes.c_int), ("bits", P)]


def p(a):
    return a.ctypes.data_as(P)


@pytest.fixture(scope="module")
def libs():
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref not built")
    return ctypes.CDLL(osmo_gmr_b200.LIB_PATH), ctypes.CDLL(REF)


def test_scrambler(libs):
    rng = np.random.default_rng(1)
    for ln in (1, 16, 96, 424, 648, 1000):
        s_in = rng.integers(-127, 128, ln).astype(np.int8)
        u_in = rng.integers(0, 2, ln).astype(np.uint8)
        outs = []
        for lib in libs:
            so, uo = np.zeros(ln, np.int8), np.zeros(ln, np.uint8)
            lib.gmr1_scramble_sbit(p(so), p(s_in), ln)
            lib.gmr1_scramble_ubit(p(uo), p(u_in), ln)
            outs.append((so, uo))
        assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()


def test_intra_interleaver(libs):
    rng = np.random.default_rng(2)
    for N in (1, 12, 14, 33, 53, 54, 80, 81):
        x = rng.integers(0, 256, 8 * N).astype(np.uint8)
        res = []
        for lib in libs:
            a, b = np.zeros(8 * N, np.uint8), np.zeros(8 * N, np.uint8)
            lib.gmr1_interleave_intra(p(a), p(x), N)
            lib.gmr1_deinterleave_intra(p(b), p(a), N)
            res.append((a, b))
        assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all() and (res[0][1] == x).all()


def test_inter_interleaver(libs):
    rng = np.random.default_rng(3)
    for N, K in ((3, 648), (2, 96), (4, 40)):
        bursts = rng.integers(0, 2, (7, K)).astype(np.uint8)
        res = []
        for lib in libs:
            tx, rx = IL(), IL()
            assert lib.gmr1_interleaver_init(ctypes.byref(tx), N, K) == 0
            assert lib.gmr1_interleaver_init(ctypes.byref(rx), N, K) == 0
            outs = []
            for b in bursts:
                epp, back = np.zeros(K, np.uint8), np.zeros(K, np.uint8)
                lib.gmr1_interleave_inter(ctypes.byref(tx), p(epp), p(np.ascontiguousarray(b)))
                lib.gmr1_deinterleave_inter(ctypes.byref(rx), p(back), p(epp))
                outs.append((epp, back))
            lib.gmr1_interleaver_fini(ctypes.byref(tx))
            lib.gmr1_interleaver_fini(ctypes.byref(rx))
            res.append(outs)
        for (a0, a1), (b0, b1) in zip(*res):
            assert (a0 == b0).all() and (a1 == b1).all()


def test_a5(libs):
    rng = np.random.default_rng(4)
    for nbits in (96, 208, 658):
        for _ in range(6):
            key = rng.integers(0, 256, 8).astype(np.uint8)
            fn = int(rng.integers(0, 1 << 19))
            res = []
            for lib in libs:
                dl, ul = np.zeros(nbits, np.uint8), np.zeros(nbits, np.uint8)
                lib.gmr1_a5(1, p(key.copy()), ctypes.c_uint32(fn), nbits, p(dl), p(ul))
                d1 = np.zeros(nbits, np.uint8)
                lib.gmr1_a5_1(p(key.copy()), ctypes.c_uint32(fn), nbits, p(d1), None)
                d0 = np.ones(nbits, np.uint8)
                lib.gmr1_a5(0, p(key.copy()), ctypes.c_uint32(fn), nbits, p(d0), None)
                res.append((dl, ul, d1, d0))
            for a, b in zip(*res):
                assert (a == b).all()
            assert (res[0][0] == res[0][2]).all() and not res[0][3].any()
