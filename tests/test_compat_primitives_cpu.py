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


class CxVec(ctypes.Structure):
    _fields_ = [("len", ctypes.c_int), ("max_len", ctypes.c_int), ("flags", ctypes.c_int), ("data", P)]


BURSTS = {"bcch": (424, 1), "dc2": (132, 1), "dc6": (432, 1), "dc12": (432, 1), "nt3_speech": (212, 1), "nt3_facch": (104, 2),
          "nt6": (434, 2), "nt9": (662, 2), "rach": (494, 1), "sdcch": (208, 4)}


def test_pi4cxpsk_mod(libs):
    """gmr1_pi4cxpsk_mod (src/sdr/pi4cxpsk.c:741-800) over every exported burst descriptor and sync sequence: the same
    1-sample-per-symbol burst as the reference (descriptor data symbols, Gray map, rotation)"""
    rng = np.random.default_rng(5)
    for name, (ebits, n_sync) in BURSTS.items():
        hard = rng.integers(0, 2, ebits).astype(np.uint8)
        for sid in range(n_sync):
            res = []
            for lib in libs:
                desc = ctypes.c_char.in_dll(lib, f"gmr1_{name}_burst")
                buf = np.zeros(1024, np.complex64)
                cv = CxVec(0, 1024, 0, buf.ctypes.data_as(P))
                rc = lib.gmr1_pi4cxpsk_mod(ctypes.c_void_p(ctypes.addressof(desc)), p(hard), sid, ctypes.byref(cv))
                assert rc == 0, (name, rc)
                res.append(buf[:cv.len].copy())
            assert res[0].shape == res[1].shape and len(res[0]) > 70, name
            assert np.abs(res[0] - res[1]).max() <= 1e-6, (name, sid)
