"""CPU: the bitsliced A5/1 core (osmo_gmr_b200/csrc/a5_bitslice.cuh - 32 keystreams per "thread", the code
a5_slice_kernel inlines) compiled for the host by tests/emu/chan_emu.cpp, against the reference's gmr1_a5
(src/l1/a5.c:57-282, oracle/_ref or the port): downlink and uplink streams of 32 random (Kc, fn) pairs, the stream
lengths the channels use (TCH3 208, FACCH3 96, TCH9 658) and one that ends inside a 32-bit block."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    emu_dir = os.path.join(ROOT, "tests", "emu")
    so, src = os.path.join(emu_dir, "libchan_emu.so"), os.path.join(emu_dir, "chan_emu.cpp")
    csrc = os.path.join(ROOT, "osmo_gmr_b200", "csrc")
    deps = [src, os.path.join(csrc, "chan_fft.cuh"), os.path.join(csrc, "a5_bitslice.cuh")]
    if not os.path.exists(so) or max(os.path.getmtime(d) for d in deps) > os.path.getmtime(so):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I" + csrc, "-o", so, src])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("nbits", [208, 96, 658, 5, 33])
def test_bitsliced_a5_equals_the_reference(emu, oracle, nbits):
    rng = np.random.default_rng(nbits)
    keys = rng.integers(0, 256, (32, 8), dtype=np.uint8)
    keys[0], keys[1] = 0, 255
    fn = rng.integers(0, 1 << 19, 32).astype(np.uint32)
    fn[:3] = [0, (1 << 19) - 1, 0x5a5a5]
    dl, ul = np.zeros((32, nbits), np.uint8), np.zeros((32, nbits), np.uint8)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    emu.a5_emu_slice(p(keys), p(fn), nbits, p(dl), p(ul))
    for i in range(32):
        d, u = oracle.a5(1, keys[i], int(fn[i]), nbits, both=True)
        assert (d == dl[i]).all() and (u == ul[i]).all(), i
    only = np.zeros((32, nbits), np.uint8)
    emu.a5_emu_slice(p(keys), p(fn), nbits, p(only), None)
    assert (only == dl).all()
