"""CPU: pins of the oracle.

(1) the oracle port (oracle/liboracle.so, our plain-C restatement) == the reference's own C sources
    (oracle/_ref, built from /root/reference when present) on identical inputs, bit for bit;
(2) whichever oracle is loaded reproduces the committed golden fixtures (tests/golden/golden.npz,
    generated from oracle/_ref by tests/golden/make_golden.py);
(3) known facts of the shim / primitives: scrambler prefix, puncturing
    budgets, CRC16 against an independent implementation, the shim DFT against numpy.fft,
    encode -> decode round trips with CRC = 0.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
import sigen
import vectors

ROOT = oracle_lib.ROOT
G = np.load(os.path.join(ROOT, "tests", "golden", "golden.npz"))


@pytest.fixture(scope="module")
def port():
    so = os.path.join(ROOT, "oracle", "liboracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    return oracle_lib.Oracle(so, "port")


@pytest.fixture(scope="module")
def ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libgmr1_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    return oracle_lib.Oracle(so, "reference")


# ------------------------------------------------------------------ (1) port == reference
def test_port_equals_reference_l1(port, ref):
    rng = np.random.default_rng(5)
    for name, nbits in (("bcch", 424), ("ccch", 432), ("xch_dc12", 432)):
        l2 = rng.integers(0, 256, 24, dtype=np.uint8)
        assert (port.encode(name, nbits, l2) == ref.encode(name, nbits, l2)).all()
        e = vectors.simple(ref, rng, name, nbits, 40)
        for i in range(40):
            a, b = port.simple_decode(name, e[i]), ref.simple_decode(name, e[i])
            assert (a[0] == b[0]).all() and a[1:] == b[1:], (name, i)
    e, ciph = vectors.facch3(ref, rng, 24, True)
    for i in range(24):
        a, b = port.facch3_decode(e[i], ciph[i]), ref.facch3_decode(e[i], ciph[i])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    e, ciph = vectors.facch9(ref, rng, 16, True)
    for i in range(16):
        a, b = port.facch9_decode(e[i], ciph[i]), ref.facch9_decode(e[i], ciph[i])
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    for mode in (0, 1, 2):
        e, ciph, _, _ = vectors.tch9(ref, rng, mode, 2, 6, True)
        for c in range(2):
            ia, ib = port.interleaver(), ref.interleaver()
            for k in range(6):
                a = port.tch9_decode(e[c * 6 + k], mode, ciph[c * 6 + k], ia)
                b = ref.tch9_decode(e[c * 6 + k], mode, ciph[c * 6 + k], ib)
                assert all(np.array_equal(x, y) for x, y in zip(a, b))
    e, masks = vectors.rach(ref, rng, 24)
    for i in range(24):
        a, b = port.rach_decode(e[i], masks[i]), ref.rach_decode(e[i], masks[i])
        assert (a[0] == b[0]).all() and a[1:] == b[1:]
    e, ciph = vectors.tch3(rng, 24, True)
    for i in range(24):
        a, b = port.tch3_decode(e[i], ciph[i], i % 2), ref.tch3_decode(e[i], ciph[i], i % 2)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    key = rng.integers(0, 256, 8, dtype=np.uint8)
    assert (port.a5(1, key, 0x12345, 658) == ref.a5(1, key, 0x12345, 658)).all()


def test_port_equals_reference_sdr(port, ref):
    rng = np.random.default_rng(6)
    for name, win, ns in (("bcch", 80, 1), ("dc6", 40, 1), ("nt3_speech", 6, 1), ("nt3_facch", 6, 2), ("nt9", 6, 2),
                          ("rach", 6, 1), ("sdcch", 40, 4), ("dc12", 40, 1)):
        hard = rng.integers(0, 2, (6, sigen.burst_ebits(name)), dtype=np.uint8)
        x = sigen.modulate(name, hard, 4, win, rng.uniform(2, max(win - 2, 3), 6), rng.uniform(-0.013, 0.013, 6),
                           rng.uniform(0, 6, 6), np.array([6.0, 10, 15, 30, 10, 15]), rng, sync_id=ns - 1)
        for i in range(6):
            a, b = port.demod(name, x[i], 4, 0.003 * i), ref.demod(name, x[i], 4, 0.003 * i)
            assert a[0] == b[0] and (a[1] == b[1]).all() and a[2:] == b[2:], (name, i)
    x = sigen.modulate("nt3_speech", rng.integers(0, 2, (1, 212), dtype=np.uint8), 4, 6, 3.0, 0.0, 1.0, 12.0, rng)[0]
    assert port.detect(["nt3_facch", "nt3_speech"], 3.0, x, 4, 0.0) == ref.detect(["nt3_facch", "nt3_speech"], 3.0, x, 4, 0.0)
    assert port.mod_order(x, 4, 0.0) == ref.mod_order(x, 4, 0.0)
    w = sigen.fcch_window(30888, 4, 9000, 0.03, 10.0, rng)
    assert port.fcch_rough(w, 4, 0.01) == ref.fcch_rough(w, 4, 0.01)
    f = sigen.fcch_window(468 + 64, 4, 35, 0.1, 12.0, rng)[32:500]
    assert port.fcch_fine(f, 4, 0.0) == ref.fcch_fine(f, 4, 0.0)
    assert port.fcch_snr(f, 4, 0.0) == ref.fcch_snr(f, 4, 0.0)
    d = sigen.modulate_symbols(sigen.dkab_symbols(1, 9, rng), 4, 6, 2.7, 0.0, 0.5, 20.0, rng)[0]
    a, b = port.dkab_demod(d, 4, 0.0, 9), ref.dkab_demod(d, 4, 0.0, 9)
    assert a[0] == b[0] and (a[1] == b[1]).all() and a[2] == b[2]


# ------------------------------------------------------------------ (2) golden fixtures
def test_golden_l1(oracle):
    for name in ("bcch", "ccch", "xch_dc12"):
        for i, e in enumerate(G[f"{name}_e"]):
            l2, crc, conv = oracle.simple_decode(name, e)
            assert (l2 == G[f"{name}_l2"][i]).all() and crc == G[f"{name}_crc"][i] and conv == G[f"{name}_conv"][i]
    for i, e in enumerate(G["facch3_e"]):
        l2, s, crc, conv = oracle.facch3_decode(e, G["facch3_ciph"][i])
        assert (l2 == G["facch3_l2"][i]).all() and (s == G["facch3_s"][i]).all() and crc == G["facch3_crc"][i]
    for i, e in enumerate(G["facch9_e"]):
        l2, sa, st, crc, conv = oracle.facch9_decode(e, G["facch9_ciph"][i])
        assert (l2 == G["facch9_l2"][i]).all() and (sa == G["facch9_sacch"][i]).all() and conv == G["facch9_conv"][i]
    for mode in (0, 1, 2):
        e = G[f"tch9_{mode}_e"]
        for c in range(2):
            il = oracle.interleaver()
            for k in range(5):
                i = c * 5 + k
                l2, _, _, conv = oracle.tch9_decode(e[i], mode, G[f"tch9_{mode}_ciph"][i], il)
                assert (l2 == G[f"tch9_{mode}_l2"][i]).all() and conv == G[f"tch9_{mode}_conv"][i]
    for i, e in enumerate(G["rach_e"]):
        r, crc, conv, c2 = oracle.rach_decode(e, G["rach_mask"][i])
        assert (r == G["rach_l2"][i]).all() and crc == G["rach_crc"][i] and c2 == list(G["rach_crc2"][i])
    for i, e in enumerate(G["tch3_e"]):
        f0, f1, s, c0, c1 = oracle.tch3_decode(e, G["tch3_ciph"][i], i % 2)
        assert (f0 == G["tch3_f0"][i]).all() and (f1 == G["tch3_f1"][i]).all() and c0 == G["tch3_c0"][i]
    assert (oracle.a5(1, G["a5_key"], 0x2a5c7, 208) == G["a5_dl"]).all()


def test_golden_sdr(oracle):
    for name in ("bcch", "dc6", "nt3_speech", "nt3_facch", "nt9", "rach"):
        for i, x in enumerate(G[f"demod_{name}_iq"]):
            rc, eb, sid, toa, fe = oracle.demod(name, x, 4, 0.0)
            assert rc == 0 and (eb == G[f"demod_{name}_ebits"][i]).all() and sid == G[f"demod_{name}_sync"][i]
            assert np.float32(toa) == G[f"demod_{name}_toa"][i] and np.float32(fe) == G[f"demod_{name}_ferr"][i]
    for i, x in enumerate(G["fcch_fine_iq"]):
        rc, toa, fe = oracle.fcch_fine(x, 4, 0.0)
        assert rc == 0 and toa == G["fcch_fine_toa"][i] and np.float32(fe) == G["fcch_fine_ferr"][i]
        assert np.float32(oracle.fcch_snr(x, 4, 0.0)[1]) == G["fcch_snr"][i]
    assert oracle.fcch_rough(G["fcch_rough_iq"], 4, 0.0)[1] == G["fcch_rough_toa"][0]


# ------------------------------------------------------------------ (3) facts
def test_scrambler_prefix(oracle):
    out = np.zeros(16, np.uint8)
    zeros = np.zeros(16, np.uint8)
    oracle.c.gmr1_scramble_ubit(out.ctypes.data_as(ctypes.c_void_p), zeros.ctypes.data_as(ctypes.c_void_p), 16)
    assert out.tolist() == [0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 1]


def test_crc16_independent(oracle):
    """the shim's osmo_crc16gen (as parameterised by gmr1_crc16, src/l1/crc.c:58-63) against an
    independent bit-serial CRC-CCITT, and a clean BCCH round trip"""
    rng = np.random.default_rng(2)
    l2 = rng.integers(0, 256, 24, dtype=np.uint8)
    bits = np.unpackbits(l2, bitorder="little")
    reg = 0
    for b in bits:
        reg ^= int(b) << 15
        reg = ((reg << 1) ^ 0x1021) & 0xFFFF if reg & 0x8000 else (reg << 1) & 0xFFFF
    want = np.array([(reg >> (15 - i)) & 1 for i in range(16)], np.uint8)
    got = np.zeros(16, np.uint8)
    P = ctypes.c_void_p
    code = ctypes.addressof(ctypes.c_char.in_dll(oracle.c, "gmr1_crc16"))
    oracle.c.osmo_crc16gen_set_bits.argtypes = [P, P, ctypes.c_int, P]
    oracle.c.osmo_crc16gen_set_bits(code, bits.ctypes.data, 192, got.ctypes.data)
    assert (got == want).all()
    hard = oracle.encode("bcch", 424, l2)
    out, crc, conv = oracle.simple_decode("bcch", np.where(hard > 0, -127, 127).astype(np.int8))
    assert crc == 0 and conv == 0 and (out == l2).all()
    assert oracle.simple_decode("bcch", np.zeros(424, np.int8))[1] == 0      # all-erased decodes to zeros: CRC(0) = 0
    rnd = rng.integers(-127, 128, 424).astype(np.int8)
    assert oracle.simple_decode("bcch", rnd)[1] != 0                         # noise fails the CRC


def test_roundtrips_crc_zero(oracle):
    rng = np.random.default_rng(3)
    for _ in range(8):
        l2 = rng.integers(0, 256, 10, dtype=np.uint8)
        l2[9] &= 0x0F
        s = rng.integers(0, 2, 32, dtype=np.uint8)
        e = oracle.facch3_encode(l2, s)
        out, so, crc, conv = oracle.facch3_decode(np.where(e > 0, -127, 127).astype(np.int8))
        assert crc == 0 and conv == 0 and (out == l2).all() and (so == s).all()
        r = rng.integers(0, 256, 18, dtype=np.uint8)
        r[17] &= 0x07
        e = oracle.rach_encode(r, 0xA5)
        out, crc, conv, c2 = oracle.rach_decode(np.where(e > 0, -127, 127).astype(np.int8), 0xA5)
        assert crc == 0 and c2 == [0, 0] and (out == r).all()


def test_puncturing_budgets():
    """received soft bits per burst fix the number of punctured positions"""
    import osmo_gmr_b200  # noqa: F401  (tables are host code; exercised through the emulation library)
    emu = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libgmr1_emu.so")) if os.path.exists(
        os.path.join(ROOT, "tests", "emu", "libgmr1_emu.so")) else None
    if emu is None:
        pytest.skip("emulation library not built yet")
    want = {4: (740, 648), 5: (732, 648), 6: (968, 648), 7: (652, 382), 8: (96, 72), 9: (624, 432)}
    for ch, (total, kept) in want.items():
        m = np.zeros(1024, np.uint8)
        n = emu.gmr1_emu_keep_mask(ch, m.ctypes.data_as(ctypes.c_void_p), 1024)
        assert n == total and int(m[:n].sum()) == kept, ch


def test_shim_dft_matches_numpy(oracle):
    rng = np.random.default_rng(4)
    x = (rng.standard_normal(117) + 1j * rng.standard_normal(117)).astype(np.complex64)
    buf = x.copy()
    P = ctypes.c_void_p
    oracle.c.fftwf_plan_dft_1d.restype = P
    oracle.c.fftwf_plan_dft_1d.argtypes = [ctypes.c_int, P, P, ctypes.c_int, ctypes.c_uint]
    oracle.c.fftwf_execute.argtypes = [P]
    oracle.c.fftwf_destroy_plan.argtypes = [P]
    plan = oracle.c.fftwf_plan_dft_1d(117, buf.ctypes.data, buf.ctypes.data, -1, 64)
    oracle.c.fftwf_execute(plan)
    oracle.c.fftwf_destroy_plan(plan)
    assert np.abs(buf - np.fft.fft(x.astype(np.complex128))).max() < 1e-4
