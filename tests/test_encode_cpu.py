"""CPU: the product's host-side channel encoders vs the reference encoders (bit-exact), and the
corrected TCH3 encoder vs the reference TCH3 *decoder* (round trip)."""
import numpy as np
import pytest


@pytest.fixture(scope="module")
def L():
    import osmo_gmr_b200
    return osmo_gmr_b200.lib()        # loading needs no GPU; encoders are host code


def test_xcch(L, oracle):
    rng = np.random.default_rng(1)
    for chan, name, nbits in ((0, "bcch", 424), (1, "ccch", 432), (2, "xch_dc12", 432)):
        l2 = rng.integers(0, 256, (32, 24), dtype=np.uint8)
        out = np.zeros((32, nbits), np.uint8)
        L.call("gmr1b200_xcch_encode_batch", chan, out, l2, 32)
        for i in range(32):
            assert (out[i] == oracle.encode(name, nbits, l2[i])).all(), (name, i)


@pytest.mark.parametrize("use_ciph", [False, True])
def test_facch3_facch9_rach(L, oracle, use_ciph):
    rng = np.random.default_rng(2)
    for _ in range(16):
        l2 = rng.integers(0, 256, 10, dtype=np.uint8)
        s = rng.integers(0, 2, 32, dtype=np.uint8)
        c = rng.integers(0, 2, 384, dtype=np.uint8) if use_ciph else None
        out = np.zeros(416, np.uint8)
        L.call("gmr1b200_facch3_encode", out, l2, s, c)
        assert (out == oracle.facch3_encode(l2, s, c)).all()

        l2 = rng.integers(0, 256, 38, dtype=np.uint8)
        sa = rng.integers(0, 2, 10, dtype=np.uint8)
        st = rng.integers(0, 2, 4, dtype=np.uint8)
        c = rng.integers(0, 2, 658, dtype=np.uint8) if use_ciph else None
        out = np.zeros(662, np.uint8)
        L.call("gmr1b200_facch9_encode", out, l2, sa, st, c)
        assert (out == oracle.facch9_encode(l2, sa, st, c)).all()

        r = rng.integers(0, 256, 18, dtype=np.uint8)
        m = int(rng.integers(0, 256))
        out = np.zeros(494, np.uint8)
        L.call("gmr1b200_rach_encode", out, r, m)
        assert (out == oracle.rach_encode(r, m)).all()


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_tch9(L, oracle, mode):
    rng = np.random.default_rng(3 + mode)
    nb = (18, 30, 60)[mode]
    il_o = oracle.interleaver()
    il = L.c.gmr1b200_tch9_interleaver_new()
    try:
        for _ in range(7):
            l2 = rng.integers(0, 256, nb, dtype=np.uint8)
            sa = rng.integers(0, 2, 10, dtype=np.uint8)
            st = rng.integers(0, 2, 4, dtype=np.uint8)
            c = rng.integers(0, 2, 658, dtype=np.uint8)
            out = np.zeros(662, np.uint8)
            L.call("gmr1b200_tch9_encode", out, l2, mode, sa, st, c, il)
            assert (out == oracle.tch9_encode(l2, mode, sa, st, c, il_o)).all()
    finally:
        L.c.gmr1b200_tch9_interleaver_free(il)


@pytest.mark.parametrize("m", [0, 1])
def test_tch3_roundtrip_through_reference_decoder(L, oracle, m):
    rng = np.random.default_rng(9)
    for _ in range(32):
        f0 = rng.integers(0, 256, 10, dtype=np.uint8)
        f1 = rng.integers(0, 256, 10, dtype=np.uint8)
        s = rng.integers(0, 2, 4, dtype=np.uint8)
        c = rng.integers(0, 2, 208, dtype=np.uint8)
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, f0, f1, s, c, m)
        soft = np.where(out > 0, -127, 127).astype(np.int8)
        g0, g1, gs, c0, c1 = oracle.tch3_decode(soft, c, m)
        assert (g0 == f0).all() and (g1 == f1).all() and (gs == s).all() and c0 == 0 and c1 == 0
