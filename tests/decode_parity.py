"""Channel-decode parity checks shared by the CPU-emulation tests and the GPU tests.
Bar: bit-exact (L2 bytes, CRC results, Viterbi metric, side outputs) - integer path."""
import numpy as np

import vectors
from decode_backends import CH


def _eq(a, b, what):
    assert a.shape == b.shape, what
    bad = np.argwhere(a != b)
    assert bad.size == 0, f"{what}: {bad.shape[0]} mismatches, first at {bad[0]}: got {a[tuple(bad[0])]} want {b[tuple(bad[0])]}"


def check_simple(backend, oracle, name, ch, nbits, n, seed):
    rng = np.random.default_rng(seed)
    e = vectors.simple(oracle, rng, name, nbits, n)
    want = [oracle.simple_decode(name, e[i]) for i in range(n)]
    got = backend.decode(ch, e)
    _eq(got["l2"], np.stack([w[0] for w in want]), f"{name} l2")
    _eq(got["crc"], np.array([w[1] for w in want], np.int32), f"{name} crc")
    _eq(got["conv"], np.array([w[2] for w in want], np.int32), f"{name} conv")
    assert (np.array([w[1] for w in want]) == 0).mean() > 0.2      # a useful share decodes cleanly


def check_facch3(backend, oracle, n, seed, use_ciph):
    rng = np.random.default_rng(seed)
    e, ciph = vectors.facch3(oracle, rng, n, use_ciph)
    want = [oracle.facch3_decode(e[i], ciph[i] if use_ciph else None) for i in range(n)]
    got = backend.decode(CH["FACCH3"], e, ciph=ciph)
    _eq(got["l2"], np.stack([w[0] for w in want]), "facch3 l2")
    _eq(got["bits_s"], np.stack([w[1] for w in want]), "facch3 bits_s")
    _eq(got["crc"], np.array([w[2] for w in want], np.int32), "facch3 crc")
    _eq(got["conv"], np.array([w[3] for w in want], np.int32), "facch3 conv")


def check_facch9(backend, oracle, n, seed, use_ciph):
    rng = np.random.default_rng(seed)
    e, ciph = vectors.facch9(oracle, rng, n, use_ciph)
    want = [oracle.facch9_decode(e[i], ciph[i] if use_ciph else None) for i in range(n)]
    got = backend.decode(CH["FACCH9"], e, ciph=ciph)
    _eq(got["l2"], np.stack([w[0] for w in want]), "facch9 l2")
    _eq(got["sacch"], np.stack([w[1] for w in want]), "facch9 sacch")
    _eq(got["status"], np.stack([w[2] for w in want]), "facch9 status")
    _eq(got["crc"], np.array([w[3] for w in want], np.int32), "facch9 crc")
    _eq(got["conv"], np.array([w[4] for w in want], np.int32), "facch9 conv")


def check_tch9(backend, oracle, mode, nchan, nburst, seed, use_ciph):
    rng = np.random.default_rng(seed)
    e, ciph, prev1, prev2 = vectors.tch9(oracle, rng, mode, nchan, nburst, use_ciph)
    want = []
    for c in range(nchan):
        il = oracle.interleaver()
        for b in range(nburst):
            i = c * nburst + b
            want.append(oracle.tch9_decode(e[i], mode, ciph[i] if use_ciph else None, il))
    got = backend.decode(CH["TCH9_2K4"] + mode, e, ciph=ciph, prev1=prev1, prev2=prev2)
    _eq(got["l2"], np.stack([w[0] for w in want]), f"tch9[{mode}] l2")
    _eq(got["sacch"], np.stack([w[1] for w in want]), f"tch9[{mode}] sacch")
    _eq(got["status"], np.stack([w[2] for w in want]), f"tch9[{mode}] status")
    _eq(got["conv"], np.array([w[3] for w in want], np.int32), f"tch9[{mode}] conv")


def check_rach(backend, oracle, n, seed):
    rng = np.random.default_rng(seed)
    e, masks = vectors.rach(oracle, rng, n)
    want = [oracle.rach_decode(e[i], masks[i]) for i in range(n)]
    got = backend.decode(CH["RACH"], e, sb_mask=masks)
    _eq(got["l2"], np.stack([w[0] for w in want]), "rach bytes")
    _eq(got["crc"], np.array([w[1] for w in want], np.int32), "rach crc")
    _eq(got["conv"], np.array([w[2] for w in want], np.int32), "rach conv")
    _eq(got["crc2"], np.array([w[3] for w in want], np.int32), "rach crc_rv")


def check_tch3(backend, oracle, n, seed, use_ciph, m):
    rng = np.random.default_rng(seed)
    e, ciph = vectors.tch3(rng, n, use_ciph)
    want = [oracle.tch3_decode(e[i], ciph[i] if use_ciph else None, m) for i in range(n)]
    got = backend.decode(CH["TCH3"], e, ciph=ciph, m=m)
    _eq(got["l2"], np.stack([w[0] for w in want]), "tch3 frame0")
    _eq(got["l2b"], np.stack([w[1] for w in want]), "tch3 frame1")
    _eq(got["bits_s"], np.stack([w[2] for w in want]), "tch3 bits_s")
    _eq(got["conv"], np.array([w[3] for w in want], np.int32), "tch3 conv0")
    _eq(got["conv1"], np.array([w[4] for w in want], np.int32), "tch3 conv1")


def check_dc12(backend, oracle, n, seed):
    rng = np.random.default_rng(seed)
    e = vectors.simple(oracle, rng, "xch_dc12", 432, n)
    want = [oracle.simple_decode("xch_dc12", e[i]) for i in range(n)]
    got = backend.decode(CH["DC12"], e)
    _eq(got["l2"], np.stack([w[0] for w in want]), "dc12 l2")
    _eq(got["crc"], np.array([w[1] for w in want], np.int32), "dc12 crc")
    _eq(got["conv"], np.array([w[2] for w in want], np.int32), "dc12 conv")
