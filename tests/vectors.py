"""Seeded test-vector generation for the channel decoders (stage 3).

Soft bits come from the oracle's own encoders (reference src/l1/*_encode) plus integer noise, and
from pure noise, so that tie-breaks, erasures (0), saturation (+-127, -128) and failing CRCs are
all exercised.
"""
import numpy as np


def soften(rng, hard, sigma, amp=64):
    s = np.where(hard > 0, -1.0, 1.0) * amp + rng.normal(0, sigma, hard.shape)
    return np.clip(np.round(s), -127, 127).astype(np.int8)


def spice(rng, e):
    """overwrite a few rows with adversarial content"""
    n = e.shape[0]
    if n >= 8:
        e[-1] = rng.integers(-128, 128, e.shape[1])
        e[-2] = 0
        e[-3] = -128
        e[-4] = 127
        e[-5, ::2] = 0
        e[-6] = rng.integers(-2, 3, e.shape[1])
    return e


def simple(oracle, rng, name, nbits, n, l2_bytes=24):
    hard = np.stack([oracle.encode(name, nbits, rng.integers(0, 256, l2_bytes, dtype=np.uint8)) for _ in range(n)])
    sig = np.where(np.arange(n)[:, None] % 3 == 0, 20.0, np.where(np.arange(n)[:, None] % 3 == 1, 45.0, 70.0))
    e = soften(rng, hard, 1.0) if False else np.clip(
        np.round(np.where(hard > 0, -64.0, 64.0) + rng.normal(0, 1, hard.shape) * sig), -127, 127).astype(np.int8)
    return spice(rng, e)


def facch3(oracle, rng, n, use_ciph):
    ciph = rng.integers(0, 2, (n, 384), dtype=np.uint8) if use_ciph else None
    hard = np.stack([oracle.facch3_encode(rng.integers(0, 256, 10, dtype=np.uint8),
                                          rng.integers(0, 2, 32, dtype=np.uint8),
                                          ciph[i] if use_ciph else None) for i in range(n)])
    return spice(rng, soften(rng, hard, 40)), ciph


def facch9(oracle, rng, n, use_ciph):
    ciph = rng.integers(0, 2, (n, 658), dtype=np.uint8) if use_ciph else None
    hard = np.stack([oracle.facch9_encode(rng.integers(0, 256, 38, dtype=np.uint8),
                                          rng.integers(0, 2, 10, dtype=np.uint8),
                                          rng.integers(0, 2, 4, dtype=np.uint8),
                                          ciph[i] if use_ciph else None) for i in range(n)])
    return spice(rng, soften(rng, hard, 45)), ciph


def tch9(oracle, rng, mode, nchan, nburst, use_ciph):
    """nchan channels x nburst consecutive bursts, channel-major; returns ebits, ciph, prev1, prev2"""
    n = nchan * nburst
    nb = (18, 30, 60)[mode]
    ciph = rng.integers(0, 2, (n, 658), dtype=np.uint8) if use_ciph else None
    hard = np.zeros((n, 662), np.uint8)
    prev1 = np.full(n, -1, np.int32)
    prev2 = np.full(n, -1, np.int32)
    for c in range(nchan):
        il = oracle.interleaver()
        for b in range(nburst):
            i = c * nburst + b
            hard[i] = oracle.tch9_encode(rng.integers(0, 256, nb, dtype=np.uint8), mode,
                                         rng.integers(0, 2, 10, dtype=np.uint8),
                                         rng.integers(0, 2, 4, dtype=np.uint8),
                                         ciph[i] if use_ciph else None, il)
            if b >= 1:
                prev1[i] = i - 1
            if b >= 2:
                prev2[i] = i - 2
    e = soften(rng, hard, 50)
    e[-1] = rng.integers(-128, 128, 662)
    return e, ciph, prev1, prev2


def rach(oracle, rng, n):
    masks = rng.integers(0, 256, n, dtype=np.uint8)
    hard = np.stack([oracle.rach_encode(rng.integers(0, 256, 18, dtype=np.uint8), masks[i] if i % 2 else 0)
                     for i in range(n)])
    return spice(rng, soften(rng, hard, 45)), masks


def tch3(rng, n, use_ciph):
    """the reference's gmr1_tch3_encode is unusable (src/l1/tch3.c:81 swaps the encoder's
    arguments), so TCH3 vectors are structured noise: random +-60 pattern plus noise"""
    e = rng.integers(-128, 128, (n, 212)).astype(np.int8)
    h = n // 2
    e[:h] = np.clip(e[:h].astype(int) // 2 + np.where(rng.integers(0, 2, (h, 212)) > 0, 60, -60), -127, 127)
    ciph = rng.integers(0, 2, (n, 208), dtype=np.uint8) if use_ciph else None
    return spice(rng, e), ciph
