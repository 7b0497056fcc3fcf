"""The C-ABI shared library loads without a GPU and exports every function and data symbol that include/gmr1_b200.h
and include/gmr1_b200_compat.h declare; without a CUDA device the entry points fail loudly (-ENODEV / -EIO with an
error text) instead of falling back to anything."""
import ctypes
import errno
import os
import re

import numpy as np
import pytest

import osmo_gmr_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    funcs, data = set(), set()
    for name in ("gmr1_b200.h", "gmr1_b200_compat.h"):
        src = open(os.path.join(ROOT, "include", name)).read()
        src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
        src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
        for stmt in src.split(";"):
            stmt = " ".join(stmt.split())
            m = re.search(r"\b(gmr1b200_\w+|gmr1_\w+)\s*\(", stmt)
            if m and "typedef" not in stmt and not stmt.startswith("struct") and "{" not in stmt:
                funcs.add(m.group(1))
            elif stmt.startswith("extern "):
                data.update(re.findall(r"\b(gmr1_\w+)\b(?=\s*(?:,|$))", stmt))
    masks = re.findall(r"^PUNCT\((\w+),", open(os.path.join(ROOT, "include", "gmr1_punct_masks.inc")).read(), re.M)
    data.update("gmr1_punct_" + m for m in masks)
    return funcs, data


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(osmo_gmr_b200.LIB_PATH)
    funcs, data = declared()
    assert len(funcs) >= 75 and len(data) >= 79, (len(funcs), len(data))
    missing = [s for s in sorted(funcs | data) if not hasattr(lib, s)]
    assert not missing, missing
    for s in ("gmr1b200_pi4cxpsk_demod_batch", "gmr1b200_bcch_decode_batch", "gmr1b200_fcch_acquire_batch",
              "gmr1b200_rx_bcch_batch", "gmr1b200_gsmtap_batch", "gmr1_pi4cxpsk_demod", "gmr1_bcch_decode",
              "gmr1_fcch_rough", "gmr1_puncturer_generate"):
        assert s in funcs
    for s in ("gmr1_bcch_burst", "gmr1_conv_k5_12", "gmr1_crc16", "gmr1_punct_k5_12_P23", "gmr1_fcch_burst"):
        assert s in data


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    lib = ctypes.CDLL(osmo_gmr_b200.LIB_PATH)
    lib.gmr1b200_last_error.restype = ctypes.c_char_p
    eb = np.zeros((4, 424), np.int8)
    l2 = np.zeros((4, 24), np.uint8)
    crc = np.zeros(4, np.int32)
    rc = lib.gmr1b200_bcch_decode_batch(l2.ctypes.data_as(ctypes.c_void_p), eb.ctypes.data_as(ctypes.c_void_p), None,
                                        crc.ctypes.data_as(ctypes.c_void_p), 4, None)
    assert rc in (-errno.ENODEV, -errno.EIO), rc
    assert len(lib.gmr1b200_last_error()) > 0
    assert not l2.any()                     # nothing was computed anywhere else
