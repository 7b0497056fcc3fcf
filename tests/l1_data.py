"""Reads the reference-named L1 data symbols (convolutional codes, CRC parameters, puncturing masks) out of a
shared library and runs gmr1_puncturer_generate on it.  Used on libgmr1_b200.so (the product) and on
oracle/_ref/libgmr1_ref.so (the reference build) by tests/test_l1_data_cpu.py and tests/golden/make_l1_data.py."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONV = ["k5_12", "k5_13", "k5_14", "k5_15", "k6_14", "k9_12", "k9_13", "k9_14", "tch3"]
MASKS = [m.group(1) for m in re.finditer(r"^PUNCT\((\w+),", open(os.path.join(ROOT, "include", "gmr1_punct_masks.inc")).read(), re.M)]


class ConvCode(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("K", ctypes.c_int), ("len", ctypes.c_int), ("term", ctypes.c_int),
                ("next_output", ctypes.c_void_p), ("next_state", ctypes.c_void_p),
                ("next_term_output", ctypes.c_void_p), ("next_term_state", ctypes.c_void_p),
                ("puncture", ctypes.POINTER(ctypes.c_int))]


class Crc8(ctypes.Structure):
    _fields_ = [("bits", ctypes.c_int), ("poly", ctypes.c_uint8), ("init", ctypes.c_uint8), ("remainder", ctypes.c_uint8)]


class Crc16(ctypes.Structure):
    _fields_ = [("bits", ctypes.c_int), ("poly", ctypes.c_uint16), ("init", ctypes.c_uint16), ("remainder", ctypes.c_uint16)]


class PunctHdr(ctypes.Structure):
    _fields_ = [("r", ctypes.c_int), ("L", ctypes.c_int), ("N", ctypes.c_int)]


# (code, data bits, termination, pre, main, post, repeat): the reference's own calls (tch3.c:48, tch9.c:61-77,
# xch_dc12.c:52) and a few more shapes
GENERATE = [("tch3", 48, 2, None, "k5_12_P12", None, 0),
            ("k5_15", 144, 0, "k5_15_P53", "k5_15_P23", "k5_15_Ps53", 41),
            ("k5_13", 240, 0, "k5_13_P15", "k5_13_P25", "k5_13_Ps15", 41),
            ("k5_12", 480, 0, "k5_12_P25", "k5_12_P23", "k5_12_Ps25", 158),
            ("k9_13", 208, 2, None, "k9_13_P1213", None, 0),
            ("k5_12", 100, 0, None, "k5_12_P311", None, 0),
            ("k5_12", 77, 0, "k5_12_P412", "k5_12_P26", None, 0),
            ("k9_12", 64, 0, None, "k9_12_P47", "k9_12_P13", 3),
            ("k5_12", 10, 0, None, "k5_13_P16", None, 0)]          # rate mismatch -> -EINVAL


def read_all(path):
    """-> dict of numpy arrays describing every symbol of the library at `path`"""
    c = ctypes.CDLL(path)
    out = {}
    for name in CONV:
        cc = ConvCode.in_dll(c, "gmr1_conv_" + name)
        ns = 1 << (cc.K - 1)
        out[f"conv_{name}_hdr"] = np.array([cc.N, cc.K, cc.len, cc.term, int(bool(cc.puncture)),
                                            int(bool(cc.next_term_output)), int(bool(cc.next_term_state))], np.int32)
        out[f"conv_{name}_out"] = np.ctypeslib.as_array((ctypes.c_uint8 * (2 * ns)).from_address(cc.next_output)).copy()
        out[f"conv_{name}_state"] = np.ctypeslib.as_array((ctypes.c_uint8 * (2 * ns)).from_address(cc.next_state)).copy()
    c8 = Crc8.in_dll(c, "gmr1_crc8")
    out["crc8"] = np.array([c8.bits, c8.poly, c8.init, c8.remainder], np.int32)
    for n in ("crc12", "crc16"):
        cr = Crc16.in_dll(c, "gmr1_" + n)
        out[n] = np.array([cr.bits, cr.poly, cr.init, cr.remainder], np.int32)
    for m in MASKS:
        h = PunctHdr.in_dll(c, "gmr1_punct_" + m)
        addr = ctypes.addressof(h) + ctypes.sizeof(PunctHdr)
        mask = np.ctypeslib.as_array((ctypes.c_uint8 * (h.L * h.N)).from_address(addr)).copy()
        out[f"punct_{m}"] = np.concatenate([np.array([h.r, h.L, h.N], np.int32), mask.astype(np.int32)])
    c.gmr1_puncturer_generate.restype = ctypes.c_int
    c.gmr1_puncturer_generate.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int]
    for i, (code, ln, term, pre, main, post, rep) in enumerate(GENERATE):
        src = ConvCode.in_dll(c, "gmr1_conv_" + code)
        cc = ConvCode(src.N, src.K, ln, term, src.next_output, src.next_state, None, None, None)
        sym = lambda m: ctypes.addressof(PunctHdr.in_dll(c, "gmr1_punct_" + m)) if m else None
        rv = c.gmr1_puncturer_generate(ctypes.addressof(cc), sym(pre), sym(main), sym(post), rep)
        lst = [rv]
        if rv == 0:
            k = 0
            while cc.puncture[k] >= 0:
                lst.append(cc.puncture[k])
                k += 1
        out[f"generate_{i}"] = np.array(lst, np.int32)
    return out
