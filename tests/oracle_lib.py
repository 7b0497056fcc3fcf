"""ctypes wrapper around the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Preference order: oracle/_ref/libgmr1_ref.so (the reference's own C sources compiled against
oracle/shim) and, when that is absent, oracle/liboracle.so (our plain-C restatement).  Both
export the reference's function names, so the wrapper is the same.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = ctypes.c_void_p


def p(a):
    return a.ctypes.data_as(P) if a is not None else None


class IL(ctypes.Structure):
    _fields_ = [("N", ctypes.c_int), ("K", ctypes.c_int), ("n", ctypes.c_int), ("bits", P)]


class CxVec(ctypes.Structure):
    _fields_ = [("len", ctypes.c_int), ("max_len", ctypes.c_int), ("flags", ctypes.c_int), ("data", P)]


class Oracle:
    def __init__(self, path, kind):
        self.path, self.kind = path, kind
        self.c = ctypes.CDLL(path)

    # ---------------- L1 encoders (used as vector generators)
    def encode(self, name, nbits, l2):
        out = np.zeros(nbits, np.uint8)
        getattr(self.c, f"gmr1_{name}_encode")(p(out), p(l2))
        return out

    def facch3_encode(self, l2, bits_s, ciph=None):
        out = np.zeros(416, np.uint8)
        self.c.gmr1_facch3_encode(p(out), p(l2), p(bits_s), p(ciph))
        return out

    def facch9_encode(self, l2, sacch, status, ciph=None):
        out = np.zeros(662, np.uint8)
        self.c.gmr1_facch9_encode(p(out), p(l2), p(sacch), p(status), p(ciph))
        return out

    def rach_encode(self, rach, sb_mask):
        out = np.zeros(494, np.uint8)
        self.c.gmr1_rach_encode(p(out), p(rach), ctypes.c_uint8(int(sb_mask)))
        return out

    def interleaver(self):
        il = IL()
        self.c.gmr1_interleaver_init(ctypes.byref(il), 3, 648)
        return il

    def tch9_encode(self, l2, mode, sacch, status, ciph, il):
        out = np.zeros(662, np.uint8)
        self.c.gmr1_tch9_encode(p(out), p(l2), mode, p(sacch), p(status), p(ciph), ctypes.byref(il))
        return out

    # ---------------- L1 decoders
    def simple_decode(self, name, e, l2_bytes=24):
        l2 = np.zeros(l2_bytes, np.uint8)
        cv = ctypes.c_int()
        crc = getattr(self.c, f"gmr1_{name}_decode")(p(l2), p(e), ctypes.byref(cv))
        return l2, crc, cv.value

    def facch3_decode(self, e, ciph=None):
        l2 = np.zeros(10, np.uint8)
        s = np.zeros(32, np.uint8)
        cv = ctypes.c_int()
        crc = self.c.gmr1_facch3_decode(p(l2), p(s), p(e), p(ciph), ctypes.byref(cv))
        return l2, s, crc, cv.value

    def facch9_decode(self, e, ciph=None):
        l2 = np.zeros(38, np.uint8)
        sa = np.zeros(10, np.int8)
        st = np.zeros(4, np.int8)
        cv = ctypes.c_int()
        crc = self.c.gmr1_facch9_decode(p(l2), p(sa), p(st), p(e), p(ciph), ctypes.byref(cv))
        return l2, sa, st, crc, cv.value

    def tch9_decode(self, e, mode, ciph, il):
        l2 = np.zeros((18, 30, 60)[mode], np.uint8)
        sa = np.zeros(10, np.int8)
        st = np.zeros(4, np.int8)
        cv = ctypes.c_int()
        self.c.gmr1_tch9_decode(p(l2), p(sa), p(st), p(e), mode, p(ciph), ctypes.byref(il), ctypes.byref(cv))
        return l2, sa, st, cv.value

    def tch3_decode(self, e, ciph, m):
        f0 = np.zeros(10, np.uint8)
        f1 = np.zeros(10, np.uint8)
        s = np.zeros(4, np.uint8)
        c0, c1 = ctypes.c_int(), ctypes.c_int()
        self.c.gmr1_tch3_decode(p(f0), p(f1), p(s), p(e), p(ciph), m, ctypes.byref(c0), ctypes.byref(c1))
        return f0, f1, s, c0.value, c1.value

    def rach_decode(self, e, sb_mask):
        r = np.zeros(18, np.uint8)
        cv = ctypes.c_int()
        c2 = (ctypes.c_int * 2)()
        crc = self.c.gmr1_rach_decode(p(r), p(e), ctypes.c_uint8(int(sb_mask)), ctypes.byref(cv), c2)
        return r, crc, cv.value, list(c2)

    # ---------------- SDR
    def _burst(self, name):
        return ctypes.addressof(ctypes.c_char.in_dll(self.c, f"gmr1_{name}_burst"))

    def _cxvec(self, x):
        x = np.ascontiguousarray(x, np.complex64)
        return CxVec(len=x.shape[0], max_len=x.shape[0], flags=0, data=x.ctypes.data), x

    def demod(self, name, window, sps, freq_shift):
        """gmr1_pi4cxpsk_demod -> (rc, ebits, sync_id, toa, freq_err)"""
        neb = {"bcch": 424, "dc2": 132, "dc6": 432, "dc12": 432, "nt3_speech": 212, "nt3_facch": 104,
               "nt6": 434, "nt9": 662, "rach": 494, "sdcch": 208}[name]
        cv, keep = self._cxvec(window)
        eb = np.zeros(neb, np.int8)
        sid, toa, fe = ctypes.c_int(-99), ctypes.c_float(), ctypes.c_float()
        fn = self.c.gmr1_pi4cxpsk_demod
        fn.argtypes = [P, P, ctypes.c_int, ctypes.c_float, P, P, P, P]
        rc = fn(self._burst(name), ctypes.addressof(cv), sps, float(freq_shift), p(eb),
                ctypes.addressof(sid), ctypes.addressof(toa), ctypes.addressof(fe))
        return rc, eb, sid.value, toa.value, fe.value

    def detect(self, names, e_toa, window, sps, freq_shift):
        """gmr1_pi4cxpsk_detect -> (rc, bt_id, sync_id, toa)"""
        cv, keep = self._cxvec(window)
        arr = (P * (len(names) + 1))(*[self._burst(n) for n in names], None)
        bt, sid, toa = ctypes.c_int(-99), ctypes.c_int(-99), ctypes.c_float()
        fn = self.c.gmr1_pi4cxpsk_detect
        fn.argtypes = [P, ctypes.c_float, P, ctypes.c_int, ctypes.c_float, P, P, P]
        rc = fn(arr, float(e_toa), ctypes.addressof(cv), sps, float(freq_shift),
                ctypes.addressof(bt), ctypes.addressof(sid), ctypes.addressof(toa))
        return rc, bt.value, sid.value, toa.value

    def _fcch(self, t):
        name = ["gmr1_fcch_burst", "gmr1_fcch3_lband_burst", "gmr1_fcch3_sband_burst"][t]
        return ctypes.addressof(ctypes.c_char.in_dll(self.c, name))

    def fcch_rough(self, window, sps, freq_shift, t=0):
        cv, keep = self._cxvec(window)
        toa = ctypes.c_int(-99999)
        fn = self.c.gmr1_fcch_rough
        fn.argtypes = [P, P, ctypes.c_int, ctypes.c_float, P]
        rc = fn(self._fcch(t), ctypes.addressof(cv), sps, float(freq_shift), ctypes.addressof(toa))
        return rc, toa.value

    def fcch_rough_multi(self, window, sps, freq_shift, N=16, t=0):
        cv, keep = self._cxvec(window)
        toa = (ctypes.c_int * N)()
        fn = self.c.gmr1_fcch_rough_multi
        fn.argtypes = [P, P, ctypes.c_int, ctypes.c_float, P, ctypes.c_int]
        rc = fn(self._fcch(t), ctypes.addressof(cv), sps, float(freq_shift), ctypes.addressof(toa), N)
        return rc, [toa[i] for i in range(max(rc, 0))]

    def fcch_fine(self, window, sps, freq_shift, t=0):
        cv, keep = self._cxvec(window)
        toa, fe = ctypes.c_int(-99999), ctypes.c_float()
        fn = self.c.gmr1_fcch_fine
        fn.argtypes = [P, P, ctypes.c_int, ctypes.c_float, P, P]
        rc = fn(self._fcch(t), ctypes.addressof(cv), sps, float(freq_shift), ctypes.addressof(toa), ctypes.addressof(fe))
        return rc, toa.value, fe.value

    def fcch_snr(self, window, sps, freq_shift, t=0):
        cv, keep = self._cxvec(window)
        snr = ctypes.c_float()
        fn = self.c.gmr1_fcch_snr
        fn.argtypes = [P, P, ctypes.c_int, ctypes.c_float, P]
        rc = fn(self._fcch(t), ctypes.addressof(cv), sps, float(freq_shift), ctypes.addressof(snr))
        return rc, snr.value

    def dkab_demod(self, window, sps, freq_shift, pp):
        cv, keep = self._cxvec(window)
        eb = np.zeros(8, np.int8)
        toa = ctypes.c_float()
        fn = self.c.gmr1_dkab_demod
        fn.argtypes = [P, ctypes.c_int, ctypes.c_float, ctypes.c_int, P, P]
        rc = fn(ctypes.addressof(cv), sps, float(freq_shift), pp, p(eb), ctypes.addressof(toa))
        return rc, eb, toa.value

    def mod_order(self, window, sps, freq_shift):
        cv, keep = self._cxvec(window)
        fn = self.c.gmr1_pi4cxpsk_mod_order
        fn.argtypes = [P, ctypes.c_int, ctypes.c_float]
        return fn(ctypes.addressof(cv), sps, float(freq_shift))

    def a5(self, n, key, fn, nbits, both=False):
        dl = np.zeros(nbits, np.uint8)
        ul = np.zeros(nbits, np.uint8)
        k = np.ascontiguousarray(key, np.uint8)
        self.c.gmr1_a5(n, p(k), ctypes.c_uint32(fn), nbits, p(dl), p(ul) if both else None)
        return (dl, ul) if both else dl

    def gsmtap(self, chan_type, fn, tn, l2):
        """gmr1_gsmtap_makemsg (src/gsmtap.c:44) -> the bytes of the message"""
        self.c.gmr1_gsmtap_makemsg.restype = ctypes.c_void_p
        l2 = np.ascontiguousarray(l2, np.uint8)
        m = self.c.gmr1_gsmtap_makemsg(ctypes.c_uint8(chan_type), ctypes.c_uint32(fn), ctypes.c_uint8(tn), p(l2), len(l2))
        n = ctypes.c_uint16.from_address(m).value                  # struct msgb { uint16_t len, alloc; uint8_t data[]; }
        out = np.ctypeslib.as_array((ctypes.c_uint8 * n).from_address(m + 4)).copy()
        self.c.msgb_free(ctypes.c_void_p(m))
        return out


    # ---------------- batch loops in C (oracle/harness.c): n windows per call, no Python per burst
    @staticmethod
    def _iq(x):
        x = np.ascontiguousarray(x, np.complex64)
        assert x.ndim == 2
        return x

    def h_xcch(self, is_ccch, x, sps=4, freq_shift=0.0):
        x = self._iq(x)
        n, wl = x.shape
        l2, crc, conv, toa = np.zeros((n, 24), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        self.c.oh_xcch(int(is_ccch), p(x), n, wl, sps, ctypes.c_float(freq_shift), p(l2), p(crc), p(conv), p(toa))
        return l2, crc, conv, toa

    def h_tch3(self, x, ciph=None, m=0, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        f0, f1, bs, conv = np.zeros((n, 10), np.uint8), np.zeros((n, 10), np.uint8), np.zeros((n, 4), np.uint8), np.zeros((n, 2), np.int32)
        ciph = None if ciph is None else np.ascontiguousarray(ciph, np.uint8)
        self.c.oh_tch3(p(x), n, wl, sps, p(ciph), m, p(f0), p(f1), p(bs), p(conv))
        return f0, f1, bs, conv

    def h_facch3(self, x, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        g = n // 4
        l2, bs, crc, sid = np.zeros((g, 10), np.uint8), np.zeros((g, 32), np.uint8), np.zeros(g, np.int32), np.zeros(4 * g, np.int32)
        self.c.oh_facch3(p(x), g, wl, sps, p(l2), p(bs), p(crc), p(sid))
        return l2, bs, crc, sid

    def h_facch9(self, x, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        l2, crc, sid = np.zeros((n, 38), np.uint8), np.zeros(n, np.int32), np.zeros(n, np.int32)
        self.c.oh_facch9(p(x), n, wl, sps, p(l2), p(crc), p(sid))
        return l2, crc, sid

    def h_tch9(self, x, n_burst, mode=2, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        nb = (18, 30, 60)[mode]
        l2, conv = np.zeros((n, nb), np.uint8), np.zeros(n, np.int32)
        self.c.oh_tch9(p(x), n // n_burst, n_burst, wl, sps, mode, p(l2), p(conv))
        return l2, conv

    def h_rach(self, x, sb_mask=None, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        r, crc = np.zeros((n, 18), np.uint8), np.zeros(n, np.int32)
        sb = None if sb_mask is None else np.ascontiguousarray(sb_mask, np.uint8)
        self.c.oh_rach(p(x), n, wl, sps, p(sb), p(r), p(crc))
        return r, crc

    def h_fcch_acquire(self, x, sps=4, freq_shift=0.0):
        x = self._iq(x)
        n, wl = x.shape
        rough, align, ferr = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32)
        self.c.oh_fcch_acquire(p(x), n, wl, sps, ctypes.c_float(freq_shift), p(rough), p(align), p(ferr))
        return rough, align, ferr

    def h_fcch_grid(self, x, shifts, sps=4):
        x = self._iq(x)
        n, wl = x.shape
        sh = np.ascontiguousarray(shifts, np.float32)
        toa = np.zeros((len(sh), n), np.int32)
        self.c.oh_fcch_grid(p(x), n, wl, sps, p(sh), len(sh), p(toa))
        return toa


def host_has_avx2():
    try:
        flags = open("/proc/cpuinfo").read()
        return all(f" {k}" in flags for k in ("avx2", "bmi2", "fma", "movbe"))
    except OSError:
        return False


def load(fast=False):
    """fast = True: the -march=x86-64-v3 build of the same sources when this host can run it (CPU timing legs);
    float results are the same in both builds (-ffp-contract=off)."""
    sfx = "_v3" if fast and host_has_avx2() else ""
    ref = os.path.join(ROOT, "oracle", "_ref", f"libgmr1_ref{sfx}.so")
    port = os.path.join(ROOT, "oracle", f"liboracle{sfx}.so")
    if not os.path.exists(ref) and os.path.isdir("/root/reference/src"):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if os.path.exists(ref):
        o = Oracle(ref, "reference")
    else:
        if not os.path.exists(port):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), os.path.basename(port)])
        o = Oracle(port, "port")
    o.flags = "-O2 -march=x86-64-v3 -ffp-contract=off" if sfx else "-O2 -ffp-contract=off"
    return o
