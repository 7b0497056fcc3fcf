"""Two ways to run the product's channel-decode code on a batch:

   EmuBackend  the __host__ __device__ decode functions compiled for the CPU (tests/emu) - logic
               check that runs without a GPU; test harness, not a product path
   GpuBackend  the real thing: CUDA kernels through the C ABI of libgmr1_b200.so; `device=True`
               passes torch CUDA tensors (device pointers), else numpy host buffers
Both return dicts of numpy arrays with the same keys so one set of parity checks serves both.
"""
import ctypes

import numpy as np

P = ctypes.c_void_p
CH = dict(BCCH=0, CCCH=1, FACCH3=2, FACCH9=3, TCH9_2K4=4, TCH9_4K8=5, TCH9_9K6=6, RACH=7, TCH3=8, DC12=9)
L2B = {0: 24, 1: 24, 2: 10, 3: 38, 4: 18, 5: 30, 6: 60, 7: 18, 8: 10, 9: 24}


class Args(ctypes.Structure):
    _fields_ = [("ebits", P), ("ciph", P), ("n", ctypes.c_int32), ("l2", P), ("l2b", P), ("conv", P), ("conv1", P),
                ("crc", P), ("crc2", P), ("bits_s", P), ("sacch", P), ("status", P), ("prev1", P), ("prev2", P),
                ("sb_mask", P), ("sb_mask0", ctypes.c_int32), ("tch3_m", ctypes.c_int32), ("t9_rows", ctypes.c_int32),
                ("n_dev", P), ("dec_scratch", P)]


def _p(a):
    return a.ctypes.data_as(P) if a is not None else None


def _outs(ch, n):
    o = dict(l2=np.zeros((n, L2B[ch]), np.uint8), conv=np.full(n, -7, np.int32))
    if ch in (0, 1, 2, 3, 7, 9):
        o["crc"] = np.full(n, -7, np.int32)
    if ch == 2:
        o["bits_s"] = np.zeros((n, 32), np.uint8)
    if ch in (3, 4, 5, 6):
        o["sacch"] = np.zeros((n, 10), np.int8)
        o["status"] = np.zeros((n, 4), np.int8)
    if ch == 7:
        o["crc2"] = np.full((n, 2), -7, np.int32)
    if ch == 8:
        o["l2b"] = np.zeros((n, 10), np.uint8)
        o["conv1"] = np.full(n, -7, np.int32)
        o["bits_s"] = np.zeros((n, 4), np.uint8)
    return o


class EmuBackend:
    name = "emu"

    def __init__(self, emu, p16=False, lut=False):
        """p16: two codewords per "thread" with packed 16-bit metrics (viterbi_p16.cuh); lut: one codeword per "thread"
        with the branch metrics through the table (viterbi_tpc.cuh, RelLut) - what the kernels run"""
        self.emu = emu
        self.p16 = p16
        self.lut = lut
        self.name = "emu-p16" if p16 else "emu-lut" if lut else "emu"
        assert ctypes.sizeof(Args) == emu.gmr1_emu_sizeof_args()

    def decode(self, ch, e, ciph=None, prev1=None, prev2=None, sb_mask=None, m=0):
        n = e.shape[0]
        o = _outs(ch, n)
        a = Args(ebits=_p(e), ciph=_p(ciph), n=n, l2=_p(o["l2"]), l2b=_p(o.get("l2b")), conv=_p(o["conv"]),
                 conv1=_p(o.get("conv1")), crc=_p(o.get("crc")), crc2=_p(o.get("crc2")), bits_s=_p(o.get("bits_s")),
                 sacch=_p(o.get("sacch")), status=_p(o.get("status")), prev1=_p(prev1), prev2=_p(prev2),
                 sb_mask=_p(sb_mask), sb_mask0=0, tch3_m=m)
        fn = self.emu.gmr1_emu_decode_p16 if self.p16 else self.emu.gmr1_emu_decode_lut if self.lut else self.emu.gmr1_emu_decode
        assert fn(ch, ctypes.byref(a)) == 0
        return o


class GpuBackend:
    def __init__(self, lib, device=False, misalign=0):
        """misalign (device only): soft bits and cipher bytes start `misalign` bytes behind a 16-byte boundary (no bulk
        copy of the tile, no word loads of the cipher bytes)"""
        self.lib = lib
        self.device = device
        self.misalign = misalign
        self.name = "gpu-dev" if device else "gpu-host"

    def decode(self, ch, e, ciph=None, prev1=None, prev2=None, sb_mask=None, m=0):
        n = e.shape[0]
        o = _outs(ch, n)
        if self.device:
            import torch
            dev = lambda a: None if a is None else torch.from_numpy(a).cuda()

            def dev_off(a):
                if a is None or not self.misalign:
                    return dev(a)
                flat = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1))
                buf = torch.zeros(flat.numel() + 32, dtype=torch.uint8, device="cuda")
                sl = buf[self.misalign:self.misalign + flat.numel()]
                sl.copy_(flat)
                assert sl.data_ptr() % 16 == self.misalign % 16
                return sl.view(torch.int8 if a.dtype == np.int8 else torch.uint8).view(a.shape)
            ins = [dev_off(e), dev_off(ciph)] + [dev(x) for x in (prev1, prev2, sb_mask)]
            e_, ciph_, prev1_, prev2_, sb_ = ins
            d = {k: dev(v) for k, v in o.items()}
        else:
            e_, ciph_, prev1_, prev2_, sb_ = e, ciph, prev1, prev2, sb_mask
            d = o
        g = d.get
        L = self.lib
        if ch in (0, 1, 9):
            fn = {0: "gmr1b200_bcch_decode_batch", 1: "gmr1b200_ccch_decode_batch", 9: "gmr1b200_xch_dc12_decode_batch"}[ch]
            L.call(fn, g("l2"), e_, g("conv"), g("crc"), n, None)
        elif ch == 2:
            L.call("gmr1b200_facch3_decode_batch", g("l2"), g("bits_s"), e_, ciph_, g("conv"), g("crc"), n, None)
        elif ch == 3:
            L.call("gmr1b200_facch9_decode_batch", g("l2"), g("sacch"), g("status"), e_, ciph_, g("conv"), g("crc"), n, None)
        elif ch in (4, 5, 6):
            L.call("gmr1b200_tch9_decode_batch", g("l2"), g("sacch"), g("status"), e_, ch - 4, ciph_, prev1_, prev2_,
                   g("conv"), n, None)
        elif ch == 7:
            L.call("gmr1b200_rach_decode_batch", g("l2"), e_, sb_, 0, g("conv"), g("crc2"), g("crc"), n, None)
        elif ch == 8:
            L.call("gmr1b200_tch3_decode_batch", g("l2"), g("l2b"), g("bits_s"), e_, ciph_, m, g("conv"), g("conv1"), n, None)
        if self.device:
            import torch
            torch.cuda.synchronize()
            o = {k: v.cpu().numpy() for k, v in d.items()}
        return o
