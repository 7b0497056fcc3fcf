"""Synthetic workloads of BASELINE.json configs 3, 4 and 5 on the GPU, shared by bench.py and the full-size parity
tests.  TEST / BENCH INFRASTRUCTURE (drives the product's C ABI; the only oracle use is in the check_* helpers).

Every workload holds its IQ in device memory (made by the library's burst synthesiser), runs the receive chain
through the batched C entry points, and can hand the FIRST n_check units of every burst type - encoded from valid
payloads, so CRC results mean something - to the CPU reference (oracle/harness.c loops around the reference's own
functions) for a unit-for-unit comparison.  The units behind the checked head carry random hard bits: the kernels'
work does not depend on the payload.

  config 3  NT3 traffic: 75 % speech bursts (TCH3, half of them A5/1-ciphered with masks made on the device),
            25 % FACCH3 bursts in aligned groups of 4, sync sequence alternating per group
  config 4  60 % NT9 (half FACCH9, half TCH9-9k6 in chains of 3 consecutive bursts per channel), 40 % RACH, plus one
            330 ms FCCH search per ARFCN over a grid of 5 frequency shifts and one fine estimate
  config 5  per ARFCN a 1 s recording slice: FCCH search window + 25 bursts (4 BCCH, 12 CCCH, 9 NT3 speech) cut
            by window offsets; processed chunk by chunk, device-resident or streamed from pinned host memory
"""
import numpy as np

SPS = 4
BT = {"bcch": 0, "dc6": 2, "nt3_speech": 4, "nt3_facch": 5, "nt9": 7, "rach": 8}
WIN = {"bcch": 80, "dc6": 40, "nt3_speech": 6, "nt3_facch": 6, "nt9": 6, "rach": 6}
FCCH_WIN = (330 * 23400 * SPS) // 1000            # 30 888 samples
GRID = np.array([-0.54, -0.27, 0.0, 0.27, 0.54], np.float32)      # +-2, +-1, 0 kHz in rad/symbol at 23.4 ksym/s


def geom(L, name):
    ln, eb = L.c.gmr1b200_burst_len(BT[name]), L.c.gmr1b200_burst_ebits(BT[name])
    return ln * SPS + WIN[name], eb


def demod_bytes(L, name):
    """algorithmic bytes per burst of the demod kernel: window in + soft bits + 16 B of metadata out"""
    wl, eb = geom(L, name)
    return 8 * wl + eb + 16


def _synth(L, torch, dev, name, hard_head, n, seed, sync_id=None, esn0=15.0, cfo=0.005):
    """n windows of burst type `name`; the first len(hard_head) carry those hard bits, the rest random ones"""
    wl, eb = geom(L, name)
    g = torch.Generator(device=dev).manual_seed(seed)
    hard = torch.randint(0, 2, (n, eb), dtype=torch.uint8, device=dev, generator=g)
    if hard_head is not None and len(hard_head):
        hard[:len(hard_head)] = torch.from_numpy(np.ascontiguousarray(hard_head)).to(dev)
    toa = torch.rand(n, device=dev, generator=g) * (WIN[name] - 3) + 1.5
    cf = (torch.rand(n, device=dev, generator=g) - 0.5) * 2 * cfo
    ph = torch.rand(n, device=dev, generator=g) * 6.28
    sid = None if sync_id is None else torch.from_numpy(np.ascontiguousarray(sync_id, np.int32)).to(dev)
    iq = torch.empty((n, wl, 2), dtype=torch.float32, device=dev)
    L.call("gmr1b200_synth_bursts", BT[name], hard, eb, sid, SPS, wl, toa, 0.0, cf, 0.0, ph, 0.0, None, esn0,
           None, 1.0, seed, iq, n * wl, None, wl, n, None)
    torch.cuda.synchronize()
    return iq


def fcch_windows_gpu(torch, dev, n, seed, cfo_max=0.27):
    """n FCCH search windows on the device: noise + one dual chirp at a random offset and carrier offset"""
    g = torch.Generator(device=dev).manual_seed(seed)
    x = 0.3 * torch.randn((n, FCCH_WIN, 2), dtype=torch.float32, device=dev, generator=g)
    pos = torch.randint(600, FCCH_WIN - 1200, (n,), device=dev, generator=g)
    cfo = (torch.rand(n, device=dev, generator=g) - 0.5) * 2 * cfo_max
    k = torch.arange(117 * SPS, device=dev, dtype=torch.float32)
    t = k / SPS - 58.5
    chirp = (2.0 ** 0.5) * torch.cos(0.32 * 2 * np.pi / 117 * t * t)
    ang = cfo[:, None] * (k[None, :] / SPS)
    idx = pos[:, None] + torch.arange(117 * SPS, device=dev)[None, :]
    rows = torch.arange(n, device=dev)[:, None].expand_as(idx)
    x[rows, idx, 0] += chirp[None, :] * torch.cos(ang)
    x[rows, idx, 1] += chirp[None, :] * torch.sin(ang)
    return x, pos.cpu().numpy(), cfo.cpu().numpy()


class Timer:
    """CUDA-event timing of callables on the current stream (kernel-only numbers, serial)"""

    def __init__(self, torch):
        self.torch, self.ms = torch, {}

    def run(self, name, fn, reps):
        t = self.torch
        fn()
        t.cuda.synchronize()
        a, b = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        t.cuda.synchronize()
        self.ms[name] = a.elapsed_time(b) / reps
        return self.ms[name]


class Config3:
    def __init__(self, L, torch, dev, arfcns=4096, frames=128, n_check=4096, seed=2000):
        self.L, self.torch, self.dev = L, torch, dev
        n = arfcns * frames
        self.n_sp, self.n_fa = n * 3 // 4, (n // 4) // 4 * 4
        self.n_check = n_check = min(n_check, self.n_sp, self.n_fa) // 4 * 4
        rng = np.random.default_rng(seed)
        e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=dev)
        # cipher: A5/1 for odd bursts, A5/0 for even ones; the masks are made on the device from (Kc, fn)
        self.keys = torch.randint(0, 256, (self.n_sp, 8), dtype=torch.uint8, device=dev)
        self.fn = torch.randint(0, 1 << 19, (self.n_sp,), dtype=torch.int32, device=dev)
        self.alg = (torch.arange(self.n_sp, device=dev) % 2).to(torch.int32)
        self.ciph = e(self.n_sp, 208)
        self.masks()
        torch.cuda.synchronize()
        ciph_head = self.ciph[:n_check].cpu().numpy()
        self.ciph_head = ciph_head
        # valid payloads for the checked head
        self.f0 = rng.integers(0, 256, (n_check, 10), dtype=np.uint8)
        self.f1 = rng.integers(0, 256, (n_check, 10), dtype=np.uint8)
        self.bs = rng.integers(0, 2, (n_check, 4), dtype=np.uint8)
        hard = np.zeros((n_check, 212), np.uint8)
        for i in range(n_check):
            L.call("gmr1b200_tch3_encode", hard[i], self.f0[i], self.f1[i], self.bs[i], ciph_head[i], 0)
        self.iq_sp = _synth(L, torch, dev, "nt3_speech", hard, self.n_sp, seed + 1)
        g_check = n_check // 4
        self.fl2 = rng.integers(0, 256, (g_check, 10), dtype=np.uint8)
        self.fl2[:, 9] &= 0x0F                               # 76 payload bits
        self.fbs = rng.integers(0, 2, (g_check, 32), dtype=np.uint8)
        hard = np.zeros((g_check, 416), np.uint8)
        for i in range(g_check):
            L.call("gmr1b200_facch3_encode", hard[i], self.fl2[i], self.fbs[i], None)
        sid = (np.arange(self.n_fa) // 4) % 2                # sync sequence alternates from one FACCH3 message to the next
        self.fa_sid = sid
        self.iq_fa = _synth(L, torch, dev, "nt3_facch", hard.reshape(-1, 104), self.n_fa, seed + 2, sync_id=sid)
        self.wl, self.eb_sp = geom(L, "nt3_speech")
        _, self.eb_fa = geom(L, "nt3_facch")
        self.ebits_sp, self.ebits_fa = e(self.n_sp, self.eb_sp, dt=torch.int8), e(self.n_fa, self.eb_fa, dt=torch.int8)
        self.o_f0, self.o_f1, self.o_bs = e(self.n_sp, 10), e(self.n_sp, 10), e(self.n_sp, 4)
        self.o_l2, self.o_bs3 = e(self.n_fa // 4, 10), e(self.n_fa // 4, 32)
        self.o_crc = e(self.n_fa // 4, dt=torch.int32)
        self.o_sid = e(self.n_fa, dt=torch.int32)

    n_bursts = property(lambda self: self.n_sp + self.n_fa)
    iq_bytes = property(lambda self: (self.n_sp + self.n_fa) * self.wl * 8)

    def masks(self, st=None):
        self.L.call("gmr1b200_a5_batch", self.alg, 0, self.keys, self.fn, 208, 208, self.ciph, None, self.n_sp, st)

    def steps(self, st=None):
        L, wl = self.L, self.wl
        return {
            "a5_masks": lambda: self.masks(st),
            "demod_nt3_speech": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt3_speech"], self.iq_sp, self.n_sp * wl,
                                               None, wl, wl, SPS, None, 0.0, self.ebits_sp, self.eb_sp, None, None, None,
                                               None, self.n_sp, st),
            "decode_tch3": lambda: L.call("gmr1b200_tch3_decode_batch", self.o_f0, self.o_f1, self.o_bs, self.ebits_sp,
                                          self.ciph, 0, None, None, self.n_sp, st),
            "demod_nt3_facch": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt3_facch"], self.iq_fa, self.n_fa * wl,
                                              None, wl, wl, SPS, None, 0.0, self.ebits_fa, self.eb_fa, self.o_sid, None,
                                              None, None, self.n_fa, st),
            "decode_facch3": lambda: L.call("gmr1b200_facch3_decode_batch", self.o_l2, self.o_bs3, self.ebits_fa, None, None,
                                            self.o_crc, self.n_fa // 4, st),
        }

    def run(self, st=None):
        for f in self.steps(st).values():
            f()

    def head(self):
        """results of the checked head on the host"""
        n, g = self.n_check, self.n_check // 4
        c = lambda t, k: t[:k].cpu().numpy()
        return dict(f0=c(self.o_f0, n), f1=c(self.o_f1, n), bs=c(self.o_bs, n), l2=c(self.o_l2, g), bs3=c(self.o_bs3, g),
                    crc=c(self.o_crc, g), sid=c(self.o_sid, n))

    def head_iq(self):
        n = self.n_check
        v = lambda t: t[:n].cpu().numpy().view(np.complex64).reshape(n, self.wl)
        return {"tch3": (v(self.iq_sp), {"ciph": self.ciph_head}), "facch3": (v(self.iq_fa), {})}

    def compare(self, head, cpu):
        """unit-for-unit comparison of the GPU results with the CPU reference's on the checked head"""
        f0, f1, bs, _ = cpu["tch3"]
        l2, bs3, crc, sid = cpu["facch3"]
        same = {"tch3": bool((head["f0"] == f0).all() and (head["f1"] == f1).all() and (head["bs"] == bs).all()),
                "facch3": bool((head["l2"] == l2).all() and (head["crc"] == crc).all() and (head["bs3"] == bs3).all()
                               and (head["sid"] == sid).all())}
        prot = (head["f0"][:, :6] == self.f0[:, :6]).all(axis=1) & (head["f1"][:, :6] == self.f1[:, :6]).all(axis=1)
        ok3 = head["crc"] == 0
        info = {"tch3_protected_bits_recovered_frac": float(prot.mean()),
                "facch3_crc_ok_frac": float(ok3.mean()),
                "facch3_payload_ok_where_crc_ok": bool((head["l2"][ok3] == self.fl2[ok3]).all()),
                "note": "the reference's never-cleared sync accumulator answers sequence 1 for every NT3-FACCH burst "
                        "(pi4cxpsk.c:207,232), so it decodes only the FACCH3 messages sent with sequence 1 (every second "
                        "one here); the GPU path reproduces that"}
        return same, info


class Config4:
    def __init__(self, L, torch, dev, arfcns=8192, per=64, n_check=4096, seed=3000, fcch=True):
        self.L, self.torch, self.dev = L, torch, dev
        n = arfcns * per
        n9 = n * 6 // 10 // 6 * 6
        self.n_f9, self.n_t9, self.n_ra = n9 // 2, n9 - n9 // 2, n - n9
        self.n_t9 = self.n_t9 // 3 * 3
        self.arfcns = arfcns
        self.n_check = n_check = min(n_check, self.n_f9, self.n_t9, self.n_ra) // 3 * 3
        rng = np.random.default_rng(seed)
        e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=dev)
        # FACCH9 head
        self.l2_f9 = rng.integers(0, 256, (n_check, 38), dtype=np.uint8)
        self.l2_f9[:, 37] &= 0x0F                            # 300 payload bits
        hard = np.zeros((n_check, 662), np.uint8)
        sa, stt = rng.integers(0, 2, (n_check, 10), dtype=np.uint8), rng.integers(0, 2, (n_check, 4), dtype=np.uint8)
        for i in range(n_check):
            L.call("gmr1b200_facch9_encode", hard[i], self.l2_f9[i], sa[i], stt[i], None)
        # every NT9 burst is sent with training sequence 1: it is the only one the reference's default sync search
        # (accumulator never cleared between candidates) can lock to
        self.iq_f9 = _synth(L, torch, dev, "nt9", hard, self.n_f9, seed + 1, sync_id=np.ones(self.n_f9, np.int32))
        # TCH9-9k6 head: chains of 3 consecutive bursts per channel through the depth-3 interleaver
        self.l2_t9 = rng.integers(0, 256, (n_check, 60), dtype=np.uint8)
        hard = np.zeros((n_check, 662), np.uint8)
        for c in range(n_check // 3):
            il = L.c.gmr1b200_tch9_interleaver_new()
            for k in range(3):
                i = 3 * c + k
                L.call("gmr1b200_tch9_encode", hard[i], self.l2_t9[i], 2, sa[i], stt[i], None, il)
            L.c.gmr1b200_tch9_interleaver_free(il)
        self.iq_t9 = _synth(L, torch, dev, "nt9", hard, self.n_t9, seed + 2, sync_id=np.ones(self.n_t9, np.int32))
        idx = torch.arange(self.n_t9, device=dev, dtype=torch.int32)
        self.prev1 = torch.where(idx % 3 >= 1, idx - 1, torch.full_like(idx, -1))
        self.prev2 = torch.where(idx % 3 >= 2, idx - 2, torch.full_like(idx, -1))
        # RACH head
        self.rach = rng.integers(0, 256, (n_check, 18), dtype=np.uint8)
        self.rach[:, 17] &= 0x07                             # 139 payload bits
        self.sb = rng.integers(0, 256, self.n_ra).astype(np.uint8)
        hard = np.zeros((n_check, 494), np.uint8)
        for i in range(n_check):
            L.call("gmr1b200_rach_encode", hard[i], self.rach[i], int(self.sb[i]))
        self.iq_ra = _synth(L, torch, dev, "rach", hard, self.n_ra, seed + 3)
        self.sb_d = torch.from_numpy(self.sb).to(dev)
        self.wl9, self.eb9 = geom(L, "nt9")
        self.wlr, self.ebr = geom(L, "rach")
        self.eb_f9, self.eb_t9 = e(self.n_f9, self.eb9, dt=torch.int8), e(self.n_t9, self.eb9, dt=torch.int8)
        self.eb_ra = e(self.n_ra, self.ebr, dt=torch.int8)
        self.o_l2f, self.o_crcf = e(self.n_f9, 38), e(self.n_f9, dt=torch.int32)
        self.o_l2t = e(self.n_t9, 60)
        self.o_rach, self.o_crcr = e(self.n_ra, 18), e(self.n_ra, dt=torch.int32)
        self.o_sidf = e(self.n_f9, dt=torch.int32)
        self.fcch = fcch
        if fcch:
            self.fw, self.f_pos, self.f_cfo = fcch_windows_gpu(torch, dev, arfcns, seed + 4, cfo_max=0.54)
            self.g_toa = e(len(GRID), arfcns, dt=torch.int32)
            self.g_peak = e(len(GRID), arfcns, dt=torch.float32)
            self.f_toa, self.f_ferr = e(arfcns, dt=torch.int32), e(arfcns, dt=torch.float32)

    n_bursts = property(lambda self: self.n_f9 + self.n_t9 + self.n_ra)
    iq_bytes = property(lambda self: (self.n_f9 + self.n_t9) * self.wl9 * 8 + self.n_ra * self.wlr * 8 +
                        (self.arfcns * FCCH_WIN * 8 if self.fcch else 0))

    def fcch_search(self, st=None):
        """rough search at every shift of the grid (the +-frequency-offset FCCH search), then the fine estimate at
        the position the strongest hypothesis found"""
        L, W, n = self.L, FCCH_WIN, self.arfcns
        if hasattr(L.c, "gmr1b200_fcch_rough_grid_batch"):
            L.call("gmr1b200_fcch_rough_grid_batch", 0, self.fw, n * W, None, W, W, SPS, GRID, len(GRID), self.g_toa,
                   self.g_peak, n, st)
        else:
            for k, fs in enumerate(GRID):
                L.call("gmr1b200_fcch_rough_batch", 0, self.fw, n * W, None, W, W, SPS, None, float(fs), self.g_toa[k],
                       self.g_peak[k], n, st)
        L.call("gmr1b200_fcch_fine_batch", 0, self.fw, n * W, None, W, SPS, None, 0.0, self.f_toa, self.f_ferr, n, st)

    def steps(self, st=None):
        L = self.L
        d = {
            "demod_nt9_facch9": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt9"], self.iq_f9, self.n_f9 * self.wl9, None,
                                               self.wl9, self.wl9, SPS, None, 0.0, self.eb_f9, self.eb9, self.o_sidf, None, None,
                                               None, self.n_f9, st),
            "decode_facch9": lambda: L.call("gmr1b200_facch9_decode_batch", self.o_l2f, None, None, self.eb_f9, None, None,
                                            self.o_crcf, self.n_f9, st),
            "demod_nt9_tch9": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["nt9"], self.iq_t9, self.n_t9 * self.wl9, None,
                                             self.wl9, self.wl9, SPS, None, 0.0, self.eb_t9, self.eb9, None, None, None, None,
                                             self.n_t9, st),
            "decode_tch9_9k6": lambda: L.call("gmr1b200_tch9_decode_batch", self.o_l2t, None, None, self.eb_t9, 2, None,
                                              self.prev1, self.prev2, None, self.n_t9, st),
            "demod_rach": lambda: L.call("gmr1b200_pi4cxpsk_demod_batch", BT["rach"], self.iq_ra, self.n_ra * self.wlr, None,
                                         self.wlr, self.wlr, SPS, None, 0.0, self.eb_ra, self.ebr, None, None, None, None,
                                         self.n_ra, st),
            "decode_rach": lambda: L.call("gmr1b200_rach_decode_batch", self.o_rach, self.eb_ra, self.sb_d, 0, None, None,
                                          self.o_crcr, self.n_ra, st),
        }
        if self.fcch:
            d["fcch_5_shift_search_and_fine"] = lambda: self.fcch_search(st)
        return d

    def run(self, st=None):
        for f in self.steps(st).values():
            f()

    def head(self):
        n = self.n_check
        c = lambda t, k=n: t[:k].cpu().numpy()
        h = dict(l2f=c(self.o_l2f), crcf=c(self.o_crcf), sidf=c(self.o_sidf), l2t=c(self.o_l2t), rach=c(self.o_rach),
                 crcr=c(self.o_crcr))
        if self.fcch:
            h["g_toa"] = self.g_toa[:, :self.n_fcch_check()].cpu().numpy()
        return h

    def n_fcch_check(self):
        return min(self.arfcns, 256)

    def head_iq(self):
        n = self.n_check
        v = lambda t, wl: t[:n].cpu().numpy().view(np.complex64).reshape(n, wl)
        d = {"facch9": (v(self.iq_f9, self.wl9), {}), "tch9": (v(self.iq_t9, self.wl9), {"n_burst": 3}),
             "rach": (v(self.iq_ra, self.wlr), {"sb_mask": self.sb[:n]})}
        if self.fcch:
            m = self.n_fcch_check()
            d["fcch_grid"] = (self.fw[:m].cpu().numpy().view(np.complex64).reshape(m, FCCH_WIN), {"shifts": GRID})
        return d

    def compare(self, head, cpu):
        l2f, crcf, sidf = cpu["facch9"]
        l2t, _ = cpu["tch9"]
        rach, crcr = cpu["rach"]
        same = {"facch9": bool((head["l2f"] == l2f).all() and (head["crcf"] == crcf).all() and (head["sidf"] == sidf).all()),
                "tch9_9k6": bool((head["l2t"] == l2t).all()),
                "rach": bool((head["rach"] == rach).all() and ((head["crcr"] != 0) == (crcr != 0)).all())}
        okf, okr = head["crcf"] == 0, head["crcr"] == 0
        # a TCH9 block comes out of the depth-3 de-interleaver complete with the third burst of its chain
        t9 = head["l2t"].reshape(-1, 3, 60)[:, 2], self.l2_t9.reshape(-1, 3, 60)[:, 0]
        info = {"facch9_crc_ok_frac": float(okf.mean()), "facch9_payload_ok_where_crc_ok": bool((head["l2f"][okf] == self.l2_f9[okf]).all()),
                "rach_crc_ok_frac": float(okr.mean()), "rach_payload_ok_where_crc_ok": bool((head["rach"][okr] == self.rach[okr]).all()),
                "tch9_first_block_recovered_frac": float((t9[0] == t9[1]).all(axis=1).mean())}
        if "fcch_grid" in cpu:
            d = np.abs(head["g_toa"].astype(np.int64) - cpu["fcch_grid"])
            same["fcch_grid"] = bool((d <= 1).all() and (d == 0).mean() >= 0.97)
        return same, info


# ---- config 5: 1 s recording slices, FCCH + 25 bursts per ARFCN ---------------------------------------------------
SLICE = 23400 * SPS                                   # samples per 1 s slice (748.8 KB)
MIX = ["bcch" if j % 8 == 0 else ("dc6" if j % 2 else "nt3_speech") for j in range(25)]      # 4 BCCH, 12 CCCH, 9 NT3
BURST_OFS = [31200 + 2400 * j for j in range(25)]     # even sample offsets behind the FCCH search window


class Config5Chunk:
    """one chunk of `c` ARFCN slices in device memory + everything to process it: the FCCH acquisition of every
    slice and the demodulation + decode of its 25 bursts, addressed in place by window offsets"""

    def __init__(self, L, torch, dev, c=1024, seed=5000):
        self.L, self.torch, self.dev, self.c = L, torch, dev, c
        rng = np.random.default_rng(seed)
        self.iq = torch.zeros((c, SLICE, 2), dtype=torch.float32, device=dev)
        fw, self.f_pos, _ = fcch_windows_gpu(torch, dev, c, seed + 1)
        self.iq[:, :FCCH_WIN] = fw
        del fw
        self.kinds = {}
        for name, chan in (("bcch", 0), ("dc6", 1), ("nt3_speech", None)):
            js = [j for j in range(25) if MIX[j] == name]
            n = c * len(js)
            wl, eb = geom(L, name)
            ofs = (np.arange(c, dtype=np.int64)[:, None] * SLICE + np.array([BURST_OFS[j] for j in js], np.int64)[None, :]).reshape(-1)
            hard = np.zeros((n, eb), np.uint8)
            l2 = None
            if chan is not None:
                l2 = rng.integers(0, 256, (n, 24), dtype=np.uint8)
                L.call("gmr1b200_xcch_encode_batch", chan, hard, l2, n)
            else:
                hard = rng.integers(0, 2, (n, eb), dtype=np.uint8)
            d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
            g = torch.Generator(device=dev).manual_seed(seed + 7 + len(js))
            toa = torch.rand(n, device=dev, generator=g) * (WIN[name] - 3) + 1.5
            ofs_d = d(ofs)
            L.call("gmr1b200_synth_bursts", BT[name], d(hard), eb, None, SPS, wl, toa, 0.0, None, 0.0, None, 0.5, None, 15.0,
                   None, 1.0, seed + len(js), self.iq, c * SLICE, ofs_d, 0, n, None)
            e = lambda *s, dt=torch.uint8: torch.empty(s, dtype=dt, device=dev)
            k = dict(n=n, per=len(js), wl=wl, eb=eb, ofs=ofs_d, chan=chan, l2_true=l2)
            if chan is not None:
                k.update(l2=[e(n, 24) for _ in range(2)], crc=[e(n, dt=torch.int32) for _ in range(2)])
            else:
                k.update(ebits=[e(n, eb, dt=torch.int8) for _ in range(2)], f0=[e(n, 10) for _ in range(2)],
                         f1=[e(n, 10) for _ in range(2)])
            self.kinds[name] = k
        torch.cuda.synchronize()
        i32 = lambda: [torch.empty(c, dtype=torch.int32, device=dev) for _ in range(2)]
        self.a_rough, self.a_align = i32(), i32()
        self.a_ferr = [torch.empty(c, dtype=torch.float32, device=dev) for _ in range(2)]
        self.bursts = sum(k["n"] for k in self.kinds.values())

    def process(self, iq, st, c=None, slot=0):
        """the receive chain over the first c slices of a chunk that lies at `iq` (this object's own tensor or a streamed
        copy); results go to result set `slot` (one per stream)"""
        L = self.L
        c = self.c if c is None else c
        L.call("gmr1b200_fcch_acquire_batch", 0, iq, self.c * SLICE, None, SLICE, FCCH_WIN, SPS, self.a_rough[slot],
               self.a_align[slot], self.a_ferr[slot], c, st)
        for name, k in self.kinds.items():
            n = c * k["per"]
            if k["chan"] is not None:
                L.call("gmr1b200_rx_xcch_batch", k["chan"], iq, self.c * SLICE, k["ofs"], 0, k["wl"], SPS, None, 0.0,
                       k["l2"][slot], k["crc"][slot], None, None, None, n, st)
            else:
                L.call("gmr1b200_pi4cxpsk_demod_batch", BT[name], iq, self.c * SLICE, k["ofs"], 0, k["wl"], SPS, None, 0.0,
                       k["ebits"][slot], k["eb"], None, None, None, None, n, st)
                L.call("gmr1b200_tch3_decode_batch", k["f0"][slot], k["f1"][slot], None, k["ebits"][slot], None, 0, None, None,
                       n, st)

    def result_buffers(self, pinned=True):
        t = self.torch
        mk = lambda x: (t.empty(x.shape, dtype=x.dtype).pin_memory() if pinned else t.empty(x.shape, dtype=x.dtype))
        d = {"align": mk(self.a_align[0]), "ferr": mk(self.a_ferr[0])}
        for name, k in self.kinds.items():
            for key in ("l2", "crc", "f0", "f1"):
                if key in k:
                    d[name + "_" + key] = mk(k[key][0])
        return d

    def results_to_host(self, host, c, slot):
        """what a caller gets back per chunk: alignment + frequency error per slice, L2 + CRC / speech frames per burst"""
        host["align"][:c].copy_(self.a_align[slot][:c], non_blocking=True)
        host["ferr"][:c].copy_(self.a_ferr[slot][:c], non_blocking=True)
        for name, k in self.kinds.items():
            n = c * k["per"]
            for key in ("l2", "crc", "f0", "f1"):
                if key in k:
                    host[name + "_" + key][:n].copy_(k[key][slot][:n], non_blocking=True)

    def sane(self):
        """the chunk decodes: FCCH found at the planted position, control-channel CRCs pass, payloads recovered"""
        ok = True
        for slot in range(2):
            ok = ok and bool((np.abs(self.a_align[slot].cpu().numpy() - self.f_pos) <= 2).mean() > 0.99)
            for name, k in self.kinds.items():
                if k["chan"] is not None:
                    crc = k["crc"][slot].cpu().numpy()
                    ok = ok and (crc == 0).mean() > 0.98
                    ok = ok and bool((k["l2"][slot].cpu().numpy()[crc == 0] == k["l2_true"][crc == 0]).all())
        return bool(ok)
