"""CPU-only: the TCH3 burst-loop state machine of the product (osmo_gmr_b200/csrc/tch3_state.cuh, compiled for the
host by tests/emu) against the reference application on a recorded call.

The driver below is the frame walk the device-side loop performs: process_bcch / rx_bcch /
rx_ccch / rx_tch3 of src/gmr1_rx.c with every *decision* taken by the product's state functions and every piece of
signal processing by the reference's own functions (the oracle), so that any difference from the application's log is
a difference in the state machine.  Bar: identical per frame - burst kinds, CRC results and Viterbi metrics, FACCH3
flushes (plain attempt, ciphered retry), speech frames, the frame the channel is released in."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import recording
import rxlog
from test_call_fixture_cpu import PLAN, REF_BIN
from test_rxsched_gpu import _acquire_like_main

SPS = 4
FRAME = SPS * 24 * 39


class Tch3State(ctypes.Structure):
    _fields_ = [("active", ctypes.c_int32), ("tn", ctypes.c_int32), ("p", ctypes.c_int32), ("ciph", ctypes.c_int32),
                ("energy_dkab", ctypes.c_float), ("energy_burst", ctypes.c_float),
                ("weak_cnt", ctypes.c_int32), ("sync_id", ctypes.c_int32), ("burst_cnt", ctypes.c_int32),
                ("bi_fn", ctypes.c_uint32 * 4)]


def _energy(w):
    """burst_energy, gmr1_rx.c:172-182: sequential float sum over the inner 30/32 of the window"""
    b = len(w) >> 5
    v = w[b:len(w) - b]
    sq = (v.real * v.real + v.imag * v.imag).astype(np.float32)
    return np.float32(np.cumsum(sq, dtype=np.float32)[-1] / np.float32(len(w)))


def _roundf(t):
    return int(np.sign(t) * np.floor(abs(t) + 0.5))


class Tch9State(ctypes.Structure):
    _fields_ = [("active", ctypes.c_int32), ("tn", ctypes.c_int32)]


def _walk(o, emu, bcch, tch, kc, csd=None):
    P = ctypes.c_void_p
    ptr = lambda a: a.ctypes.data_as(P)
    emu.gmr1_emu_tch3_gate.argtypes = [P, ctypes.c_float]
    emu.gmr1_emu_tch3_dkab_result.argtypes = [P, ctypes.c_float, ctypes.c_int]
    emu.gmr1_emu_tch3_init.argtypes = [P, P, P, ctypes.c_float]
    emu.gmr1_emu_tch3_facch_flush_before.argtypes = [P, ctypes.c_int]
    emu.gmr1_emu_tch3_facch_store.argtypes = [P, P, P, ctypes.c_int, ctypes.c_uint32]
    emu.gmr1_emu_tch3_flush_first_try_ciphered.argtypes = [P]
    emu.gmr1_emu_tch3_flush_wants_retry.argtypes = [P, ctypes.c_int]
    emu.gmr1_emu_tch3_flush_done.argtypes = [P, P, ctypes.c_int, ctypes.c_int]
    assert emu.gmr1_emu_tch3_sizeof() == ctypes.sizeof(Tch3State)
    emu.gmr1_emu_tch9_init_from_facch3.argtypes = [P, P, ctypes.c_int]
    emu.gmr1_emu_tch9_avg_magnitude.argtypes = [P]
    st = Tch3State()
    sp = ctypes.addressof(st)
    t9 = Tch9State()
    il = o.interleaver()
    store = np.zeros(416, np.int8)
    (align, ferr), = _acquire_like_main(o, bcch)
    ferr = np.float32(ferr)
    fn, bcch_energy = 0, np.float32(np.nan)
    frames = []

    def flush(rec):
        masks = lambda: np.concatenate([o.a5(1, kc, int(st.bi_fn[i]), 96) for i in range(4)])
        ciph = masks() if emu.gmr1_emu_tch3_flush_first_try_ciphered(sp) else None
        l2, _, crc, conv = o.facch3_decode(store, ciph)
        rec["flush"].append((crc, conv))
        retried = bool(emu.gmr1_emu_tch3_flush_wants_retry(sp, crc))
        if retried:
            l2, _, crc, conv = o.facch3_decode(store, masks())
            rec["flush"].append((crc, conv))
        good = emu.gmr1_emu_tch3_flush_done(sp, ptr(store), crc, int(retried))
        if csd is not None and emu.gmr1_emu_tch9_init_from_facch3(ctypes.addressof(t9), ptr(l2), good):   # :436-441
            o.c.gmr1_interleaver_init(ctypes.byref(il), 3, 648)

    while True:
        rec = {"fn": fn, "kind": None, "crc": None, "conv": None, "tch": None, "flush": [], "assigned": None, "end": False}
        m = fn & 7                                              # sa_sirfn_delay stays 0: no SI1 in the fixture
        if m == 2:                                              # rx_bcch :747-803
            begin = align - 40
            if begin + 1016 <= len(bcch):
                w = bcch[begin:begin + 1016]
                rc, eb, _, toa, fe = o.demod("bcch", w, SPS, -ferr)
                if rc == 0:
                    bcch_energy = _energy(w)
                    l2, crc, conv = o.simple_decode("bcch", eb)
                    rec.update(kind="bcch", crc=crc, conv=conv)
                    if crc == 0:
                        align += _roundf(toa) - 40
                        ferr = np.float32(ferr + np.float32(fe))
        elif m != 0:                                            # rx_ccch :805-851
            begin = align - 20
            if begin + 976 <= len(bcch):
                w = bcch[begin:begin + 976]
                min_energy = np.float32(bcch_energy / np.float32(2.0))
                if not (_energy(w) < min_energy):
                    rc, eb, _, _, _ = o.demod("dc6", w, SPS, -ferr)
                    if rc == 0:
                        l2, crc, conv = o.simple_decode("ccch", eb)
                        rec.update(kind="ccch", crc=crc, conv=conv)
                        if crc == 0 and l2[1] == 0x06 and l2[2] == 0x3f:
                            emu.gmr1_emu_tch3_init(sp, ptr(store), ptr(l2), float(min_energy))
                            rec["assigned"] = st.tn
        if st.active:                                           # rx_tch3 :538-600
            begin = align + SPS * st.tn * 39 - 3
            if begin + 474 <= len(bcch):
                w = tch[begin:begin + 474]
                be = _energy(w)
                if emu.gmr1_emu_tch3_gate(sp, float(be)) == 1:
                    rv, _, _ = o.dkab_demod(w, SPS, -ferr, st.p)
                    rec["tch"] = "dkab"
                    rec["end"] = bool(emu.gmr1_emu_tch3_dkab_result(sp, float(be), rv))
                else:
                    rc, bt, _, _ = o.detect(["nt3_facch", "nt3_speech"], 3.0, w, SPS, -ferr)
                    assert rc >= 0
                    if bt == 0:
                        rec["tch"], rec["bi"] = "facch3", fn & 3
                        rc, eb, sid, _, _ = o.demod("nt3_facch", w, SPS, -ferr)
                        rec["sync_id"] = sid
                        if emu.gmr1_emu_tch3_facch_flush_before(sp, sid):
                            flush(rec)
                        if emu.gmr1_emu_tch3_facch_store(sp, ptr(store), ptr(eb), sid, fn):
                            flush(rec)
                    else:
                        rec["tch"] = "tch3"
                        rc, eb, _, _, _ = o.demod("nt3_speech", w, SPS, -ferr)
                        ciph = o.a5(st.ciph, kc, fn, 208)
                        f0, f1, _, c0, c1 = o.tch3_decode(eb, ciph, 0)
                        rec.update(frame0=bytes(f0), frame1=bytes(f1), conv0=c0, conv1=c1)
        if t9.active:                                           # rx_tch9 :281-355
            begin = align + SPS * t9.tn * 39 - 3
            if begin + 1410 <= len(bcch):
                w = csd[begin:begin + 1410]
                rc, eb, sid, _, _ = o.demod("nt9", w, SPS, -ferr)
                ciph = o.a5(1, kc, fn, 658)
                rec["csd_sync"] = sid
                if emu.gmr1_emu_tch9_is_facch9(sid):
                    rec["csd"] = "facch9"
                    _, _, _, crc, conv = o.facch9_decode(eb, ciph)
                    rec.update(csd_crc=crc, csd_conv=conv)
                else:
                    rec["csd"] = "tch9"
                    _, _, _, conv = o.tch9_decode(eb, 2, ciph, il)
                    rec.update(conv9=conv, avg=emu.gmr1_emu_tch9_avg_magnitude(ptr(eb)))
        frames.append(rec)
        fn += 1
        align += FRAME
        if align + 2 * FRAME > len(bcch):
            break
    return frames


# a call whose FACCH3 codewords arrive in pieces: two stray FACCH3 bursts between speech bursts, so that the next
# codeword completes the count of four half way (a garbage flush, plain and ciphered attempt), then silence inside the
# call (weak frames that do not add up to a release), DKABs and bursts again
PLAN_RAGGED = "sssss" + "sfsf" + "ffff" + "dd-s-d" + "--s" + "ffff" + "-" * 12


@pytest.mark.parametrize("plan", [PLAN, PLAN_RAGGED])
@pytest.mark.parametrize("key", [None, "0123456789abcdef"])
def test_state_machine_follows_the_reference_application(oracle, emu, tmp_path, key, plan):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()

    def enc_speech(f0, f1, bs, c):
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, np.ascontiguousarray(f0), np.ascontiguousarray(f1),
               np.ascontiguousarray(bs), c, 0)
        return out

    kc = np.frombuffer(bytes.fromhex(key), np.uint8) if key else np.zeros(8, np.uint8)
    b, t, _ = recording.make_call(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2),
                                  enc_speech, lambda l2, bs, c: oracle.facch3_encode(l2, bs, c), plan, tn=7, p=3,
                                  ass_frame=3, kc=kc if key else None, a5=lambda k, fn, n: oracle.a5(1, k, fn, n), seed=5)
    pb, pt = str(tmp_path / "bcch.cfile"), str(tmp_path / "tch.cfile")
    b.tofile(pb)
    t.tofile(pt)
    r = subprocess.run([REF_BIN, "4", pb, pt] + ([key] if key else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = rxlog.parse(r.stderr.split("\n"))
    got = _walk(oracle, emu, b, t, kc)
    assert len(got) == len(ref)
    keys = ("fn", "kind", "crc", "conv", "tch", "flush", "assigned", "end", "bi", "sync_id", "frame0", "frame1", "conv0", "conv1")
    for g, e in zip(got, ref):
        for k in keys:
            assert g.get(k) == e.get(k), (g["fn"], k, g.get(k), e.get(k))
    assert sum(e["tch"] == "tch3" for e in ref) == plan.count("s") and any(e["end"] for e in ref)
    assert sum(len(e["flush"]) for e in ref) >= 3
    if plan is PLAN_RAGGED:
        # a codeword put together from two half codewords that no attempt can decode, and one from three good quarters
        # (rate 1/4: each burst carries one generator's output) that decodes with a large metric
        assert any(e["flush"] and e["flush"][-1][0] != 0 for e in ref)
        assert any(e["flush"] and e["flush"][-1] != (0, 0) and e["flush"][-1][0] == 0 and e["flush"][-1][1] > 1000 for e in ref)


def test_tch9_hand_off_follows_the_reference_application(oracle, emu, tmp_path):
    """the same walk through an ASSIGNMENT COMMAND 1 on the FACCH3 into the TCH9 loop (rx_tch9_init / rx_tch9,
    gmr1_rx.c:264-355): the reference application, given the third recording, starts the TCH9 loop in the frame the
    command's codeword completes, on the timeslot it names, and reports the same Viterbi metric and soft-bit magnitude
    for every frame from there on (bursts, silence, interleaver refill) as the product's state functions driving the
    reference's signal processing."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()

    def enc_speech(f0, f1, bs, c):
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, np.ascontiguousarray(f0), np.ascontiguousarray(f1),
               np.ascontiguousarray(bs), c, 0)
        return out

    key = "0123456789abcdef"
    kc = np.frombuffer(bytes.fromhex(key), np.uint8)
    il_tx = oracle.interleaver()
    rng = np.random.default_rng(3)

    def enc_tch9(fn):
        return oracle.tch9_encode(rng.integers(0, 256, 60, dtype=np.uint8), 2, rng.integers(0, 2, 10, dtype=np.uint8),
                                  rng.integers(0, 2, 4, dtype=np.uint8), oracle.a5(1, kc, fn, 658), il_tx)

    plan = "sssss" + "ffff" + "ssds" + "s" * 12 + "-" * 12
    b, t, _, c = recording.make_call(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2),
                                     enc_speech, lambda l2, bs, cc: oracle.facch3_encode(l2, bs, cc), plan, tn=7, p=3,
                                     ass_frame=3, kc=kc, a5=lambda k, fn, n: oracle.a5(1, k, fn, n), seed=5,
                                     csd=(0, 12, "tttttt-tttt", enc_tch9))
    paths = [str(tmp_path / f"{nm}.cfile") for nm in ("bcch", "tch", "csd")]
    for path, x in zip(paths, (b, t, c)):
        x.tofile(path)
    r = subprocess.run([REF_BIN, "4", paths[0], paths[1], key, paths[2]], capture_output=True, text=True, timeout=300)
    if os.path.exists("/tmp/csd.data"):              # the reference application dumps the TCH9 blocks there (gmr1_rx.c:340-346)
        os.unlink("/tmp/csd.data")
    assert r.returncode == 0, r.stderr[-2000:]
    ref = rxlog.parse(r.stderr.split("\n"))
    got = _walk(oracle, emu, b, t, kc, csd=c)
    assert len(got) == len(ref)
    keys = ("fn", "kind", "crc", "conv", "tch", "flush", "assigned", "end", "frame0", "frame1",
            "csd", "csd_sync", "conv9", "avg", "csd_crc", "csd_conv")
    for g, e in zip(got, ref):
        for k in keys:
            assert g.get(k) == e.get(k), (g["fn"], k, g.get(k), e.get(k))
    first = min(e["fn"] for e in ref if e.get("csd"))
    assert first == 11 and all(e.get("csd") for e in ref if e["fn"] >= first)      # codeword 0 = frames 8..11
    good = [e["conv9"] for e in ref if e.get("conv9") is not None and e["conv9"] < 50]
    assert len(good) >= 6                                                          # blocks decoded through the cipher and the interleaver


def test_state_functions_compile_for_the_device(tmp_path):
    """the same header through nvcc for sm_100a (cross-compiled, no GPU needed): every state function instantiated in
    a kernel, and no fused multiply-add in the running averages (the reference rounds the two products separately)"""
    import shutil
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    src = tmp_path / "t3.cu"
    src.write_text('''
#include "viterbi_tpc.cuh"
#include "tch3_state.cuh"
using namespace gmr1;
__global__ void k(Tch3State *s, Tch9State *s9, int8_t *eb, const int8_t *b, const uint8_t *l2, const float *be, int *out, int n)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	tch3_init(s[i], eb + 416 * i, l2 + 24 * i, be[i]);
	int r = tch3_gate(s[i], be[i]);
	r += tch3_dkab_result(s[i], be[i], r & 1);
	r += tch3_facch_flush_before(s[i], r & 1);
	r += tch3_facch_store(s[i], eb + 416 * i, b + 104 * i, 1, (uint32_t)i);
	r += tch3_flush_first_try_ciphered(s[i]) + tch3_flush_wants_retry(s[i], r & 1);
	r += tch3_flush_done(s[i], eb + 416 * i, r & 1, true);
	r += tch9_init_from_facch3(s9[i], l2 + 24 * i, true) + tch9_is_facch9(r & 1) + tch9_avg_magnitude(b + 662 * i);
	out[i] = r;
}
''')
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "osmo_gmr_b200", "csrc")
    obj = str(tmp_path / "t3.o")
    subprocess.check_call([nvcc, "-std=c++17", "--expt-relaxed-constexpr", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-I" + csrc, "-c", str(src), "-o", obj])
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    assert "FMUL" in sass and "FFMA" not in sass


@pytest.mark.parametrize("seed", range(1, 9))
def test_random_calls_follow_the_reference_application(oracle, emu, tmp_path, seed):
    """random traffic plans, timeslots, DKAB positions, Es/N0 (10 / 14 / 22 dB) and carrier offsets, ciphered: whatever
    the reference application makes of such a call (including what it misclassifies while its thresholds settle), the
    walk driven by the product's state functions makes the same of it, frame for frame"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()

    def enc_speech(f0, f1, bs, c):
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, np.ascontiguousarray(f0), np.ascontiguousarray(f1),
               np.ascontiguousarray(bs), c, 0)
        return out

    key = "a1b2c3d4e5f60718"
    kc = np.frombuffer(bytes.fromhex(key), np.uint8)
    rng = np.random.default_rng(seed)
    plan = "".join(rng.choice(list("sssfd-"), 34)) + "-" * 11
    esn0, cfo = float(rng.choice([10.0, 14.0, 22.0])), float(rng.uniform(-400, 400))
    b, t, _ = recording.make_call(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2),
                                  enc_speech, lambda l2, bs, c: oracle.facch3_encode(l2, bs, c), plan,
                                  tn=int(rng.integers(0, 21)), p=int(rng.integers(0, 40)), ass_frame=3, kc=kc,
                                  a5=lambda k, fn, n: oracle.a5(1, k, fn, n), seed=seed, esn0_db=esn0, cfo_hz=cfo)
    pb, pt = str(tmp_path / "bcch.cfile"), str(tmp_path / "tch.cfile")
    b.tofile(pb)
    t.tofile(pt)
    r = subprocess.run([REF_BIN, "4", pb, pt, key], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    ref = rxlog.parse(r.stderr.split("\n"))
    got = _walk(oracle, emu, b, t, kc)
    assert len(got) == len(ref) and any(e["tch"] for e in ref)
    keys = ("fn", "kind", "crc", "conv", "tch", "flush", "assigned", "end", "bi", "sync_id", "frame0", "frame1", "conv0", "conv1")
    for g, e in zip(got, ref):
        for k in keys:
            assert g.get(k) == e.get(k), (g["fn"], k, g.get(k), e.get(k))
