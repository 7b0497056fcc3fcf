"""The whole receiver frame loop on the device (gmr1b200_rx_call_batch, csrc/rx_call.cu: control channels, the TCH3
burst loop behind an IMMEDIATE ASSIGNMENT, the TCH9 loop behind an ASSIGNMENT COMMAND 1) against the reference's own
application (oracle/_ref/gmr1_rx = src/gmr1_rx.c, run once per call) on recorded calls made by tests/recording.py:
plain and ciphered, a regular and a ragged plan, eight random calls, and a call that is handed over to a TCH9.

Bar, per channel and frame: the same frame numbers, control-channel burst kinds and CRCs, the same classification of
the traffic window (DKAB found / missed, FACCH3 burst, speech burst), the same release frame, the same FACCH3 sync
ids, the same number of FACCH3 decode attempts with the same CRC results (plain attempt, ciphered retry), the same
speech frames; Viterbi metrics agree to the soft-bit tolerance of the demodulator (the float stage is within +-1 LSB
of the reference's soft bits, which moves a metric by a few units at most).  Several calls run in ONE batch, so the
per-channel state, the compacted lists and the device-side counts are exercised with channels in different phases."""
import os
import subprocess

import numpy as np
import pytest

import recording
import rxlog
from test_call_fixture_cpu import PLAN, REF_BIN
from test_rxsched_gpu import _acquire_like_main
from test_tch3_state_emu import PLAN_RAGGED

pytestmark = pytest.mark.gpu
SPS = 4
TCH = {0: None, 1: "dkab", 2: "dkab", 3: "facch3", 4: "tch3"}
CONV_TOL = 8


def _enc_speech(L):
    def f(f0, f1, bs, c):
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, np.ascontiguousarray(f0), np.ascontiguousarray(f1),
               np.ascontiguousarray(bs), c, 0)
        return out
    return f


def _reference(tmp_path, tag, b, t, key, c=None):
    paths = [str(tmp_path / f"{tag}_{nm}.cfile") for nm in ("bcch", "tch", "csd")]
    b.tofile(paths[0])
    t.tofile(paths[1])
    cmd = [REF_BIN, "4", paths[0], paths[1]]
    if key or c is not None:
        cmd.append(key or "0000000000000000")
    if c is not None:
        c.tofile(paths[2])
        cmd.append(paths[2])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    if os.path.exists("/tmp/csd.data"):              # the reference application dumps the TCH9 blocks there (gmr1_rx.c:340-346)
        os.unlink("/tmp/csd.data")
    assert r.returncode == 0, r.stderr[-2000:]
    return rxlog.parse(r.stderr.split("\n"))


def _run_batch(L, oracle, calls, F=80):
    """calls: list of (bcch, tch, kc bytes or None, csd or None) -> per call the frame records of the device walk"""
    n = len(calls)
    parts, rec_ofs, tch_ofs, csd_ofs, rec_len, align0, ferr0 = [], [], [], [], [], [], []
    pos = 0
    for b, t, kc, c in calls:
        (a, fe), = _acquire_like_main(oracle, b)
        assert len(t) == len(b) and (c is None or len(c) == len(b))
        rec_ofs.append(pos)
        tch_ofs.append(pos + len(b))
        csd_ofs.append(pos + 2 * len(b) if c is not None else -1)
        rec_len.append(len(b))
        align0.append(a)
        ferr0.append(fe)
        parts += [b, t] + ([c] if c is not None else [])
        pos += len(b) * (3 if c is not None else 2)
    iq = np.ascontiguousarray(np.concatenate(parts).astype(np.complex64)).view(np.float32)
    kcs = np.stack([np.zeros(8, np.uint8) if kc is None else np.asarray(kc, np.uint8) for _, _, kc, _ in calls])
    i32 = lambda *s: np.zeros(s, np.int32)
    kind, fn, crc, conv, nfr = i32(n, F), i32(n, F), i32(n, F), i32(n, F), i32(n)
    l2 = np.zeros((n, F, 24), np.uint8)
    trec, tdat = i32(n, F, 12), np.zeros((n, F, 20), np.uint8)
    crec, cdat = i32(n, F, 6), np.zeros((n, F, 60), np.uint8)
    L.call("gmr1b200_rx_call_batch", iq, len(iq) // 2, np.array(rec_ofs, np.int64), np.array(rec_len, np.int32),
           np.array(tch_ofs, np.int64), np.array(csd_ofs, np.int64), kcs, np.array(align0, np.int32),
           np.array(ferr0, np.float32), SPS, n, F, kind, fn, crc, conv, l2, nfr, trec, tdat, crec, cdat, None)
    # SURVEY 8f N4, the AMBE hand-off: the speech frames of every call as the vocoder's input stream, built on the device
    # from the records above; checked here against those records, by the callers against the reference's log
    voice, flag = np.zeros((n, 2 * F, 10), np.uint8), np.zeros((n, 2 * F), np.uint8)
    n_voice, first_fn = i32(n), i32(n)
    L.call("gmr1b200_tch3_voice_stream_batch", trec, tdat, fn, nfr, n, F, voice, flag, n_voice, first_fn, None)
    _run_batch.voice = []
    for i in range(n):
        act = [f for f in range(nfr[i]) if trec[i, f, 0] != 0 or trec[i, f, 1]]
        if not act:
            assert n_voice[i] == 0 and first_fn[i] == -1 and (flag[i] == 2).all() and not voice[i].any()
            _run_batch.voice.append(b"")
            continue
        f0, f1 = act[0], act[-1]
        assert n_voice[i] == 2 * (f1 - f0 + 1) and first_fn[i] == fn[i, f0]
        for f in range(f0, f1 + 1):
            sp = trec[i, f, 0] == 4
            for h in (0, 1):
                k = 2 * (f - f0) + h
                assert flag[i, k] == (0 if sp else 1)
                assert bytes(voice[i, k]) == (bytes(tdat[i, f, 10 * h:10 * h + 10]) if sp else bytes(10))
        assert (flag[i, n_voice[i]:] == 2).all() and not voice[i, n_voice[i]:].any()
        _run_batch.voice.append(b"".join(bytes(voice[i, k]) for k in range(n_voice[i]) if flag[i, k] == 0))
    out = []
    for i in range(n):
        frames = []
        for f in range(nfr[i]):
            r, cr = trec[i, f], crec[i, f]
            rec = {"fn": int(fn[i, f]), "kind": {0: None, 1: "bcch", 2: "ccch"}[int(kind[i, f])],
                   "crc": int(crc[i, f]) if kind[i, f] else None, "conv": int(conv[i, f]) if kind[i, f] else None,
                   "tch": TCH[int(r[0])], "dkab_found": int(r[0]) == 1, "end": bool(r[1]), "flush": [],
                   "assigned": int(r[10]) if r[10] >= 0 else None}
            if r[0] == 3:
                rec["sync_id"], rec["bi"] = int(r[2]), rec["fn"] & 3
                rec["flush"] = [(int(r[4 + 2 * k]), int(r[5 + 2 * k])) for k in range(int(r[3]))]
                if r[11]:
                    rec["facch3_l2"] = bytes(tdat[i, f, :10])
            if r[0] == 4:
                rec.update(frame0=bytes(tdat[i, f, :10]), frame1=bytes(tdat[i, f, 10:]), conv0=int(r[8]), conv1=int(r[9]))
            if cr[0]:
                rec["csd"] = {1: "facch9", 2: "tch9"}[int(cr[0])]
                rec["csd_sync"] = int(cr[1])
                if cr[0] == 1:
                    rec.update(csd_crc=int(cr[2]), csd_conv=int(cr[3]))
                else:
                    rec.update(conv9=int(cr[3]), avg=int(cr[4]), block=bytes(cdat[i, f]))
            frames.append(rec)
        out.append(frames)
    return out


def _compare(got, ref, tag):
    assert len(got) == len(ref), (tag, len(got), len(ref))
    near = lambda a, b: a is not None and b is not None and abs(a - b) <= CONV_TOL
    for g, e in zip(got, ref):
        at = (tag, e["fn"])
        for k in ("fn", "kind", "crc", "tch", "assigned", "end"):
            assert g.get(k) == e.get(k), (at, k, g.get(k), e.get(k))
        if e["kind"]:
            assert near(g["conv"], e["conv"]), (at, g["conv"], e["conv"])
        if e["tch"] == "facch3":
            assert g["sync_id"] == e["sync_id"] and g["bi"] == e["bi"], at
            assert len(g["flush"]) == len(e["flush"]), (at, g["flush"], e["flush"])
            for (gc, gv), (ec, ev) in zip(g["flush"], e["flush"]):
                assert gc == ec and near(gv, ev), (at, g["flush"], e["flush"])
        if e["tch"] == "tch3":
            assert near(g["conv0"], e["conv0"]) and near(g["conv1"], e["conv1"]), at
            # 48 convolutionally protected bits per frame arrive exactly; the 32 class-2 bits are hard decisions on
            # single soft bits, where the +-1 LSB of the float stage can turn a soft bit of 0 / -1
            assert g["frame0"][:6] == e["frame0"][:6] and g["frame1"][:6] == e["frame1"][:6], at
            diff = sum(bin(a ^ b).count("1") for a, b in zip(g["frame0"] + g["frame1"], e["frame0"] + e["frame1"]))
            assert diff <= 1, (at, diff)
        if e.get("csd"):
            assert g.get("csd") == e["csd"] and g["csd_sync"] == e["csd_sync"], (at, g.get("csd"), e["csd"])
            if e["csd"] == "facch9":
                assert g["csd_crc"] == e["csd_crc"] and near(g["csd_conv"], e["csd_conv"]), at
            else:
                assert near(g["conv9"], e["conv9"]) and abs(g["avg"] - e["avg"]) <= 1, (at, g["conv9"], e["conv9"])
        else:
            assert not g.get("csd"), at


def test_calls_follow_the_reference_application(gpu_lib, oracle, tmp_path):
    """plain / ciphered x regular / ragged plan, and eight random ciphered calls, all in one batch of 12 channels"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    L = gpu_lib
    enc = (lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2), _enc_speech(L),
           lambda l2, bs, c: oracle.facch3_encode(l2, bs, c))
    a5 = lambda k, fn, n: oracle.a5(1, k, fn, n)
    calls, refs, tags = [], [], []
    for plan in (PLAN, PLAN_RAGGED):
        for key in (None, "0123456789abcdef"):
            kc = np.frombuffer(bytes.fromhex(key), np.uint8) if key else None
            b, t, _ = recording.make_call(*enc, plan, tn=7, p=3, ass_frame=3, kc=kc, a5=a5, seed=5)
            tag = f"plan{len(plan)}_{'c' if key else 'p'}"
            calls.append((b, t, kc, None))
            refs.append(_reference(tmp_path, tag, b, t, key))
            tags.append(tag)
    key = "a1b2c3d4e5f60718"
    kc = np.frombuffer(bytes.fromhex(key), np.uint8)
    for seed in range(1, 9):
        rng = np.random.default_rng(seed)
        plan = "".join(rng.choice(list("sssfd-"), 34)) + "-" * 11
        esn0, cfo = float(rng.choice([10.0, 14.0, 22.0])), float(rng.uniform(-400, 400))
        b, t, _ = recording.make_call(*enc, plan, tn=int(rng.integers(0, 21)), p=int(rng.integers(0, 40)), ass_frame=3,
                                      kc=kc, a5=a5, seed=seed, esn0_db=esn0, cfo_hz=cfo)
        calls.append((b, t, kc, None))
        refs.append(_reference(tmp_path, f"rnd{seed}", b, t, key))
        tags.append(f"rnd{seed}")
    got = _run_batch(L, oracle, calls)
    for g, e, tag in zip(got, refs, tags):
        _compare(g, e, tag)
    # the vocoder streams (what gmr1_ambe_decode would read): the reference's speech frames in order; the protected
    # 6 bytes of every frame exactly (the class-2 bits carry the float stage's +-1 LSB, see _compare)
    for stream, e in zip(_run_batch.voice, refs):
        want = [fr for rec in e if rec["tch"] == "tch3" for fr in (rec["frame0"], rec["frame1"])]
        assert len(stream) == 10 * len(want) and len(want) > 0
        for k, fr in enumerate(want):
            assert stream[10 * k:10 * k + 6] == fr[:6]
    ref0 = refs[0]
    assert sum(e["tch"] == "tch3" for e in ref0) == PLAN.count("s") and any(e["end"] for e in ref0)
    assert sum(len(e["flush"]) for r in refs for e in r) >= 12
    assert any(len(e["flush"]) == 2 for r in refs for e in r)                      # a ciphered retry happened
    # a good FACCH3 message is delivered where the reference's last attempt passed its CRC
    for g, e in zip(got[1], refs[1]):
        assert ("facch3_l2" in g) == bool(e["flush"] and e["flush"][-1][0] == 0)


def test_tch9_hand_off_follows_the_reference_application(gpu_lib, oracle, tmp_path):
    """through an ASSIGNMENT COMMAND 1 on the FACCH3 into the TCH9 loop: same first frame and timeslot, FACCH9 / TCH9
    split, Viterbi metric and soft-bit magnitude of every TCH9 block through the cipher and the depth-3 interleaver;
    batched next to a call without a third recording"""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    L = gpu_lib
    enc = (lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2), _enc_speech(L),
           lambda l2, bs, c: oracle.facch3_encode(l2, bs, c))
    key = "0123456789abcdef"
    kc = np.frombuffer(bytes.fromhex(key), np.uint8)
    a5 = lambda k, fn, n: oracle.a5(1, k, fn, n)
    il_tx = oracle.interleaver()
    rng = np.random.default_rng(3)
    blocks = {}

    def enc_tch9(fn):
        blocks[fn] = rng.integers(0, 256, 60, dtype=np.uint8)
        return oracle.tch9_encode(blocks[fn], 2, rng.integers(0, 2, 10, dtype=np.uint8),
                                  rng.integers(0, 2, 4, dtype=np.uint8), oracle.a5(1, kc, fn, 658), il_tx)

    plan = "sssss" + "ffff" + "ssds" + "s" * 12 + "-" * 12
    b, t, _, c = recording.make_call(*enc, plan, tn=7, p=3, ass_frame=3, kc=kc, a5=a5, seed=5,
                                     csd=(0, 12, "tttttt-tttt", enc_tch9))
    ref = _reference(tmp_path, "t9", b, t, key, c)
    b2, t2, _ = recording.make_call(*enc, PLAN, tn=7, p=3, ass_frame=3, kc=kc, a5=a5, seed=6)
    ref2 = _reference(tmp_path, "t9b", b2, t2, key)
    got = _run_batch(L, oracle, [(b, t, kc, c), (b2, t2, kc, None)])
    _compare(got[0], ref, "t9")
    _compare(got[1], ref2, "t9b")
    first = min(e["fn"] for e in ref if e.get("csd"))
    assert first == 11 and all(g.get("csd") for g in got[0] if g["fn"] >= first)
    good = [g for g in got[0] if g.get("conv9") is not None and g["conv9"] < 50]
    assert len(good) >= 6
    # the blocks that come out of the interleaver are the ones that went in two bursts earlier
    sent = [blocks[k] for k in sorted(blocks)]
    hits = sum(any(bytes(s) == g["block"] for s in sent) for g in good)
    assert hits >= 6
