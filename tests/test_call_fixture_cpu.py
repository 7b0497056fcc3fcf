"""The TCH3 call fixture (tests/recording.make_call) is understood by the reference application: gmr1_rx on the
BCCH + traffic recordings follows the IMMEDIATE ASSIGNMENT, classifies every traffic burst as sent (speech / FACCH3 /
DKAB), returns the speech frames and FACCH3 messages that were encoded, discovers the ciphering from the first FACCH3
codeword (the unciphered attempt fails, the retry passes, gmr1_rx.c:417-430) and releases the channel after ten
silent frames.  CPU only; this is the recording the device-side TCH3 burst loop is tested on."""
import os
import subprocess

import numpy as np
import pytest

import recording
import rxlog

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "gmr1_rx")
PLAN = "sssssssss" + "ffff" + "ssdd" + "ffff" + "sdsd" + "-" * 12          # IMM.ASS in frame 3: 'f' groups on fn & 3 == 0


@pytest.mark.parametrize("key", [None, "0123456789abcdef"])
def test_reference_follows_the_call(oracle, tmp_path, key):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()                         # host-side encoder (the reference's gmr1_tch3_encode is unusable)

    def enc_speech(f0, f1, bs, c):
        out = np.zeros(212, np.uint8)
        L.call("gmr1b200_tch3_encode", out, np.ascontiguousarray(f0), np.ascontiguousarray(f1),
               np.ascontiguousarray(bs), c, 0)
        return out

    kc = np.frombuffer(bytes.fromhex(key), np.uint8) if key else None
    b, t, truth = recording.make_call(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2),
                                      enc_speech, lambda l2, bs, c: oracle.facch3_encode(l2, bs, c), PLAN, tn=7, p=3,
                                      ass_frame=3, kc=kc, a5=lambda k, fn, n: oracle.a5(1, k, fn, n), seed=5)
    pb, pt = str(tmp_path / "bcch.cfile"), str(tmp_path / "tch.cfile")
    b.tofile(pb)
    t.tofile(pt)
    r = subprocess.run([REF_BIN, "4", pb, pt] + ([key] if key else []), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    fr = {f["fn"]: f for f in rxlog.parse(r.stderr.split("\n"))}
    assert fr[3]["assigned"] == 7 and all(fr[f]["assigned"] is None for f in fr if f != 3)
    n_speech = n_exact = 0
    for f, kind, payload in truth:
        if kind == "tch3":
            assert fr[f]["tch"] == "tch3", (f, fr[f])
            f0, f1 = payload
            assert fr[f]["frame0"][:6] == bytes(f0[:6]) and fr[f]["frame1"][:6] == bytes(f1[:6]), f     # protected bits
            n_speech += 1
            n_exact += fr[f]["frame0"] == bytes(f0) and fr[f]["frame1"] == bytes(f1)
        elif kind == "facch3":
            assert [fr[g]["tch"] for g in range(f - 3, f + 1)] == ["facch3"] * 4 and [fr[g]["bi"] for g in range(f - 3, f + 1)] == [0, 1, 2, 3]
            crcs = [c for c, _ in fr[f]["flush"]]
            # first ciphered codeword: the plain attempt fails and the retry passes; afterwards one attempt
            first = f == min(g for g, k, _ in truth if k == "facch3")
            assert crcs == ([1, 0] if (key and first) else [0]), (f, fr[f]["flush"])
        elif kind == "dkab":
            assert fr[f]["tch"] == "dkab", (f, fr[f])
    assert n_speech == PLAN.count("s") and n_exact >= 0.9 * n_speech
    last = 3 + len(PLAN.rstrip("-")) - 1
    ends = [f for f in fr if fr[f]["end"]]
    assert ends == [last + 10], ends                 # weak_cnt++ > 8 on the tenth silent frame (gmr1_rx.c:566-569)
    assert all(fr[f]["tch"] is None for f in fr if f > last + 10)
