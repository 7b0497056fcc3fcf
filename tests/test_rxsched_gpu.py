"""Receiver frame loop for several channels at once: gmr1b200_rx_bcch_batch against the
reference's own application (oracle/_ref/gmr1_rx = src/gmr1_rx.c linked to the reference C libraries) run
once per recording.  Every channel must walk the same frames with the same frame numbers, find the same
burst kinds behind the energy gate, and report the same CRC result for every burst; Viterbi metrics agree to
the soft-bit tolerance, the final alignment to one sample."""
import os
import re
import subprocess

import numpy as np
import pytest

import recording

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "gmr1_rx")
SPS = 4
START_DISCARD = 8000                              # gmr1_rx.c:56


def _reference_runs(path):
    """parse gmr1_rx's log into one list of frames per process_bcch() call"""
    r = subprocess.run([REF_BIN, str(SPS), path], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    runs = []
    for l in r.stderr.split("\n"):
        m = re.match(r"\[\+\] Processing BCCH @(\d+) .*freq_err = (-?[\d.]+) Hz", l)
        if m:
            runs.append({"align": int(m.group(1)), "freq_hz": float(m.group(2)), "frames": []})
            continue
        if not runs:
            continue
        m = re.match(r"\[-\]  FN:\s*(-?\d+)", l)
        if m:
            runs[-1]["frames"].append({"fn": int(m.group(1)), "kind": 0, "crc": None, "conv": None})
            continue
        if l.startswith("[.]   BCCH"):
            runs[-1]["frames"][-1]["kind"] = 1
        elif l.startswith("[.]   CCCH"):
            runs[-1]["frames"][-1]["kind"] = 2
        m = re.match(r"\[\+\] TCH3 assigned on TN (\d+)", l)
        if m:
            runs[-1].setdefault("ass", []).append((len(runs[-1]["frames"]) - 1, int(m.group(1))))
        m = re.match(r"crc=(-?\d+), conv=(-?\d+)", l)
        if m:
            runs[-1]["frames"][-1]["crc"] = int(m.group(1))
            runs[-1]["frames"][-1]["conv"] = int(m.group(2))
    return runs


def _acquire_like_main(o, x):
    """main() / fcch_single_init / fcch_multi_process of gmr1_rx.c:606-741 on the oracle's functions: the
    (align, freq_err) every process_bcch() call starts from, with freq_err as the exact float"""
    blen = 117 * SPS
    align = START_DISCARD
    rc, toa = o.fcch_rough(x[align:align + (330 * 23400 * SPS) // 1000], SPS, 0.0)
    assert rc == 0
    align += toa
    rc, toa, ferr = o.fcch_fine(x[align:align + blen], SPS, 0.0)
    assert rc == 0
    align += toa
    ferr = np.float32(ferr)
    base = max(align - blen, 0)
    n, mtoa = o.fcch_rough_multi(x[base:base + (650 * 23400 * SPS) // 1000], SPS, float(-ferr), 16)
    assert n >= 1
    out, ref_snr, ref_fe = [], 0.0, 0.0
    for i in range(n):
        rc, t, fe = o.fcch_fine(x[base + mtoa[i]:base + mtoa[i] + blen], SPS, float(-ferr))
        assert rc == 0
        rc, snr = o.fcch_snr(x[base + mtoa[i] + t:base + mtoa[i] + t + blen], SPS, float(-(ferr + np.float32(fe))))
        if i == 0:
            ref_snr, ref_fe = snr, fe
        elif snr < 2.0 or snr < ref_snr / 6.0 or 23400.0 * abs(ref_fe - fe) / (2 * np.pi) > 500.0:
            continue
        out.append((base + mtoa[i] + t, float(ferr)))
    return out


@pytest.fixture(params=[0, 1], ids=["paced", "lockstep"])
def walk_mode(request, gpu_lib):
    """both schedules of the frame walk (gmr1b200_set_rx_lockstep): paced by each channel's BCCH bursts (default) and
    frame by frame in lock step"""
    prev = gpu_lib.call("gmr1b200_set_rx_lockstep", request.param)
    yield request.param
    gpu_lib.call("gmr1b200_set_rx_lockstep", prev)


def test_rx_bcch_batch_vs_gmr1_rx(gpu_lib, oracle, tmp_path, walk_mode):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    cases = [dict(esn0_db=15.0, cfo_hz=300.0, seed=1), dict(esn0_db=8.0, cfo_hz=-450.0, seed=2, seconds=2.6),
             dict(esn0_db=12.0, cfo_hz=120.0, seed=3, tdma_si1=True), dict(esn0_db=4.0, cfo_hz=-60.0, seed=4, frac=0.81),
             dict(esn0_db=20.0, cfo_hz=700.0, seed=5, tdma_si1=True, seconds=3.0, start=12345)]
    recs, tasks, ref = [], [], []
    for ci, kw in enumerate(cases):
        x, _ = recording.make(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2), **kw)
        path = str(tmp_path / f"rec{ci}.cfile")
        x.tofile(path)
        runs = _reference_runs(path)
        acq = _acquire_like_main(oracle, x)
        assert [r["align"] for r in runs] == [a for a, _ in acq]            # the harness mirrors main()
        for r, (a, fe) in zip(runs, acq):
            assert abs(r["freq_hz"] - 23400.0 * fe / (2 * np.pi)) <= 0.06
            tasks.append((ci, a, fe))
            ref.append(r)
        recs.append(x)
    rec_len = np.array([len(x) for x in recs], np.int32)
    rec_ofs = np.concatenate([[0], np.cumsum(rec_len[:-1])]).astype(np.int64)
    iq = np.ascontiguousarray(np.concatenate(recs)).view(np.float32)
    n = len(tasks)
    t_ofs = np.array([rec_ofs[c] for c, _, _ in tasks], np.int64)
    t_len = np.array([rec_len[c] for c, _, _ in tasks], np.int32)
    align0 = np.array([a for _, a, _ in tasks], np.int32)
    ferr0 = np.array([f for _, _, f in tasks], np.float32)
    F = 96
    kind = np.zeros((n, F), np.int32)
    fn = np.zeros((n, F), np.int32)
    crc = np.zeros((n, F), np.int32)
    conv = np.zeros((n, F), np.int32)
    l2 = np.zeros((n, F, 24), np.uint8)
    nfr = np.zeros(n, np.int32)
    align1 = np.zeros(n, np.int32)
    ferr1 = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_rx_bcch_batch", iq, len(iq) // 2, t_ofs, t_len, align0, ferr0, SPS, n, F,
                 kind, fn, crc, conv, l2, nfr, align1, ferr1, None)
    n_burst = n_ok = 0
    for i, r in enumerate(ref):
        fr = r["frames"]
        assert nfr[i] == len(fr), (i, nfr[i], len(fr))
        for f, e in enumerate(fr):
            assert fn[i, f] == e["fn"], (i, f)
            rk = e["kind"] if e["crc"] is not None else 0       # "[.]   BCCH" is printed before the window is mapped
            assert kind[i, f] == rk, (i, f, kind[i, f], rk)
            if rk:
                assert crc[i, f] == e["crc"], (i, f)
                assert abs(int(conv[i, f]) - e["conv"]) <= 8, (i, f)
                n_burst += 1
                n_ok += e["crc"] == 0
    assert n_burst >= 200 and n_ok >= 0.8 * n_burst
    assert any(r["frames"][0]["fn"] != r["frames"][-1]["fn"] - len(r["frames"]) + 1 for r in ref)   # an SI1 re-timed a channel

    # the same walk on a caller-owned stream with device-resident buffers: identical results
    import torch
    dev = torch.device("cuda", 0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        t_in = [d(iq), d(t_ofs), d(t_len), d(align0), d(ferr0)]
        outs = [torch.zeros_like(d(a)) for a in (kind, fn, crc, conv, l2, nfr, align1, ferr1)]
        gpu_lib.call("gmr1b200_rx_bcch_batch", t_in[0], len(iq) // 2, t_in[1], t_in[2], t_in[3], t_in[4], SPS, n, F,
                     *outs, st.cuda_stream)
    st.synchronize()
    got = [o.cpu().numpy() for o in outs]
    assert (got[5] == nfr).all() and (got[6] == align1).all() and (got[7] == ferr1).all()
    assert (got[0] == kind).all() and (got[2] == crc).all()          # the call initialises these two for every frame slot
    valid = np.arange(F)[None, :] < nfr[:, None]                     # fn / conv / l2 are written for walked frames only
    assert (got[1][valid] == fn[valid]).all()
    dec = valid & (kind > 0)
    assert (got[3][dec] == conv[dec]).all() and (got[4][dec] == l2[dec]).all()



def _multi_cell_recordings(oracle):
    """recordings with the FCCH / BCCH / CCCH frames of one, two and three cells at different timings inside the
    320 ms FCCH period; the third cell of the last one is 800 Hz away from the strongest and must be filtered"""
    enc = (lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2))
    mk = lambda amp, **kw: amp * recording.make(*enc, seconds=1.6, esn0_db=40.0, **kw)[0].astype(np.complex64)
    noise = lambda seed, n: (0.12 * (np.random.default_rng(seed).standard_normal((n, 2)) @ np.array([1, 1j]))).astype(np.complex64)
    recs = []
    a = mk(1.0, cfo_hz=200.0, seed=11, start=9000)
    recs.append(a + noise(1, len(a)))
    b = mk(0.7, cfo_hz=260.0, seed=12, start=9000 + 11000)
    recs.append(a + b + noise(2, len(a)))
    c = mk(0.8, cfo_hz=-600.0, seed=13, start=9000 + 17000)
    d = mk(0.6, cfo_hz=150.0, seed=14, start=9000 + 5000, frac=0.6)
    recs.append(a + c + d + noise(3, len(a)))
    return recs


def test_fcch_multi_batch_vs_gmr1_rx(gpu_lib, oracle, tmp_path):
    """gmr1b200_fcch_acquire_batch + gmr1b200_fcch_multi_batch against main() / fcch_single_init / fcch_multi_process of
    the reference application (src/gmr1_rx.c:606-741): the same number of process_bcch() calls per recording, starting
    at the same alignments (in the reference's order, strongest first; +-1 sample as for every FCCH TOA)."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    recs = _multi_cell_recordings(oracle)
    ref = []
    for ci, x in enumerate(recs):
        path = str(tmp_path / f"multi{ci}.cfile")
        x.tofile(path)
        ref.append([r["align"] for r in _reference_runs(path)])
    assert [len(r) for r in ref] == [1, 2, 2], ref            # the generator is understood by the reference
    n = len(recs)
    rec_len = np.array([len(x) for x in recs], np.int32)
    rec_ofs = np.concatenate([[0], np.cumsum(rec_len[:-1])]).astype(np.int64)
    iq = np.ascontiguousarray(np.concatenate(recs)).view(np.float32)
    W330 = (330 * 23400 * SPS) // 1000
    align = np.zeros(n, np.int32)
    ferr = np.zeros(n, np.float32)
    gpu_lib.call("gmr1b200_fcch_acquire_batch", 0, iq, len(iq) // 2, rec_ofs + START_DISCARD, 0, W330, SPS,
                 None, align, ferr, n, None)
    align += START_DISCARD
    M = 4
    cnt = np.full(n, -99, np.int32)
    cal = np.zeros((n, M), np.int32)
    snr = np.zeros((n, M), np.float32)
    cfe = np.zeros((n, M), np.float32)
    gpu_lib.call("gmr1b200_fcch_multi_batch", 0, iq, len(iq) // 2, rec_ofs, rec_len, align, ferr, SPS, n, M,
                 cnt, cal, snr, cfe, None)
    for i in range(n):
        assert cnt[i] == len(ref[i]), (i, cnt[i], cal[i], ref[i])
        for a, b in zip(cal[i, :cnt[i]].tolist(), ref[i]):
            assert abs(a - b) <= 1, (i, cal[i], ref[i])
        assert (snr[i, :cnt[i]] >= 2.0).all() and abs(cfe[i, 0]) < 0.01          # residual of the primary acquisition: a few Hz
    # the same with the recordings in device memory
    import torch
    cnt2 = np.full(n, -99, np.int32)
    cal2 = np.zeros((n, M), np.int32)
    d_iq = torch.from_numpy(iq).to("cuda:0")
    gpu_lib.call("gmr1b200_fcch_multi_batch", 0, d_iq, len(iq) // 2, rec_ofs, rec_len, align, ferr, SPS, n, M,
                 cnt2, cal2, None, None, None)
    assert (cnt2 == cnt).all() and all((cal2[i, :cnt[i]] == cal[i, :cnt[i]]).all() for i in range(n))
    # a recording with fewer than 650 ms behind the window start: the reference's "Not enough samples"
    short_len = np.array([align[0] + 50000], np.int32)
    c1 = np.zeros(1, np.int32)
    gpu_lib.call("gmr1b200_fcch_multi_batch", 0, iq, len(iq) // 2, rec_ofs[:1], short_len, align[:1], ferr[:1], SPS, 1, M,
                 c1, cal2[:1], None, None, None)
    assert c1[0] == -22


def test_rx_bcch_ass_batch_vs_gmr1_rx(gpu_lib, oracle, tmp_path, walk_mode):
    """the TCH3 hand-off of the frame loop (rx_ccch -> rx_tch3_init, src/gmr1_rx.c:836-841,362-381): IMMEDIATE
    ASSIGNMENTs on CCCH bursts are seen in the same frames with the same timeslot as by the reference application
    ("[+] TCH3 assigned on TN"), DKAB position as sent, energy thresholds from the last BCCH window; the walk itself
    is identical to gmr1b200_rx_bcch_batch."""
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref/gmr1_rx not built (needs /root/reference at build time)")
    enc = (lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2))
    plans = [{5: (7, 3), 13: (21, 10)}, None, {4: (0, 63), 6: (31, 0), 7: (12, 33)}]
    recs, ref, align0, ferr0 = [], [], [], []
    for ci, plan in enumerate(plans):
        x, _ = recording.make(*enc, seconds=1.8, esn0_db=14.0, cfo_hz=100.0 * (ci + 1), seed=30 + ci, imm_ass=plan)
        path = str(tmp_path / f"ass{ci}.cfile")
        x.tofile(path)
        runs = _reference_runs(path)
        assert len(runs) == 1
        ref.append(runs[0])
        a, fe = _acquire_like_main(oracle, x)[0]
        assert a == runs[0]["align"]
        align0.append(a)
        ferr0.append(fe)
        recs.append(x)
        assert [(f, tn) for f, tn in runs[0].get("ass", [])] == sorted((f, tn) for f, (tn, _) in (plan or {}).items())
    n = len(recs)
    rec_len = np.array([len(x) for x in recs], np.int32)
    rec_ofs = np.concatenate([[0], np.cumsum(rec_len[:-1])]).astype(np.int64)
    iq = np.ascontiguousarray(np.concatenate(recs)).view(np.float32)
    F = 48
    mk = lambda *shape, dt=np.int32: np.zeros(shape, dt)
    outs = [mk(n, F), mk(n, F), mk(n, F), mk(n, F), mk(n, F, 24, dt=np.uint8), mk(n), mk(n), mk(n, dt=np.float32)]
    outs2 = [np.zeros_like(a) for a in outs]
    tch3 = np.full((n, 4), 77, np.int32)
    en = np.full((n, 2), 77.0, np.float32)
    args = (iq, len(iq) // 2, rec_ofs, rec_len, np.array(align0, np.int32), np.array(ferr0, np.float32), SPS, n, F)
    gpu_lib.call("gmr1b200_rx_bcch_ass_batch", *args, *outs, tch3, en, None)
    gpu_lib.call("gmr1b200_rx_bcch_batch", *args, *outs2, None)
    kind, nfr = outs[0], outs[5]
    assert (outs2[0] == kind).all() and (outs2[2] == outs[2]).all() and (outs2[5] == nfr).all() and (outs2[6] == outs[6]).all()
    for i, (plan, r) in enumerate(zip(plans, ref)):
        assert nfr[i] == len(r["frames"])
        if not plan:
            assert tch3[i].tolist() == [0, 0, 0, -1] and (en[i] == 0.0).all()
            continue
        f_last, tn_last = r["ass"][-1]                       # a later IMM.ASS re-initialises the state
        assert tch3[i].tolist() == [1, tn_last, plan[f_last][1], f_last], (i, tch3[i], r["ass"])
        fb = max(f for f in range(f_last) if f % 8 == 2)     # last BCCH frame before the assignment
        b = 9000 + fb * recording.FRAME - 40
        w = recs[i][b:b + 1016]
        e_bcch = float((np.abs(w[31:1016 - 31]) ** 2).sum() / 1016)
        assert abs(en[i, 0] - 0.75 * e_bcch / 2) <= 0.02 * en[i, 0], (i, en[i], e_bcch)
        assert en[i, 1] == np.float32(en[i, 0] / np.float32(8.0))
