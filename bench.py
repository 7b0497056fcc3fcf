#!/usr/bin/env python
"""bench.py - decoded bursts/s of the GMR-1 receive hot path (FCCH acquisition + pi/4-CxPSK demod + Viterbi/CRC
decode) on B200.  Headline workload = config 2 of BASELINE.json: batched BCCH / DC6(CCCH) bursts, 1024 ARFCNs x 256
bursts per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the hot path over the whole batch: FCCH acquisition of every ARFCN (rough over a 330 ms
window + fine), then demod BCCH, decode BCCH, demod DC6, decode CCCH (6 kernel launches).
  value            bursts/s over EXACTLY K steps with the IQ resident in HBM (CUDA events, max over ranks)
  sustained        the same loop repeated until >= --min-seconds have passed (the K-step region is ~20 ms)
  roofline         the demod kernels: algorithmic bytes / CUDA-event time of their launches (serial pass) / measured peak
  e2e              the same step through the C ABI with pinned HOST IQ in and host L2/CRC out, copies inside the timed
                   region; e2e.h2d_ceiling_gbs = the same bytes on the same threads / streams with no kernels
  cpu_baseline     the reference's own C path (oracle/_ref, -O2 -march=x86-64-v3) on the host cores, compiled loops
                   (oracle/harness.c), on the head of the same IQ; also the L2 / CRC parity check
  configs          BASELINE configs 3 and 4 at full size (N = 1 only): bursts/s, per-kernel times, their own demod
                   roofline, and a 4 096-burst-per-type comparison with the CPU reference
  wideband         the headline receive work fed from one wideband int16 recording of all ARFCNs through the GPU
                   channeliser, end to end from pinned host memory and device-resident (N = 1 only)
  sweep            BASELINE config 5: ARFCN count 1k .. 64k in total, sharded over the ranks, 1 s recording slice per
                   ARFCN (FCCH search + 25 bursts), device-resident and streamed from pinned host memory
Multi-GPU: ARFCNs are independent, each rank owns its own ARFCNs (weak scaling), no data-path collective;
torch.distributed is used only for the barrier and the max-over-ranks of the time.  --single-process (with --gpus N,
not under torchrun) runs the end-to-end step of all N GPUs from ONE process through the library's device pool.
`--impl reference` times the reference's own C path on the host cores on a bounded sample of the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPS = 4
WIN = {"bcch": 80, "dc6": 40}                     # search windows of gmr1_rx.c:759,809 (20*sps, 10*sps)
LEN = {"bcch": 234, "dc6": 234}
EBITS = {"bcch": 424, "dc6": 432}
BT = {"bcch": 0, "dc6": 2}
CHAN = {"bcch": 0, "dc6": 1}                      # gmr1b200_xcch_encode_batch channel id
# algorithmic bytes per burst of the demod kernel: window in + ebits + 16 B metadata out
DEMOD_BYTES = {k: 8 * (LEN[k] * SPS + WIN[k]) + EBITS[k] + 16 for k in WIN}
SNR_GRID = np.array([6.0, 10.0, 15.0, 30.0], np.float32)
FCCH_WIN = (330 * 23400 * SPS) // 1000            # 30 888 samples


def wlen(kind):
    return LEN[kind] * SPS + WIN[kind]


def burst_params(n, kind, seed):
    rng = np.random.default_rng(seed)
    return dict(
        l2=rng.integers(0, 256, (n, 24), dtype=np.uint8),
        toa=rng.uniform(2.0, WIN[kind] - 2.0, n).astype(np.float32),
        cfo=rng.uniform(-0.0134, 0.0134, n).astype(np.float32),     # +-50 Hz residual after FCCH
        phase=rng.uniform(0, 2 * np.pi, n).astype(np.float32),
        esn0=SNR_GRID[rng.integers(0, 4, n)],
    )


def fcch_windows(n_arfcn, seed, lo=0, hi=None):
    """One 330 ms FCCH search window per ARFCN (gmr1_rx.c:612): noise + one dual chirp at a random offset with
    up to +-1 kHz of carrier offset.  Returns (pos, cfo) for all ARFCNs and a generator of (lo, hi, complex64)."""
    rng = np.random.default_rng(seed)
    pos = rng.integers(600, FCCH_WIN - 1200, n_arfcn)
    cfo = rng.uniform(-0.27, 0.27, n_arfcn)
    k = np.arange(117 * SPS)
    t = k / SPS - 58.5
    chirp = np.sqrt(2.0) * np.cos(0.32 * 2 * np.pi / 117 * t * t)

    def blocks():
        for b0 in range(0, n_arfcn, 128):
            b1 = min(n_arfcn, b0 + 128)
            x = 0.3 * (rng.standard_normal((b1 - b0, FCCH_WIN)) + 1j * rng.standard_normal((b1 - b0, FCCH_WIN)))
            for i in range(b0, b1):
                x[i - b0, pos[i]:pos[i] + 117 * SPS] += chirp * np.exp(1j * cfo[i] * k / SPS)
            yield b0, b1, x.astype(np.complex64)
    return pos, cfo, blocks()


# ------------------------------------------------------------------------------------------------
# CPU reference arm (also the cpu_baseline leg): the ONLY place bench.py executes oracle/
# ------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    """one chunk of one burst type through the reference's C functions (compiled loops, oracle/harness.c)"""
    path, kind, lo, hi, extra = job
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    o = oracle_lib.load(fast=True)
    x = np.ascontiguousarray(np.load(path, mmap_mode="r")[lo:hi])
    sl = lambda key: None if extra.get(key) is None else np.ascontiguousarray(extra[key][lo:hi])
    t0 = time.perf_counter()
    if kind == "fcch":                      # rough + fine acquisition, as Workload.fcch() does on the GPU
        _, align, ferr = o.h_fcch_acquire(x, SPS, 0.0)
        res = (align, ferr)
    elif kind == "fcch_grid":
        res = (o.h_fcch_grid(x, extra["shifts"], SPS),)
    elif kind in ("bcch", "dc6"):
        res = o.h_xcch(kind == "dc6", x, SPS, 0.0)[:2]
    elif kind == "tch3":
        res = o.h_tch3(x, sl("ciph"), 0, SPS)
    elif kind == "facch3":
        res = o.h_facch3(x, SPS)
    elif kind == "facch9":
        res = o.h_facch9(x, SPS)
    elif kind == "tch9":
        res = o.h_tch9(x, extra["n_burst"], 2, SPS)
    elif kind == "rach":
        res = o.h_rach(x, sl("sb_mask"), SPS)
    else:
        raise ValueError(kind)
    return kind, lo, res, time.perf_counter() - t0, o.kind, o.flags


def cpu_reference_pass(files, cores, pool=None):
    """files: {kind: (npy path, n units, unit granularity, extra)}; runs the reference C path over everything on
    `cores` processes.  Returns (wall seconds, {kind: tuple of result arrays}, oracle kind, compiler flags)."""
    jobs = []
    for kind, (path, n, gran, extra) in files.items():
        per = max(gran, ((n + cores - 1) // cores + gran - 1) // gran * gran)
        if kind.startswith("fcch"):
            per = max(1, min(per, 8))           # short jobs, scheduled first, so they spread over the cores
        jobs += [(path, kind, lo, min(n, lo + per), extra) for lo in range(0, n, per)]
    jobs.sort(key=lambda j: not j[1].startswith("fcch"))
    own = pool is None
    if own:
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_noop, range(cores))                       # start the workers outside the timing
    t0 = time.perf_counter()
    res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    if own:
        pool.close()
    out = {}
    for kind in files:
        parts = sorted((r for r in res if r[0] == kind), key=lambda r: r[1])
        axis = 1 if kind == "fcch_grid" else 0
        out[kind] = tuple(np.concatenate([p[2][i] for p in parts], axis=axis) for i in range(len(parts[0][2])))
    return wall, out, res[0][4], res[0][5]


def _cpu_noop(i):
    return i


def shm_dir():
    d = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    return d


def save_shm(tag, x):
    path = os.path.join(shm_dir(), f"gmr1_bench_{tag}_{os.getpid()}.npy")
    np.save(path, x)
    return path


def run_reference_arm(args):
    """bench.py --impl reference: pure CPU (no kernel of this repo runs): the reference's C functions, one process per
    host core, on a numpy-generated bounded sample of the config-2 workload (same burst formats, SNR grid, offsets and
    burst : acquisition ratio as the GPU arm; rates are per burst, so the sample size does not enter the ratio)"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    import sigen
    o = oracle_lib.load(fast=True)
    cores = os.cpu_count() or 1
    per_kind = min(32768, 1024 * cores)            # bounded sample: ~0.14 ms of C work per burst and core
    files = {}
    for kind in ("bcch", "dc6"):
        p = burst_params(per_kind, kind, 77 + BT[kind])
        chan = "bcch" if kind == "bcch" else "ccch"
        hard = np.stack([o.encode(chan, EBITS[kind], p["l2"][i]) for i in range(per_kind)])
        rng = np.random.default_rng(5)
        xs = [sigen.modulate(kind, hard[i:i + 512], SPS, WIN[kind], p["toa"][i:i + 512], p["cfo"][i:i + 512],
                             p["phase"][i:i + 512], p["esn0"][i:i + 512], rng) for i in range(0, per_kind, 512)]
        files[kind] = (save_shm("ref_" + kind, np.concatenate(xs)), per_kind, 1, {})
    n_f = max(1, 2 * per_kind // 256)               # one FCCH acquisition per 256 bursts, as in config 2
    _, _, blocks = fcch_windows(n_f, 176)
    files["fcch"] = (save_shm("ref_fcch", np.concatenate([x for _, _, x in blocks])), n_f, 1, {})
    nb = 2 * per_kind
    times = []
    pool = mp.get_context("spawn").Pool(cores)
    pool.map(_cpu_noop, range(cores))
    for step in range(args.warmup + args.steps):
        wall, _, okind, flags = cpu_reference_pass(files, cores, pool)
        if step >= args.warmup:
            times.append(wall)
    pool.close()
    for path, _, _, _ in files.values():
        os.unlink(path)
    t = float(np.mean(times))
    val = nb / t
    line = {
        "impl": "reference", "metric": "decoded bursts/s (FCCH sync+demod+Viterbi)", "value": val, "unit": "bursts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": "config2: BCCH + DC6/CCCH pi/4-CQPSK demod + K5 r1/2 Viterbi + CRC16, sps 4",
                   "bursts_per_step": nb,
                   "sample": f"{per_kind} BCCH + {per_kind} DC6 bursts + {n_f} FCCH windows per step: a bounded SAMPLE of the "
                             "GPU arm's workload (same formats, 6/10/15/30 dB grid, offsets, burst : acquisition ratio), made by "
                             "the numpy generator tests/sigen.py because no GPU code may run in this arm; rates are per burst"},
        "cpu_baseline": {"value": val, "unit": "bursts/s", "cores": cores, "kind": okind, "flags": flags,
                         "sample": f"{nb} bursts + {n_f} FCCH acquisitions per step, one process per core, compiled loops "
                                   "(oracle/harness.c) around the reference's own functions"},
        "e2e": {"value": val, "unit": "bursts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))



# ------------------------------------------------------------------------------------------------
# BASELINE config 5: ARFCN sweep, 1 s recording slice per ARFCN, device-resident and host-streamed
# ------------------------------------------------------------------------------------------------
def run_sweep(L, torch, dist, dev, world, rank, barrier, totals, chunk_arfcns):
    """Every rank processes total / world ARFCNs of each sweep point in chunks of <= chunk_arfcns slices (748.8 KB of
    IQ each): FCCH rough + fine once per slice, then its 25 bursts (4 BCCH, 12 CCCH, 9 NT3 speech) demodulated and
    decoded in place by window offsets.  'resident': the chunk lies in HBM; 'streamed': every chunk is copied from pinned
    host memory first (two streams, two device buffers, copy of chunk i+1 overlaps the kernels of chunk i) and its
    results go back to pinned host memory.  The pinned ring holds ONE chunk of distinct IQ that every copy re-reads
    (the bytes over PCIe are real, the content repeats).  Times are max over ranks."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import workloads as wl
    C = chunk_arfcns
    ck = wl.Config5Chunk(L, torch, dev, C, seed=5000 + rank)
    host = torch.empty(ck.iq.shape, dtype=torch.float32).pin_memory()
    host.copy_(ck.iq)
    dbuf = [torch.empty_like(ck.iq), torch.empty_like(ck.iq)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    res_host = [ck.result_buffers(pinned=True) for _ in range(2)]
    for s in range(2):
        ck.process(ck.iq, streams[s].cuda_stream, C, s)
    torch.cuda.synchronize()
    sane = ck.sane()
    out = []
    for total in totals:
        n_local = total // world
        if n_local < 1:
            continue
        chunks = [C] * (n_local // C) + ([n_local % C] if n_local % C else [])

        def resident():
            for i, c in enumerate(chunks):
                ck.process(ck.iq, streams[i % 2].cuda_stream, c, i % 2)

        def streamed():
            for i, c in enumerate(chunks):
                s = i % 2
                with torch.cuda.stream(streams[s]):
                    dbuf[s][:c].copy_(host[:c], non_blocking=True)
                ck.process(dbuf[s], streams[s].cuda_stream, c, s)
                with torch.cuda.stream(streams[s]):
                    ck.results_to_host(res_host[s], c, s)

        times = {}
        for name, fn in (("resident", resident), ("streamed", streamed)):
            fn()
            barrier()
            t0 = time.perf_counter()
            fn()
            for s in streams:
                s.synchronize()
            barrier()
            t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[name] = float(t[0])
        n_all = n_local * world
        out.append({"arfcns": n_all, "bursts": n_all * 25, "acquisitions": n_all,
                    "resident_bursts_per_s": n_all * 25 / times["resident"], "resident_ms": 1e3 * times["resident"],
                    "streamed_bursts_per_s": n_all * 25 / times["streamed"], "streamed_ms": 1e3 * times["streamed"],
                    "streamed_h2d_gbs": n_all * wl.SLICE * 8 / times["streamed"] / 1e9})
    return {"what": "config 5: per ARFCN a 1 s slice (748.8 KB): FCCH rough + fine, then 25 bursts (4 BCCH, 12 CCCH, 9 NT3 "
                    f"speech / TCH3) cut by window offsets; chunks of {C} ARFCNs; total ARFCNs sharded over {world} rank(s)",
            "chunk_decodes_correctly": bool(sane), "points": out}


# ------------------------------------------------------------------------------------------------
# BASELINE configs 3 and 4 at full size (one GPU): throughput, demod roofline, parity of the head vs the CPU reference
# ------------------------------------------------------------------------------------------------
def run_configs(L, torch, dev, peak_gbs, cores, scale, reps, n_check):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import workloads as wl
    out = {}
    pool = None
    if cores:
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_noop, range(cores))
    for cid in ("3", "4"):
        if cid == "3":
            w = wl.Config3(L, torch, dev, arfcns=max(8, int(4096 * scale)), frames=128, n_check=n_check)
            dem = [("demod_nt3_speech", "nt3_speech", w.n_sp), ("demod_nt3_facch", "nt3_facch", w.n_fa)]
            name = f"config3: NT3 traffic, {w.n_sp} speech bursts (TCH3, half A5/1) + {w.n_fa} FACCH3 bursts"
        else:
            w = wl.Config4(L, torch, dev, arfcns=max(8, int(8192 * scale)), per=64, n_check=n_check)
            dem = [("demod_nt9_facch9", "nt9", w.n_f9), ("demod_nt9_tch9", "nt9", w.n_t9), ("demod_rach", "rach", w.n_ra)]
            name = (f"config4: {w.n_f9} FACCH9 + {w.n_t9} TCH9-9k6 (chains of 3) over NT9, {w.n_ra} RACH, "
                    f"{w.arfcns} x 5-shift FCCH search + fine")
        tm = wl.Timer(torch)
        steps = w.steps()
        for k, f in steps.items():
            tm.run(k, f, reps)
        tm.run("whole_chain", w.run, reps)
        byt = sum(wl.demod_bytes(L, bt) * n for _, bt, n in dem)
        dem_ms = sum(tm.ms[k] for k, _, _ in dem)
        rec = {"workload": name, "bursts": w.n_bursts, "iq_bytes": int(w.iq_bytes),
               "bursts_per_s": w.n_bursts / tm.ms["whole_chain"] * 1e3, "ms": {k: round(v, 4) for k, v in tm.ms.items()},
               "roofline": {"bound": "hbm", "kernel": "demod kernels of this config", "achieved": byt / dem_ms / 1e6,
                            "peak": peak_gbs, "unit": "GB/s", "frac": byt / dem_ms / 1e6 / peak_gbs,
                            "per_format": {k: wl.demod_bytes(L, bt) * n / tm.ms[k] / 1e6 / peak_gbs for k, bt, n in dem}}}
        if cid == "3":
            rec["tch3_acs_per_s"] = w.n_sp * 12288 / tm.ms["decode_tch3"] * 1e3
        else:
            rec["fcch_searches_per_s"] = w.arfcns * 5 / tm.ms["fcch_5_shift_search_and_fine"] * 1e3
            rec["fcch_share_of_chain"] = tm.ms["fcch_5_shift_search_and_fine"] / tm.ms["whole_chain"]
        if pool is not None:
            head = w.head()
            files = {}
            for kind, (x, extra) in w.head_iq().items():
                gran = {"facch3": 4, "tch9": 3}.get(kind, 1)
                files[kind] = (save_shm(f"cfg{cid}_{kind}", x), len(x), gran, extra)
            wall, cpu, okind, flags = cpu_reference_pass(files, cores, pool)
            for path, _, _, _ in files.values():
                os.unlink(path)
            same, info = w.compare(head, cpu)
            units = sum(n for k, (_, n, _, _) in files.items() if not k.startswith("fcch"))
            rec["parity_vs_cpu_reference"] = {"units_per_type": w.n_check, "identical": same, "all_identical": all(same.values()),
                                              "oracle": okind, **info}
            rec["cpu_baseline"] = {"value": units / wall, "unit": "bursts/s", "cores": cores, "kind": okind, "flags": flags,
                                   "sample": f"head of the GPU workload: {w.n_check} units per burst type"
                                             + (", 256 x 5 FCCH searches" if cid == "4" else "")}
        out[cid] = rec
        del w
        torch.cuda.empty_cache()
    if pool is not None:
        pool.close()
    return out


# ------------------------------------------------------------------------------------------------
# the end-to-end step of several GPUs from ONE process (library device pool, csrc/api_multi.cu)
# ------------------------------------------------------------------------------------------------
def capture_shares(n_wide, n_arfcn, world, rank):
    """one wideband recording shared by `world` ranks: rank r feeds the time slice [lo, hi) of its n_wide samples (slices
    of n_slice samples, the last one ragged) and receives the ARFCNs a with a mod world == r"""
    n_slice = -(-n_wide // world)
    lo = min(n_wide, rank * n_slice)
    return {"n_slice": n_slice, "lo": lo, "hi": min(n_wide, lo + n_slice), "own": np.arange(rank, n_arfcn, world)}


def burst_rows(own, per):
    """rows of a [n_arfcn * per] burst table (ARFCN-major) that belong to the ARFCNs `own`"""
    return (np.asarray(own)[:, None] * per + np.arange(per)[None, :]).reshape(-1)


def run_wideband(L, torch, dev, n_arfcn, per_arfcn, steps, seed=4242, world=1, dist=None, shared=False, rank=0, fmt=1):
    """SURVEY 8f N3 in front of the headline workload: the SAME receive work (one FCCH acquisition per ARFCN + per_arfcn
    BCCH / DC6 bursts per ARFCN, demod + decode), but the input is ONE wideband recording of all ARFCNs (int16 I/Q at
    n_arfcn x 31.25 kS/s, what an SDR front end delivers) instead of one 4x-oversampled complex-float stream per ARFCN:
    pinned host recording -> H2D -> gmr1b200_channelize (filter bank + RRC resampler, streams stay in HBM) -> FCCH
    acquire, demod, decode on window offsets into the streams -> host L2 / CRC.  The recording is made once on the device
    (transmit-pulse bursts of every ARFCN interpolated, mixed to their carriers, summed, AWGN, int16).
    shared (world > 1): ONE recording for the whole job (every rank makes the same one from the same seed and keeps a
    1 / world time slice of it in pinned host memory): per step every rank copies its slice to its GPU, an NCCL
    all-gather over NVLink gives every GPU the whole recording - the one real exchange step of this system, SURVEY 8f
    N3 - every GPU runs the bank and resamples / acquires / demodulates / decodes ITS ARFCNs (a mod world == rank):
    strong scaling of one capture.
    fmt: 1 = int16 I/Q (the default), 2 = int8 I/Q (8-bit front ends: a quarter of the complex-float bytes per sample)."""
    import ctypes
    assert fmt in (1, 2) and not (shared and fmt != 1)
    wdt, sbytes = (torch.int16, 4) if fmt == 1 else (torch.int8, 2)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    n_b = {"bcch": per_arfcn // 2, "dc6": per_arfcn - per_arfcn // 2}
    lead = 64
    slot = wlen("bcch") + wlen("dc6")
    slen = lead + FCCH_WIN + (per_arfcn // 2) * slot + 64
    h = ctypes.c_void_p()
    assert L.c.gmr1b200_chan_create(n_arfcn, SPS, ctypes.byref(h)) == 0

    class Info(ctypes.Structure):
        _fields_ = [("n_chans", ctypes.c_int32), ("sps", ctypes.c_int32), ("n_taps", ctypes.c_int32),
                    ("taps_per_branch", ctypes.c_int32), ("n_taps_resamp", ctypes.c_int32), ("fft_stages", ctypes.c_int32),
                    ("samp_rate", ctypes.c_double), ("mid_rate", ctypes.c_double), ("resamp", ctypes.c_double),
                    ("delay_out", ctypes.c_double)]
    info = Info()
    L.c.gmr1b200_chan_info(h, ctypes.byref(info))
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    streams = torch.zeros((n_arfcn, slen, 2), dtype=torch.float32, device=dev)
    # FCCH chirp of every ARFCN inside its 330 ms head (clean: the noise is added to the wideband sum)
    rng = np.random.default_rng(seed)
    fpos = rng.integers(600, FCCH_WIN - 1200, n_arfcn)
    fcfo = rng.uniform(-0.27, 0.27, n_arfcn)
    k = torch.arange(117 * SPS, device=dev, dtype=torch.float32)
    t = k / SPS - 58.5
    chirp = (2.0 ** 0.5) * torch.cos(0.32 * 2 * np.pi / 117 * t * t)
    ang = d(fcfo.astype(np.float32))[:, None] * (k[None, :] / SPS)
    idx = d(fpos + lead)[:, None] + torch.arange(117 * SPS, device=dev)[None, :]
    rows = torch.arange(n_arfcn, device=dev)[:, None].expand_as(idx)
    streams[rows, idx, 0] = chirp[None, :] * torch.cos(ang)
    streams[rows, idx, 1] = chirp[None, :] * torch.sin(ang)
    par, ofs = {}, {}
    for kind, first in (("bcch", 0), ("dc6", wlen("bcch"))):
        n = n_arfcn * n_b[kind]
        pr = burst_params(n, kind, seed + BT[kind])
        hard = np.zeros((n, EBITS[kind]), np.uint8)
        L.call("gmr1b200_xcch_encode_batch", CHAN[kind], hard, pr["l2"], n)
        o = (np.arange(n_arfcn, dtype=np.int64)[:, None] * slen + lead + FCCH_WIN + first +
             np.arange(n_b[kind], dtype=np.int64)[None, :] * slot).reshape(-1)
        L.call("gmr1b200_synth_bursts_tx", BT[kind], d(hard), EBITS[kind], None, SPS, wlen(kind), d(pr["toa"]), 0.0,
               d(pr["cfo"]), 0.0, d(pr["phase"]), 0.0, None, 200.0, None, 1.0, seed, streams, n_arfcn * slen, d(o), 0, n, None)
        par[kind], ofs[kind] = pr, o
    n_wide = ((slen - 3) * 625 * n_arfcn) // (468 * SPS)
    wide = torch.empty((n_wide, 2), dtype=wdt, device=dev)
    esn0, gain = 15.0, 0.5 / (4.0 * np.sqrt(n_arfcn))
    torch.cuda.synchronize()
    g0 = time.perf_counter()
    L.call("gmr1b200_synth_wideband", h.value, streams, slen, slen, None, n_arfcn, esn0, gain, seed, wide, fmt, n_wide, None)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - g0
    del streams
    peak_i16 = int(wide.to(torch.int32).abs().max())
    shared = shared and world > 1
    share = capture_shares(n_wide, n_arfcn, world if shared else 1, rank if shared else 0)
    own = share["own"]                     # the ARFCNs this rank receives
    n_own = len(own)
    own_idx = d(own.astype(np.int32)) if shared else None
    if shared:
        n_slice, lo_s, hi_s = share["n_slice"], share["lo"], share["hi"]
        wide_full = torch.zeros((world * n_slice, 2), dtype=torch.int16, device=dev)
        my_slice = torch.empty((n_slice, 2), dtype=torch.int16, device=dev)
        host_wide = torch.zeros((n_slice, 2), dtype=torch.int16).pin_memory()
        host_wide[:hi_s - lo_s].copy_(wide[lo_s:hi_s])
        wide_full[:n_wide].copy_(wide)
        wide = wide_full[:n_wide]
    else:
        host_wide = torch.empty((n_wide, 2), dtype=wdt).pin_memory()
        host_wide.copy_(wide)
    n_out = int(L.c.gmr1b200_chan_out_len(h, n_wide))
    out = torch.empty((n_own, n_out, 2), dtype=torch.float32, device=dev)
    dly = int(round(info.delay_out))
    # window offsets inside the channelised streams: row stride n_out, everything delay_out later
    res, sel = {}, {}
    for kind in ("bcch", "dc6"):
        sel[kind] = burst_rows(own, n_b[kind])             # this rank's bursts
        n = len(sel[kind])
        a = np.repeat(own, n_b[kind]).astype(np.int64)
        i_loc = np.repeat(np.arange(n_own, dtype=np.int64), n_b[kind])
        o = ofs[kind][sel[kind]] - a * slen + i_loc * n_out + dly
        res[kind] = dict(ofs=d(o), eb=torch.empty((n, EBITS[kind]), dtype=torch.int8, device=dev),
                         l2=torch.empty((n, 24), dtype=torch.uint8, device=dev), crc=torch.empty(n, dtype=torch.int32, device=dev),
                         h_l2=torch.empty((n, 24), dtype=torch.uint8).pin_memory(), h_crc=torch.empty(n, dtype=torch.int32).pin_memory())
    f_ofs = d(np.arange(n_own, dtype=np.int64) * n_out + lead + dly)
    f_toa = torch.empty(n_own, dtype=torch.int32, device=dev)
    f_align = torch.empty(n_own, dtype=torch.int32, device=dev)
    f_ferr = torch.empty(n_own, dtype=torch.float32, device=dev)
    h_f = torch.empty((2, n_own), dtype=torch.float32).pin_memory()
    st = torch.cuda.Stream(device=dev)
    sh = st.cuda_stream

    def receive(timers=None, src=None):
        marks = []
        mark = lambda name: (marks.append((name, ev())), marks[-1][1].record(st)) if timers is not None else None
        mark("start")
        L.call("gmr1b200_channelize", h.value, wide if src is None else src, fmt, n_wide, own_idx, n_own, out, n_out, sh)
        mark("channelize")
        L.call("gmr1b200_fcch_acquire_batch", 0, out, n_own * n_out, f_ofs, 0, FCCH_WIN, SPS, f_toa, f_align, f_ferr, n_own, sh)
        mark("fcch")
        for kind in ("bcch", "dc6"):
            r = res[kind]
            n = r["crc"].numel()
            L.call("gmr1b200_pi4cxpsk_demod_batch", BT[kind], out, n_own * n_out, r["ofs"], 0, wlen(kind), SPS, None, 0.0,
                   r["eb"], EBITS[kind], None, None, None, None, n, sh)
            L.call("gmr1b200_bcch_decode_batch" if kind == "bcch" else "gmr1b200_ccch_decode_batch", r["l2"], r["eb"], None,
                   r["crc"], n, sh)
        mark("demod_decode")
        if timers is not None:
            timers.append(marks)

    def e2e():
        # the pinned HOST recording goes straight into the C entry point: it copies it in pieces on its own copy stream
        # while the bank and the resampler of the pieces before run (csrc/api_chan.cu)
        if shared:                         # this rank's time slice up, all-gather over NVLink, then the device-resident call
            with torch.cuda.stream(st):
                my_slice.copy_(host_wide, non_blocking=True)
                dist.all_gather_into_tensor(wide_full.view(torch.int32), my_slice.view(torch.int32))    # one I/Q pair = one int32
            receive()
        else:
            receive(src=host_wide)
        with torch.cuda.stream(st):
            for kind in ("bcch", "dc6"):
                res[kind]["h_l2"].copy_(res[kind]["l2"], non_blocking=True)
                res[kind]["h_crc"].copy_(res[kind]["crc"], non_blocking=True)
            h_f[0].copy_(f_align.to(torch.float32), non_blocking=True)
            h_f[1].copy_(f_ferr, non_blocking=True)
        st.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def rank_max(v):                       # a time of the whole job = the slowest rank's
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    for _ in range(3):
        receive()
    barrier()
    launches0 = L.kernel_launches()
    timers = []
    a, b = ev(), ev()
    a.record(st)
    for _ in range(steps):
        receive(timers)
    b.record(st)
    barrier()
    launches = (L.kernel_launches() - launches0) / steps
    dev_ms = rank_max(a.elapsed_time(b) / steps)
    part = {}
    for marks in timers:
        for (_, e0), (name, e1) in zip(marks[:-1], marks[1:]):
            part[name] = part.get(name, 0.0) + e0.elapsed_time(e1) / steps
    e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e()
    barrier()
    e2e_ms = rank_max(1e3 * (time.perf_counter() - t0) / steps)
    # the same with TWO recordings in flight (a receiver that runs on consecutive captures): the second recording's
    # copy runs under the first one's FCCH / demod / decode tail; every recording has its own streams and buffers
    pipe = None
    if not shared:
        def make_slot():
            r2 = {k: dict(ofs=res[k]["ofs"], eb=torch.empty_like(res[k]["eb"]), l2=torch.empty_like(res[k]["l2"]),
                          crc=torch.empty_like(res[k]["crc"]), h_l2=torch.empty_like(res[k]["h_l2"]).pin_memory(),
                          h_crc=torch.empty_like(res[k]["h_crc"]).pin_memory()) for k in res}
            return dict(out=torch.empty_like(out), res=r2, f_toa=torch.empty_like(f_toa), f_align=torch.empty_like(f_align),
                        f_ferr=torch.empty_like(f_ferr), h_f=torch.empty_like(h_f).pin_memory(), st=torch.cuda.Stream(device=dev))
        slots = [dict(out=out, res=res, f_toa=f_toa, f_align=f_align, f_ferr=f_ferr, h_f=h_f, st=st), make_slot()]

        def issue(S):
            q = S["st"].cuda_stream
            L.call("gmr1b200_channelize", h.value, host_wide, fmt, n_wide, own_idx, n_own, S["out"], n_out, q)
            L.call("gmr1b200_fcch_acquire_batch", 0, S["out"], n_own * n_out, f_ofs, 0, FCCH_WIN, SPS, S["f_toa"], S["f_align"],
                   S["f_ferr"], n_own, q)
            for kind in ("bcch", "dc6"):
                r = S["res"][kind]
                n = r["crc"].numel()
                L.call("gmr1b200_pi4cxpsk_demod_batch", BT[kind], S["out"], n_own * n_out, r["ofs"], 0, wlen(kind), SPS, None, 0.0,
                       r["eb"], EBITS[kind], None, None, None, None, n, q)
                L.call("gmr1b200_bcch_decode_batch" if kind == "bcch" else "gmr1b200_ccch_decode_batch", r["l2"], r["eb"], None,
                       r["crc"], n, q)
            with torch.cuda.stream(S["st"]):
                for kind in ("bcch", "dc6"):
                    S["res"][kind]["h_l2"].copy_(S["res"][kind]["l2"], non_blocking=True)
                    S["res"][kind]["h_crc"].copy_(S["res"][kind]["crc"], non_blocking=True)
                S["h_f"][0].copy_(S["f_align"].to(torch.float32), non_blocking=True)
                S["h_f"][1].copy_(S["f_ferr"], non_blocking=True)

        for S in slots:
            issue(S)
        barrier()
        n_pipe = max(steps, 4)
        t0 = time.perf_counter()
        issue(slots[0])
        for i in range(1, n_pipe):
            issue(slots[i % 2])
            slots[(i - 1) % 2]["st"].synchronize()
        slots[(n_pipe - 1) % 2]["st"].synchronize()
        barrier()
        pipe_ms = rank_max(1e3 * (time.perf_counter() - t0) / n_pipe)
        same = all(bool((slots[1]["res"][k]["h_l2"] == res[k]["h_l2"]).all()) and bool((slots[1]["res"][k]["h_crc"] == res[k]["h_crc"]).all())
                   for k in res)
        pipe = {"value": world * sum(n_arfcn * n_b[kk] for kk in n_b) / (pipe_ms * 1e-3), "unit": "bursts/s", "ms_per_recording": pipe_ms,
                "recordings_in_flight": 2, "h2d_gbs_per_gpu": n_wide * sbytes / (pipe_ms * 1e-3) / 1e9, "same_results_in_both_slots": same,
                "how": "the e2e step issued for recording i + 1 before recording i is waited for (own stream and buffers each)"}
        del slots
    # what came out: payloads against what was sent, FCCH positions against where the chirps were put
    nb = sum(n_arfcn * n_b[kk] for kk in n_b)                  # bursts of the whole recording
    nb_own = sum(len(sel[kk]) for kk in n_b)
    ok = sum(int((res[kk]["h_crc"] == 0).sum()) for kk in n_b)
    good = sum(int(((res[kk]["h_l2"].numpy() == par[kk]["l2"][sel[kk]]).all(axis=1) & (res[kk]["h_crc"].numpy() == 0)).sum()) for kk in n_b)
    wrong = sum(int(((res[kk]["h_l2"].numpy() != par[kk]["l2"][sel[kk]]).any(axis=1) & (res[kk]["h_crc"].numpy() == 0)).sum()) for kk in n_b)
    toa_err = h_f[0].numpy() - (fpos[own] + info.delay_out - dly)
    jobs = 1 if shared else world          # recordings processed per step by the whole job
    n_steps = n_wide // (n_arfcn // 2)
    bank_bytes = n_wide * sbytes + n_steps * n_arfcn * 8
    rs_bytes = n_steps * n_own * 8 + n_own * n_out * 8
    L.c.gmr1b200_chan_destroy(h)
    return {
        "iq_format": "int16" if fmt == 1 else "int8",
        "what": "the headline receive work fed from ONE wideband int16 / int8 recording of all ARFCNs through the GPU channeliser "
                "(replaces the PFB mode of utils/gmr1_rx_sdr.py) instead of one complex-float stream per ARFCN",
        "n_gpus": world, "scaling": "strong" if shared else "weak",
        "per_gpu": ("ONE recording for the whole job: every rank copies a 1 / n_gpus time slice from pinned host memory, NCCL "
                    "all-gather over NVLink, every GPU runs the bank and receives the ARFCNs a mod n_gpus == rank; value = "
                    "the recording's bursts / slowest rank's time; the check fields are rank 0's share") if shared else
                   ("every rank has its own wideband recording of n_arfcn ARFCNs (weak scaling; value = all ranks' bursts / "
                    "slowest rank's time; the check fields are rank 0's)"),
        "arfcns": n_arfcn, "bursts": nb, "fcch_acquisitions": n_arfcn, "recording_seconds": n_wide / info.samp_rate,
        "wideband_rate_msps": info.samp_rate / 1e6, "bank_taps": info.n_taps, "rrc_taps": info.n_taps_resamp,
        "e2e": {"value": jobs * nb / (e2e_ms * 1e-3), "unit": "bursts/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(host_wide.numel() * sbytes // 2), "d2h_bytes_per_step": int(nb_own * 28 + n_own * 8),
                "nvlink_allgather_bytes_per_gpu": int(wide_full.numel() * 2) if shared else 0,
                "per_arfcn_cf32_bytes_equivalent": int(n_arfcn * n_out * 8),
                "h2d_gbs_per_gpu": n_wide * sbytes / (e2e_ms * 1e-3) / 1e9,
                "how": "pinned host int16 recording -> gmr1b200_channelize (H2D in pieces under the bank + resampler "
                       "kernels) -> fcch_acquire + demod + decode on offsets into the device-resident streams -> host "
                       "L2 / CRC / alignments"},
        "e2e_pipelined": pipe,
        "device_resident": {"bursts_per_s": jobs * nb / (dev_ms * 1e-3), "ms_per_step": dev_ms, "ms": {k_: round(v, 4) for k_, v in part.items()},
                            "channelizer_input_msps": n_wide / (part["channelize"] * 1e-3) / 1e6,
                            "channelizer_algorithmic_gbs": (bank_bytes + rs_bytes) / (part["channelize"] * 1e-3) / 1e9,
                            "realtime_factor": (n_wide / info.samp_rate) / (dev_ms * 1e-3), "launches_per_step": launches},
        "crc_ok_frac": ok / nb_own, "payload_correct_frac": good / nb_own, "crc_ok_but_payload_wrong": wrong,
        "fcch_found_frac": float((np.abs(toa_err) <= 2).mean()),
        "esn0_db": esn0, "int16_peak" if fmt == 1 else "int8_peak": peak_i16, "generation_s": gen_s,
    }



def run_pool_e2e(L, torch, W, host_iq, host_fcch, n_gpus, steps):
    """host IQ of n_gpus x (this GPU's ARFCNs) in pinned memory -> gmr1b200_pool_fcch_acquire + gmr1b200_pool_rx_xcch
    (BCCH, DC6): ARFCN a is processed on device a mod n_gpus, host-side gather of L2 / CRC.  The content of ARFCN a
    is that of local ARFCN a mod 1024, so the results must equal the device-resident pass replicated."""
    have = torch.cuda.device_count()
    if have < n_gpus:
        return {"skipped": f"{have} GPU(s) visible, {n_gpus} asked for"}
    per = {k: W.n[k] // W.n_arfcn for k in W.iq}
    rep = lambda t: t.repeat((n_gpus,) + (1,) * (t.dim() - 1)).pin_memory()
    iq = {k: rep(host_iq[k]) for k in W.iq}
    fc = rep(host_fcch)
    n_arfcn = W.n_arfcn * n_gpus
    l2 = {k: torch.empty((W.n[k] * n_gpus, 24), dtype=torch.uint8).pin_memory() for k in W.iq}
    crc = {k: torch.empty(W.n[k] * n_gpus, dtype=torch.int32).pin_memory() for k in W.iq}
    align = torch.empty(n_arfcn, dtype=torch.int32).pin_memory()
    ferr = torch.empty(n_arfcn, dtype=torch.float32).pin_memory()
    pool = L.pool_create(list(range(n_gpus)), streams_per_dev=3, chunk_bytes=64 << 20)

    def step():
        L.call("gmr1b200_pool_fcch_acquire", pool, 0, fc, n_arfcn, FCCH_WIN, SPS, None, align, ferr)
        for k in W.iq:
            L.call("gmr1b200_pool_rx_xcch", pool, CHAN[k], iq[k], n_arfcn, per[k], wlen(k), SPS, 0.0, l2[k], crc[k], None, None)

    try:
        step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
    finally:
        L.pool_destroy(pool)
    same = all(bool((l2[k].view(n_gpus, -1) == W.l2[k].cpu().view(1, -1)).all()) and
               bool((crc[k].view(n_gpus, -1) == W.crc[k].cpu().view(1, -1)).all()) for k in W.iq)
    nb = sum(W.n.values()) * n_gpus
    h2d = sum(t.numel() * 4 for t in iq.values()) + fc.numel() * 4
    return {"devices": n_gpus, "value": nb / dt, "unit": "bursts/s", "ms_per_step": 1e3 * dt,
            "h2d_bytes_per_step": int(h2d), "h2d_gbs_total": h2d / dt / 1e9, "same_results_as_device_path": same,
            "how": "one process, gmr1b200_pool_*: one feeder thread + 3 streams per GPU, ARFCN a on device a mod G, 64 MB chunks"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_evt = index, [], threading.Event()
        # NVML in-process when the bindings are there (a sample per 2 ms: the timed region is tens of ms long),
        # else one nvidia-smi call per sample (~50 ms each).  Initialised here, before the timed region starts.
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[index]) if vis and vis.split(",")[index].strip().isdigit() else index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                    pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            self.nvml = (pynvml, h, pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM), get_reasons, bits)
        except Exception:
            self.nvml = None

    def run(self):
        if self.nvml:
            pynvml, h, mx, get_reasons, bits = self.nvml
            try:
                while not self.stop_evt.is_set():
                    r = get_reasons(h)
                    self.samples.append([str(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)), str(mx)] +
                                        ["Active" if r & b else "Not Active" for b in bits])
                    self.stop_evt.wait(0.002)
                return
            except Exception:
                pass
        while not self.stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [s.strip() for s in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_evt.wait(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(s[2 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


class Workload:
    """device-resident synthetic recording of this rank's ARFCNs + all result buffers"""

    def __init__(self, L, torch, n_arfcn, per_arfcn, seed, dev):
        self.L, self.torch = L, torch
        self.n = {"bcch": n_arfcn * per_arfcn // 2, "dc6": n_arfcn * per_arfcn // 2}
        self.iq, self.eb, self.l2, self.crc, self.par = {}, {}, {}, {}, {}
        self.sid, self.toa = {}, {}
        for kind in ("bcch", "dc6"):
            n, wl = self.n[kind], wlen(kind)
            p = burst_params(n, kind, seed + BT[kind])
            hard = np.zeros((n, EBITS[kind]), np.uint8)
            L.call("gmr1b200_xcch_encode_batch", CHAN[kind], hard, p["l2"], n)
            iq = torch.empty((n, wl, 2), dtype=torch.float32, device=dev)
            d = lambda a: torch.from_numpy(a).to(dev)
            L.call("gmr1b200_synth_bursts", BT[kind], d(hard), EBITS[kind], None, SPS, wl, d(p["toa"]), 0.0,
                   d(p["cfo"]), 0.0, d(p["phase"]), 0.0, d(p["esn0"]), 0.0, None, 1.0, seed * 7919 + BT[kind],
                   iq, n * wl, None, wl, n, None)
            torch.cuda.synchronize()
            self.iq[kind], self.par[kind] = iq, p
            self.eb[kind] = torch.empty((n, EBITS[kind]), dtype=torch.int8, device=dev)
            self.l2[kind] = torch.empty((n, 24), dtype=torch.uint8, device=dev)
            self.crc[kind] = torch.empty(n, dtype=torch.int32, device=dev)
            self.sid[kind] = torch.empty(n, dtype=torch.int32, device=dev)
            self.toa[kind] = torch.empty(n, dtype=torch.float32, device=dev)

        # one 330 ms FCCH search window per ARFCN (gmr1_rx.c:612): noise + one dual chirp at a random
        # offset with up to +-1 kHz of carrier offset; acquired once per step (rough + fine)
        self.n_arfcn = n_arfcn
        self.fcch_pos, self.fcch_cfo, blocks = fcch_windows(n_arfcn, seed + 99)
        fw = torch.empty((n_arfcn, FCCH_WIN, 2), dtype=torch.float32, device=dev)
        for lo, hi, x in blocks:
            fw[lo:hi] = torch.from_numpy(x.view(np.float32).reshape(hi - lo, FCCH_WIN, 2)).to(dev)
        self.fcch_iq = fw
        self.fcch_toa = torch.empty(n_arfcn, dtype=torch.int32, device=dev)
        self.fcch_align = torch.empty(n_arfcn, dtype=torch.int32, device=dev)
        self.fcch_ferr = torch.empty(n_arfcn, dtype=torch.float32, device=dev)

    def total(self):
        return self.n["bcch"] + self.n["dc6"]

    def fcch(self, stream, torch_stream):
        """FCCH acquisition of every ARFCN: rough TOA over the 330 ms window, then the fine timing /
        frequency estimate on the 117-symbol burst found (fcch_single_init, gmr1_rx.c:606-639)"""
        n = self.n_arfcn
        self.L.call("gmr1b200_fcch_acquire_batch", 0, self.fcch_iq, n * FCCH_WIN, None, FCCH_WIN, FCCH_WIN, SPS,
                    self.fcch_toa, self.fcch_align, self.fcch_ferr, n, stream)

    def demod(self, kind, stream, iq=None, lo=0, hi=None, eb=None):
        hi = self.n[kind] if hi is None else hi
        wl = wlen(kind)
        iq = self.iq[kind][lo:hi] if iq is None else iq
        eb = self.eb[kind][lo:hi] if eb is None else eb
        self.L.call("gmr1b200_pi4cxpsk_demod_batch", BT[kind], iq, (hi - lo) * wl, None, wl, wl, SPS, None, 0.0,
                    eb, EBITS[kind], self.sid[kind][lo:hi], self.toa[kind][lo:hi], None, None, hi - lo, stream)

    def decode(self, kind, stream, lo=0, hi=None, eb=None, l2=None, crc=None):
        hi = self.n[kind] if hi is None else hi
        eb = self.eb[kind][lo:hi] if eb is None else eb
        l2 = self.l2[kind][lo:hi] if l2 is None else l2
        crc = self.crc[kind][lo:hi] if crc is None else crc
        fn = "gmr1b200_bcch_decode_batch" if kind == "bcch" else "gmr1b200_ccch_decode_batch"
        self.L.call(fn, l2, eb, None, crc, hi - lo, stream)


def run_gpu_arm(args):
    # the JSON line must be the only thing on stdout: library chatter (e.g. NCCL's version banner) goes to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()
    L.init(local_rank)
    if args.demod_generic:
        L.call("gmr1b200_set_demod_generic", 1)

    W = Workload(L, torch, args.arfcns, args.bursts_per_arfcn, 1000 + rank, dev)
    nb = W.total()
    stream = torch.cuda.Stream(device=dev)
    sh = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident: value + roofline of the demod kernel
    ev = lambda: torch.cuda.Event(enable_timing=True)

    stream2 = torch.cuda.Stream(device=dev)
    stream3 = torch.cuda.Stream(device=dev)
    three = args.streams >= 3
    two = args.streams >= 2

    def step(timers=None):
        """one pass over the batch.  --streams 2: the BCCH and DC6/CCCH halves are independent, each
        runs demod -> decode on its own stream so that the decode (integer-ALU-bound) of one half
        shares the SMs with the demod of the other; --streams 3 (default): the FCCH acquisitions, a third
        independent piece of work, get their own stream too.  The per-kernel event timers are only
        meaningful with --streams 1 (serial), which is how the roofline pass below is run."""
        if timers is not None:
            f0, f1 = ev(), ev()
            f0.record(stream)
        sf = stream3 if (three and timers is None) else stream
        W.fcch(sf.cuda_stream, sf)
        if timers is not None:
            f1.record(stream)
            timers.append(("fcch", f0, f1))
        for kind, st in (("bcch", stream), ("dc6", stream2 if (two and timers is None) else stream)):
            if timers is not None:
                a, b = ev(), ev()
                a.record(st)
            W.demod(kind, st.cuda_stream)
            if timers is not None:
                b.record(st)
                timers.append((kind, a, b))
            W.decode(kind, st.cuda_stream)
            if timers is not None:
                c = ev()
                c.record(st)
                timers.append(("decode_" + kind, b, c))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = L.kernel_launches()
    t_start, t_end = ev(), ev()
    stream2.wait_stream(stream)
    stream3.wait_stream(stream)
    t_start.record(stream)
    stream2.wait_event(t_start)
    stream3.wait_event(t_start)
    for _ in range(args.steps):
        step()
    stream.wait_stream(stream2)
    stream.wait_stream(stream3)
    t_end.record(stream)
    barrier()
    launches = L.kernel_launches() - launches0
    ms_total = t_start.elapsed_time(t_end)
    # roofline pass: the same K steps serially on one stream with CUDA events around every demod launch
    timers = []
    r_start, r_end = ev(), ev()
    r_start.record(stream)
    for _ in range(args.steps):
        step(timers)
    r_end.record(stream)
    barrier()
    ms_serial = r_start.elapsed_time(r_end)
    sampler.stop_evt.set()
    sampler.join()
    # the same loop until >= --min-seconds have passed (the K-step region above is the contract's number; at ~1 ms per
    # step it is only ~20 ms long, so the long run sits next to it)
    s_steps, s_ms = 0, 0.0
    sampler2 = ClockSampler(local_rank)
    sampler2.start()
    if args.min_seconds > 0:
        per = max(ms_total / args.steps, 1e-3)
        s_steps = int(np.ceil(1e3 * args.min_seconds / per))
        s0, s1 = ev(), ev()
        stream2.wait_stream(stream)
        stream3.wait_stream(stream)
        s0.record(stream)
        stream2.wait_event(s0)
        stream3.wait_event(s0)
        for _ in range(s_steps):
            step()
        stream.wait_stream(stream2)
        stream.wait_stream(stream3)
        s1.record(stream)
        barrier()
        s_ms = s0.elapsed_time(s1)
    sampler2.stop_evt.set()
    sampler2.join()

    # ---------------- end to end: pinned host IQ in, host L2/CRC out, through the C ABI
    chunks = args.e2e_chunks
    # the pinned staging buffers are allocated (first touched) from the CPUs next to this rank's GPU, so that with
    # several ranks per host every H2D stream reads memory of its own NUMA node; the affinity is restored afterwards
    old_affinity = None
    try:
        import pynvml
        pynvml.nvmlInit()
        old_affinity = os.sched_getaffinity(0)
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        if not os.sched_getaffinity(0):
            os.sched_setaffinity(0, old_affinity)
    except Exception:
        pass
    host_iq = {k: torch.empty(W.iq[k].shape, dtype=torch.float32).pin_memory() for k in W.iq}
    host_l2 = {k: torch.empty((W.n[k], 24), dtype=torch.uint8).pin_memory() for k in W.iq}
    host_crc = {k: torch.empty(W.n[k], dtype=torch.int32).pin_memory() for k in W.iq}
    for k in W.iq:
        host_iq[k].copy_(W.iq[k])
    host_fcch = torch.empty(W.fcch_iq.shape, dtype=torch.float32).pin_memory()
    host_fcch.copy_(W.fcch_iq)
    host_fcch_out = torch.empty((2, W.n_arfcn), dtype=torch.float32).pin_memory()
    torch.cuda.synchronize()
    if old_affinity:
        try:
            os.sched_setaffinity(0, old_affinity)
        except Exception:
            pass
    jobs = []
    for k in ("bcch", "dc6"):
        per = (W.n[k] + chunks - 1) // chunks
        jobs += [(k, lo, min(W.n[k], lo + per)) for lo in range(0, W.n[k], per)]
    n_thr = 3
    streams = [torch.cuda.Stream(device=dev) for _ in range(n_thr)]

    def e2e_thread(t):
        torch.cuda.set_device(local_rank)
        s = streams[t].cuda_stream
        if t == 0:
            # FCCH windows of every ARFCN: host -> device, acquire, TOA / frequency error back to the host
            with torch.cuda.stream(streams[0]):
                W.fcch_iq.copy_(host_fcch, non_blocking=True)
            W.fcch(s, streams[0])
            with torch.cuda.stream(streams[0]):
                host_fcch_out[0].copy_(W.fcch_align.to(torch.float32), non_blocking=True)
                host_fcch_out[1].copy_(W.fcch_ferr, non_blocking=True)
            streams[0].synchronize()
        for k, lo, hi in jobs[t::n_thr]:
            # host IQ -> (library stages it) -> demod -> ebits stay on the device -> decode -> host L2/CRC
            W.demod(k, s, iq=host_iq[k][lo:hi], lo=lo, hi=hi)
            W.decode(k, s, lo=lo, hi=hi, l2=host_l2[k][lo:hi], crc=host_crc[k][lo:hi])

    def e2e_step():
        th = [threading.Thread(target=e2e_thread, args=(t,)) for t in range(n_thr)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    # the ceiling under the end-to-end number: the same bytes from the same pinned buffers on the same threads and
    # streams, copies only (no kernels, no results)
    def h2d_thread(t):
        torch.cuda.set_device(local_rank)
        with torch.cuda.stream(streams[t]):
            if t == 0:
                W.fcch_iq.copy_(host_fcch, non_blocking=True)
            for k, lo, hi in jobs[t::n_thr]:
                W.iq[k][lo:hi].copy_(host_iq[k][lo:hi], non_blocking=True)
        streams[t].synchronize()

    def threads_step(fn):
        th = [threading.Thread(target=fn, args=(t,)) for t in range(n_thr)]
        for t in th:
            t.start()
        for t in th:
            t.join()

    threads_step(h2d_thread)
    barrier()
    p_steps = 3
    t0 = time.perf_counter()
    for _ in range(p_steps):
        threads_step(h2d_thread)
    barrier()
    h2d_ms = 1e3 * (time.perf_counter() - t0) / p_steps

    for _ in range(2):
        e2e_step()
    barrier()
    e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(e_steps):
        e2e_step()
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / e_steps

    # ---------------- reduce over ranks (time = max)
    t = torch.tensor([ms_total, e2e_ms, h2d_ms, s_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, h2d_ms, s_ms = (float(v) for v in t)
    ms_step = ms_total / args.steps

    crc_ok = float(sum(int((W.crc[k] == 0).sum()) for k in W.crc)) / nb
    # the e2e pass must have produced the same answers as the device-resident pass
    e2e_same = all(bool((host_l2[k] == W.l2[k].cpu()).all()) and bool((host_crc[k] == W.crc[k].cpu()).all())
                   for k in W.l2)

    # ---------------- the whole end-to-end step of several GPUs from ONE process through the library's device pool
    pool_e2e = None
    if args.single_process and world == 1 and args.gpus > 1:
        pool_e2e = run_pool_e2e(L, torch, W, host_iq, host_fcch, args.gpus, min(args.steps, 5))

    # ---------------- config 5 sweep (all ranks)
    sweep = None
    del host_iq, host_fcch
    if not args.no_sweep:
        totals = [1024 * k for k in (1, 2, 4, 8, 16, 32, 64)]
        if args.sweep_max:
            totals = [t for t in totals if t <= args.sweep_max]
        sweep = run_sweep(L, torch, dist if world > 1 else None, dev, world, rank, barrier, totals, args.sweep_chunk)

    # ---------------- the same work from one wideband recording per GPU through the channeliser (every rank)
    wideband = None
    if not args.no_wideband:
        torch.cuda.empty_cache()
        wideband = run_wideband(L, torch, dev, args.arfcns, args.bursts_per_arfcn, min(args.steps, 5), seed=4242 + rank,
                                world=world, dist=dist if world > 1 else None)
        if world == 1:                     # the same recording as int8 I/Q (8-bit front ends): half the bytes again
            torch.cuda.empty_cache()
            w8 = run_wideband(L, torch, dev, args.arfcns, args.bursts_per_arfcn, min(args.steps, 5), seed=4242 + rank, fmt=2)
            wideband["int8_recording"] = {k: w8[k] for k in ("iq_format", "e2e", "e2e_pipelined", "device_resident", "crc_ok_frac",
                                                             "payload_correct_frac", "crc_ok_but_payload_wrong",
                                                             "fcch_found_frac", "int8_peak")}
        if world > 1:                      # and ONE recording shared by all GPUs (all-gather over NVLink): strong scaling
            torch.cuda.empty_cache()
            wideband["shared_capture"] = run_wideband(L, torch, dev, args.arfcns, args.bursts_per_arfcn, min(args.steps, 5),
                                                      seed=4242, world=world, dist=dist, shared=True, rank=rank)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (demod), timed live with CUDA events
    fcc = [(k, a, b) for k, a, b in timers if k == "fcch"]
    timers = [t for t in timers if t[0] != "fcch"]
    fcch_ms = sum(a.elapsed_time(b) for _, a, b in fcc) / max(1, len(fcc))
    toa_err = W.fcch_align.cpu().numpy() - W.fcch_pos
    ferr_hz = W.fcch_ferr.cpu().numpy() * 23400.0 / (2 * np.pi) - W.fcch_cfo * 23400.0 / (2 * np.pi)
    fcch = {"acquisitions_per_step": W.n_arfcn, "ms_per_step": fcch_ms,
            "acquisitions_per_s": W.n_arfcn / (fcch_ms * 1e-3),
            "window_bytes": FCCH_WIN * 8, "achieved_gbs": W.n_arfcn * (FCCH_WIN * 8 + 3752) / (fcch_ms * 1e-3) / 1e9,
            "found_frac": float((np.abs(toa_err) <= 2).mean()), "freq_err_rms_hz": float(np.sqrt((ferr_hz ** 2).mean())),
            "what": "gmr1_fcch_rough over a 330 ms window + gmr1_fcch_fine, once per ARFCN per step"}
    dem = [(k, a, b) for k, a, b in timers if not k.startswith("decode_")]
    dec = [(k[7:], a, b) for k, a, b in timers if k.startswith("decode_")]
    dem_ms = sum(a.elapsed_time(b) for _, a, b in dem)
    dem_bytes = sum(DEMOD_BYTES[k] * W.n[k] for k, _, _ in dem)
    dec_ms = sum(a.elapsed_time(b) for _, a, b in dec)
    dec_cw = sum(W.n[k] for k, _, _ in dec)
    timers = dem
    achieved = dem_bytes / (dem_ms * 1e-3) / 1e9
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    try:
        mp_ = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(mp_["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    # ncu counters of the dominant kernels (DRAM bytes, warp-instructions per launch): read from the committed summary
    # of tools/kernel_counters.py, and only used when that summary was taken on the build that is being timed
    traffic, counters, build = None, None, L.version().split("build ")[-1]
    try:
        kc = json.load(open(os.path.join(ROOT, "profiles", "kernel_counters.json")))
        if kc.get("build") == build:
            counters = kc["kernels"]
            d = [v for k, v in counters.items() if k.startswith("demod")]
            traffic = sum(v["dram_bytes"] for v in d) / len(d)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "demod_kernel (BCCH + DC6 launches)", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel_share_of_step": dem_ms / ms_serial, "measured_in": "serial pass (1 stream), same K steps",
                "serial_ms_per_step": ms_serial / args.steps,
                "bytes_per_launch": dem_bytes / len(timers), "ms_per_launch": dem_ms / len(timers),
                "traffic_source": ("profiles/kernel_counters.json (ncu, this build)" if traffic else
                                   "none: profiles/kernel_counters.json is absent or from another build"),
                "per_format": {k: DEMOD_BYTES[k] * W.n[k] / (sum(a.elapsed_time(b) for kk, a, b in dem if kk == k) /
                                                               sum(1 for kk, _, _ in dem if kk == k) * 1e-3) / 1e9 / peak
                               for k in W.n}}

    # Viterbi kernel: integer-ALU-bound; 212 trellis steps x 16 states per BCCH/CCCH codeword
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    alu_peak = sms * 128 * 1.965e9            # int32 lane-ops/s at the max SM clock (128 lanes per SM)
    viterbi = {"kernel": "decode_tpc_kernel<BCCH|CCCH>", "codewords_per_s": dec_cw / (dec_ms * 1e-3),
               "acs_state_updates_per_s": 3392 * dec_cw / (dec_ms * 1e-3), "ms_per_launch": dec_ms / len(dec),
               "thread_instr_per_state_update": None, "frac_of_int32_issue_peak": None,
               "note": "ACS rate is measured here; instructions per state update come from ncu smsp__inst_executed of "
                       "this build (profiles/kernel_counters.json) and are omitted when that file is from another build; "
                       "peak = SMs x 128 lanes x 1.965 GHz"}
    if counters:
        d = [v for k, v in counters.items() if k.startswith("decode")]
        if d:
            ipu = sum(v["warp_inst"] for v in d) * 32.0 / (sum(v["units"] for v in d) * 3392)
            viterbi["thread_instr_per_state_update"] = ipu
            viterbi["frac_of_int32_issue_peak"] = ipu * 3392 * dec_cw / (dec_ms * 1e-3) / alu_peak

    # ---------------- CPU baseline leg (rank 0, N = 1): reference C path on a bounded sample + parity
    cpu, configs = None, None
    cores = os.cpu_count() or 1
    if world == 1 and not args.no_cpu_baseline:
        m = min(W.n["bcch"], max(1024, 4096 * cores))      # ~20 core-seconds of reference C work
        n_f = min(W.n_arfcn, max(1, 2 * m // args.bursts_per_arfcn))
        files = {}
        for k in ("bcch", "dc6"):
            x = W.iq[k][:m].cpu().numpy().view(np.complex64).reshape(m, wlen(k))
            files[k] = (save_shm(k, x), m, 1, {})
        files["fcch"] = (save_shm("fcch", W.fcch_iq[:n_f].cpu().numpy().view(np.complex64).reshape(n_f, FCCH_WIN)), n_f, 1, {})
        pool = mp.get_context("spawn").Pool(cores)
        pool.map(_cpu_noop, range(cores))
        wall, out, okind, flags = cpu_reference_pass(files, cores, pool)
        # one core alone on a slice of the same sample (BASELINE.md run B2: bursts/s/core)
        m1 = min(m, 2048)
        one = {k: (files[k][0], m1, m1, {}) for k in ("bcch", "dc6")}
        one["fcch"] = (files["fcch"][0], max(1, 2 * m1 // args.bursts_per_arfcn), 8, {})
        wall1, _, _, _ = cpu_reference_pass(one, 1, pool)
        pool.close()
        for path, _, _, _ in files.values():
            os.unlink(path)
        acq = out.pop("fcch")
        g_toa = W.fcch_align[:n_f].cpu().numpy()
        fcch["toa_identical_to_reference"] = bool((acq[0] == g_toa).all())
        fcch["freq_err_max_abs_diff_vs_reference_rad_per_sym"] = float(
            np.abs(acq[1] - W.fcch_ferr[:n_f].cpu().numpy()).max())
        same = all(bool((out[k][0] == W.l2[k][:m].cpu().numpy()).all()) and
                   bool((out[k][1] == W.crc[k][:m].cpu().numpy()).all()) for k in out)
        cpu = {"value": 2 * m / wall, "unit": "bursts/s", "cores": cores, "kind": okind, "flags": flags,
               "sample": f"first {m} BCCH + {m} DC6 bursts and {n_f} FCCH windows of the GPU workload, one process "
                         f"per core, compiled loops (oracle/harness.c), {wall:.2f} s wall",
               "single_core_bursts_per_s": 2 * m1 / wall1,
               "l2_crc_identical_to_gpu": same}

    # ---------------- BASELINE configs 3 and 4 at full size (N = 1)
    if world == 1 and not args.no_configs:
        configs = run_configs(L, torch, dev, peak, 0 if args.no_cpu_baseline else cores, args.config_scale, 5, 4096)

    h2d_bytes = int(sum(W.iq[k].numel() * 4 for k in W.iq) + W.fcch_iq.numel() * 4)
    line = {
        "metric": "decoded bursts/s (FCCH sync+demod+Viterbi)", "value": nb * world / (ms_step * 1e-3),
        "unit": "bursts/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i32", "data": "synthetic",
        "config": {"workload": "config2: batched BCCH + DC6/CCCH bursts (pi/4-CQPSK demod + K5 r1/2 Viterbi + CRC16), "
                               f"{args.arfcns} ARFCNs x {args.bursts_per_arfcn} bursts per GPU, sps 4",
                   "bursts_per_gpu": nb, "iq_bytes_per_gpu": int(sum(W.iq[k].numel() * 4 for k in W.iq)),
                   "fcch_iq_bytes_per_gpu": int(W.fcch_iq.numel() * 4),
                   "l2_flush": "inputs (2.1 GB) larger than L2", "esn0_db": [6, 10, 15, 30],
                   "parallelism": f"arfcn-sharded x{world}, no collective",
                   "streams_per_gpu": args.streams},
        "crc_ok_frac": crc_ok,
        "sustained": {"steps": s_steps, "seconds": s_ms * 1e-3,
                      "value": (nb * world * s_steps / (s_ms * 1e-3)) if s_ms > 0 else None, "unit": "bursts/s",
                      "clocks": sampler2.summary(),
                      "what": "the same step loop run for >= --min-seconds, after the K-step region and the roofline pass "
                              "(a second of back-to-back steps reaches the board's power cap: sw_power_cap is expected here)"},
        "e2e": {"value": nb * world / (e2e_ms * 1e-3), "unit": "bursts/s",
                "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": int(nb * 28 + W.n_arfcn * 8), "ms_per_step": e2e_ms,
                "h2d_gbs_per_gpu": h2d_bytes / (e2e_ms * 1e-3) / 1e9,
                "h2d_ceiling_gbs_per_gpu": h2d_bytes / (h2d_ms * 1e-3) / 1e9,
                "frac_of_ceiling": h2d_ms / e2e_ms,
                "ceiling_how": "the same pinned buffers, chunks, threads and streams with the copies only (no kernels), "
                               "all ranks at once, max over ranks",
                "single_process_pool": pool_e2e,
                "how": f"{len(jobs)} chunks on {n_thr} host threads/streams through gmr1b200_*_batch with pinned "
                       "host IQ in and host L2/CRC out", "same_results_as_device_path": e2e_same},
        "gpu_launches": int(launches),
        "roofline": roofline,
        "viterbi": viterbi,
        "fcch": fcch,
        "cpu_baseline": cpu,
        "configs": configs,
        "wideband": wideband,
        "sweep": sweep,
        "build": L.version(),
        "clocks": sampler.summary(),
    }
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--arfcns", type=int, default=1024)
    ap.add_argument("--bursts-per-arfcn", type=int, default=256)
    ap.add_argument("--e2e-chunks", type=int, default=8)
    ap.add_argument("--streams", type=int, default=3, choices=[1, 2, 3])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3 and 4")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-5 ARFCN sweep")
    ap.add_argument("--no-wideband", action="store_true", help="skip the wideband-recording (channeliser) leg")
    ap.add_argument("--config-scale", type=float, default=1.0, help="fraction of the ARFCN counts of configs 3 / 4")
    ap.add_argument("--sweep-max", type=int, default=0, help="largest total ARFCN count of the sweep (0: 65536)")
    ap.add_argument("--sweep-chunk", type=int, default=1024, help="ARFCN slices per streamed chunk")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="length of the sustained run (0: skip)")
    ap.add_argument("--single-process", action="store_true",
                    help="with --gpus N (not under torchrun): additionally run the end-to-end step of all N GPUs from "
                         "this one process through the library's device pool")
    ap.add_argument("--demod-generic", action="store_true",
                    help="A/B: force the generic demodulation kernel instead of the per-format ones")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
