"""Drop-in test (BASELINE.json config 1): the reference's OWN receiver application, src/gmr1_rx.c,
compiled unmodified and linked against libgmr1_b200.so (tests/dropin/Makefile), runs on a synthetic
single-ARFCN recording next to the same application linked against the reference's C libraries
(oracle/_ref/gmr1_rx).  Both must acquire the same FCCH, walk the same frames, and report the same
CRC result for every BCCH / CCCH burst; Viterbi metrics and frequency estimates agree to tolerance.
"""
import os
import re
import subprocess

import numpy as np
import pytest

import recording

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "gmr1_rx")
GPU_BIN = os.path.join(ROOT, "tests", "dropin", "_build", "gmr1_rx_b200")


def _run(binary, path):
    r = subprocess.run([binary, "4", path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stderr.split("\n")


@pytest.mark.parametrize("esn0,cfo", [(15.0, 300.0), (8.0, -450.0)])
def test_gmr1_rx_unmodified_on_gpu_library(gpu_lib, oracle, tmp_path, esn0, cfo):
    if not (os.path.exists(REF_BIN) and os.path.exists(GPU_BIN)):
        pytest.skip("drop-in binaries not built (need /root/reference at build time)")
    x, truth = recording.make(lambda l2: oracle.encode("bcch", 424, l2), lambda l2: oracle.encode("ccch", 432, l2),
                              esn0_db=esn0, cfo_hz=cfo, seed=int(esn0))
    path = str(tmp_path / "cfg1.cfile")
    x.tofile(path)
    ref, gpu = _run(REF_BIN, path), _run(GPU_BIN, path)

    def parse(lines):
        crc = [(int(m.group(1)), int(m.group(2))) for m in (re.match(r"crc=(-?\d+), conv=(-?\d+)", l) for l in lines) if m]
        kinds = [l.strip() for l in lines if l.startswith("[.]   ")]
        fn = [l for l in lines if l.startswith("[-]  FN:")]
        acq = [m for m in (re.match(r"\[\+\] Processing BCCH @(\d+) .*freq_err = (-?[\d.]+) Hz", l) for l in lines) if m]
        return crc, kinds, fn, acq

    crc_r, kinds_r, fn_r, acq_r = parse(ref)
    crc_g, kinds_g, fn_g, acq_g = parse(gpu)
    assert len(acq_r) == len(acq_g) >= 1
    for a, b in zip(acq_r, acq_g):
        assert abs(int(a.group(1)) - int(b.group(1))) <= 1                  # FCCH alignment (samples)
        assert abs(float(a.group(2)) - float(b.group(2))) <= 0.5            # frequency error (Hz)
    assert fn_r == fn_g and kinds_r == kinds_g                              # same frames, same burst kinds
    assert len(crc_r) == len(crc_g) >= 40
    assert [c for c, _ in crc_r] == [c for c, _ in crc_g]                   # identical CRC outcomes
    assert max(abs(a - b) for (_, a), (_, b) in zip(crc_r, crc_g)) <= 8     # Viterbi metric: soft bits +-1 LSB
    assert sum(c == 0 for c, _ in crc_g) >= 0.9 * len(crc_g)
