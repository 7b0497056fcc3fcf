import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def emu():
    """CPU build of the product's __host__ __device__ decode code (test harness only)."""
    import ctypes
    so = os.path.join(ROOT, "tests", "emu", "libgmr1_emu.so")
    src = os.path.join(ROOT, "tests", "emu", "decode_emu.cpp")
    csrc = os.path.join(ROOT, "osmo_gmr_b200", "csrc")
    deps = [src] + [os.path.join(csrc, f) for f in os.listdir(csrc)]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas",
                               "-I" + csrc, "-o", so, src, os.path.join(csrc, "gmr1_tables.cpp")])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def gpu_lib():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import osmo_gmr_b200
    L = osmo_gmr_b200.lib()      # raises if the extension is missing: no CPU fallback
    L.init(0)
    return L


@pytest.fixture
def port_noquirk(monkeypatch):
    """the oracle port with the sync accumulator reset per candidate (GMR1_ORACLE_SYNC_RESET)"""
    import os
    import subprocess
    import oracle_lib
    so = os.path.join(oracle_lib.ROOT, "oracle", "liboracle.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(oracle_lib.ROOT, "oracle"), "liboracle.so"])
    monkeypatch.setenv("GMR1_ORACLE_SYNC_RESET", "1")
    return oracle_lib.Oracle(so, "port")
