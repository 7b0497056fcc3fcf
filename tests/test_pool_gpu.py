"""The device pool (include/gmr1_b200.h: gmr1b200_pool_*, csrc/api_multi.cu): host IQ of many ARFCNs sharded
`arfcn mod G` over the pool's devices, chunked through per-device streams, results gathered into the caller's host
arrays - must return exactly what the single-device batched entry points return on the same windows.  A pool may
name the same CUDA device several times, so the sharding / strided-copy / gather logic with G = 2 and 3 is covered
on a one-GPU box; with more GPUs present the same test also runs over distinct devices."""
import numpy as np
import pytest

import sigen

pytestmark = pytest.mark.gpu


def _bursts(oracle, n_arfcn, per, rng):
    n = n_arfcn * per
    l2 = rng.integers(0, 256, (n, 24), dtype=np.uint8)
    hard = np.stack([oracle.encode("bcch", 424, l2[i]) for i in range(n)])
    x = sigen.modulate("bcch", hard, 4, 80, rng.uniform(2, 78, n), rng.uniform(-0.01, 0.01, n),
                       rng.uniform(0, 6.28, n), np.full(n, 12.0), rng)
    return l2, np.ascontiguousarray(x)


@pytest.mark.parametrize("members", [1, 2, 3, "all"])
def test_pool_rx_xcch_and_fcch(gpu_lib, oracle, members):
    import torch
    L = gpu_lib
    rng = np.random.default_rng(42)
    n_arfcn, per = 11, 6                     # ragged on purpose: 11 ARFCNs over 2 / 3 members, several chunks
    l2_true, x = _bursts(oracle, n_arfcn, per, rng)
    n, wl = x.shape
    iq = x.view(np.float32).reshape(n, 2 * wl)
    ref_l2, ref_crc = np.zeros((n, 24), np.uint8), np.zeros(n, np.int32)
    ref_conv, ref_toa = np.zeros(n, np.int32), np.zeros(n, np.float32)
    L.call("gmr1b200_rx_xcch_batch", 0, iq, n * wl, None, wl, wl, 4, None, 0.0, ref_l2, ref_crc, ref_conv, ref_toa,
           None, n, None)
    assert (ref_crc == 0).mean() > 0.9 and (ref_l2[ref_crc == 0] == l2_true[ref_crc == 0]).all()
    if members == "all":
        devs = list(range(torch.cuda.device_count()))
    else:
        devs = [i % torch.cuda.device_count() for i in range(members)]
    # chunk_bytes so small that every member needs several chunks (2 ARFCNs per chunk)
    cur = torch.cuda.current_device()
    pool = L.pool_create(devs, streams_per_dev=2, chunk_bytes=max(1 << 20, 2 * per * wl * 8 + 64))
    assert torch.cuda.current_device() == cur        # the pool visits every device but leaves the caller's current one
    try:
        assert L.call("gmr1b200_pool_size", pool) == len(devs)
        pin = torch.from_numpy(iq).pin_memory()
        l2 = torch.zeros((n, 24), dtype=torch.uint8).pin_memory()
        crc = torch.full((n,), -7, dtype=torch.int32).pin_memory()
        conv = torch.zeros(n, dtype=torch.int32).pin_memory()
        toa = torch.zeros(n, dtype=torch.float32).pin_memory()
        L.call("gmr1b200_pool_rx_xcch", pool, 0, pin, n_arfcn, per, wl, 4, 0.0, l2, crc, conv, toa)
        assert (l2.numpy() == ref_l2).all() and (crc.numpy() == ref_crc).all()
        assert (conv.numpy() == ref_conv).all() and (toa.numpy() == ref_toa).all()
        # pageable host memory works too (staged copies)
        l2b, crcb = np.zeros((n, 24), np.uint8), np.full(n, -7, np.int32)
        L.call("gmr1b200_pool_rx_xcch", pool, 0, iq, n_arfcn, per, wl, 4, 0.0, l2b, crcb, None, None)
        assert (l2b == ref_l2).all() and (crcb == ref_crc).all()
        # FCCH acquisition of 7 search windows
        W = (330 * 23400 * 4) // 1000
        fw = np.stack([sigen.fcch_window(W, 4, 4000 + 913 * k, 0.05 * (k - 3), 10.0, rng) for k in range(7)])
        fiq = np.ascontiguousarray(fw).view(np.float32).reshape(7, 2 * W)
        r0, a0, f0 = np.zeros(7, np.int32), np.zeros(7, np.int32), np.zeros(7, np.float32)
        L.call("gmr1b200_fcch_acquire_batch", 0, fiq, 7 * W, None, W, W, 4, r0, a0, f0, 7, None)
        r1, a1, f1 = np.zeros(7, np.int32), np.zeros(7, np.int32), np.zeros(7, np.float32)
        L.call("gmr1b200_pool_fcch_acquire", pool, 0, fiq, 7, W, 4, r1, a1, f1)
        assert (r0 == r1).all() and (a0 == a1).all() and (f0 == f1).all()
        assert np.abs(a1 - (4000 + 913 * np.arange(7))).max() <= 2
    finally:
        L.pool_destroy(pool)
    assert torch.cuda.current_device() == cur


def test_pool_errors(gpu_lib):
    L = gpu_lib
    with pytest.raises(RuntimeError, match="no such device"):
        L.pool_create([99])
    with pytest.raises(RuntimeError, match="bad argument"):
        L.pool_create([0], streams_per_dev=0)
