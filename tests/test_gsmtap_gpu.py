"""GPU parity: gmr1b200_gsmtap_batch vs the reference's gmr1_gsmtap_makemsg (src/gsmtap.c:44-71) - the GSMTAP
records of a batch of decoded units, byte for byte (integer work: exact), host and device pointers, explicit and
default channel type / frame number / timeslot, padded record stride."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,length,stride_pad", [(1, 24, 0), (257, 24, 0), (1000, 10, 6), (333, 38, 3), (64, 0, 0)])
def test_gsmtap_records(gpu_lib, oracle, n, length, stride_pad):
    rng = np.random.default_rng(n + length)
    ct = rng.integers(0, 32, n, dtype=np.uint8)
    fn = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    tn = rng.integers(0, 24, n, dtype=np.uint8)
    l2_stride = length + 5
    l2 = rng.integers(0, 256, (n, l2_stride), dtype=np.uint8)
    ostride = 16 + length + stride_pad
    out = np.full((n, ostride), 0xee, np.uint8)
    gpu_lib.call("gmr1b200_gsmtap_batch", ct, 0, fn, 0, tn, 0, l2, l2_stride, length, out, ostride, n, None)
    for i in range(n):
        want = oracle.gsmtap(int(ct[i]), int(fn[i]), int(tn[i]), l2[i, :length])
        assert len(want) == 16 + length and (out[i, :16 + length] == want).all(), i

    # defaults: one channel type / timeslot, consecutive frame numbers
    out2 = np.zeros((n, ostride), np.uint8)
    gpu_lib.call("gmr1b200_gsmtap_batch", None, 7, None, 0xfffffff0, None, 3, l2, l2_stride, length, out2, ostride, n, None)
    for i in (0, n // 2, n - 1):
        want = oracle.gsmtap(7, (0xfffffff0 + i) & 0xffffffff, 3, l2[i, :length])
        assert (out2[i, :16 + length] == want).all()


def test_gsmtap_device_pointers_after_decode(gpu_lib, oracle):
    """decode -> GSMTAP records without the L2 leaving the device in between"""
    import torch
    from decode_backends import CH
    rng = np.random.default_rng(3)
    n = 512
    l2 = rng.integers(0, 256, (n, 24), dtype=np.uint8)
    eb = np.stack([oracle.encode("bcch", 424, l2[i]) for i in range(n)])
    soft = np.where(eb == 0, 100, -100).astype(np.int8)
    dev = torch.device("cuda:0")
    d_soft = torch.from_numpy(soft).to(dev)
    d_l2 = torch.zeros((n, 24), dtype=torch.uint8, device=dev)
    d_crc = torch.zeros(n, dtype=torch.int32, device=dev)
    d_out = torch.zeros((n, 40), dtype=torch.uint8, device=dev)
    gpu_lib.call("gmr1b200_bcch_decode_batch", d_l2, d_soft, None, d_crc, n, None)
    gpu_lib.call("gmr1b200_gsmtap_batch", None, 1, None, 100, None, 0, d_l2, 24, 24, d_out, 40, n, None)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    assert (d_crc.cpu().numpy() == 0).all()
    for i in range(0, n, 37):
        assert (out[i] == oracle.gsmtap(1, 100 + i, 0, l2[i])).all()
