"""Error conventions at the drop-in boundary: the reference-named symbols of libgmr1_b200.so must
return what the reference's functions return for the inputs the reference itself rejects or flags - hard errors
as negative errno, soft conditions as positive values - and the batched entry points must reject malformed
batches with -EINVAL and a message, doing no work.  The compat handle is the oracle's python wrapper pointed at
libgmr1_b200.so: the very same calls go to both libraries."""
import errno

import numpy as np
import pytest

import oracle_lib
import sigen
from osmo_gmr_b200.lib import Gmr1Error, LIB_PATH

pytestmark = pytest.mark.gpu
SPS = 4


@pytest.fixture(scope="module")
def compat(gpu_lib):
    return oracle_lib.Oracle(LIB_PATH, "b200")


def test_reference_named_symbols_follow_the_reference(compat, oracle):
    rng = np.random.default_rng(7)
    noise = lambda n: (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    # gmr1_fcch_fine insists on exactly burst_len*sps samples (fcch.c:546-551)
    for n in (117 * SPS - 4, 117 * SPS + 4):
        x = noise(n)
        assert compat.fcch_fine(x, SPS, 0.0)[0] == oracle.fcch_fine(x, SPS, 0.0)[0] == -errno.EINVAL
    # gmr1_fcch_rough_multi needs 650 ms of signal (fcch.c:355-356)
    x = noise((600 * 23400 * SPS) // 1000)
    assert compat.fcch_rough_multi(x, SPS, 0.0)[0] == oracle.fcch_rough_multi(x, SPS, 0.0)[0] == -errno.EINVAL
    # gmr1_dkab_demod: positive return for "not a DKAB" (dkab.c:138), 0 for a DKAB
    w = noise(117 * SPS + 6)
    r_ref, r_gpu = oracle.dkab_demod(w, SPS, 0.0, 5)[0], compat.dkab_demod(w, SPS, 0.0, 5)[0]
    assert r_ref == r_gpu and r_ref >= 0
    # all-zero window: nothing correlates; both report the same (negative) code and leave no soft bits
    z = np.zeros(234 * SPS + 80, np.complex64)
    rc_ref, eb_ref, _, _, _ = oracle.demod("bcch", z, SPS, 0.0)
    rc_gpu, eb_gpu, _, _, _ = compat.demod("bcch", z, SPS, 0.0)
    assert rc_ref == rc_gpu and rc_ref < 0 and not eb_gpu.any()
    # a good burst through the same wrapper: success is 0 on both sides
    hard = oracle.encode("bcch", 424, rng.integers(0, 256, 24, dtype=np.uint8))
    x = sigen.modulate("bcch", hard[None, :], SPS, 80, 40.3, 0.004, 1.0, 20.0, rng)[0]
    assert oracle.demod("bcch", x, SPS, 0.0)[0] == compat.demod("bcch", x, SPS, 0.0)[0] == 0


def test_batched_entry_points_reject_malformed_batches(gpu_lib):
    L = gpu_lib
    n, wl = 4, 234 * SPS + 80
    iq = np.zeros((n, wl, 2), np.float32)
    eb = np.zeros((n, 424), np.int8)
    l2 = np.zeros((n, 24), np.uint8)
    before = L.kernel_launches()

    def rejected(name, *args):
        with pytest.raises(Gmr1Error) as e:
            L.call(name, *args)
        assert f"rc={-errno.EINVAL}" in str(e.value) and len(str(e.value)) > 20      # code and a message

    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, wl, 0, None, 0.0, eb, 424, None, None, None, None, n, None)   # sps 0
    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, wl, 17, None, 0.0, eb, 424, None, None, None, None, n, None)  # sps 17
    rejected("gmr1b200_pi4cxpsk_demod_batch", 99, iq, n * wl, None, wl, wl, SPS, None, 0.0, eb, 424, None, None, None, None, n, None)  # burst type
    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, wl, SPS, None, 0.0, eb, 100, None, None, None, None, n, None)  # ebits rows too short
    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, 234 * SPS - 1, SPS, None, 0.0, eb, 424, None, None, None, None, n, None)  # window < burst
    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl - 1, None, wl, wl, SPS, None, 0.0, eb, 424, None, None, None, None, n, None)  # windows exceed iq
    rejected("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, wl, SPS, None, 0.0, None, 424, None, None, None, None, n, None)  # no output
    rejected("gmr1b200_bcch_decode_batch", l2, None, None, None, n, None)                                                             # no input
    rejected("gmr1b200_bcch_decode_batch", l2, eb, None, None, -1, None)                                                              # negative n
    rejected("gmr1b200_fcch_rough_batch", 0, iq, n * wl, None, wl, 100, SPS, None, 0.0, np.zeros(n, np.int32), None, n, None)         # window < FCCH burst
    rejected("gmr1b200_fcch_rough_batch", 7, iq, n * wl, None, wl, wl, SPS, None, 0.0, np.zeros(n, np.int32), None, n, None)          # FCCH type
    rejected("gmr1b200_gsmtap_batch", None, 1, None, 0, None, 0, l2, 24, 24, np.zeros((n, 39), np.uint8), 39, n, None)   # record stride < 16 + len
    rejected("gmr1b200_gsmtap_batch", None, 1, None, 0, None, 0, l2, 20, 24, np.zeros((n, 40), np.uint8), 40, n, None)   # l2 stride < len
    rejected("gmr1b200_a5_batch", None, 1, np.zeros((n, 8), np.uint8), np.zeros(n, np.uint32), 208, 100, np.zeros((n, 208), np.uint8), None, n, None)  # stride < nbits
    rejected("gmr1b200_rx_bcch_batch", iq, n * wl, np.zeros(n, np.int64), np.full(n, wl, np.int32), np.zeros(n, np.int32), None, SPS, n, 0,
             np.zeros((n, 1), np.int32), np.zeros((n, 1), np.int32), np.zeros((n, 1), np.int32), np.zeros((n, 1), np.int32),
             np.zeros((n, 1, 24), np.uint8), np.zeros(n, np.int32), None, None, None)                                                 # max_frames 0
    i32 = lambda *shape: np.zeros(shape, np.int32)
    walk_out = (i32(n, 1), i32(n, 1), i32(n, 1), i32(n, 1), np.zeros((n, 1, 24), np.uint8), i32(n), None, None)
    rejected("gmr1b200_rx_bcch_ass_batch", iq, n * wl, np.zeros(n, np.int64), np.full(n, wl, np.int32), i32(n), None, SPS, n, 1,
             *walk_out, None, None, None)                                                                                             # no hand-off record
    multi = (np.zeros(n, np.int64), np.full(n, wl, np.int32), i32(n), None)
    rejected("gmr1b200_fcch_multi_batch", 0, iq, n * wl, *multi, SPS, n, 17, i32(n), i32(n, 17), None, None, None)                    # more than 16 candidates
    rejected("gmr1b200_fcch_multi_batch", 3, iq, n * wl, *multi, SPS, n, 4, i32(n), i32(n, 4), None, None, None)                      # FCCH type
    rejected("gmr1b200_fcch_multi_batch", 0, iq, n * wl, *multi, SPS, n, 4, None, i32(n, 4), None, None, None)                        # no count output
    assert L.kernel_launches() == before                                             # nothing was launched
    # recordings shorter than the 650 ms window are flagged one by one (gmr1_rx.c:661-665)
    cnt = np.full(n, 5, np.int32)
    assert L.call("gmr1b200_fcch_multi_batch", 0, iq, n * wl, *multi, SPS, n, 4, cnt, i32(n, 4), None, None, None) == 0
    assert (cnt == -errno.EINVAL).all()              # decided on the device: the rows are rejected by the prep kernel
    before = L.kernel_launches()
    assert L.call("gmr1b200_fcch_multi_batch", 0, iq, n * wl, *multi, SPS, 0, 4, cnt, i32(n, 4), None, None, None) == 0
    # empty batches are fine and do nothing
    assert L.call("gmr1b200_bcch_decode_batch", l2, eb, None, None, 0, None) == 0
    assert L.call("gmr1b200_pi4cxpsk_demod_batch", 0, iq, n * wl, None, wl, wl, SPS, None, 0.0, eb, 424, None, None, None, None, 0, None) == 0
    assert L.kernel_launches() == before


def test_caller_descriptors_must_be_consistent(gpu_lib):
    """gmr1b200_pi4cxpsk_demod_desc_batch: a descriptor whose ebits disagrees with its data symbols, whose data chunks
    overlap or whose sync symbols are not phase indices is rejected before anything is launched; host-readable window
    offsets are range-checked (the reference has no such path: its descriptors are compiled in, sdr/nb.c)."""
    import ctypes
    L = gpu_lib

    class Desc(ctypes.Structure):            # include/gmr1_b200.h: struct gmr1b200_burst_desc
        _fields_ = [("rotation", ctypes.c_float), ("nbits", ctypes.c_int32), ("len", ctypes.c_int32),
                    ("ebits", ctypes.c_int32), ("n_sync", ctypes.c_int32), ("n_chunk", ctypes.c_int32 * 4),
                    ("s_pos", ctypes.c_int16 * 6 * 4), ("s_len", ctypes.c_int16 * 6 * 4),
                    ("s_sym", ctypes.c_uint8 * 32 * 6 * 4), ("n_data", ctypes.c_int32),
                    ("d_pos", ctypes.c_int16 * 6), ("d_len", ctypes.c_int16 * 6)]

    L.c.gmr1b200_burst_desc_get.argtypes = [ctypes.c_int, ctypes.c_void_p]
    good = Desc()
    assert L.c.gmr1b200_burst_desc_get(0, ctypes.addressof(good)) == 0 and good.ebits == 424
    n, wl = 2, 234 * SPS + 80
    iq = np.zeros((n, wl, 2), np.float32)
    eb = np.zeros((n, 424), np.int8)

    def run(d, ofs=None):
        return L.c.gmr1b200_pi4cxpsk_demod_desc_batch(ctypes.addressof(d), iq.ctypes.data, n * wl,
                                                      None if ofs is None else ofs.ctypes.data, wl, wl, SPS, None, 0.0,
                                                      eb.ctypes.data, 424, None, None, None, None, n, None)

    def variant(f):
        d = Desc.from_buffer_copy(bytes(good))
        f(d)
        return d

    assert run(good) == 0
    before = L.kernel_launches()
    bad = [variant(lambda d: setattr(d, "ebits", 400)),                     # fewer soft bits than data symbols
           variant(lambda d: d.d_len.__setitem__(0, d.d_len[0] - 1)),       # data symbols no longer add up to ebits
           variant(lambda d: d.s_sym[0][0].__setitem__(0, 7)),              # not a phase index
           variant(lambda d: d.d_pos.__setitem__(1, d.d_pos[0]))]           # second data chunk on top of the first
    for d in bad:
        assert run(d) == -errno.EINVAL and b"descriptor" in L.c.gmr1b200_last_error()
    assert run(good, np.array([0, wl + 1], np.int64)) == -errno.EINVAL      # second window runs past iq_len
    assert run(good, np.array([-1, wl], np.int64)) == -errno.EINVAL
    assert L.kernel_launches() == before
    assert run(good, np.array([wl, 0], np.int64)) == 0


def test_round2_entry_points_reject_malformed_calls(gpu_lib):
    """the entry points added around the path (channeliser, vocoder stream, FCCH grid, kernel switches): -EINVAL with
    a message for malformed calls, empty batches are no-ops, switches report the previous setting"""
    import ctypes
    L = gpu_lib
    h = ctypes.c_void_p()
    for n_chans, sps in ((15, 4), (0, 4), (4098, 4), (2 * 37, 4), (16, 0), (16, 17)):      # odd, empty, too many, prime > 31, sps
        assert L.c.gmr1b200_chan_create(n_chans, sps, ctypes.byref(h)) == -errno.EINVAL
        assert b"chan_create" in L.c.gmr1b200_last_error()
    assert L.c.gmr1b200_chan_create(16, 4, None) == -errno.EINVAL
    assert L.c.gmr1b200_chan_create(16, 4, ctypes.byref(h)) == 0
    assert L.c.gmr1b200_chan_out_len(h, -1) == -errno.EINVAL and L.c.gmr1b200_chan_out_len(h, 7) == 0
    x = np.zeros((100, 2), np.int16)
    out = np.zeros((1, 64, 2), np.float32)
    L.call("gmr1b200_channelize", h.value, x, 1, 7, None, 1, out, 64, None)          # shorter than one bank step: no output
    L.call("gmr1b200_channelize", h.value, x, 1, 100, None, 0, out, 64, None)        # no stream wanted
    assert not out.any()
    for args in ((None, x, 1, 100, None, 1, out, 64, None), (h.value, None, 1, 100, None, 1, out, 64, None),
                 (h.value, x, 1, 100, None, 17, out, 64, None), (h.value, x, 1, -1, None, 1, out, 64, None)):
        with pytest.raises(Gmr1Error) as e:
            L.call("gmr1b200_channelize", *args)
        assert f"rc={-errno.EINVAL}" in str(e.value)
    L.c.gmr1b200_chan_destroy(h)
    L.c.gmr1b200_chan_destroy(None)
    # vocoder stream
    F = 4
    rec, dat = np.zeros((2, F, 12), np.int32), np.zeros((2, F, 20), np.uint8)
    fn, nfr = np.zeros((2, F), np.int32), np.zeros(2, np.int32)
    voice, flag, nv = np.zeros((2, 2 * F, 10), np.uint8), np.zeros((2, 2 * F), np.uint8), np.full(2, -1, np.int32)
    L.call("gmr1b200_tch3_voice_stream_batch", rec, dat, fn, nfr, 0, F, voice, flag, nv, None, None)       # n = 0
    assert (nv == -1).all()
    L.call("gmr1b200_tch3_voice_stream_batch", rec, dat, fn, nfr, 2, F, voice, flag, nv, None, None)       # no call at all
    assert (nv == 0).all() and (flag == 2).all()
    for args in ((None, dat, fn, nfr, 2, F, voice, flag, nv, None, None), (rec, dat, fn, nfr, 2, 0, voice, flag, nv, None, None),
                 (rec, dat, fn, nfr, -1, F, voice, flag, nv, None, None), (rec, dat, fn, nfr, 2, F, voice, None, nv, None, None)):
        with pytest.raises(Gmr1Error) as e:
            L.call("gmr1b200_tch3_voice_stream_batch", *args)
        assert f"rc={-errno.EINVAL}" in str(e.value)
    # FCCH grid: shift count, missing outputs, window shorter than the burst
    W = 117 * SPS + 40
    iq = np.zeros((1, W, 2), np.float32)
    toa = np.zeros((17, 1), np.int32)
    g = np.zeros(17, np.float32)
    for args in ((0, iq, W, None, W, W, SPS, g, 17, toa, None, 1, None), (0, iq, W, None, W, W, SPS, g, 0, toa, None, 1, None),
                 (0, iq, W, None, W, W, SPS, None, 2, toa, None, 1, None), (0, iq, W, None, W, W, SPS, g, 2, None, None, 1, None),
                 (0, iq, W, None, W, 100, SPS, g, 2, toa, None, 1, None)):
        with pytest.raises(Gmr1Error) as e:
            L.call("gmr1b200_fcch_rough_grid_batch", *args)
        assert f"rc={-errno.EINVAL}" in str(e.value)
    # switches: return the previous setting
    for name, default in (("gmr1b200_set_fcch_fft", 1), ("gmr1b200_set_chan_generic", 0), ("gmr1b200_set_demod_generic", 0)):
        fn_ = getattr(L.c, name)
        assert fn_(1 - default) == default and fn_(default) == 1 - default and fn_(default) == default
