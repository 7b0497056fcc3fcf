"""ctypes binding of libgmr1_b200.so (see include/gmr1_b200.h for the contract)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMR1B200_LIB") or os.path.join(_HERE, "libgmr1_b200.so")   # env: A/B builds only

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float


def build(verbose=False):
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", _HERE, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("building libgmr1_b200.so failed")
    return LIB_PATH


def _ptr(x):
    """numpy array / torch tensor / int / None -> address (host or device, the library detects)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):          # torch tensor
        assert x.is_contiguous()
        return x.data_ptr()
    if hasattr(x, "ctypes"):            # numpy array
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    raise TypeError(type(x))


class Gmr1Error(RuntimeError):
    pass


class Lib:
    """One loaded libgmr1_b200.so.  Methods mirror the C entry points one to one."""

    # name -> argtypes (restype is int unless listed in _RESTYPE)
    _SIG = {
        "gmr1b200_init": [_I],
        "gmr1b200_bcch_decode_batch": [_P, _P, _P, _P, _I, _P],
        "gmr1b200_ccch_decode_batch": [_P, _P, _P, _P, _I, _P],
        "gmr1b200_facch3_decode_batch": [_P, _P, _P, _P, _P, _P, _I, _P],
        "gmr1b200_facch9_decode_batch": [_P, _P, _P, _P, _P, _P, _P, _I, _P],
        "gmr1b200_tch3_decode_batch": [_P, _P, _P, _P, _P, _I, _P, _P, _I, _P],
        "gmr1b200_tch9_decode_batch": [_P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _P],
        "gmr1b200_tch9_decode_rows_batch": [_P, _P, _I, _P, _I, _P],
        "gmr1b200_pi4cxpsk_demod_desc_batch": [_P, _P, _L, _P, _L, _I, _I, _P, _F, _P, _I, _P, _P, _P, _P, _I, _P],
        "gmr1b200_fcch_rough_multi": [_I, _P, _L, _I, _F, _P, _I, _P],
        "gmr1b200_rach_decode_batch": [_P, _P, _P, _I, _P, _P, _P, _I, _P],
        "gmr1b200_xch_dc12_decode_batch": [_P, _P, _P, _P, _I, _P],
        "gmr1b200_xcch_encode_batch": [_I, _P, _P, _I],
        "gmr1b200_facch3_encode": [_P, _P, _P, _P],
        "gmr1b200_facch9_encode": [_P, _P, _P, _P, _P],
        "gmr1b200_tch9_encode": [_P, _P, _I, _P, _P, _P, _P],
        "gmr1b200_rach_encode": [_P, _P, _I],
        "gmr1b200_tch3_encode": [_P, _P, _P, _P, _P, _I],
        "gmr1b200_fcch_multi_batch": [_I, _P, _L, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P],
        "gmr1b200_fcch_rough_batch": [_I, _P, _L, _P, _L, _I, _I, _P, _F, _P, _P, _I, _P],
        "gmr1b200_fcch_fine_batch": [_I, _P, _L, _P, _L, _I, _P, _F, _P, _P, _I, _P],
        "gmr1b200_rx_xcch_batch": [_I, _P, _L, _P, _L, _I, _I, _P, _F, _P, _P, _P, _P, _P, _I, _P],
        "gmr1b200_rx_bcch_batch": [_P, _L, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
        "gmr1b200_rx_bcch_ass_batch": [_P, _L, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
        "gmr1b200_a5_batch": [_P, _I, _P, _P, _I, _I, _P, _P, _I, _P],
        "gmr1b200_gsmtap_batch": [_P, _I, _P, ctypes.c_uint32, _P, _I, _P, _I, _I, _P, _I, _I, _P],
        "gmr1b200_fcch_rough_grid_batch": [_I, _P, _L, _P, _L, _I, _I, _P, _I, _P, _P, _I, _P],
        "gmr1b200_fcch_acquire_batch": [_I, _P, _L, _P, _L, _I, _I, _P, _P, _P, _I, _P],
        "gmr1b200_fcch_snr_batch": [_I, _P, _L, _P, _L, _I, _P, _F, _P, _I, _P],
        "gmr1b200_dkab_demod_batch": [_P, _L, _P, _L, _I, _I, _P, _F, _P, _I, _P, _P, _P, _I, _P],
        "gmr1b200_pi4cxpsk_mod_order_batch": [_P, _L, _P, _L, _I, _I, _P, _F, _P, _I, _P],
        "gmr1b200_synth_bursts": [_I, _P, _I, _P, _I, _I, _P, _F, _P, _F, _P, _F, _P, _F, _P, _F, ctypes.c_uint64,
                                  _P, _L, _P, _L, _I, _P],
        "gmr1b200_synth_bursts_tx": [_I, _P, _I, _P, _I, _I, _P, _F, _P, _F, _P, _F, _P, _F, _P, _F, ctypes.c_uint64,
                                     _P, _L, _P, _L, _I, _P],
        "gmr1b200_chan_create": [_I, _I, _P],
        "gmr1b200_chan_info": [_P, _P],
        "gmr1b200_chan_taps": [_P, _P, _I, _P, _I],
        "gmr1b200_channelize": [_P, _P, _I, _L, _P, _I, _P, _L, _P],
        "gmr1b200_chan_stream_create": [_P, _P, _I, _P],
        "gmr1b200_chan_stream_push": [_P, _P, _I, _L, _P, _L, _P, _P],
        "gmr1b200_synth_wideband": [_P, _P, _L, _L, _P, _I, _F, _F, ctypes.c_uint64, _P, _I, _L, _P],
        "gmr1b200_set_sync_accumulator_reset": [_I],
        "gmr1b200_set_demod_generic": [_I],
        "gmr1b200_set_chan_generic": [_I],
        "gmr1b200_set_fcch_fft": [_I],
        "gmr1b200_set_a5_bitslice": [_I],
        "gmr1b200_set_rx_lockstep": [_I],
        "gmr1b200_rx_call_batch": [_P, _L, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P],
        "gmr1b200_tch3_voice_stream_batch": [_P, _P, _P, _P, _I, _I, _P, _P, _P, _P, _P],
        "gmr1b200_pool_create": [_P, _I, _I, _L, _P],
        "gmr1b200_pool_size": [_P],
        "gmr1b200_pool_rx_xcch": [_P, _I, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P],
        "gmr1b200_pool_fcch_acquire": [_P, _I, _P, _I, _I, _I, _P, _P, _P],
        "gmr1b200_burst_len": [_I],
        "gmr1b200_burst_ebits": [_I],
        "gmr1b200_pi4cxpsk_demod_batch": [_I, _P, _L, _P, _L, _I, _I, _P, _F, _P, _I, _P, _P, _P, _P, _I, _P],
        "gmr1b200_pi4cxpsk_detect_batch": [_P, _I, _P, _F, _P, _L, _P, _L, _I, _I, _P, _F, _P, _P, _P, _I, _P],
    }

    def __init__(self, path=LIB_PATH):
        if not os.path.exists(path):
            raise ImportError(
                f"{path} not found: build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no CPU fallback.")
        self.path = path
        self.c = ctypes.CDLL(path)
        for name, args in self._SIG.items():
            fn = getattr(self.c, name)
            fn.argtypes = args
            fn.restype = _I
        self.c.gmr1b200_tch9_interleaver_new.restype = _P
        self.c.gmr1b200_tch9_interleaver_free.argtypes = [_P]
        self.c.gmr1b200_tch9_interleaver_free.restype = None
        self.c.gmr1b200_last_error.restype = ctypes.c_char_p
        self.c.gmr1b200_version.restype = ctypes.c_char_p
        self.c.gmr1b200_kernel_launches.restype = ctypes.c_uint64
        self.c.gmr1b200_pool_destroy.argtypes = [_P]
        self.c.gmr1b200_pool_destroy.restype = None
        self.c.gmr1b200_host_alloc.argtypes = [ctypes.c_size_t]
        self.c.gmr1b200_host_alloc.restype = _P
        self.c.gmr1b200_host_free.argtypes = [_P]
        self.c.gmr1b200_host_free.restype = None
        self.c.gmr1b200_chan_destroy.argtypes = [_P]
        self.c.gmr1b200_chan_destroy.restype = None
        self.c.gmr1b200_chan_out_len.argtypes = [_P, _L]
        self.c.gmr1b200_chan_out_len.restype = _L
        self.c.gmr1b200_chan_stream_destroy.argtypes = [_P]
        self.c.gmr1b200_chan_stream_destroy.restype = None
        self.c.gmr1b200_chan_stream_max_out.argtypes = [_P, _L]
        self.c.gmr1b200_chan_stream_max_out.restype = _L

    # -- helpers
    def _chk(self, rc, what):
        if rc < 0:
            raise Gmr1Error(f"{what}: rc={rc} ({self.c.gmr1b200_last_error().decode()})")
        return rc

    def version(self):
        return self.c.gmr1b200_version().decode()

    def kernel_launches(self):
        return int(self.c.gmr1b200_kernel_launches())

    def init(self, device=0):
        return self._chk(self.c.gmr1b200_init(device), "init")

    def pool_create(self, devices, streams_per_dev=3, chunk_bytes=64 << 20):
        """gmr1b200_pool_create -> opaque handle (int); devices: list of CUDA ordinals"""
        arr = (ctypes.c_int * len(devices))(*devices)
        out = ctypes.c_void_p()
        self._chk(self.c.gmr1b200_pool_create(ctypes.cast(arr, _P), len(devices), streams_per_dev, chunk_bytes,
                                              ctypes.cast(ctypes.byref(out), _P)), "pool_create")
        return out.value

    def pool_destroy(self, pool):
        self.c.gmr1b200_pool_destroy(pool)

    def call(self, name, *args):
        """Raw call by C name with pointer-like python objects converted."""
        fn = getattr(self.c, name)
        conv = [(_ptr(a) if t is _P else a) for a, t in zip(args, fn.argtypes)]
        return self._chk(fn(*conv), name)


_lib = None


def lib():
    """Process-wide library handle (loaded on first use)."""
    global _lib
    if _lib is None:
        _lib = Lib()
    return _lib
