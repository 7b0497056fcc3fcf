"""osmo_gmr_b200 - thin Python handle on libgmr1_b200.so (the B200-native GMR-1 receive path).

The product is the CUDA/C++ shared library and its C ABI (include/gmr1_b200.h); this package
only loads it with ctypes so that tests and bench.py can call the same entry points a C
caller (e.g. the reference's gmr1_rx.c) would.  There is no Python or CPU compute path here:
if the library is missing, import fails loudly.
"""
from .lib import Lib, lib, LIB_PATH, build  # noqa: F401
