"""iteres_b200 -- B200-native (sm_100a) implementation of the iteres hot path.

The product is the C-ABI library iteres_b200/csrc/libiteres_gpu.so (see include/iteres_gpu.h) and
the `iteres` command line tool built next to it; this package is only the ctypes binding that tests
and bench.py use.  There is no CPU fallback: importing works anywhere, but every scan needs the
CUDA library and a device and raises otherwise."""
from .capi import Index, ScanOpts, Trace, Profile, lib, lib_path, default_opts, ItxError  # noqa: F401
