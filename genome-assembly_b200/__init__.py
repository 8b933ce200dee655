"""gbin — B200-native k-mer binning (process_read -> two-level table -> prune) behind a C ABI.

Python here is a thin ctypes mirror of ``include/gbin.h`` for tests, the benchmark and the
multi-GPU launcher; the product is ``libgbin.so`` (C host code + hand-written sm_100a kernels).
"""
from . import binding, synth  # noqa: F401,E402
from .binding import Binner, GbinError, HostTable, load_library, read_file_fgets  # noqa: F401,E402
