"""bedops_b200 -- B200-native sorted-interval engine behind the BEDOPS command lines.

The product is `lib/libbedkit.so` (hand-written sm_100a CUDA behind the C ABI in include/bedkit.h) and the three
drop-in command-line tools in `bin/`.  This package is the thin ctypes binding used by tests and bench.py; it has
no CPU implementation of anything and raises if the CUDA library is missing.
"""
from ._lib import BedKit, BedKitError, lib_path, tool_path, load_library  # noqa: F401
