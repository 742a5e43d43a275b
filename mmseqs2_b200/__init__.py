"""mmseqs2_b200 -- B200-native implementation of MMseqs2's alignment hot path behind a C ABI.

Python here is plumbing only (ctypes over mmseqs2_b200/libb200align.so, used by tests and bench.py);
the product is the shared library declared in include/b200_align.h + include/b200_host.h.
There is no CPU fallback: importing the device API without the built library raises.
"""
from .api import (  # noqa: F401
    B200Error,
    Context,
    MultiContext,
    QueryProfile,
    SubMatrix,
    lib_path,
    load_library,
)
