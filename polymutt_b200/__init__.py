"""polymutt_b200 — B200-native per-site family-likelihood engine for polymutt's calling path.

The product is the C-ABI shared library ``polymutt_b200/lib/libpolymutt_b200.so`` (hand-written
sm_100a CUDA kernels behind ``include/polymutt_b200.h``) and the drop-in executable
``polymutt_b200/bin/polymutt-b200``.  This package is only a thin ctypes binding used by the tests and
``bench.py``; it has no CPU implementation of the likelihood path and raises if the library is
missing.
"""
from .capi import (  # noqa: F401
    Engine,
    LibraryNotBuilt,
    PedigreeArrays,
    Params,
    load_library,
    lib_path,
    SITE_RESULT_DTYPE,
    PERSON_RESULT_DTYPE,
    SITE_HDR_DTYPE,
    PERSON_SITE_DTYPE,
)
from .pmpk import load_pmpk, Pmpk  # noqa: F401

__all__ = [
    "Engine", "LibraryNotBuilt", "PedigreeArrays", "Params", "load_library", "lib_path",
    "SITE_RESULT_DTYPE", "PERSON_RESULT_DTYPE", "SITE_HDR_DTYPE", "PERSON_SITE_DTYPE", "load_pmpk", "Pmpk",
]
