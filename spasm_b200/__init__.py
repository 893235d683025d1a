"""spasm-b200: SpaSM's echelonization hot path, rebuilt for B200 (sm_100a) behind SpaSM's own C ABI.

The product is the shared library `spasm_b200/lib/libspasm_b200.so` (host C + CUDA, built in-tree by
`python -m spasm_b200.build`).  This Python package is only the host-side mirror used by the tests
and the benchmark: it binds the C ABI with ctypes and offers the reference's operator names.

There is no CPU fallback: if the shared library is missing `lib()` raises, and inside the library
every GPU entry point aborts with errx(1, ...) when no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os

from . import abi, host, synthetic  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libspasm_b200.so")

_lib = None

# extra C-ABI entry points of include/spasm_b200.h (instrumentation for tests / bench)
EXTRA_PROTOTYPES = {
    "spasm_b200_device_count": (C.c_int, []),
    "spasm_b200_set_device": (None, [C.c_int]),
    "spasm_b200_reset_stats": (None, []),
    "spasm_b200_get_stats": (None, [C.c_void_p]),
    "spasm_b200_version": (C.c_char_p, []),
}


class MissingExtension(RuntimeError):
    pass


def lib() -> C.CDLL:
    """The product library.  Raises MissingExtension (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MissingExtension(
                f"{LIB_PATH} is missing: build it with `python -m spasm_b200.build` "
                "(there is deliberately no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        abi.bind(L)
        for name, (res, args) in EXTRA_PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib
