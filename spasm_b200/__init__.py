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
    "spasm_b200_set_verbose": (None, [C.c_int]),
    "spasm_b200_trim": (None, []),
    "spasm_b200_upload_csr": (C.c_void_p, [abi.CsrP]),
    "spasm_b200_free_csr": (None, [C.c_void_p]),
    "spasm_b200_echelonize_resident": (C.c_int, [C.c_void_p, abi.OptsP, C.POINTER(C.c_double)]),
    "spasm_b200_flush_l2": (None, []),
    "spasm_b200_comm_unique_id": (None, [C.c_void_p]),
    "spasm_b200_comm_init": (None, [C.c_int, C.c_int, C.c_void_p]),
    "spasm_b200_comm_destroy": (None, []),
    "spasm_b200_comm_world": (C.c_int, []),
    "spasm_b200_comm_result_root": (None, [C.c_int]),
    "spasm_b200_gemm_sub": (None, [C.c_int64, C.c_int, C.c_int, C.c_int, abi.i32_p, abi.i32_p, abi.i32_p, C.c_int]),
    "spasm_b200_gemm_time": (C.c_double, [C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spasm_b200_prng_stream": (None, [C.c_int64, C.c_uint64, C.c_uint32, C.c_int, abi.i32_p]),
    "spasm_b200_last_pivot_pairs": (C.c_int, [abi.c_int_p, abi.c_int_p, abi.c_int_p]),
}


class Stats(C.Structure):
    """ctypes mirror of struct spasm_b200_stats (include/spasm_b200.h)."""
    _fields_ = [("kernel_launches", C.c_int64), ("ms_pivots", C.c_double), ("ms_pivots_greedy", C.c_double),
                ("ms_solve", C.c_double), ("ms_dense", C.c_double), ("ms_dense_gemm", C.c_double),
                ("ms_total_echelonize", C.c_double), ("ms_device_echelonize", C.c_double), ("ms_k_greedy", C.c_double),
                ("ms_k_panel_solve", C.c_double), ("solve_bytes", C.c_double), ("solve_rows", C.c_int64),
                ("solve_batches", C.c_int64), ("solve_traffic_model", C.c_double), ("gemm_fieldops", C.c_double),
                ("gemm_int8_ops", C.c_double), ("greedy_edges", C.c_int64), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("nccl_bytes", C.c_int64), ("nrounds", C.c_int), ("found_FL", C.c_int * 64), ("found_FLcol", C.c_int * 64),
                ("found_greedy", C.c_int * 64), ("density", C.c_double * 64), ("finish", C.c_int), ("nblocks", C.c_int),
                ("block_Sn", C.c_int * 4096), ("block_Sm", C.c_int * 4096), ("block_rr", C.c_int * 4096),
                ("block_w", C.c_int * 4096), ("dag_depth", C.c_int)]


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
        L = C.CDLL(LIB_PATH)        # RTLD_LOCAL: the reference library used by the tests exports the same names
        abi.bind(L)
        for name, (res, args) in EXTRA_PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib
