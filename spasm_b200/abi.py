"""ctypes description of the SpaSM C ABI (include/spasm.h == reference src/spasm.h).

The same description binds two different shared objects:
  * spasm_b200/lib/libspasm_b200.so -- this repository's CUDA implementation (the product);
  * oracle/_ref/libspasm_ref.so     -- the reference's own C sources compiled in place (tests only).
Nothing here computes anything; it only declares layouts and prototypes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

i64 = C.c_int64
i32 = C.c_int32
c_int_p = C.POINTER(C.c_int)
i64_p = C.POINTER(i64)
i32_p = C.POINTER(i32)


class Field(C.Structure):          # include/spasm.h: struct spasm_field_struct
    _fields_ = [("p", i64), ("halfp", i64), ("mhalfp", i64), ("dinvp", C.c_double)]


class Csr(C.Structure):            # include/spasm.h: struct spasm_csr
    _fields_ = [("nzmax", i64), ("n", C.c_int), ("m", C.c_int), ("p", i64_p), ("j", c_int_p), ("x", i32_p),
                ("field", Field * 1)]


class Triplet(C.Structure):        # include/spasm.h: struct spasm_triplet
    _fields_ = [("nzmax", i64), ("nz", i64), ("n", C.c_int), ("m", C.c_int), ("i", c_int_p), ("j", c_int_p),
                ("x", i32_p), ("field", Field * 1)]


class Lu(C.Structure):             # include/spasm.h: struct spasm_lu
    _fields_ = [("r", C.c_int), ("complete", C.c_bool), ("L", C.POINTER(Csr)), ("U", C.POINTER(Csr)),
                ("qinv", c_int_p), ("p", c_int_p), ("Ltmp", C.POINTER(Triplet))]


class EchelonizeOpts(C.Structure):  # include/spasm.h: struct echelonize_opts
    _fields_ = [("enable_greedy_pivot_search", C.c_bool), ("enable_tall_and_skinny", C.c_bool),
                ("enable_dense", C.c_bool), ("enable_GPLU", C.c_bool), ("L", C.c_bool), ("complete", C.c_bool),
                ("min_pivot_proportion", C.c_double), ("max_round", C.c_int), ("sparsity_threshold", C.c_double),
                ("dense_block_size", C.c_int), ("low_rank_ratio", C.c_double), ("tall_and_skinny_ratio", C.c_double),
                ("low_rank_start_weight", C.c_double)]


class RankCertificate(C.Structure):   # include/spasm.h: struct spasm_rank_certificate
    _fields_ = [("r", C.c_int), ("prime", i64), ("hash", C.c_uint8 * 32), ("i", c_int_p), ("j", c_int_p), ("x", i32_p), ("y", i32_p)]


class Sha256Ctx(C.Structure):
    _fields_ = [("h", C.c_uint32 * 8), ("Nl", C.c_uint32), ("Nh", C.c_uint32), ("data", C.c_uint32 * 16),
                ("num", C.c_uint32), ("md_len", C.c_uint32)]


class PrngCtx(C.Structure):
    _fields_ = [("block", C.c_uint32 * 11), ("hash", C.c_uint32 * 8), ("prime", C.c_uint32), ("mask", C.c_uint32),
                ("counter", C.c_int), ("i", C.c_int), ("field", Field * 1)]


SPASM_DOUBLE, SPASM_FLOAT, SPASM_I64 = 0, 1, 2

CsrP = C.POINTER(Csr)
TripletP = C.POINTER(Triplet)
LuP = C.POINTER(Lu)
OptsP = C.POINTER(EchelonizeOpts)
CertP = C.POINTER(RankCertificate)

# name -> (restype, argtypes); every symbol include/spasm.h declares
PROTOTYPES = {
    "spasm_field_init": (None, [i64, C.POINTER(Field)]),
    "spasm_ZZp_init": (i32, [C.POINTER(Field), i64]),
    "spasm_ZZp_add": (i32, [C.POINTER(Field), i32, i32]),
    "spasm_ZZp_sub": (i32, [C.POINTER(Field), i32, i32]),
    "spasm_ZZp_mul": (i32, [C.POINTER(Field), i32, i32]),
    "spasm_ZZp_inverse": (i32, [C.POINTER(Field), i32]),
    "spasm_ZZp_axpy": (i32, [C.POINTER(Field), i32, i32, i32]),
    "spasm_SHA256_init": (None, [C.POINTER(Sha256Ctx)]),
    "spasm_SHA256_update": (None, [C.POINTER(Sha256Ctx), C.c_void_p, C.c_size_t]),
    "spasm_SHA256_final": (None, [C.c_void_p, C.POINTER(Sha256Ctx)]),
    "spasm_prng_seed": (None, [C.c_void_p, i64, C.c_uint32, C.POINTER(PrngCtx)]),
    "spasm_prng_seed_simple": (None, [i64, C.c_uint64, C.c_uint32, C.POINTER(PrngCtx)]),
    "spasm_prng_u32": (C.c_uint32, [C.POINTER(PrngCtx)]),
    "spasm_prng_ZZp": (i32, [C.POINTER(PrngCtx)]),
    "spasm_wtime": (C.c_double, []),
    "spasm_nnz": (i64, [CsrP]),
    "spasm_malloc": (C.c_void_p, [i64]),
    "spasm_calloc": (C.c_void_p, [i64, i64]),
    "spasm_realloc": (C.c_void_p, [C.c_void_p, i64]),
    "spasm_csr_alloc": (CsrP, [C.c_int, C.c_int, i64, i64, C.c_bool]),
    "spasm_csr_realloc": (None, [CsrP, i64]),
    "spasm_csr_resize": (None, [CsrP, C.c_int, C.c_int]),
    "spasm_csr_free": (None, [CsrP]),
    "spasm_triplet_alloc": (TripletP, [C.c_int, C.c_int, i64, i64, C.c_bool]),
    "spasm_triplet_realloc": (None, [TripletP, i64]),
    "spasm_triplet_free": (None, [TripletP]),
    "spasm_dm_alloc": (C.c_void_p, [C.c_int, C.c_int]),
    "spasm_dm_free": (None, [C.c_void_p]),
    "spasm_lu_free": (None, [LuP]),
    "spasm_human_format": (None, [i64, C.c_char_p]),
    "spasm_get_num_threads": (C.c_int, []),
    "spasm_get_thread_num": (C.c_int, []),
    "spasm_add_entry": (None, [TripletP, C.c_int, C.c_int, i64]),
    "spasm_triplet_transpose": (None, [TripletP]),
    "spasm_compress": (CsrP, [TripletP]),
    "spasm_triplet_load": (TripletP, [C.c_void_p, i64, C.c_void_p]),
    "spasm_triplet_save": (None, [TripletP, C.c_void_p]),
    "spasm_csr_save": (None, [CsrP, C.c_void_p]),
    "spasm_transpose": (CsrP, [CsrP, C.c_int]),
    "spasm_xApy": (None, [i32_p, CsrP, i32_p]),
    "spasm_Axpy": (None, [CsrP, i32_p, i32_p]),
    # small-grain host verifiers (csrc/host/verify.c; reference: src/spasm.h:199-223, :268)
    "spasm_scatter": (None, [CsrP, C.c_int, i32, i32_p]),
    "spasm_dfs": (C.c_int, [C.c_int, CsrP, C.c_int, c_int_p, c_int_p, c_int_p, c_int_p]),
    "spasm_reach": (C.c_int, [CsrP, CsrP, C.c_int, C.c_int, c_int_p, c_int_p]),
    "spasm_sparse_triangular_solve": (C.c_int, [CsrP, CsrP, C.c_int, c_int_p, i32_p, c_int_p]),
    "spasm_dense_back_solve": (None, [CsrP, i32_p, i32_p, c_int_p]),
    "spasm_dense_forward_solve": (C.c_bool, [CsrP, i32_p, i32_p, c_int_p]),
    "spasm_pvec": (None, [c_int_p, i32_p, i32_p, C.c_int]),
    "spasm_ipvec": (None, [c_int_p, i32_p, i32_p, C.c_int]),
    "spasm_pinv": (c_int_p, [c_int_p, C.c_int]),
    "spasm_permute": (CsrP, [CsrP, c_int_p, c_int_p, C.c_int]),
    "spasm_random_permutation": (c_int_p, [C.c_int]),
    "spasm_range_pvec": (None, [c_int_p, C.c_int, C.c_int, c_int_p]),
    "spasm_submatrix": (CsrP, [CsrP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "spasm_kernel_from_rref": (CsrP, [CsrP, c_int_p]),
    "spasm_pivots_extract_structural": (C.c_int, [CsrP, c_int_p, LuP, c_int_p, OptsP]),
    "spasm_schur_estimate_density": (C.c_double, [CsrP, c_int_p, C.c_int, CsrP, c_int_p, C.c_int]),
    "spasm_schur": (CsrP, [CsrP, c_int_p, C.c_int, LuP, C.c_double, TripletP, c_int_p, c_int_p]),
    "spasm_schur_dense": (None, [CsrP, c_int_p, C.c_int, c_int_p, LuP, C.c_void_p, C.c_int, c_int_p, c_int_p]),
    "spasm_schur_dense_randomized": (None, [CsrP, c_int_p, C.c_int, CsrP, c_int_p, C.c_void_p, C.c_int, c_int_p,
                                            C.c_int, C.c_int]),
    "spasm_ffpack_rref": (C.c_int, [i64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "spasm_ffpack_LU": (C.c_int, [i64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_size_t),
                                  C.POINTER(C.c_size_t)]),
    "spasm_datatype_read": (i32, [C.c_void_p, C.c_size_t, C.c_int]),
    "spasm_datatype_write": (None, [C.c_void_p, C.c_size_t, C.c_int, i32]),
    "spasm_datatype_size": (C.c_size_t, [C.c_int]),
    "spasm_datatype_choose": (C.c_int, [i64]),
    "spasm_datatype_name": (C.c_char_p, [C.c_int]),
    "spasm_echelonize_init_opts": (None, [OptsP]),
    "spasm_echelonize": (LuP, [CsrP, OptsP]),
    "spasm_rref": (CsrP, [LuP, c_int_p]),
    "spasm_kernel": (CsrP, [LuP]),
    "spasm_solve": (C.c_bool, [LuP, i32_p, i32_p]),
    "spasm_gesv": (CsrP, [LuP, CsrP, C.POINTER(C.c_bool)]),
    "spasm_certificate_rank_create": (CertP, [CsrP, C.c_void_p, LuP]),
    "spasm_certificate_rank_verify": (C.c_bool, [CsrP, C.c_void_p, CertP]),
    "spasm_rank_certificate_save": (None, [CertP, C.c_void_p]),
    "spasm_rank_certificate_load": (C.c_bool, [C.c_void_p, CertP]),
    "spasm_factorization_verify": (C.c_bool, [CsrP, LuP, C.c_uint64]),
}


def bind(lib: C.CDLL, names=None) -> C.CDLL:
    """Attach prototypes; raises AttributeError naming the first missing symbol."""
    for name in (names or PROTOTYPES):
        fn = getattr(lib, name)       # AttributeError if the library does not export it
        fn.restype, fn.argtypes = PROTOTYPES[name]
    return lib


# ------------------------------------------------------------------ numpy <-> C helpers

def as_int_p(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int_p)


def as_i32_p(a: np.ndarray):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(i32_p)


def as_i64_p(a: np.ndarray):
    assert a.dtype == np.int64 and a.flags.c_contiguous
    return a.ctypes.data_as(i64_p)


def csr_to_numpy(A) -> dict:
    """Copy a C `struct spasm_csr *` into numpy arrays."""
    a = A.contents
    n, m = a.n, a.m
    p = np.ctypeslib.as_array(a.p, shape=(n + 1,)).copy() if n >= 0 else np.zeros(1, np.int64)
    nnz = int(p[n])
    if nnz > 0:
        j = np.ctypeslib.as_array(a.j, shape=(nnz,)).copy().astype(np.int32)
        x = np.ctypeslib.as_array(a.x, shape=(nnz,)).copy().astype(np.int32) if a.x else None
    else:
        j = np.zeros(0, np.int32)
        x = np.zeros(0, np.int32)
    return {"n": n, "m": m, "p": p.astype(np.int64), "j": j, "x": x, "prime": int(a.field[0].p)}
