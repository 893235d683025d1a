"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python loaders for
  * oracle/liboracle.so          -- our single-threaded CPU restatement (oracle.c + ffpack_restate.c)
  * oracle/_ref/libspasm_ref.so  -- the reference's own C sources, compiled where they lie under
                                    /root/reference by oracle/Makefile (plus ffpack_restate.c for
                                    the one C++ file that needs FFLAS-FFPACK).  Prebuilt; the GPU box
                                    has no /root/reference and just loads the .so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under spasm_b200/ does (tests/test_no_oracle_in_product.py checks it).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = "/root/reference"

i64 = C.c_int64
i32 = C.c_int32
MAX_ROUNDS = 64
MAX_BLOCKS = 4096


class OCsr(C.Structure):
    _fields_ = [("n", C.c_int), ("m", C.c_int), ("nzmax", i64), ("p", C.POINTER(i64)), ("j", C.POINTER(C.c_int)),
                ("x", C.POINTER(i32)), ("prime", i64)]


class OOpts(C.Structure):
    _fields_ = [("enable_greedy_pivot_search", C.c_int), ("enable_tall_and_skinny", C.c_int), ("enable_dense", C.c_int),
                ("enable_GPLU", C.c_int), ("min_pivot_proportion", C.c_double), ("max_round", C.c_int),
                ("sparsity_threshold", C.c_double), ("dense_block_size", C.c_int), ("low_rank_ratio", C.c_double),
                ("tall_and_skinny_ratio", C.c_double), ("low_rank_start_weight", C.c_double)]


class OLu(C.Structure):
    _fields_ = [("rank", C.c_int), ("U", C.POINTER(OCsr)), ("qinv", C.POINTER(C.c_int)),
                ("nrounds", C.c_int), ("found_FL", C.c_int * MAX_ROUNDS), ("found_FLcol", C.c_int * MAX_ROUNDS),
                ("found_greedy", C.c_int * MAX_ROUNDS), ("pair_start", C.c_int * (MAX_ROUNDS + 1)), ("npairs", C.c_int),
                ("pair_row", C.POINTER(C.c_int)), ("pair_col", C.POINTER(C.c_int)), ("density", C.c_double * MAX_ROUNDS),
                ("finish", C.c_int), ("nblocks", C.c_int), ("block_Sn", C.c_int * MAX_BLOCKS),
                ("block_Sm", C.c_int * MAX_BLOCKS), ("block_rr", C.c_int * MAX_BLOCKS), ("block_w", C.c_int * MAX_BLOCKS),
                ("tsolve_bytes", C.c_double), ("tsolve_rows", i64), ("greedy_edges", i64), ("dense_fieldops", C.c_double),
                ("seconds_pivots", C.c_double), ("seconds_schur", C.c_double), ("seconds_dense", C.c_double),
                ("seconds_total", C.c_double)]


_lib = None
_ref = None
_libc = C.CDLL(None)


def build(with_ref: bool | None = None) -> None:
    """Compile liboracle.so, and oracle/_ref when the reference tree is present (this container only)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if with_ref is None:
        with_ref = os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))
    if with_ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)
        if os.path.exists(os.path.join(os.path.dirname(HERE), "spasm_b200", "lib", "libspasm_b200.so")):
            # the reference's unmodified tools linked against the product library
            subprocess.run(["make", "-s", "-C", HERE, "b200_tools"], check=True)
            # ... and its unmodified test programs (tests/test_reference_tests.py)
            subprocess.run(["make", "-s", "-C", HERE, "b200_tests"], check=True)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(with_ref=False)
        L = C.CDLL(path)
        P = C.POINTER
        L.oracle_default_opts.argtypes = [P(OOpts)]
        L.oracle_csr_alloc.restype = P(OCsr)
        L.oracle_csr_alloc.argtypes = [C.c_int, C.c_int, i64, i64]
        L.oracle_csr_free.argtypes = [P(OCsr)]
        L.oracle_compress.restype = P(OCsr)
        L.oracle_compress.argtypes = [C.c_int, C.c_int, i64, P(C.c_int), P(C.c_int), P(i64), i64]
        L.oracle_transpose.restype = P(OCsr)
        L.oracle_transpose.argtypes = [P(OCsr)]
        L.oracle_echelonize.restype = P(OLu)
        L.oracle_echelonize.argtypes = [P(OCsr), P(OOpts)]
        L.oracle_lu_free.argtypes = [P(OLu)]
        L.oracle_rref.restype = P(OCsr)
        L.oracle_rref.argtypes = [P(OCsr), P(C.c_int), P(C.c_int)]
        L.oracle_kernel.restype = P(OCsr)
        L.oracle_kernel.argtypes = [P(OCsr), P(C.c_int)]
        L.oracle_pivots_find.restype = C.c_int
        L.oracle_pivots_find.argtypes = [P(OCsr), C.c_int, P(C.c_int), P(C.c_int), P(C.c_int), P(i64)]
        L.oracle_tsolve.restype = C.c_int
        L.oracle_tsolve.argtypes = [P(OCsr), P(OCsr), C.c_int, P(C.c_int), P(i32), P(C.c_int)]
        L.oracle_schur_dense.argtypes = [P(OCsr), P(C.c_int), C.c_int, P(OCsr), P(C.c_int), P(i32), P(C.c_int)]
        L.oracle_schur_dense_randomized.argtypes = [P(OCsr), P(C.c_int), C.c_int, P(OCsr), P(C.c_int), P(i32), P(C.c_int),
                                                    C.c_int, C.c_int]
        L.oracle_schur.restype = P(OCsr)
        L.oracle_schur.argtypes = [P(OCsr), P(C.c_int), C.c_int, P(OCsr), P(C.c_int), C.c_double]
        L.oracle_schur_estimate_density.restype = C.c_double
        L.oracle_schur_estimate_density.argtypes = [P(OCsr), P(C.c_int), C.c_int, P(OCsr), P(C.c_int), C.c_int]
        L.oracle_dense_rref_i32.restype = C.c_int
        L.oracle_dense_rref_i32.argtypes = [i64, C.c_int, C.c_int, P(i32), P(C.c_int)]
        L.oracle_zp_mul.restype = i32
        L.oracle_zp_mul.argtypes = [i64, i32, i32]
        L.oracle_zp_inverse.restype = i32
        L.oracle_zp_inverse.argtypes = [i64, i32]
        L.oracle_zp_axpy.restype = i32
        L.oracle_zp_axpy.argtypes = [i64, i32, i32, i32]
        L.oracle_sha256.argtypes = [C.c_void_p, i64, C.c_void_p]
        L.oracle_prng_stream.argtypes = [i64, C.c_uint64, C.c_uint32, C.c_int, P(i32)]
        _lib = L
    return _lib


def ref_path() -> str:
    return os.path.join(HERE, "_ref", "libspasm_ref.so")


def ref():
    """The reference library (ABI of include/spasm.h), or None when it was not built."""
    global _ref
    if _ref is None:
        if not os.path.exists(ref_path()):
            return None
        from spasm_b200 import abi      # ABI description only -- no product code runs through it
        _ref = abi.bind(C.CDLL(ref_path()))
    return _ref


_ref_det = None


def ref_det():
    """The reference's sources with src/spasm_pivots.c compiled without OpenMP (greedy search in row order = the
    one-thread parity target) and everything else with OpenMP: fast enough for the full-size BASELINE configs."""
    global _ref_det
    if _ref_det is None:
        path = os.path.join(HERE, "_ref", "libspasm_ref_det.so")
        if not os.path.exists(path):
            return None
        from spasm_b200 import abi
        _ref_det = abi.bind(C.CDLL(path))
    return _ref_det


def reset_rand() -> None:
    """glibc rand() is never seeded by the reference (seed 1); restore that state before each run."""
    _libc.srand(1)


# ------------------------------------------------------------------ numpy helpers

def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def csr_numpy(ptr) -> dict:
    """Copy a `struct ocsr *` into numpy arrays (no ownership taken)."""
    a = ptr.contents
    p = np.ctypeslib.as_array(a.p, shape=(a.n + 1,)).copy()
    nnz = int(p[a.n])
    if nnz:
        j = np.ctypeslib.as_array(a.j, shape=(nnz,)).astype(np.int32)
        x = np.ctypeslib.as_array(a.x, shape=(nnz,)).astype(np.int32)
    else:
        j = np.zeros(0, np.int32)
        x = np.zeros(0, np.int32)
    return {"n": a.n, "m": a.m, "p": p, "j": j, "x": x, "prime": int(a.prime)}


class Matrix:
    """Owning wrapper around a `struct ocsr *`."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().oracle_csr_free(self.ptr)
            self.ptr = None

    @property
    def n(self):
        return self.ptr.contents.n

    @property
    def m(self):
        return self.ptr.contents.m

    @property
    def prime(self):
        return self.ptr.contents.prime

    def numpy(self) -> dict:
        return csr_numpy(self.ptr)


def compress(trip) -> Matrix:
    """Triplets (spasm_b200.synthetic.Triplets or anything with n, m, prime, i, j, x) -> oracle CSR."""
    ti = np.ascontiguousarray(trip.i, dtype=np.int32)
    tj = np.ascontiguousarray(trip.j, dtype=np.int32)
    tx = np.ascontiguousarray(trip.x, dtype=np.int64)
    return Matrix(lib().oracle_compress(trip.n, trip.m, len(ti), _ip(ti), _ip(tj), tx.ctypes.data_as(C.POINTER(i64)), trip.prime))


def from_numpy(d: dict) -> Matrix:
    """CSR dict (n, m, p, j, x, prime) -> oracle CSR (copy)."""
    nnz = int(d["p"][d["n"]])
    ptr = lib().oracle_csr_alloc(d["n"], d["m"], max(nnz, 1), d["prime"])
    a = ptr.contents
    C.memmove(a.p, np.ascontiguousarray(d["p"], np.int64).ctypes.data, 8 * (d["n"] + 1))
    if nnz:
        C.memmove(a.j, np.ascontiguousarray(d["j"], np.int32).ctypes.data, 4 * nnz)
        C.memmove(a.x, np.ascontiguousarray(d["x"], np.int32).ctypes.data, 4 * nnz)
    return Matrix(ptr)


def default_opts(**kw) -> OOpts:
    o = OOpts()
    lib().oracle_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Echelon:
    """Result of oracle_echelonize with numpy copies of everything the parity tests compare."""

    def __init__(self, ptr):
        f = ptr.contents
        self._ptr = ptr
        self.rank = f.rank
        self.U = csr_numpy(f.U)
        m = self.U["m"]
        self.qinv = np.ctypeslib.as_array(f.qinv, shape=(m,)).copy() if m else np.zeros(0, np.int32)
        self.nrounds = f.nrounds
        r = min(f.nrounds, MAX_ROUNDS)
        self.found = [(f.found_FL[k], f.found_FLcol[k], f.found_greedy[k]) for k in range(r)]
        self.density = [f.density[k] for k in range(r)]
        self.pair_start = [f.pair_start[k] for k in range(r + 1)]
        if f.npairs:
            self.pair_row = np.ctypeslib.as_array(f.pair_row, shape=(f.npairs,)).copy()
            self.pair_col = np.ctypeslib.as_array(f.pair_col, shape=(f.npairs,)).copy()
        else:
            self.pair_row = np.zeros(0, np.int32)
            self.pair_col = np.zeros(0, np.int32)
        self.finish = f.finish
        nb = min(f.nblocks, MAX_BLOCKS)
        self.blocks = [(f.block_Sn[k], f.block_Sm[k], f.block_rr[k], f.block_w[k]) for k in range(nb)]
        self.stats = {k: getattr(f, k) for k in ("tsolve_bytes", "tsolve_rows", "greedy_edges", "dense_fieldops",
                                                   "seconds_pivots", "seconds_schur", "seconds_dense", "seconds_total")}

    def __del__(self):
        if getattr(self, "_ptr", None):
            lib().oracle_lu_free(self._ptr)
            self._ptr = None


def echelonize(A: Matrix, opts: OOpts | None = None) -> Echelon:
    reset_rand()
    return Echelon(lib().oracle_echelonize(A.ptr, C.byref(opts) if opts is not None else None))


def rref(U: dict, qinv: np.ndarray):
    Um = from_numpy(U)
    q = np.ascontiguousarray(qinv, np.int32)
    Rq = np.zeros(max(U["m"], 1), np.int32)
    R = Matrix(lib().oracle_rref(Um.ptr, _ip(q), _ip(Rq)))
    return R.numpy(), Rq[:U["m"]]


def kernel(U: dict, qinv: np.ndarray):
    Um = from_numpy(U)
    q = np.ascontiguousarray(qinv, np.int32)
    return Matrix(lib().oracle_kernel(Um.ptr, _ip(q))).numpy()


# ------------------------------------------------------------------ canonical artefacts

def canonical(M: dict) -> dict:
    """Rows sorted by the column of their FIRST entry (the pivot for U/R, the free column for a kernel
    basis); inside a row the first entry stays first and the others are sorted by column."""
    n, p, j, x = M["n"], M["p"], M["j"], M["x"]
    lens = np.diff(p)
    assert (lens > 0).all(), "canonical form needs non-empty rows"
    first_col = j[p[:-1]]
    order = np.argsort(first_col, kind="stable")
    newp = np.zeros(n + 1, np.int64)
    newp[1:] = np.cumsum(lens[order])
    nj = np.empty_like(j)
    nx = np.empty_like(x)
    # vectorised: key = (new row, is_not_first, column)
    row_of = np.repeat(np.arange(n), lens)
    is_first = np.zeros(len(j), bool)
    is_first[p[:-1]] = True
    rank_of_row = np.empty(n, np.int64)
    rank_of_row[order] = np.arange(n)
    key = np.lexsort((j, ~is_first, rank_of_row[row_of]))
    nj[:] = j[key]
    nx[:] = x[key]
    return {"n": n, "m": M["m"], "p": newp, "j": nj, "x": nx, "prime": M.get("prime", 0)}


def canonical_hash(M: dict) -> str:
    """sha256 over (n, m, row pointers, columns, values) of the canonical form, little endian."""
    if M["n"] == 0:
        c = {"n": 0, "m": M["m"], "p": np.zeros(1, np.int64), "j": np.zeros(0, np.int32), "x": np.zeros(0, np.int32)}
    else:
        c = canonical(M)
    h = hashlib.sha256()
    h.update(np.array([c["n"], c["m"]], np.int64).tobytes())
    h.update(np.ascontiguousarray(c["p"], np.int64).tobytes())
    h.update(np.ascontiguousarray(c["j"], np.int32).tobytes())
    h.update(np.ascontiguousarray(c["x"], np.int32).tobytes())
    return h.hexdigest()


def pivot_columns(qinv: np.ndarray) -> np.ndarray:
    return np.nonzero(np.asarray(qinv) >= 0)[0].astype(np.int32)
