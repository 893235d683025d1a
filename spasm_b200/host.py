"""Host-side mirror of the reference's operator interface, on top of the C ABI (include/spasm.h).

Every function takes the bound shared library as first argument, so the very same calls drive
  * the product   : `spasm_b200.lib()`  (CUDA, sm_100a)
  * the reference : `oracle.ref()`      (tests only)
and the parity tests read like the reference's own test programs (tests/echelonize.c, tests/kernel.c ...):
load -> compress -> echelonize -> rref / kernel.

Names and argument meaning follow the reference: spasm_compress, spasm_echelonize, spasm_rref,
spasm_kernel, spasm_schur_dense, ...  Errors follow it too: the C side aborts with errx(1, ...).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class CsrHandle:
    """A `struct spasm_csr *` allocated by the library (malloc memory), freed with spasm_csr_free."""

    def __init__(self, lib, ptr, owned=True):
        self.lib, self.ptr, self.owned = lib, ptr, owned

    def __del__(self):
        if getattr(self, "owned", False) and self.ptr:
            self.lib.spasm_csr_free(self.ptr)
            self.ptr = None

    n = property(lambda self: self.ptr.contents.n)
    m = property(lambda self: self.ptr.contents.m)
    prime = property(lambda self: int(self.ptr.contents.field[0].p))
    nnz = property(lambda self: int(self.lib.spasm_nnz(self.ptr)))

    def numpy(self) -> dict:
        return abi.csr_to_numpy(self.ptr)


class LuHandle:
    """A `struct spasm_lu *` returned by spasm_echelonize."""

    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            self.lib.spasm_lu_free(self.ptr)
            self.ptr = None

    @property
    def rank(self) -> int:
        return self.ptr.contents.U.contents.n

    @property
    def U(self) -> dict:
        return abi.csr_to_numpy(self.ptr.contents.U)

    @property
    def qinv(self) -> np.ndarray:
        m = self.ptr.contents.U.contents.m
        return np.ctypeslib.as_array(self.ptr.contents.qinv, shape=(m,)).copy() if m else np.zeros(0, np.int32)

    @property
    def L(self) -> dict | None:
        """n x rank factor of the L mode (opts.L / opts.complete), None otherwise"""
        return abi.csr_to_numpy(self.ptr.contents.L) if self.ptr.contents.L else None

    @property
    def p(self) -> np.ndarray | None:
        """p[k] = row of the input behind row k of U (L mode)"""
        r = self.rank
        if not self.ptr.contents.p:
            return None
        return np.ctypeslib.as_array(self.ptr.contents.p, shape=(r,)).copy() if r else np.zeros(0, np.int32)


def compress(lib, trip) -> CsrHandle:
    """spasm_triplet_alloc + spasm_add_entry (bulk) + spasm_compress  (reference: tools/rank.c:83-90).

    `trip` has n, m, prime and 0-based arrays i, j, x in file order.  Values are reduced into the
    balanced range and multiples of p dropped, exactly as spasm_add_entry does one entry at a time
    (reference: src/spasm_triplet.c:7-24)."""
    nz = len(trip.i)
    T = lib.spasm_triplet_alloc(trip.n, trip.m, max(nz, 1), trip.prime, True)
    t = T.contents
    p = int(trip.prime)
    x = np.asarray(trip.x, dtype=np.int64)
    # balanced representative, C-style truncated remainder then one correction (src/spasm_ZZp.c:17-30)
    r = np.fmod(x, p)
    half, mhalf = p // 2, p // 2 - p + 1
    r = np.where(r > half, r - p, np.where(r < mhalf, r + p, r))
    keep = r != 0
    ti = np.ascontiguousarray(np.asarray(trip.i, np.int32)[keep])
    tj = np.ascontiguousarray(np.asarray(trip.j, np.int32)[keep])
    tx = np.ascontiguousarray(r[keep].astype(np.int32))
    k = len(ti)
    if k:
        C.memmove(t.i, ti.ctypes.data, 4 * k)
        C.memmove(t.j, tj.ctypes.data, 4 * k)
        C.memmove(t.x, tx.ctypes.data, 4 * k)
    t.nz = k
    A = lib.spasm_compress(T)
    lib.spasm_triplet_free(T)
    return CsrHandle(lib, A)


def from_numpy(lib, d: dict) -> CsrHandle:
    nnz = int(d["p"][d["n"]])
    A = lib.spasm_csr_alloc(d["n"], d["m"], max(nnz, 1), d["prime"], True)
    a = A.contents
    C.memmove(a.p, np.ascontiguousarray(d["p"], np.int64).ctypes.data, 8 * (d["n"] + 1))
    if nnz:
        C.memmove(a.j, np.ascontiguousarray(d["j"], np.int32).ctypes.data, 4 * nnz)
        C.memmove(a.x, np.ascontiguousarray(d["x"], np.int32).ctypes.data, 4 * nnz)
    return CsrHandle(lib, A)


def default_opts(lib, **kw) -> abi.EchelonizeOpts:
    o = abi.EchelonizeOpts()
    lib.spasm_echelonize_init_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def echelonize(lib, A: CsrHandle, opts: abi.EchelonizeOpts | None = None) -> LuHandle:
    """spasm_echelonize (reference: src/spasm_echelonize.c:473)."""
    if opts is None:
        opts = default_opts(lib)
    return LuHandle(lib, lib.spasm_echelonize(A.ptr, C.byref(opts)))


def factorization_verify(lib, A: CsrHandle, fact: LuHandle, seed: int) -> bool:
    """spasm_factorization_verify (reference: src/spasm_certificate.c:164): x*A == (x*L)*U for a random x on the pivotal rows"""
    return bool(lib.spasm_factorization_verify(A.ptr, fact.ptr, seed))


def solve(lib, fact: LuHandle, b: np.ndarray):
    """spasm_solve (reference: src/spasm_solve.c:13): x with x*A == b, and whether it exists"""
    n = fact.ptr.contents.L.contents.n
    b = np.ascontiguousarray(b, np.int32)
    x = np.zeros(max(n, 1), np.int32)
    ok = lib.spasm_solve(fact.ptr, b.ctypes.data_as(abi.i32_p), x.ctypes.data_as(abi.i32_p))
    return x[:n], bool(ok)


def rank(lib, A: CsrHandle, opts=None) -> int:
    """What tools/rank prints (reference: tools/rank.c:100-103)."""
    return echelonize(lib, A, opts).rank


def rref(lib, fact: LuHandle):
    """spasm_rref (reference: src/spasm_rref.c:22). Returns (R, Rqinv)."""
    m = fact.ptr.contents.U.contents.m
    Rqinv = np.zeros(max(m, 1), np.int32)
    R = CsrHandle(lib, lib.spasm_rref(fact.ptr, abi.as_int_p(Rqinv)))
    return R, Rqinv[:m]


def kernel(lib, fact: LuHandle) -> CsrHandle:
    """spasm_kernel (reference: src/spasm_kernel.c:9)."""
    return CsrHandle(lib, lib.spasm_kernel(fact.ptr))


def transpose(lib, A: CsrHandle) -> CsrHandle:
    return CsrHandle(lib, lib.spasm_transpose(A.ptr, 1))


def new_fact(lib, A: CsrHandle):
    """An empty `struct spasm_lu` the way spasm_echelonize sets it up (reference: src/spasm_echelonize.c:493-514),
    for driving spasm_pivots_extract_structural / spasm_schur* directly like tests/schur.c does."""
    n, m, prime = A.n, A.m, A.prime
    U = lib.spasm_csr_alloc(n, m, max(A.nnz, 1), prime, True)
    U.contents.n = 0
    qinv = (C.c_int * max(m, 1))(*([-1] * max(m, 1)))
    fact = abi.Lu()
    fact.r = 0
    fact.complete = False
    fact.L = None
    fact.U = U
    fact.qinv = C.cast(qinv, abi.c_int_p)
    fact.p = None
    fact.Ltmp = None
    fact._keep = qinv          # keep the Python-owned qinv alive
    return fact


def pivots_extract_structural(lib, A: CsrHandle, opts=None):
    """spasm_pivots_extract_structural (reference: src/spasm_pivots.c:369).
    Returns (npiv, p, fact) with fact.U holding the npiv pivotal rows."""
    if opts is None:
        opts = default_opts(lib)
    fact = new_fact(lib, A)
    p = np.zeros(max(A.n, 1), np.int32)
    npiv = lib.spasm_pivots_extract_structural(A.ptr, None, C.byref(fact), abi.as_int_p(p), C.byref(opts))
    return npiv, p[:A.n], fact


def fact_pairs(fact, p, npiv):
    """(row of A, pivot column) of each structural pivot, from U (pivot = first entry of a U row)."""
    U = abi.csr_to_numpy(fact.U)
    cols = U["j"][U["p"][:npiv]]
    return np.asarray(p[:npiv], np.int32), cols.astype(np.int32)


def schur_dense(lib, A: CsrHandle, p: np.ndarray, n: int, fact) -> tuple[np.ndarray, np.ndarray]:
    """spasm_schur_dense (reference: src/spasm_schur.c:257) with datatype chosen like the reference does."""
    m = A.m
    Sm = m - fact.U.contents.n
    dt = lib.spasm_datatype_choose(A.prime)
    npdt = {abi.SPASM_DOUBLE: np.float64, abi.SPASM_FLOAT: np.float32, abi.SPASM_I64: np.int64}[dt]
    S = np.zeros((max(n, 1), max(Sm, 1)), npdt)
    q = np.zeros(max(Sm, 1), np.int32)
    p_out = np.zeros(max(n, 1), np.int32)
    pp = np.ascontiguousarray(p, np.int32)
    lib.spasm_schur_dense(A.ptr, abi.as_int_p(pp), n, None, C.byref(fact), S.ctypes.data_as(C.c_void_p), dt,
                          abi.as_int_p(q), abi.as_int_p(p_out))
    return S[:n, :Sm].astype(np.int64), q[:Sm]


def ffpack_rref(lib, prime: int, M: np.ndarray):
    """spasm_ffpack_rref on an int matrix; returns (rank, qinv, packed output as int64)."""
    n, m = M.shape
    dt = lib.spasm_datatype_choose(prime)
    npdt = {abi.SPASM_DOUBLE: np.float64, abi.SPASM_FLOAT: np.float32, abi.SPASM_I64: np.int64}[dt]
    A = np.array(M, dtype=npdt, order="C", copy=True)       # the call works in place
    qinv = (C.c_size_t * max(m, 1))()
    r = lib.spasm_ffpack_rref(prime, n, m, A.ctypes.data_as(C.c_void_p), m, dt, qinv)
    return r, np.array(list(qinv)[:m], np.int64), A.astype(np.int64)
