"""The L side of the factorization (opts->L / opts->complete; 8f-4: rank certificates) on the GPU.

What the reference's own tests pin (tests/lu.c, solve.c, gesv.c, rank_cert.c -- run unmodified against the library by
tests/test_reference_tests.py on the 32 fixtures x 6 moduli) is VALIDITY: A == L*U on the pivotal rows (all rows when
`complete`), L lower triangular along fact->p (spasm_solve substitutes backwards over it), certificates verify.  The
factors themselves are not canonical (FFPACK's PLUQ is not vendored; oracle/ffpack_restate.c restates it).  Here the
same properties are checked on BASELINE-shaped inputs that reach the code paths no fixture reaches (several dense
blocks, several sparse Schur rounds, the sparse finisher, a 31-bit prime), and the canonical artefacts of the L-mode run
(rank, pivot columns, RREF, kernel) are compared with the reference's own sources run in L mode (oracle/_ref)."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle
import util
from spasm_b200 import abi, host, synthetic

pytestmark = pytest.mark.gpu


def vec_times_csr(x, M, p):
    """(x * M) mod p, exact: |x|, |M.x| < 2^31 so every product fits an int64 and is reduced before the sum"""
    rows = np.repeat(np.arange(M["n"]), np.diff(M["p"]))
    prod = (x.astype(np.int64)[rows] * M["x"].astype(np.int64)) % p
    out = np.zeros(M["m"], np.int64)
    np.add.at(out, M["j"], prod)
    return out % p


def check_factorization(product, t, complete, **opts):
    A = host.compress(product, t)
    oracle.reset_rand()
    o = host.default_opts(product, **opts)
    if complete:
        o.complete = True
    else:
        o.L = True
    f = host.echelonize(product, A, o)
    assert o.L and not o.enable_tall_and_skinny          # written back like the reference (echelonize.c:487-490)
    a, U, L, p, qinv = A.numpy(), f.U, f.L, f.p, f.qinv
    r, n, m, prime = f.rank, t.n, t.m, t.prime
    assert L is not None and L["n"] == n and L["m"] == r and len(p) == r
    assert bool(f.ptr.contents.complete) == bool(complete)
    util.check_echelon_form(U, qinv)
    assert len(set(p.tolist())) == r and ((p >= 0) & (p < n)).all()
    # the reference's probabilistic check, its three seeds (tests/lu.c:52-53,112)
    for seed in (1337, 21011984, 42):
        assert host.factorization_verify(product, A, f, seed)
    # x*A == (x*L)*U, exact, for a random x on the rows that must be factored
    rng = np.random.default_rng(7)
    x = rng.integers(0, prime, n)
    if not complete:
        mask = np.zeros(n, bool)
        mask[p] = True
        x = np.where(mask, x, 0)
    assert (vec_times_csr(x, a, prime) == vec_times_csr(vec_times_csr(x, L, prime), U, prime)).all()
    # L is lower triangular along p: row p[k] holds a non-zero entry on column k and nothing to its right
    rows_of = np.repeat(np.arange(n), np.diff(L["p"]))
    pos = np.full(n, -1, np.int64)
    pos[p] = np.arange(r)
    piv = pos[rows_of] >= 0
    assert (L["j"][piv] <= pos[rows_of][piv]).all()
    diag = np.zeros(r, bool)
    diag[L["j"][piv][L["j"][piv] == pos[rows_of][piv]]] = True
    assert diag.all()
    # spasm_solve: a right-hand side in the row space is solved, one outside is refused (tests/solve.c:52-88)
    y = rng.integers(0, prime, n)
    b = vec_times_csr(y, a, prime)
    b = np.where(b > prime // 2, b - prime, b).astype(np.int32)
    sol, ok = host.solve(product, f, b)
    assert ok
    back = vec_times_csr(sol, a, prime)
    assert (back == (b.astype(np.int64) % prime)).all()
    if r < m:
        bogus = rng.integers(1, prime, m)
        bogus = np.where(bogus > prime // 2, bogus - prime, bogus).astype(np.int32)
        _, ok = host.solve(product, f, bogus)
        assert not ok
    return f, A


CASES = [("config1", 0.1, {}), ("config2T", 0.03, {}), ("config4", 0.02, {}), ("config5", 0.05, {}),
         ("config3", 0.02, {"sparsity_threshold": 0.01}),
         ("config1", 0.04, {"sparsity_threshold": 2.0}),                                    # sparse rounds, then the sparse finisher
         ("config1", 0.06, {"sparsity_threshold": 2.0, "max_round": 6}),
         ("config4", 0.01, {"enable_dense": False}),                                        # the sparse finisher at once
         ("config1", 0.03, {"dense_block_size": 64})]                                       # many dense blocks


def _input(name, scale):
    return synthetic.config2(scale).transposed() if name == "config2T" else synthetic.CONFIGS[name](scale)


@pytest.mark.parametrize("complete", [False, True], ids=["L", "complete"])
@pytest.mark.parametrize("name,scale,opts", CASES, ids=[f"{c[0]}@{c[1]}" + "".join(f"-{k}={v}" for k, v in c[2].items()) for c in CASES])
def test_factorization_with_L(product, name, scale, opts, complete):
    t = _input(name, scale)
    f, A = check_factorization(product, t, complete, **opts)
    # canonical artefacts of the L-mode run against the reference's own sources in L mode
    Rm, _ = host.rref(product, f)
    Km = host.kernel(product, f)
    got = {"rank": f.rank, "pivot_columns": hashlib.sha256(oracle.pivot_columns(f.qinv).tobytes()).hexdigest(),
           "rref": oracle.canonical_hash(Rm.numpy()), "kernel": oracle.canonical_hash(Km.numpy()), "kernel_dim": Km.n}
    try:
        oracle.ref()
    except Exception:
        pytest.skip("oracle/_ref not built")
    want = util.run_reference(t, **opts, **({"complete": True} if complete else {"L": True}))
    assert got == want


@pytest.mark.parametrize("batch", ["5", "1024"])
def test_sparse_finisher_with_L_in_batches(product, monkeypatch, batch):
    """echelonize_GPLU with L (src/spasm_echelonize.c:119-141): no early abort, every row processed"""
    monkeypatch.setenv("SPASM_B200_GPLU_BATCH", batch)
    for t in (synthetic.config4(0.01), synthetic.config1(0.03)):
        check_factorization(product, t, True, enable_dense=False)


def test_rank_certificate_round_trip(product, tmp_path):
    """spasm_certificate_rank_create / _verify / _save / _load (src/spasm_certificate.c) on a rank-deficient input;
    a tampered certificate and a different matrix are refused."""
    t = synthetic.config4(0.01)
    A = host.compress(product, t)
    oracle.reset_rand()
    o = host.default_opts(product)
    o.L = True
    f = host.echelonize(product, A, o)
    digest = (C.c_uint8 * 32)(*hashlib.sha256(t.to_sms()).digest())
    proof = product.spasm_certificate_rank_create(A.ptr, digest, f.ptr)
    assert proof.contents.r == f.rank
    assert product.spasm_certificate_rank_verify(A.ptr, digest, proof)
    # text round trip
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    path = str(tmp_path / "proof.txt").encode()
    fh = libc.fopen(path, b"w")
    product.spasm_rank_certificate_save(proof, C.c_void_p(fh))
    libc.fclose(fh)
    loaded = abi.RankCertificate()
    fh = libc.fopen(path, b"r")
    assert product.spasm_rank_certificate_load(C.c_void_p(fh), C.byref(loaded))
    libc.fclose(fh)
    assert loaded.r == f.rank and loaded.prime == t.prime
    assert product.spasm_certificate_rank_verify(A.ptr, digest, C.byref(loaded))
    # tampering
    if f.rank > 0:
        loaded.x[0] = (loaded.x[0] + 1) if loaded.x[0] < t.prime // 2 else 0
        assert not product.spasm_certificate_rank_verify(A.ptr, digest, C.byref(loaded))
    t2 = synthetic.config4(0.01)
    t2.x = t2.x.copy()
    t2.x[0] += 1
    A2 = host.compress(product, t2)
    assert not product.spasm_certificate_rank_verify(A2.ptr, digest, proof)


@pytest.mark.parametrize("prime", [257, 42013, 2147483629])
def test_dense_LU_boundary(product, prime):
    """spasm_ffpack_LU (src/spasm_ffpack.cpp:88-96) in the layout its consumers read (echelonize.c:276-312,
    tests/dense_lu_ffpack.c:88-121): P*A*Q = L*U on a rank-deficient block"""
    rng = np.random.default_rng(prime)
    n, m, r0 = 150, 210, 97
    M = (rng.integers(0, prime, (n, r0)).astype(object) @ rng.integers(0, prime, (r0, m)).astype(object)) % prime
    M = np.array(M, dtype=np.int64)
    M[:, 5] = 0                                   # an empty column inside the profile
    bal = np.where(M > prime // 2, M - prime, M)
    datatype = product.spasm_datatype_choose(prime)
    np_type = {abi.SPASM_DOUBLE: np.float64, abi.SPASM_FLOAT: np.float32, abi.SPASM_I64: np.int64}[datatype]
    buf = np.ascontiguousarray(bal.astype(np_type))
    P = (C.c_size_t * n)()
    Q = (C.c_size_t * m)()
    r = product.spasm_ffpack_LU(prime, n, m, buf.ctypes.data_as(C.c_void_p), m, datatype, P, Q)
    assert r == r0
    P, Q = np.array(P[:]), np.array(Q[:])
    assert sorted(P.tolist()) == list(range(n)) and sorted(Q.tolist()) == list(range(m))
    out = buf.astype(np.int64).astype(object)
    Lm = np.zeros((n, r), object)
    Um = np.zeros((r, m), object)
    for i in range(n):
        for j in range(min(i + 1, r)):
            Lm[i, j] = out[i, j]
    for i in range(r):
        Um[i, i] = 1
        Um[i, i + 1:] = out[i, i + 1:]
    assert (((Lm @ Um) - M.astype(object)[P][:, Q]) % prime == 0).all()
    assert all(Lm[i, i] % prime != 0 for i in range(r))
