"""CPU tier: the host consumers of a factorization with L (csrc/host/solve.c: spasm_solve, spasm_gesv,
spasm_factorization_verify, rank certificates -- SURVEY.md 8f-1 / 8f-4) against the reference's OWN functions
(oracle/_ref = src/spasm_solve.c, src/spasm_certificate.c compiled where they lie).

The factorization is produced on the CPU by the reference (opts.L, restated FFPACK) -- struct layouts are the ABI, so
the very same `struct spasm_lu *` is handed to both libraries: solutions, certificates and the saved certificate text
must be identical, and each library must accept the other's certificate.  (On the GPU the factorization comes from
spasm_echelonize itself: tests/test_gpu_lu.py and the reference's test programs.)"""
import ctypes as C
import hashlib

import numpy as np
import pytest

import oracle
import util
from spasm_b200 import abi, host

FIXTURES = ["small", "medium", "singular", "rectangular_h", "rectangular_l", "mat364", "BIOMD0000000424.int.mpl", "G2", "cc", "dm2"]
PRIMES = [257, 42013, 4294967291]


def _libc():
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    return libc


def _fixture(name, prime):
    return util.golden_input({"kind": "fixture", "name": name, "prime": prime})


def _factor(R, t, **kw):
    A = host.compress(R, t)
    oracle.reset_rand()
    o = host.default_opts(R, **kw)
    o.L = True
    f = host.echelonize(R, A, o)
    return A, f


def _cert_fields(proof):
    c = proof.contents
    r = c.r
    take = lambda ptr: np.ctypeslib.as_array(ptr, shape=(max(r, 1),))[:r].copy()
    return r, int(c.prime), bytes(c.hash), take(c.i), take(c.j), take(c.x), take(c.y)


@pytest.mark.parametrize("prime", PRIMES)
@pytest.mark.parametrize("name", FIXTURES)
def test_consumers_match_the_reference(product, name, prime, tmp_path):
    R = oracle.ref()
    if R is None:
        pytest.skip("oracle/_ref not built")
    t = _fixture(name, prime)
    A, f = _factor(R, t)
    n, m, r = t.n, t.m, f.rank
    # spasm_factorization_verify: same verdict (true) for the reference's seeds
    for seed in (42, 1337, 21011984):
        assert R.spasm_factorization_verify(A.ptr, f.ptr, seed)
        assert product.spasm_factorization_verify(A.ptr, f.ptr, seed)
    # spasm_solve: same solution, same verdict, for a right-hand side in the row space and for a random one
    rng = np.random.default_rng(1)
    a = A.numpy()
    rows_of = np.repeat(np.arange(a["n"]), np.diff(a["p"]))
    y = rng.integers(-(prime // 2), prime // 2 + 1, max(n, 1))      # balanced: |y * x| < 2^62 for the 32-bit prime too
    b = np.zeros(max(m, 1), np.int64)
    np.add.at(b, a["j"], (y[rows_of] * a["x"].astype(np.int64)) % prime)
    b %= prime
    for rhs in (np.where(b > prime // 2, b - prime, b), np.where((z := rng.integers(0, prime, max(m, 1))) > prime // 2, z - prime, z)):
        rhs = np.ascontiguousarray(rhs, np.int32)
        x1 = np.zeros(max(n, 1), np.int32)
        x2 = np.zeros(max(n, 1), np.int32)
        ok1 = R.spasm_solve(f.ptr, rhs.ctypes.data_as(abi.i32_p), x1.ctypes.data_as(abi.i32_p))
        ok2 = product.spasm_solve(f.ptr, rhs.ctypes.data_as(abi.i32_p), x2.ctypes.data_as(abi.i32_p))
        assert bool(ok1) == bool(ok2)
        assert (x1 == x2).all()
    # spasm_gesv on A itself: same X (canonical: rows sorted by column), same ok flags
    ok1 = (C.c_bool * max(n, 1))()
    ok2 = (C.c_bool * max(n, 1))()
    X1 = host.CsrHandle(R, R.spasm_gesv(f.ptr, A.ptr, ok1))
    X2 = host.CsrHandle(product, product.spasm_gesv(f.ptr, A.ptr, ok2))
    assert list(ok1) == list(ok2)
    assert _rows(X1.numpy()) == _rows(X2.numpy())
    # certificates: identical, cross-accepted, identical text
    digest = (C.c_uint8 * 32)(*hashlib.sha256(t.to_sms()).digest())
    p1 = R.spasm_certificate_rank_create(A.ptr, digest, f.ptr)
    p2 = product.spasm_certificate_rank_create(A.ptr, digest, f.ptr)
    c1, c2 = _cert_fields(p1), _cert_fields(p2)
    assert c1[:3] == c2[:3]
    for u, v in zip(c1[3:], c2[3:]):
        assert (u == v).all()
    assert R.spasm_certificate_rank_verify(A.ptr, digest, p2) and product.spasm_certificate_rank_verify(A.ptr, digest, p1)
    libc = _libc()
    texts = []
    for lib, proof, tag in ((R, p1, "ref"), (product, p2, "b200")):
        path = str(tmp_path / f"{tag}.cert").encode()
        fh = libc.fopen(path, b"w")
        lib.spasm_rank_certificate_save(proof, C.c_void_p(fh))
        libc.fclose(fh)
        texts.append(open(path, "rb").read())
    assert texts[0] == texts[1]
    # our loader reads the file back into a certificate that verifies
    loaded = abi.RankCertificate()
    fh = libc.fopen(str(tmp_path / "ref.cert").encode(), b"r")
    assert product.spasm_rank_certificate_load(C.c_void_p(fh), C.byref(loaded))
    libc.fclose(fh)
    assert product.spasm_certificate_rank_verify(A.ptr, digest, C.byref(loaded))
    # a wrong certificate is refused by both
    if r > 0:
        loaded.y[0] = loaded.y[0] + 1 if loaded.y[0] < prime // 2 else 0
        assert not product.spasm_certificate_rank_verify(A.ptr, digest, C.byref(loaded))
        assert not R.spasm_certificate_rank_verify(A.ptr, digest, C.byref(loaded))


def _rows(M):
    out = []
    for i in range(M["n"]):
        lo, hi = M["p"][i], M["p"][i + 1]
        out.append(sorted(zip(M["j"][lo:hi].tolist(), M["x"][lo:hi].tolist())))
    return out


def test_solve_on_a_multi_stage_factorization(product):
    """a factorization with several sparse rounds and the sparse finisher (reference, CPU), consumed by our solve"""
    R = oracle.ref()
    if R is None:
        pytest.skip("oracle/_ref not built")
    from spasm_b200 import synthetic
    t = synthetic.config1(0.02)
    A, f = _factor(R, t, sparsity_threshold=2.0)
    assert product.spasm_factorization_verify(A.ptr, f.ptr, 7)
    rng = np.random.default_rng(3)
    prime = t.prime
    a = A.numpy()
    rows_of = np.repeat(np.arange(a["n"]), np.diff(a["p"]))
    y = rng.integers(0, prime, t.n)
    b = np.zeros(t.m, np.int64)
    np.add.at(b, a["j"], (y[rows_of] * a["x"].astype(np.int64)) % prime)
    b %= prime
    rhs = np.ascontiguousarray(np.where(b > prime // 2, b - prime, b), np.int32)
    x = np.zeros(t.n, np.int32)
    assert product.spasm_solve(f.ptr, rhs.ctypes.data_as(abi.i32_p), x.ctypes.data_as(abi.i32_p))
    back = np.zeros(t.m, np.int64)
    np.add.at(back, a["j"], (x.astype(np.int64)[rows_of] * a["x"].astype(np.int64)) % prime)
    assert ((back - b) % prime == 0).all()
