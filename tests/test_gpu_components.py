"""The individual entry points of the hot path against the oracle (reference: tests/schur.c, tests/schur_dense.c,
tests/dense_rref_ffpack.c), through the C ABI on a B200."""
import ctypes as C

import numpy as np
import pytest

import oracle
import util
from spasm_b200 import abi, host, synthetic

pytestmark = pytest.mark.gpu


def _inputs():
    return [synthetic.config1(0.05), synthetic.config2(0.02).transposed(), synthetic.config4(0.01), synthetic.config5(0.04),
            synthetic.config3(0.01)]


def _oracle_pivots(t):
    A = oracle.compress(t)
    pinv = np.zeros(max(t.n, 1), np.int32)
    qinv = np.zeros(max(t.m, 1), np.int32)
    counts = (C.c_int * 3)()
    edges = C.c_int64()
    oracle.lib().oracle_pivots_find(A.ptr, 1, oracle._ip(pinv), oracle._ip(qinv), counts, C.byref(edges))
    return A, pinv[:t.n], qinv[:t.m], list(counts)


@pytest.mark.parametrize("t", _inputs(), ids=lambda t: t.name)
def test_pivots_extract_structural(product, t):
    """same (row, column) pairs as the sequential reference; U rows are the pivotal rows scaled to a unit pivot"""
    Ao, pinv, qinv, counts = _oracle_pivots(t)
    A = host.compress(product, t)
    npiv, p, fact = host.pivots_extract_structural(product, A)
    assert npiv == sum(counts)
    rows, cols = host.fact_pairs(fact, p, npiv)
    assert set(zip(rows.tolist(), cols.tolist())) == {(int(qinv[j]), j) for j in range(t.m) if qinv[j] >= 0}
    assert sorted(p.tolist()) == list(range(t.n))
    assert (np.diff(p[npiv:]) > 0).all()                       # non-pivotal rows by increasing index (pivots.c:353-357)
    U = abi.csr_to_numpy(fact.U)
    a = A.numpy()
    for k in range(0, npiv, max(1, npiv // 50)):
        i, j = int(rows[k]), int(cols[k])
        ra = dict(zip(a["j"][a["p"][i]:a["p"][i + 1]].tolist(), a["x"][a["p"][i]:a["p"][i + 1]].tolist()))
        ru = dict(zip(U["j"][U["p"][k]:U["p"][k + 1]].tolist(), U["x"][U["p"][k]:U["p"][k + 1]].tolist()))
        assert U["j"][U["p"][k]] == j and ru[j] == 1
        inv = pow(ra[j] % t.prime, -1, t.prime)
        assert all((ru[c] - ra[c] * inv) % t.prime == 0 for c in ra)
    # U is triangular in its row order: an entry on a pivotal column belongs to a later row
    Uq = np.ctypeslib.as_array(fact.qinv, shape=(t.m,))
    owner = Uq[U["j"]]
    row_of = np.repeat(np.arange(U["n"]), np.diff(U["p"]))
    first = np.zeros(len(U["j"]), bool)
    first[U["p"][:-1]] = True
    assert ((owner < 0) | first | (owner > row_of)).all()
    product.spasm_csr_free(fact.U)


@pytest.mark.parametrize("t", _inputs(), ids=lambda t: t.name)
def test_schur_dense_and_density_and_sparse(product, t):
    Ao, pinv, qinv, counts = _oracle_pivots(t)
    A = host.compress(product, t)
    npiv, p, fact = host.pivots_extract_structural(product, A)
    U = abi.csr_to_numpy(fact.U)
    Uq = np.ctypeslib.as_array(fact.qinv, shape=(t.m,)).copy()
    Uo = oracle.from_numpy(U)
    rows = np.ascontiguousarray(p[npiv:npiv + 200], np.int32)
    n = len(rows)
    Sm = t.m - npiv
    # dense block (reference: tests/schur_dense.c)
    S, q = host.schur_dense(product, A, rows, n, fact)
    So = np.zeros((max(n, 1), max(Sm, 1)), np.int32)
    qo = np.zeros(max(t.m, 1), np.int32)
    oracle.lib().oracle_schur_dense(Ao.ptr, oracle._ip(rows), n, Uo.ptr, oracle._ip(Uq), So.ctypes.data_as(C.POINTER(C.c_int32)), oracle._ip(qo))
    assert np.array_equal(q, qo[:Sm])
    assert np.array_equal(S, So[:n, :Sm].astype(np.int64))
    # density estimate: same rand() draws, same value (reference: src/spasm_schur.c:11-44)
    rest = np.ascontiguousarray(p[npiv:], np.int32)
    if len(rest):
        oracle.reset_rand()
        d_gpu = product.spasm_schur_estimate_density(A.ptr, abi.as_int_p(rest), len(rest), fact.U, fact.qinv, 100)
        oracle.reset_rand()
        d_cpu = oracle.lib().oracle_schur_estimate_density(Ao.ptr, oracle._ip(rest), len(rest), Uo.ptr, oracle._ip(Uq), 100)
        assert d_gpu == d_cpu
    # sparse Schur complement (reference: tests/schur.c): same rows, same entries IN THE SAME ORDER (DFS pattern order)
    p_out = np.zeros(max(n, 1), np.int32)
    Sg = host.CsrHandle(product, product.spasm_schur(A.ptr, abi.as_int_p(rows), n, C.byref(fact), 1.0, None, None, abi.as_int_p(p_out))).numpy()
    Sc = oracle.Matrix(oracle.lib().oracle_schur(Ao.ptr, oracle._ip(rows), n, Uo.ptr, oracle._ip(Uq), 1.0)).numpy()
    assert np.array_equal(Sg["p"], Sc["p"]) and np.array_equal(p_out[:n], rows)
    assert np.array_equal(Sg["j"], Sc["j"]) and np.array_equal(Sg["x"], Sc["x"])
    assert (Uq[Sg["j"]] < 0).all()                             # no entry of S under a pivot (tests/schur.c:60-70)
    # randomized block: same rand() + PRNG stream -> identical block (reference: src/spasm_schur.c:346-413)
    for N, w in ((17, 5), (4, 0)):
        dt = product.spasm_datatype_choose(t.prime)
        npdt = {abi.SPASM_DOUBLE: np.float64, abi.SPASM_FLOAT: np.float32, abi.SPASM_I64: np.int64}[dt]
        Sr = np.zeros((N, max(Sm, 1)), npdt)
        q2 = np.zeros(max(t.m, 1), np.int32)
        oracle.reset_rand()
        product.spasm_schur_dense_randomized(A.ptr, abi.as_int_p(rest), len(rest), fact.U, fact.qinv, Sr.ctypes.data_as(C.c_void_p), dt,
                                             abi.as_int_p(q2), N, w)
        Sro = np.zeros((N, max(Sm, 1)), np.int32)
        oracle.reset_rand()
        oracle.lib().oracle_schur_dense_randomized(Ao.ptr, oracle._ip(rest), len(rest), Uo.ptr, oracle._ip(Uq),
                                                   Sro.ctypes.data_as(C.POINTER(C.c_int32)), oracle._ip(qo), N, w)
        assert np.array_equal(Sr[:, :Sm].astype(np.int64), Sro[:, :Sm].astype(np.int64))
    product.spasm_csr_free(fact.U)


@pytest.mark.parametrize("prime", [3, 257, 42013, 189812507, 2147483629, 4294967291])
@pytest.mark.parametrize("shape", [(1, 1), (5, 9), (40, 40), (130, 70), (64, 300), (257, 129)])
def test_dense_rref_matches_oracle(product, prime, shape):
    """spasm_ffpack_rref boundary (reference: src/spasm_ffpack.cpp:78-86, tests/dense_rref_ffpack.c): rank, column rank
    profile and the packed RREF values, against the restated FFPACK of the oracle; rank-deficient inputs included."""
    n, m = shape
    rng = np.random.default_rng(prime % 1000 + n * m)
    half = prime // 2
    r = max(1, min(n, m) * 2 // 3)
    Lf = rng.integers(-half, half + 1, size=(n, r)).astype(object)
    Rf = rng.integers(-half, half + 1, size=(r, m)).astype(object)
    Rf[:, : m // 4] = 0                                        # leading zero columns: pivots do not start at column 0
    M = (Lf.dot(Rf)) % prime
    M = np.where(M > half, M - prime, M).astype(np.int64)
    rank, qinv, packed = host.ffpack_rref(product, prime, M)
    W = np.ascontiguousarray(M, np.int32)
    pivcol = np.zeros(max(m, 1), np.int32)
    ro = oracle.lib().oracle_dense_rref_i32(prime, n, m, W.ctypes.data_as(C.POINTER(C.c_int32)), oracle._ip(pivcol))
    assert rank == ro
    assert np.array_equal(qinv[:rank], pivcol[:rank])
    assert sorted(qinv.tolist()) == list(range(m))
    for i in range(rank):
        for k in range(rank, m):
            assert packed[i, k] == W[i, qinv[k]]


def test_device_prng_known_answers(product):
    """the SHA-256 counter-mode streams generated on the device (prng.cu) against tests/Expected/prng of the reference"""
    from test_oracle_pinned import PRNG_KAT
    for prime, seed, seq, want in PRNG_KAT:
        out = np.zeros(10, np.int32)
        product.spasm_b200_prng_stream(prime, seed, seq, 10, abi.as_i32_p(out))
        assert out.tolist() == want
    # long streams against the oracle's implementation, including a 32-bit prime (rejection sampling path)
    for prime in (3, 42013, 2147483629, 4294967291):
        out = np.zeros(3000, np.int32)
        product.spasm_b200_prng_stream(prime, 12345, 0, 3000, abi.as_i32_p(out))
        ref = (C.c_int32 * 3000)()
        oracle.lib().oracle_prng_stream(prime, 12345, 0, 3000, ref)
        assert out.tolist() == list(ref)


@pytest.mark.parametrize("prime", [3, 257, 42013, 65537, 189812507, 2147483629, 4278124287])
@pytest.mark.parametrize("shape", [(128, 128, 64), (200, 300, 100), (1000, 517, 1000), (77, 1031, 33), (130, 64, 4100)])
def test_modular_product_tensor_cores_vs_cuda_cores_vs_exact(product, prime, shape):
    """C -= A*B mod p: the tcgen05 int8 limb-split kernel, the CUDA-core kernel and exact Python integers agree
    bit for bit (1 limb for p=3, 2 for 42013, 3 for 65537, 4 for the 31/32-bit primes; ragged tile edges)."""
    M, N, K = shape
    rng = np.random.default_rng(prime % 997 + M + N + K)
    half, mhalf = prime // 2, prime // 2 - prime + 1
    A = rng.integers(mhalf, half + 1, size=(M, K), dtype=np.int64)
    B = rng.integers(mhalf, half + 1, size=(K, N), dtype=np.int64)
    C0 = rng.integers(mhalf, half + 1, size=(M, N), dtype=np.int64)
    if K > 40:                                   # extreme values exercise the limb carries
        A[0, :] = half
        B[:, 0] = mhalf
        A[1, :] = mhalf
    exact = None
    if M * N * K <= 8_000_000:               # exact Python integers are slow: big shapes compare the two kernels only
        exact = (C0.astype(object) - A.astype(object).dot(B.astype(object))) % prime
        exact = np.where(exact > half, exact - prime, exact).astype(np.int64)
    out = {}
    for use_tensor in (0, 1):
        Cc = np.ascontiguousarray(C0, np.int32)
        product.spasm_b200_gemm_sub(prime, M, N, K, abi.as_i32_p(Cc), abi.as_i32_p(np.ascontiguousarray(A, np.int32)),
                                    abi.as_i32_p(np.ascontiguousarray(B, np.int32)), use_tensor)
        out[use_tensor] = Cc.astype(np.int64)
    if exact is not None:
        assert np.array_equal(out[0], exact), "CUDA-core product differs from exact arithmetic"
    assert np.array_equal(out[1], out[0]), "tensor-core product differs from the CUDA-core product"


def test_out_of_order_greedy_search_is_deterministic_under_repetition():
    """The out-of-order greedy kernel must pick the pivots of the ordered kernel (which follows the reference row by
    row) every time.  Regression test for a barrier divergence (a warp skipped the collection of the pending rows and
    its barriers when thread 0 had already reset the shared flag) that showed up about once per 50 runs on wide
    matrices with 8 resident CTAs per SM.  Runs in a fresh process because the shadow check is an environment knob:
    SPASM_B200_GREEDY_SHADOW=1 makes the library run both kernels and abort if they differ."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPASM_B200_GREEDY_SHADOW="1", C3SCALE="1.0")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "pivot_stress.py"), "c3", "120"], env=env, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l and l[0].isdigit()]
    assert len(lines) == 120 and all(l.endswith("same") for l in lines), out.stdout[-2000:]


def _echelon_arrays(product, A, **kw):
    oracle.reset_rand()
    f = host.echelonize(product, A, host.default_opts(product, **kw))
    U = f.U
    return f, (f.rank, U["p"].tobytes(), U["j"].tobytes(), U["x"].tobytes(), f.qinv.tobytes())


def test_staged_download_of_large_results_is_transparent(product, monkeypatch):
    """Results above 2 MB leave the device through pinned staging buffers and several copy threads (download_bulk);
    forcing that path on a small result (threshold 4 KB, chunks still 32 MB) and forbidding it must give the same
    arrays, for the echelon form, the RREF and the kernel basis."""
    t = synthetic.config3(0.02)
    A = host.compress(product, t)
    out = {}
    for tag, mb in (("plain", "1000000"), ("staged", "0.004")):
        monkeypatch.setenv("SPASM_B200_BULK_MB", mb)
        f, arrays = _echelon_arrays(product, A, sparsity_threshold=0.01)
        R, _ = host.rref(product, f)
        K = host.kernel(product, f)
        Rn, Kn = R.numpy(), K.numpy()
        out[tag] = (arrays, Rn["p"].tobytes(), Rn["j"].tobytes(), Rn["x"].tobytes(), Kn["p"].tobytes(), Kn["j"].tobytes(), Kn["x"].tobytes())
    assert out["plain"] == out["staged"]


def test_masked_solve_gives_the_same_rref(product, monkeypatch):
    """spasm_rref through the occupancy-masked dataflow solve (opt-in, SPASM_B200_MASKED_SOLVE=1) and through the
    plain solve: identical matrices, entry order included."""
    for t, kw in ((synthetic.config1(0.1), {}), (synthetic.config4(0.01), {}), (synthetic.config2(0.02).transposed(), {})):
        A = host.compress(product, t)
        f, _ = _echelon_arrays(product, A, **kw)
        res = {}
        for tag in ("plain", "masked"):
            if tag == "masked":
                monkeypatch.setenv("SPASM_B200_MASKED_SOLVE", "1")
            else:
                monkeypatch.delenv("SPASM_B200_MASKED_SOLVE", raising=False)
            R, Rqinv = host.rref(product, f)
            Rn = R.numpy()
            res[tag] = (Rn["p"].tobytes(), Rn["j"].tobytes(), Rn["x"].tobytes(), np.asarray(Rqinv).tobytes())
        monkeypatch.delenv("SPASM_B200_MASKED_SOLVE", raising=False)
        assert res["plain"] == res["masked"]
