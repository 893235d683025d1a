"""The oracle (oracle/oracle.c) pinned against everything the reference provides for this path:
  * its known-answer vectors (tests/Expected/prng, tests/Expected/hash; values copied below with their source),
  * the arithmetic properties of tests/GFp.c,
  * the REFERENCE ITSELF on the reference's fixture matrices and on small BASELINE configs:
      - committed golden artefacts produced by oracle/_ref (tests/golden/make_golden.py),
      - a live raw comparison (U, qinv, R, K array for array) when oracle/_ref is present.
CPU only."""
import ctypes as C
import os

import numpy as np
import pytest

import util
import oracle
from spasm_b200 import host, synthetic

CASES = util.golden_cases()

# reference: tests/Expected/prng (prime, seed, seq -> first 10 outputs of spasm_prng_ZZp)
PRNG_KAT = [
    (257, 0, 0, [8, -73, -2, 43, -53, 50, 5, -113, 36, -94]),
    (257, 0, 1, [-17, 114, 108, -21, 35, 16, -83, 101, -21, -98]),
    (257, 1, 0, [80, -66, 0, 34, -103, 35, -84, -40, 31, -33]),
    (257, 1, 1, [62, -79, 98, 100, -28, 119, -22, -48, -105, 46]),
    (65537, 0xdead00000000beef, 0, [-27920, 12457, 2164, -12603, 4399, 6513, -15681, -222, -8152, 8565]),
]
# reference: tests/Expected/hash (messages of tests/sha.c:20-23)
SHA_KAT = [
    (b"", "e3b0c44298fc1c149afbf4c8996fb92427ae41e4649b934ca495991b7852b855"),
    (b"X", "4b68ab3847feda7d6c62c1fbcbeebfa35eab7351ed5e78f4ddadea5df64b8015"),
    (b"Hello World", "a591a6d40bf420404a011733cfb7b190d62c65bf0bcda32b57b277d9ad9f146e"),
    (b"abcdefghijklmnopqrstuvwxyz0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZ+-*/=", "1bd9ce49f8515f8f5bf9ea86c6ddfb52084a8242600f39a449607a8d9c69adec"),
]
# reference: tests/GFp.c:59-69
GFP_PRIMES = [2, 3, 257, 65537, 67108859, 189812507, 0x7fffffff, 3037000493, 0xfffffffb]


@pytest.mark.parametrize("prime,seed,seq,want", PRNG_KAT)
def test_prng_known_answers(prime, seed, seq, want):
    out = (C.c_int32 * 10)()
    oracle.lib().oracle_prng_stream(prime, seed, seq, 10, out)
    assert list(out) == want


@pytest.mark.parametrize("msg,want", SHA_KAT)
def test_sha256_known_answers(msg, want):
    h = (C.c_ubyte * 32)()
    oracle.lib().oracle_sha256(msg, len(msg), h)
    assert bytes(h).hex() == want


@pytest.mark.parametrize("p", GFP_PRIMES)
def test_field_properties(p):
    """reference: tests/GFp.c -- x * x^-1 == 1, results stay in the balanced range, axpy is a*x+y."""
    L = oracle.lib()
    rng = np.random.default_rng(p)
    half, mhalf = p // 2, p // 2 - p + 1
    xs = list(range(1, min(p, 300))) if p < 70000 else []
    xs += [int(v) for v in rng.integers(1, p, size=300)]
    for x in xs:
        xb = x - p if x > half else x
        inv = L.oracle_zp_inverse(p, xb)
        assert mhalf <= inv <= half
        assert (inv * xb - 1) % p == 0
        a = int(rng.integers(mhalf, half + 1))
        y = int(rng.integers(mhalf, half + 1))
        r = L.oracle_zp_axpy(p, a, xb, y)
        assert mhalf <= r <= half and (r - (a * xb + y)) % p == 0


@pytest.mark.parametrize("case", CASES, ids=util.case_id)
def test_oracle_matches_reference_golden(case):
    if case["expected"] is None:
        pytest.skip("empty matrix")
    got = util.run_oracle(util.golden_input(case), **case["opts"])
    util.assert_same(got, case["expected"], keys=("rank", "pivot_columns", "rref", "kernel", "kernel_dim"), what=util.case_id(case))


def _same_csr(a, b):
    return a["n"] == b["n"] and a["m"] == b["m"] and all(np.array_equal(a[k], b[k]) for k in "pjx")


LIVE = [c for c in CASES if c["expected"] is not None and (c["kind"] == "synthetic" or c["prime"] in (257, 42013, 4294967291))]


@pytest.mark.skipif(not os.path.exists(oracle.ref_path()), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", LIVE, ids=util.case_id)
def test_oracle_raw_equal_to_live_reference(case):
    """Array-for-array: compress, U, qinv, rref, kernel of oracle.c == the reference compiled in place (1 thread)."""
    R = oracle.ref()
    t = util.golden_input(case)
    Ao, Ar = oracle.compress(t), host.compress(R, t)
    assert _same_csr(Ao.numpy(), Ar.numpy())
    eo = oracle.echelonize(Ao, oracle.default_opts(**case["opts"]))
    oracle.reset_rand()
    fr = host.echelonize(R, Ar, host.default_opts(R, **case["opts"]))
    assert _same_csr(eo.U, fr.U) and np.array_equal(eo.qinv, fr.qinv)
    Ro, Rqo = oracle.rref(eo.U, eo.qinv)
    Rr, Rqr = host.rref(R, fr)
    assert _same_csr(Ro, Rr.numpy()) and np.array_equal(Rqo, Rqr)
    assert _same_csr(oracle.kernel(eo.U, eo.qinv), host.kernel(R, fr).numpy())


@pytest.mark.skipif(not os.path.exists(oracle.ref_path()), reason="oracle/_ref not built")
@pytest.mark.parametrize("opts", [dict(sparsity_threshold=2.0), dict(sparsity_threshold=2.0, max_round=6),
                                  dict(enable_dense=False, enable_tall_and_skinny=False)], ids=["sparse-rounds", "six-rounds", "gplu"])
def test_oracle_multi_round_and_gplu_paths(opts):
    """Second pivot rounds on a non-empty sparse Schur complement and the GPLU finisher: no fixture reaches them
    (SURVEY section 4), so they are forced through the options."""
    R = oracle.ref()
    for t in (synthetic.config1(0.02), synthetic.config4(0.01)):
        Ao, Ar = oracle.compress(t), host.compress(R, t)
        eo = oracle.echelonize(Ao, oracle.default_opts(**opts))
        oracle.reset_rand()
        fr = host.echelonize(R, Ar, host.default_opts(R, **opts))
        assert _same_csr(eo.U, fr.U) and np.array_equal(eo.qinv, fr.qinv)


def test_quirks_with_cancelling_duplicates():
    """Repeated (i,j) entries that cancel mod p: spasm_compress re-reads stale slots (src/spasm_triplet.c:36-57) and
    the greedy search may pick the last entry of a row (src/spasm_pivots.c:232-238).  The oracle reproduces both."""
    if not os.path.exists(oracle.ref_path()):
        pytest.skip("oracle/_ref not built")
    R = oracle.ref()
    t = synthetic.uniform_rows(1714, 473, 11, seed=20240302, values="small", distinct=False)
    Ao, Ar = oracle.compress(t), host.compress(R, t)
    assert _same_csr(Ao.numpy(), Ar.numpy())
    eo = oracle.echelonize(Ao, oracle.default_opts(sparsity_threshold=0.01))
    oracle.reset_rand()
    fr = host.echelonize(R, Ar, host.default_opts(R, sparsity_threshold=0.01))
    assert _same_csr(eo.U, fr.U)
