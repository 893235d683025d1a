"""CPU-side checks of the product: the shared library loads and exports every symbol include/*.h declares, the
host-side L1 code (field, sha256, prng, compress, I/O) behaves like the reference, and nothing under spasm_b200/
touches the oracle.  No GPU compute call is made here."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle
import spasm_b200
import util
from spasm_b200 import abi, host, synthetic
from test_oracle_pinned import GFP_PRIMES, PRNG_KAT, SHA_KAT

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("spasm.h", "spasm_b200.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(spasm_[A-Za-z0-9_]+)\s*\(", text):
            names.add(m.group(1))
    # static inline helpers are not exported
    return sorted(names - {"spasm_get_prime", "spasm_max", "spasm_min", "spasm_row_weight"})


def test_library_exports_every_declared_symbol():
    L = C.CDLL(spasm_b200.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"not exported: {missing}"
    assert len(declared_symbols()) > 60


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", spasm_b200.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_missing_extension_fails_loudly(monkeypatch):
    monkeypatch.setattr(spasm_b200, "_lib", None)
    monkeypatch.setattr(spasm_b200, "LIB_PATH", "/nonexistent/libspasm_b200.so")
    with pytest.raises(spasm_b200.MissingExtension):
        spasm_b200.lib()


def test_product_never_touches_the_oracle():
    """The product path must not import, link or execute anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "spasm_b200")):
        for fn in files:
            if fn.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, fn), errors="ignore").read()
                for ln in text.splitlines():
                    if re.search(r"^\s*(import|from)\s+oracle\b|liboracle|oracle/_ref|libspasm_ref", ln) and "test" not in ln and '"""' not in ln and "#" not in ln.split("oracle")[0] and "--" not in ln:
                        bad.append((fn, ln.strip()))
    assert not bad, bad
    out = subprocess.run(["ldd", spasm_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "spasm_ref" not in out


@pytest.mark.parametrize("prime,seed,seq,want", PRNG_KAT)
def test_host_prng_known_answers(product, prime, seed, seq, want):
    ctx = abi.PrngCtx()
    product.spasm_prng_seed_simple(prime, seed, seq, C.byref(ctx))
    assert [product.spasm_prng_ZZp(C.byref(ctx)) for _ in range(10)] == want


@pytest.mark.parametrize("msg,want", SHA_KAT)
def test_host_sha256_known_answers(product, msg, want):
    ctx = abi.Sha256Ctx()
    product.spasm_SHA256_init(C.byref(ctx))
    for k in range(0, len(msg), 7):              # uneven chunks exercise the pending-block logic
        product.spasm_SHA256_update(C.byref(ctx), msg[k:k + 7], len(msg[k:k + 7]))
    md = (C.c_ubyte * 32)()
    product.spasm_SHA256_final(md, C.byref(ctx))
    assert bytes(md).hex() == want


@pytest.mark.parametrize("p", GFP_PRIMES)
def test_host_field_matches_oracle(product, p):
    F = (abi.Field * 1)()
    product.spasm_field_init(p, F)
    O = oracle.lib()
    rng = np.random.default_rng(p + 1)
    half, mhalf = p // 2, p // 2 - p + 1
    for _ in range(400):
        a, x, y = (int(v) for v in rng.integers(mhalf, half + 1, size=3))
        assert product.spasm_ZZp_mul(F, a, x) == O.oracle_zp_mul(p, a, x)
        assert product.spasm_ZZp_axpy(F, a, x, y) == O.oracle_zp_axpy(p, a, x, y)
        if a % p:
            assert product.spasm_ZZp_inverse(F, a) == O.oracle_zp_inverse(p, a)
        assert (product.spasm_ZZp_add(F, a, x) - (a + x)) % p == 0 and mhalf <= product.spasm_ZZp_add(F, a, x) <= half
        assert (product.spasm_ZZp_sub(F, a, x) - (a - x)) % p == 0


def _same(a, b):
    return a["n"] == b["n"] and a["m"] == b["m"] and all(np.array_equal(a[k], b[k]) for k in "pjx")


@pytest.mark.parametrize("case", [c for c in util.golden_cases() if c["kind"] == "fixture" and c["prime"] in (3, 42013, 4294967291)],
                         ids=util.case_id)
def test_compress_matches_oracle_on_fixtures(product, case):
    t = util.golden_input(case)
    assert _same(host.compress(product, t).numpy(), oracle.compress(t).numpy())


def test_compress_quirk_and_duplicates(product):
    t = synthetic.uniform_rows(300, 40, 9, seed=5, values="small", distinct=False)     # many repeated, cancelling entries
    assert _same(host.compress(product, t).numpy(), oracle.compress(t).numpy())
    t = synthetic.config1(0.01)
    assert _same(host.compress(product, t).numpy(), oracle.compress(t).numpy())


def test_sms_io_roundtrip(product, tmp_path):
    """spasm_triplet_load / spasm_csr_save (reference: src/spasm_io.c) through real FILE* handles."""
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    t = synthetic.config1(0.01)
    src = tmp_path / "a.sms"
    src.write_bytes(t.to_sms())
    f = libc.fopen(str(src).encode(), b"r")
    digest = (C.c_ubyte * 32)()
    T = product.spasm_triplet_load(f, t.prime, digest)
    libc.fclose(f)
    import hashlib
    assert bytes(digest).hex() == hashlib.sha256(t.to_sms()).hexdigest()
    A = host.CsrHandle(product, product.spasm_compress(T))
    product.spasm_triplet_free(T)
    assert _same(A.numpy(), oracle.compress(t).numpy())
    dst = tmp_path / "b.sms"
    f = libc.fopen(str(dst).encode(), b"w")
    product.spasm_csr_save(A.ptr, f)
    libc.fclose(f)
    back = util.load_sms(str(dst), t.prime)
    assert _same(host.compress(product, back).numpy(), A.numpy())
    # MatrixMarket input (reference: src/spasm_io.c:28-52, :80-98)
    mm = tmp_path / "c.mtx"
    a = A.numpy()
    rows = np.repeat(np.arange(a["n"]), np.diff(a["p"]))
    with open(mm, "w") as g:
        g.write("%%MatrixMarket matrix coordinate integer general\n% comment\n")
        g.write(f'{a["n"]} {a["m"]} {len(a["j"])}\n')
        for i, j, x in zip(rows, a["j"], a["x"]):
            g.write(f"{i + 1} {j + 1} {x}\n")
    f = libc.fopen(str(mm).encode(), b"r")
    T = product.spasm_triplet_load(f, t.prime, None)
    libc.fclose(f)
    B = host.CsrHandle(product, product.spasm_compress(T))
    product.spasm_triplet_free(T)
    assert _same(B.numpy(), a)


def test_transpose_and_human_format(product):
    t = synthetic.config1(0.01)
    A = host.compress(product, t)
    At = host.transpose(product, A)
    Ao = oracle.compress(t)            # keep it alive across the call
    want = oracle.Matrix(oracle.lib().oracle_transpose(Ao.ptr)).numpy()
    assert _same(At.numpy(), want)
    buf = C.create_string_buffer(16)
    for n, s in ((999, b"999"), (1000, b"1.0k"), (1234567, b"1.2m"), (3 * 10**9, b"3.0g"), (5 * 10**13, b"50.0t")):
        product.spasm_human_format(n, buf)
        assert buf.value == s


def test_datatype_helpers(product):
    assert product.spasm_datatype_choose(8191) == abi.SPASM_FLOAT
    assert product.spasm_datatype_choose(42013) == abi.SPASM_DOUBLE
    assert product.spasm_datatype_choose(189812531) == abi.SPASM_DOUBLE
    assert product.spasm_datatype_choose(2147483629) == abi.SPASM_I64
    assert [product.spasm_datatype_size(k) for k in (abi.SPASM_DOUBLE, abi.SPASM_FLOAT, abi.SPASM_I64)] == [8, 4, 8]
    assert product.spasm_datatype_name(abi.SPASM_I64) == b"i64"


def test_default_options_match_reference(product):
    """reference: src/spasm_echelonize.c:9-28"""
    o = host.default_opts(product)
    assert (o.enable_greedy_pivot_search, o.enable_tall_and_skinny, o.enable_dense, o.enable_GPLU, o.L, o.complete) == (True, True, True, True, False, False)
    assert (o.min_pivot_proportion, o.max_round, o.sparsity_threshold, o.tall_and_skinny_ratio) == (0.1, 3, 0.05, 5)
    assert (o.dense_block_size, o.low_rank_ratio, o.low_rank_start_weight) == (1000, 0.5, -1)


def test_tensor_core_path_is_really_tcgen05():
    """SASS evidence (B200_PROFILING.md): tcgen05.mma -> UTC*MMA (UTCIMMA for kind::i8), tcgen05.ld -> LDTM,
    tcgen05.commit -> UTCBAR, cp.async.bulk (the operand pipeline of the packed kernel) -> UBLKCP in the built library;
    no legacy HMMA/IMMA tensor path."""
    sass = subprocess.run(["cuobjdump", "-sass", spasm_b200.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCIMMA" in sass, "no tcgen05.mma kind::i8 in the SASS"
    assert "LDTM" in sass and "UTCBAR" in sass
    assert "UBLKCP" in sass, "no bulk-copy (TMA engine) operand pipeline in the SASS"
    assert " IMMA." not in sass and " HMMA." not in sass


def test_bench_cli_and_clock_sampler_work_without_a_gpu():
    """bench.py must at least parse, print its usage, and its clock sampler must degrade to "no samples" (not crash)
    on a machine without NVML / nvidia-smi: the driver imports and launches it unconditionally."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--help"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "--impl" in out.stdout and "--gpus" in out.stdout
    code = ("import sys; sys.path.insert(0, %r); import bench; s = bench.ClockSampler(0); s.sample(); d = s.summary(); "
            "assert set(d) >= {'sm_mhz', 'sm_max_mhz', 'reasons', 'samples'}; print('ok')" % root)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-1000:]


def test_sha256_matches_hashlib_for_every_length_and_chunking(product):
    """The SHA-256 of the library (portable code or the SHA-NI path picked at run time) against hashlib: every length up
    to 200 bytes (all padding cases), block boundaries, and random splits of the input over several update calls."""
    import hashlib
    import random
    product.spasm_SHA256_init.argtypes = [C.c_void_p]
    product.spasm_SHA256_update.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    product.spasm_SHA256_final.argtypes = [C.c_void_p, C.c_void_p]
    rnd = random.Random(7)
    for n in list(range(0, 200)) + [1000, 4095, 4096, 4097, 70001]:
        data = bytes(rnd.getrandbits(8) for _ in range(n))
        ctx = C.create_string_buffer(256)
        product.spasm_SHA256_init(ctx)
        at = 0
        while at < n:
            k = rnd.randint(1, 130)
            product.spasm_SHA256_update(ctx, data[at:at + k], len(data[at:at + k]))
            at += k
        out = C.create_string_buffer(32)
        product.spasm_SHA256_final(out, ctx)
        assert out.raw == hashlib.sha256(data).digest(), n


def test_host_verifiers_equal_the_reference_row_by_row():
    """csrc/host/verify.c against the reference's own functions (oracle/_ref) on the same echelon form: the one-row
    sparse triangular solve must return the same pattern IN THE SAME ORDER (reach = DFS post-order, src/spasm_reach.c)
    and the same values; permutations, sub-matrix and kernel_from_rref the same arrays."""
    import oracle
    R_ = oracle.ref()
    if R_ is None:
        pytest.skip("oracle/_ref not built")
    from spasm_b200 import abi, host, synthetic
    L = spasm_b200.lib()
    t = synthetic.config1(0.02)
    e = oracle.echelonize(oracle.compress(t), oracle.default_opts())
    U, qinv = e.U, np.ascontiguousarray(e.qinv, np.int32)
    m = U["m"]
    out = {}
    for tag, lib_ in (("ref", R_), ("b200", L)):
        Uh, Ah = host.from_numpy(lib_, U), host.compress(lib_, t)
        rows = []
        xj = np.zeros(3 * m, np.int32)
        x = np.zeros(m, np.int32)
        for k in range(0, Ah.n, 7):
            top = lib_.spasm_sparse_triangular_solve(Uh.ptr, Ah.ptr, k, abi.as_int_p(xj), abi.as_i32_p(x), abi.as_int_p(qinv))
            pat = xj[top:m].copy()
            rows.append((top, pat.tobytes(), x[pat].tobytes()))
        perm = np.random.default_rng(5).permutation(Ah.n).astype(np.int32)
        pinv_ptr = lib_.spasm_pinv(abi.as_int_p(perm), Ah.n)
        pinv = np.ctypeslib.as_array(pinv_ptr, shape=(Ah.n,)).copy()
        Pm = host.CsrHandle(lib_, lib_.spasm_permute(Ah.ptr, abi.as_int_p(perm), None, 1)).numpy()
        Sm = host.CsrHandle(lib_, lib_.spasm_submatrix(Ah.ptr, 3, Ah.n // 2, 5, Ah.m - 7, 1)).numpy()
        out[tag] = (rows, pinv.tobytes(), tuple(Pm[k].tobytes() for k in "pjx"), tuple(Sm[k].tobytes() for k in "pjx"))
    assert out["ref"] == out["b200"]
