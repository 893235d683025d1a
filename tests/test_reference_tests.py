"""The reference's OWN test programs (tests/*.c of /root/reference), compiled where they lie against include/spasm.h
and linked against libspasm_b200.so by `make -C oracle b200_tests` (in the build container; the binaries travel to
the GPU box inside oracle/_ref/).  They are run the way tests/CMakeLists.txt:20-26,46-61 runs them: fixture on stdin,
`--modulus P`, failure = non-zero exit or "not ok" on stdout.

CPU tier: the programs that only exercise host code of the boundary (field, PRNG, SHA-256, permutations, transpose,
spmv, sub-matrix, the one-row triangular solves).  GPU tier: echelonize, kernel, schur, schur_dense,
dense_rref_ffpack, sparse_utsolve, sparse_lu_usolve and the L path (lu, solve, gesv, rank_cert, dense_lu_ffpack:
spasm_echelonize with opts->L / opts->complete, spasm_ffpack_LU) -- they call spasm_echelonize & co. (CUDA) and CHECK
the result with the host verifiers of csrc/host/verify.c and solve.c, i.e. with the reference's own row-by-row
arithmetic.

The known answers in tests/golden/expected/ are the reference's tests/Expected/{prng,hash,gaxpy.1,submatrix.1}."""
import os
import subprocess

import numpy as np
import pytest

import util
from spasm_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref")
EXPECTED = os.path.join(ROOT, "tests", "golden", "expected")
MODULI = (3, 257, 65537, 67108859, 189812507, 4294967291)       # tests/CMakeLists.txt:46-61
FIXTURES = None


def fixture_names():
    global FIXTURES
    if FIXTURES is None:
        f = np.load(os.path.join(util.GOLDEN_DIR, "fixtures.npz"))
        FIXTURES = sorted({k.rsplit(".", 1)[0] for k in f.files})
    return FIXTURES


def sms_of(name: str) -> bytes:
    f = np.load(os.path.join(util.GOLDEN_DIR, "fixtures.npz"))
    n, m = (int(v) for v in f[f"{name}.shape"])
    return synthetic.Triplets(n, m, 0, f[f"{name}.i"], f[f"{name}.j"], f[f"{name}.x"], name).to_sms()


def run(prog: str, *args, stdin: bytes = b"", timeout=120):
    path = os.path.join(BIN, "b200_test_" + prog)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/b200_test_* not built (make -C oracle b200_tests needs /root/reference)")
    env = dict(os.environ, OMP_NUM_THREADS="2")
    # one visible device per test process: the CUDA start-up of a fresh process grows with the number of GPUs it enumerates
    vis = [d for d in os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",") if d.strip()]
    env["CUDA_VISIBLE_DEVICES"] = vis[0] if vis else "0"
    return subprocess.run([path, *args], input=stdin, capture_output=True, env=env, timeout=timeout)


def check(r, what):
    assert r.returncode == 0, f"{what}: exit {r.returncode}\n{r.stdout[-600:].decode(errors='replace')}\n{r.stderr[-600:].decode(errors='replace')}"
    assert b"not ok" not in r.stdout, f"{what}:\n{r.stdout[-800:].decode(errors='replace')}"


# ---------------------------------------------------------------- host-only programs (no GPU)

def test_GFp():
    check(run("GFp", timeout=600), "GFp")


@pytest.mark.parametrize("prog,expected", [("prng", "prng"), ("sha", "hash")])
def test_known_answer_programs(prog, expected):
    r = run(prog)
    check(r, prog)
    with open(os.path.join(EXPECTED, expected), "rb") as f:
        assert r.stdout == f.read()


def test_vec_perm():
    check(run("vec_perm"), "vec_perm")


@pytest.mark.parametrize("fixture", ["small", "upper_trapeze"])
def test_mat_perm(fixture):
    check(run("mat_perm", "--modulus", "65537", stdin=sms_of(fixture)), "mat_perm " + fixture)


def test_transpose_on_every_fixture():
    for name in fixture_names():
        check(run("transpose", "--modulus", "257", stdin=sms_of(name)), "transpose " + name)


@pytest.mark.parametrize("prog,fixture,expected", [("spmv", "m1", "gaxpy.1"), ("submatrix", "singular", "submatrix.1")])
def test_expected_output_programs(prog, fixture, expected):
    r = run(prog, stdin=sms_of(fixture))
    check(r, prog)
    with open(os.path.join(EXPECTED, expected), "rb") as f:
        assert r.stdout == f.read()


@pytest.mark.parametrize("prog", ["dense_usolve", "sparse_usolve"])
def test_upper_triangular_solves(prog):
    for name in ("u1", "upper_trapeze"):
        for p in MODULI:
            check(run(prog, "--modulus", str(p), stdin=sms_of(name)), f"{prog} {name} mod {p}")


@pytest.mark.parametrize("prog", ["dense_lsolve", "sparse_lsolve"])
def test_lower_triangular_solves(prog):
    """tests/CMakeLists.txt:188-189 (spasm_dense_back_solve, the solve behind spasm_solve's L step)"""
    for name in ("l1", "lower_trapeze"):
        for p in MODULI:
            check(run(prog, "--modulus", str(p), stdin=sms_of(name)), f"{prog} {name} mod {p}")


# ---------------------------------------------------------------- programs that run the CUDA path

FULL_SWEEP = bool(os.environ.get("SPASM_B200_FULL_SWEEP"))


def _sweep(prog, jobs):
    import concurrent.futures
    inputs = {name: sms_of(name) for name in fixture_names()}

    def one(job):
        name, p = job
        for attempt in range(2):
            r = run(prog, "--modulus", str(p), stdin=inputs[name], timeout=300)
            if r.returncode == 0 and b"not ok" not in r.stdout:
                return None
            if b"no usable CUDA device" not in r.stderr:
                break                       # a real failure; a start-up refusal (many contexts created at once) is retried once
        return (name, p, r.returncode, r.stdout[-200:], r.stderr[-300:])

    # every run is its own process with its own CUDA context (~1 s of start-up each): six at a time share the GPU
    with concurrent.futures.ThreadPoolExecutor(max_workers=6) as ex:
        failures = [f for f in ex.map(one, jobs) if f is not None]
    assert not failures, f"{len(failures)} failing runs, first: {failures[:3]}"


def _jobs(prog_index: int, per_fixture: int):
    """32 fixtures x 6 moduli (tests/CMakeLists.txt spasm_run_tests_mod) is 192 processes, two minutes of GPU-box time per
    program, all of it CUDA start-up.  By default only `echelonize` runs the full product; the other programs run every
    fixture on `per_fixture` of the six moduli, rotated with the fixture and the program so that each program meets all
    six moduli and each fixture meets all six across the programs.  SPASM_B200_FULL_SWEEP=1 runs the full product for
    every program (tools/r2_job12.sh; green for all twelve programs in round 2)."""
    names = fixture_names()
    if FULL_SWEEP or per_fixture >= len(MODULI):
        return [(name, p) for name in names for p in MODULI]
    return [(name, MODULI[(k + prog_index + t * 3) % len(MODULI)]) for k, name in enumerate(names) for t in range(per_fixture)]


# program -> moduli per fixture in the default run.  L path: tests/CMakeLists.txt:192,218-225.
GPU_PROGRAMS = {"echelonize": 6, "kernel": 2, "lu": 2, "schur": 1, "schur_dense": 1, "dense_rref_ffpack": 1, "sparse_utsolve": 1,
                "sparse_lu_usolve": 1, "solve": 1, "gesv": 1, "rank_cert": 1, "dense_lu_ffpack": 1}


@pytest.mark.gpu
@pytest.mark.parametrize("prog", list(GPU_PROGRAMS))
def test_reference_program_on_the_fixtures(prog):
    _sweep(prog, _jobs(list(GPU_PROGRAMS).index(prog), GPU_PROGRAMS[prog]))
