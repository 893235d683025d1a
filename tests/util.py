"""Shared helpers of the parity tests: run the product / the oracle / the compiled reference on the same
triplets and reduce the results to the canonical artefacts of SURVEY.md section 8c."""
from __future__ import annotations

import ctypes as C
import hashlib
import os

import numpy as np

import oracle
import spasm_b200
from spasm_b200 import abi, host, synthetic

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


from spasm_b200 import Stats  # noqa: E402  (ctypes mirror of struct spasm_b200_stats)


def product_stats(L) -> Stats:
    s = Stats()
    L.spasm_b200_get_stats(C.byref(s))
    return s


def product_pairs(L):
    s = product_stats(L)
    n = L.spasm_b200_last_pivot_pairs(None, None, None)
    rows = np.zeros(max(n, 1), np.int32)
    cols = np.zeros(max(n, 1), np.int32)
    starts = np.zeros(s.nrounds + 2, np.int32)
    L.spasm_b200_last_pivot_pairs(abi.as_int_p(rows), abi.as_int_p(cols), abi.as_int_p(starts))
    return rows[:n], cols[:n], starts[:s.nrounds + 1]


def pairs_digest(rows, cols) -> str:
    """order-independent digest of a set of (row, column) pivot pairs"""
    if len(rows) == 0:
        return hashlib.sha256(b"").hexdigest()
    key = np.sort(np.asarray(rows, np.int64) * (1 << 32) + np.asarray(cols, np.int64))
    return hashlib.sha256(key.tobytes()).hexdigest()


def summarize(rank, qinv, R, K, pair_rows, pair_cols, pair_start, found, finish, blocks) -> dict:
    """The canonical, comparable description of one echelonization."""
    out = {
        "rank": int(rank),
        "pivot_columns": hashlib.sha256(oracle.pivot_columns(qinv).tobytes()).hexdigest(),
        "rref": oracle.canonical_hash(R),
        "kernel": oracle.canonical_hash(K) if K is not None else None,
        "kernel_dim": int(K["n"]) if K is not None else None,
        "found": [list(map(int, f)) for f in found],
        "pairs_per_round": [pairs_digest(pair_rows[pair_start[r]:pair_start[r + 1]], pair_cols[pair_start[r]:pair_start[r + 1]])
                            for r in range(len(pair_start) - 1)],
        "finish": int(finish),
        "blocks": [list(map(int, b)) for b in blocks],
    }
    return out


def run_oracle(trip, **opts) -> dict:
    A = oracle.compress(trip)
    e = oracle.echelonize(A, oracle.default_opts(**opts))
    R, _ = oracle.rref(e.U, e.qinv)
    K = oracle.kernel(e.U, e.qinv)
    return summarize(e.rank, e.qinv, R, K, e.pair_row, e.pair_col, e.pair_start, e.found, e.finish, e.blocks)


def run_reference(trip, **opts) -> dict:
    """The compiled reference (oracle/_ref).  It exposes no trace, so pairs come from
    spasm_pivots_extract_structural on round 0 only and found/blocks are left out."""
    R_ = oracle.ref()
    A = host.compress(R_, trip)
    oracle.reset_rand()
    f = host.echelonize(R_, A, host.default_opts(R_, **opts))
    Rm, _ = host.rref(R_, f)
    Km = host.kernel(R_, f)
    return {"rank": f.rank,
            "pivot_columns": hashlib.sha256(oracle.pivot_columns(f.qinv).tobytes()).hexdigest(),
            "rref": oracle.canonical_hash(Rm.numpy()), "kernel": oracle.canonical_hash(Km.numpy()), "kernel_dim": Km.n}


def run_product(L, trip, **opts) -> dict:
    A = host.compress(L, trip)
    oracle.reset_rand()        # glibc rand() is part of the reference's result (never seeded there)
    L.spasm_b200_reset_stats()
    f = host.echelonize(L, A, host.default_opts(L, **opts))
    s = product_stats(L)
    rows, cols, starts = product_pairs(L)
    Rm, _ = host.rref(L, f)
    Km = host.kernel(L, f)
    found = [(s.found_FL[r], s.found_FLcol[r], s.found_greedy[r]) for r in range(s.nrounds)]
    blocks = [(s.block_Sn[k], s.block_Sm[k], s.block_rr[k], s.block_w[k]) for k in range(min(s.nblocks, 4096))]
    out = summarize(f.rank, f.qinv, Rm.numpy(), Km.numpy(), rows, cols, starts, found, s.finish, blocks)
    out["_U"] = f.U
    out["_qinv"] = f.qinv
    return out


def check_echelon_form(U: dict, qinv: np.ndarray):
    """reference: tests/echelonize.c:30-50 -- each row non-empty, first entry is a unit pivot on a fresh column."""
    p, j, x = U["p"], U["j"], U["x"]
    assert (np.diff(p) > 0).all()
    first = p[:-1]
    assert (x[first] == 1).all()
    pc = j[first]
    assert len(np.unique(pc)) == len(pc)
    assert (qinv[pc] == np.arange(U["n"])).all()
    assert (np.asarray(qinv) >= 0).sum() == U["n"]


COMPARED = ("rank", "pivot_columns", "rref", "kernel", "kernel_dim", "found", "finish", "blocks")


def assert_same(got: dict, want: dict, keys=COMPARED, what=""):
    for k in keys:
        if k in want and want[k] is not None:
            assert got[k] == want[k], f"{what}: {k} differs: {got[k]} vs {want[k]}"


def load_sms(path: str, prime: int) -> synthetic.Triplets:
    rows = []
    with open(path) as f:
        hdr = f.readline().split()
        n, m = int(hdr[0]), int(hdr[1])
        for line in f:
            a = line.split()
            if len(a) < 3:
                continue
            i, j, x = int(a[0]), int(a[1]), int(a[2])
            if i == 0 and j == 0 and x == 0:
                break
            rows.append((i - 1, j - 1, x))
    arr = np.array(rows, dtype=np.int64).reshape(-1, 3)
    return synthetic.Triplets(n, m, prime, arr[:, 0].astype(np.int32), arr[:, 1].astype(np.int32), arr[:, 2].copy(),
                              os.path.basename(path))


# ------------------------------------------------------------------ golden vectors (tests/golden/make_golden.py)

def golden_cases():
    import json
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)["cases"]


_fixtures = None


def golden_input(case) -> synthetic.Triplets:
    global _fixtures
    if case["kind"] == "fixture":
        if _fixtures is None:
            _fixtures = np.load(os.path.join(GOLDEN_DIR, "fixtures.npz"))
        b = case["name"]
        n, m = (int(v) for v in _fixtures[f"{b}.shape"])
        return synthetic.Triplets(n, m, case["prime"], _fixtures[f"{b}.i"], _fixtures[f"{b}.j"], _fixtures[f"{b}.x"], b)
    if case["name"] == "config2T":
        return synthetic.config2(case["scale"]).transposed()
    return synthetic.CONFIGS[case["name"]](case["scale"])


def case_id(case) -> str:
    extra = "".join(f"-{k}={v}" for k, v in case["opts"].items())
    return f'{case["name"]}{("@" + str(case["scale"])) if "scale" in case else ""}-p{case["prime"]}{extra}'
