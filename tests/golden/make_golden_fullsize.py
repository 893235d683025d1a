"""Regenerates tests/golden/golden_fullsize.json: the BASELINE configs AT FULL SIZE, answered by the reference itself.

Run in the build container only (needs oracle/_ref/libspasm_ref_det.so, i.e. /root/reference):

    python tests/golden/make_golden_fullsize.py [case ...]          # hours of CPU for the RREFs; cases are appended

The library used is the "deterministic-parallel" build of the reference's own sources (oracle/Makefile:
src/spasm_pivots.c compiled without OpenMP, so the greedy search commits in row order like OMP_NUM_THREADS=1 -- the
parity target -- and everything else with OpenMP).  At these sizes a one-thread run of the reference takes hours
(1524 s for config 1 with real FFPACK) and the GPU box is charged while its host cores work, so the
answers are computed here once and committed; tests/test_gpu_fullsize.py compares the CUDA path with them.

Recorded per case: rank, number and digest of the round-0 structural (row, column) pivot pairs (from
the reference's own spasm_pivots_extract_structural), sha256 of the final pivot-column set, sha256 of the canonical
RREF and kernel basis, kernel dimension.  `rref: "identity"` marks full column rank (RREF = identity on every column:
the reference's row-by-row spasm_rref needs > 20 min to say so on config 2; the hash is computed from the identity).
"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import oracle  # noqa: E402
import util  # noqa: E402
from spasm_b200 import host, synthetic  # noqa: E402

CASES = {
    "config2T": ("config2T", 1.0, {}, "identity"),
    "config1": ("config1", 1.0, {}, "full"),
    "config4": ("config4", 1.0, {}, "full"),
    "config5": ("config5", 1.0, {}, "full"),
    "config4@0.2": ("config4", 0.2, {}, "full"),          # the full-size RREF (394 M entries) takes the reference more than three hours
    "config3@0.25": ("config3", 0.25, {"sparsity_threshold": 0.01}, "identity"),
    "config3@0.5": ("config3", 0.5, {"sparsity_threshold": 0.01}, "identity"),
}
OUT = os.environ.get("GOLDEN_OUT", os.path.join(HERE, "golden_fullsize.json"))     # several generators in parallel: one file each, merged by hand


def identity_hash(m: int, prime: int) -> str:
    ident = {"n": m, "m": m, "p": np.arange(m + 1, dtype=np.int64), "j": np.arange(m, dtype=np.int32), "x": np.ones(m, np.int32), "prime": prime}
    return oracle.canonical_hash(ident)


def run(key):
    name, scale, opts, rref_mode = CASES[key]
    t = synthetic.config2(scale).transposed() if name == "config2T" else synthetic.CONFIGS[name](scale)
    R_ = oracle.ref_det()
    A = host.compress(R_, t)
    out = {"name": name, "scale": scale, "opts": opts, "prime": t.prime, "n": t.n, "m": t.m}
    t0 = time.time()
    npiv, p, fact = host.pivots_extract_structural(R_, A, host.default_opts(R_, **opts))
    rows, cols = host.fact_pairs(fact, p, npiv)
    out["npairs_round0"] = int(npiv)
    out["pairs_round0"] = util.pairs_digest(rows, cols)
    out["seconds_pivots"] = time.time() - t0
    oracle.reset_rand()
    t0 = time.time()
    f = host.echelonize(R_, A, host.default_opts(R_, **opts))
    out["seconds_echelonize"] = time.time() - t0
    out["rank"] = int(f.rank)
    out["pivot_columns"] = hashlib.sha256(oracle.pivot_columns(f.qinv).tobytes()).hexdigest()
    if rref_mode == "identity":
        assert f.rank == t.m, "identity shortcut needs full column rank"
        out["rref"] = identity_hash(t.m, t.prime)
        out["rref_how"] = "identity"
    else:
        t0 = time.time()
        Rm, _ = host.rref(R_, f)
        out["seconds_rref"] = time.time() - t0
        out["rref"] = oracle.canonical_hash(Rm.numpy())
        out["rref_nnz"] = int(Rm.nnz)
        out["rref_how"] = "spasm_rref"
    t0 = time.time()
    Km = host.kernel(R_, f)
    out["seconds_kernel"] = time.time() - t0
    out["kernel"] = oracle.canonical_hash(Km.numpy())
    out["kernel_dim"] = int(Km.n)
    return out


def main():
    keys = sys.argv[1:] or list(CASES)
    data = {"generator": "tests/golden/make_golden_fullsize.py", "reference": "oracle/_ref/libspasm_ref_det.so", "cases": {}}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(2)
    for key in keys:
        os.dup2(devnull, 2)
        try:
            res = run(key)
        finally:
            os.dup2(saved, 2)
        data["cases"][key] = res
        with open(OUT, "w") as f:
            json.dump(data, f, indent=1)
        print(key, res, flush=True)


if __name__ == "__main__":
    main()
