"""The BASELINE configs AT FULL SIZE against the reference's own answers (tests/golden/golden_fullsize.json, produced in
the build container by the reference's sources -- tests/golden/make_golden_fullsize.py explains why they are computed
there and committed instead of being recomputed on the GPU box).

Compared, bit for bit: rank, the number and the digest of the round-0 structural (row, column) pivot pairs, the final
pivot-column set, the canonical RREF hash, the canonical kernel hash and the kernel dimension.  These sizes exercise
code paths the small goldens cannot reach: the 128 x 128 GEMM tiling (K >= 4096), panel_capacity chunking, the staged
bulk download, rows longer than the in-kernel caches, thousands of resolution windows of the greedy search.
"""
import hashlib
import json
import os

import pytest

import oracle
import util
from spasm_b200 import host, synthetic

pytestmark = pytest.mark.gpu

PATH = os.path.join(util.GOLDEN_DIR, "golden_fullsize.json")
GOLD = json.load(open(PATH))["cases"] if os.path.exists(PATH) else {}


@pytest.mark.parametrize("key", sorted(GOLD) or ["none"])
def test_full_size_config_against_the_reference(product, key):
    if key == "none":
        pytest.skip("tests/golden/golden_fullsize.json not generated")
    g = GOLD[key]
    t = synthetic.config2(g["scale"]).transposed() if g["name"] == "config2T" else synthetic.CONFIGS[g["name"]](g["scale"])
    assert (t.n, t.m, t.prime) == (g["n"], g["m"], g["prime"])
    A = host.compress(product, t)
    oracle.reset_rand()
    product.spasm_b200_reset_stats()
    f = host.echelonize(product, A, host.default_opts(product, **g["opts"]))
    assert f.rank == g["rank"]
    rows, cols, starts = util.product_pairs(product)
    assert starts[1] - starts[0] == g["npairs_round0"]
    assert util.pairs_digest(rows[starts[0]:starts[1]], cols[starts[0]:starts[1]]) == g["pairs_round0"], "structural pivots differ from the reference"
    assert hashlib.sha256(oracle.pivot_columns(f.qinv).tobytes()).hexdigest() == g["pivot_columns"]
    util.check_echelon_form(f.U, f.qinv)
    Km = host.kernel(product, f)
    assert Km.n == g["kernel_dim"]
    assert oracle.canonical_hash(Km.numpy()) == g["kernel"]
    del Km
    Rm, _ = host.rref(product, f)
    if "rref_nnz" in g:
        assert Rm.nnz == g["rref_nnz"]
    assert oracle.canonical_hash(Rm.numpy()) == g["rref"]
