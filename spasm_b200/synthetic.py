"""Fixed-seed synthetic inputs of the shapes named in BASELINE.json / SURVEY.md section 8d.

Every generator returns a `Triplets` (0-based row, column, value arrays in *file
order*, plus shape and prime).  File order matters: `spasm_compress` keeps the
entries of a CSR row in file order and the pivot search takes the first
eligible entry of a row (reference: src/spasm_triplet.c:139-147,
src/spasm_pivots.c:104-116), so the same triplets in the same order are fed to
the reference oracle and to the B200 library.
"""
from __future__ import annotations

import dataclasses
import io

import numpy as np

DEFAULT_PRIME = 42013          # reference: tools/common.c:26
BIG_PRIME = 2147483629         # BASELINE.json config 5


@dataclasses.dataclass
class Triplets:
    n: int
    m: int
    prime: int
    i: np.ndarray      # int32
    j: np.ndarray      # int32
    x: np.ndarray      # int64 (unreduced, as they would appear in an SMS file)
    name: str = ""

    @property
    def nz(self) -> int:
        return int(self.i.shape[0])

    def transposed(self) -> "Triplets":
        """What tools/rank.c:84-88 does when n < m (swap the roles of i and j)."""
        return Triplets(self.m, self.n, self.prime, self.j.copy(), self.i.copy(), self.x.copy(), self.name + "^T")

    def with_prime(self, prime: int) -> "Triplets":
        return dataclasses.replace(self, prime=prime)

    def to_sms(self) -> bytes:
        """SMS text (reference: src/spasm_io.c:164-180)."""
        buf = io.BytesIO()
        buf.write(f"{self.n} {self.m} M\n".encode())
        body = np.stack([self.i.astype(np.int64) + 1, self.j.astype(np.int64) + 1, self.x.astype(np.int64)], axis=1)
        np.savetxt(buf, body, fmt="%d")
        buf.write(b"0 0 0\n")
        return buf.getvalue()


def _values(rng, count, prime, kind):
    if kind == "uniform":          # uniform on [1, p)
        return rng.integers(1, prime, size=count, dtype=np.int64)
    if kind == "pm1":              # boundary-matrix like
        return rng.choice(np.array([-1, 1], dtype=np.int64), size=count)
    if kind == "small":            # [-5, 5] \ {0}
        v = rng.integers(1, 6, size=count, dtype=np.int64)
        return v * rng.choice(np.array([-1, 1], dtype=np.int64), size=count)
    raise ValueError(kind)


def uniform_rows(n, m, nnz_per_row, prime=DEFAULT_PRIME, seed=20240229, values="uniform", name="", distinct=False) -> Triplets:
    """Each row: `nnz_per_row` column indices i.i.d. uniform on [0, m), row-major file order.
    distinct=False: repeats allowed (they are summed by spasm_compress), as SURVEY 8d states for config 1.
    distinct=True : a repeated column inside a row is re-drawn, like real boundary / GL7d matrices which
    never list an entry twice.  (Repeated entries that cancel mod p trigger two reference quirks --
    src/spasm_triplet.c:36-57 and src/spasm_pivots.c:232-238, see DESIGN.md -- which are covered by
    dedicated tests instead of being baked into the benchmark inputs.)"""
    rng = np.random.default_rng(seed)
    if np.isscalar(nnz_per_row):
        counts = np.full(n, int(nnz_per_row), dtype=np.int64)
    else:
        counts = np.asarray(nnz_per_row, dtype=np.int64)
    total = int(counts.sum())
    i = np.repeat(np.arange(n, dtype=np.int32), counts)
    j = rng.integers(0, m, size=total, dtype=np.int64)
    if distinct:
        for _ in range(64):
            key = i.astype(np.int64) * m + j
            order = np.argsort(key, kind="stable")
            dup_sorted = np.zeros(total, bool)
            dup_sorted[1:] = key[order][1:] == key[order][:-1]
            dup = np.zeros(total, bool)
            dup[order] = dup_sorted
            if not dup.any():
                break
            j[dup] = rng.integers(0, m, size=int(dup.sum()), dtype=np.int64)
    x = _values(rng, total, prime, values)
    return Triplets(n, m, prime, i, j.astype(np.int32), x, name)


def config1(scale=1.0, prime=DEFAULT_PRIME) -> Triplets:
    """BASELINE config 1: 20000 x 20000, 5 nnz/row, uniform values."""
    n = max(8, int(round(20000 * scale)))
    return uniform_rows(n, n, 5, prime, seed=20240229, values="uniform", name=f"config1(scale={scale})")


def config2(scale=1.0, prime=DEFAULT_PRIME) -> Triplets:
    """BASELINE config 2: mk13.b5 *shape* 135135 x 270270, 6 nnz/row, values +-1.
    tools/rank transposes it (n < m) before compressing; use `.transposed()` for that."""
    n = max(8, int(round(135135 * scale)))
    m = max(16, int(round(270270 * scale)))
    return uniform_rows(n, m, 6, prime, seed=20240301, values="pm1", name=f"config2(scale={scale})", distinct=True)


def config3(scale=1.0, prime=DEFAULT_PRIME) -> Triplets:
    """BASELINE config 3: GL7d14 *shape* 171375 x 47271, ~11 nnz/row, small integer values.
    Meant to be run with sparsity_threshold (--dense-threshold) 0.01."""
    n = max(8, int(round(171375 * scale)))
    m = max(8, int(round(47271 * scale)))
    return uniform_rows(n, m, 11, prime, seed=20240302, values="small", name=f"config3(scale={scale})", distinct=True)


def planted(n, prime=DEFAULT_PRIME, seed=20240303, band=1000, extra=4, random_fraction=0.01, dup_fraction=0.002,
            name="") -> Triplets:
    """BASELINE config 4 family: a unit upper-triangular band matrix (diagonal plus `extra`
    entries within the next `band` columns) whose rows are shuffled, with a fraction of the rows replaced
    by uniform random rows (-> a small dense Schur complement) and a fraction replaced by copies of other
    rows (-> rank deficiency, non-trivial kernel)."""
    rng = np.random.default_rng(seed)
    rows_i, rows_j, rows_x = [], [], []
    # band part: row r has entries (r, r) = 1 and `extra` entries in (r, r + band]
    base = np.arange(n, dtype=np.int64)
    off = rng.integers(1, band + 1, size=(n, extra), dtype=np.int64)
    cols = np.minimum(base[:, None] + off, n - 1)
    jj = np.concatenate([base[:, None], cols], axis=1)              # n x (extra+1)
    xx = np.concatenate([np.ones((n, 1), dtype=np.int64), _values(rng, n * extra, prime, "uniform").reshape(n, extra)], axis=1)
    # random rows replace some band rows
    n_rand = int(n * random_fraction)
    rand_rows = rng.choice(n, size=n_rand, replace=False)
    jj[rand_rows] = rng.integers(0, n, size=(n_rand, extra + 1), dtype=np.int64)
    xx[rand_rows] = _values(rng, n_rand * (extra + 1), prime, "uniform").reshape(n_rand, extra + 1)
    # duplicated rows: row a becomes a copy of row b
    n_dup = int(n * dup_fraction)
    if n_dup > 0:
        dst = rng.choice(n, size=n_dup, replace=False)
        src = rng.integers(0, n, size=n_dup)
        jj[dst] = jj[src]
        xx[dst] = xx[src]
    perm = rng.permutation(n)                                       # shuffle the rows
    jj = jj[perm]
    xx = xx[perm]
    i = np.repeat(np.arange(n, dtype=np.int32), extra + 1)
    return Triplets(n, n, prime, i, jj.reshape(-1).astype(np.int32), xx.reshape(-1), name or f"planted(n={n})")


def config4(scale=1.0, prime=DEFAULT_PRIME) -> Triplets:
    """BASELINE config 4: 500k x 500k planted matrix for kernel + RREF."""
    return planted(max(64, int(round(500000 * scale))), prime, name=f"config4(scale={scale})")


def config5(scale=1.0) -> Triplets:
    """BASELINE config 5: config-1 shape with the 31-bit prime 2147483629 (4 int8 limbs in the dense update)."""
    t = config1(scale, prime=BIG_PRIME)
    t.name = f"config5(scale={scale})"
    return t


CONFIGS = {"config1": config1, "config2": config2, "config3": config3, "config4": config4, "config5": config5}
