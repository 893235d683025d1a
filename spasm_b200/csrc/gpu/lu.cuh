/*
 * Dense LU pieces of the L path (lu.cu): row-pivoted LU of a full-column-rank block and the re-expression of reduced
 * pivot rows as unit upper triangular rows.  reference: src/spasm_ffpack.cpp:52-75, src/spasm_echelonize.c:228-313.
 */
#pragma once
#include "common.cuh"
#include "zp.cuh"

namespace sb {

/* in-place LU with row pivoting of C (n x k, rank k): prow[t] = row used at step t (first unused row with a non-zero
 * entry on column t).  Row prow[t]: columns <= t = its row of Lc (diagonal at column t), columns > t = row t of Uc
 * (unit diagonal implied).  The other rows hold their rows of Lc. */
void dense_lu_fullcol(i32 *C, int n, int k, int ld, const Zp &F, std::vector<int> &prow);

/* out (k x width, ldo) = Uc * R,  Uc read from the factored C, R (k x width, ldr) the reduced pivot rows */
void dense_lu_rows(const i32 *C, int ldc, int k, const std::vector<int> &prow, const i32 *R, int ldr, int width, i32 *out, int ldo, const Zp &F);

}  // namespace sb
