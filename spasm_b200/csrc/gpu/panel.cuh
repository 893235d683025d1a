/*
 * Right-hand-side panels for the batched solve: building them from sparse rows
 * (or random combinations of rows) and reading results back out.
 * Layout in HBM: X[node][r], one vector of `ld` int32 per node (column of the
 * matrix), r = index of the right-hand side inside the batch, ld % 4 == 0.
 */
#pragma once
#include "common.cuh"
#include "zp.cuh"

namespace sb {

struct Panel {
	DevBuf<i32> X;
	int nnodes = 0, ld = 0, R = 0;
	/* optional occupancy mask for very sparse batches (spasm_rref): bit g of node c = the group of 4 right-hand sides
	 * 4g..4g+3 of X[c] may be non-zero.  mw words per node; the scatter sets the bits of the right-hand sides, the
	 * solve propagates them (OR over the dependencies) and computes the marked groups only, the conversion back to
	 * sparse rows reads the marked groups only. */
	DevBuf<unsigned> mask;
	int mw = 0;
	bool masked = false;
	void shape(int nnodes_, int R_, bool masked_ = false);     /* (re)allocate and zero */
};

/* how many right-hand sides fit the panel budget (GB of HBM, SPASM_B200_PANEL_GB overrides) for this many nodes.
 * A solve costs about the same whatever its width (it is bound by the depth of the pivot DAG), so callers that own
 * the whole device (rref, kernel) ask for a large panel. */
int panel_capacity(int nnodes, double gb = 16.0);

/* X[B.j[e]][r] += B.x[e] for every entry e of row rows[r] (r < R); skip_first drops the first entry of each row */
void panel_scatter_rows(const DevCsr &B, const int *d_rows, int R, Panel &P, const Zp &F, bool skip_first, int r_off = 0);
/* X[.][k] += coef[k*w+t] * A[rows[k*w+t]]   for k < N, t < w */
void panel_scatter_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w, Panel &P, const Zp &F, int r_off = 0);
/* transposed variant for spasm_kernel: right-hand side r = column cols[r] of U, i.e. X[i][r] = U[i][cols[r]] */
void panel_scatter_columns(const DevCsr &U, const int *d_colslot /* size m: slot of a column or -1 */, Panel &P);

/* S[r*ldS + c] = X[q[c]][r]  (row-major dense block, reference: src/spasm_schur.c:205-233 "gather") */
void panel_gather_dense(const Panel &P, const int *d_q, int Sm, i32 *S, int ldS, int r_off = 0, int count = -1);      /* right-hand sides r_off .. r_off + count */
/* number of non-zero entries on nodes with flag[node] < 0 (non-pivotal columns) over the whole panel */
i64 panel_count_nonzero(const Panel &P, const int *d_flag, int R_limit = -1);      /* first R_limit right-hand sides only */

/* CSR of the panel restricted to nodes with flag[node] < 0, entries of a row by increasing node.
 * If d_first != NULL, row r starts with the extra entry (first_col[r], first_val) (used by rref / kernel). */
void panel_to_csr(const Panel &P, const int *d_flag, const int *d_first_col, i32 first_val, const int *d_node_to_col,
                  DevBuf<i64> &Sp, DevBuf<int> &Sj, DevBuf<i32> &Sx, i64 &nnz, bool count_only = false);

/* fill Sj/Sx of rows counted by panel_to_csr(count_only) with the entries in the reference's DFS pattern order */
void panel_emit_reference_order(const Panel &P, const DevCsr &B, const int *d_rows, const DevCsr &U, const int *d_qinv,
                                int dag_depth, const i64 *d_Sp, int *d_Sj, i32 *d_Sx);

/* 8 * sum over right-hand sides of nnz(U rows reached): the dominant term of SURVEY 8d's algorithmic bytes */
double panel_reached_bytes(const Panel &P, const int *d_qinv, const i64 *d_Up);

}  // namespace sb
