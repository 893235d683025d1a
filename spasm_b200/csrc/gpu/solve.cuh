/*
 * Batched sparse triangular solve ("panel solve"), the B200 formulation of
 * spasm_sparse_triangular_solve (reference: src/spasm_triangular.c:109-146,
 * src/spasm_reach.c, src/spasm_scatter.c).
 *
 * The reference solves x*U = B[k] one row at a time: a DFS finds the reach of
 * B[k], then the reached pivot rows are scattered into a dense vector with
 * random read-modify-writes.  Here R right-hand sides are solved together in
 * PULL form over a column-major panel X (one 4*R-byte vector per column):
 *
 *     X[c][:]  =  B[c][:]  -  sum over rows i of U with an entry u on column c
 *                             (c not the pivot of i) of   u * X[pivot_col(i)][:]
 *
 * Every column is written exactly once, by one thread group, from columns that
 * are final: no atomics, no per-row DFS, only streaming 16-byte loads.  Columns
 * are scheduled by the levels of the pivot DAG (level(c) = 1 + max level of the
 * pivot columns it depends on).  On exit X[c][r] holds, for a pivotal column,
 * the elimination coefficient (what the reference leaves in x[j], i.e. the L
 * entry) and, for a non-pivotal column, the Schur complement entry.  Values are
 * the same field elements the reference computes (exact arithmetic mod p).
 *
 * The engine is generic over a "dependency graph": node -> list of (source
 * node, coefficient).  Two instances are built from U:
 *   forward    nodes = columns       (schur, schur_dense, randomized, rref, density)
 *   transposed nodes = rows of U     (spasm_kernel solves against U^t, reference: src/spasm_kernel.c:54)
 */
#pragma once
#include "common.cuh"
#include "zp.cuh"

namespace sb {

struct DepGraph {
	int nnodes = 0;
	i64 ndeps = 0;
	DevBuf<i64> ptr;          /* nnodes + 1 */
	DevBuf<int> src;          /* ndeps */
	DevBuf<i32> val;          /* ndeps */
	/* schedule */
	int nlevels = 0;
	std::vector<int> level_ptr_h;   /* nlevels + 1 offsets into order[]; level 0 = nodes without dependency */
	DevBuf<int> level_ptr;
	DevBuf<int> order;        /* nodes sorted by (level, index) */
	DevBuf<int> level;        /* level of each node */
	i64 scheduled_deps = 0;
	/* dataflow schedule: reverse adjacency (node -> nodes that depend on it) and, per node, the number of its
	 * dependencies that are not sources (level > 0) */
	DevBuf<i64> rptr;
	DevBuf<int> rdst;
	DevBuf<int> pending0;
	DevBuf<int> seeds;        /* nodes of level >= 1 all of whose dependencies are sources */
	int nseeds = 0, nscheduled = 0;
	/* false after a lazy schedule: level[] / order[] / level_ptr / nlevels are produced by the first dataflow solve
	 * (level[c] = 1 + max level of its dependencies is a by-product of the pass) and completed by depgraph_finish_levels */
	bool levels_known = true;
};

/* deps of column c: (pivot column of row i, U[i][c]) for every row i of U holding c outside its pivot.
 * U must have the pivot as first entry of each row (reference convention, src/spasm_pivots.c:6-8). */
void depgraph_forward(const DevCsr &U, DepGraph &G);
/* deps of row i: (row holding the pivot of column c, U[i][c]) for every entry of row i on a pivotal column c other than its own pivot */
void depgraph_transposed(const DevCsr &U, const int *d_qinv, DepGraph &G);
/* Kahn levels + deterministic order.  Aborts if the graph has a cycle (U not triangular).
 * lazy = true skips the Kahn pass (a full traversal of the DAG, 14 ms on the 3989-level graph of config 2): only what the
 * dataflow solve needs is built, and the levels come out of the first solve (a cycle is then reported by that solve). */
void depgraph_schedule(DepGraph &G, bool lazy = false);
/* order / level boundaries / nlevels from a filled level[] */
void depgraph_finish_levels(DepGraph &G);

/* X is nnodes x ld (ld multiple of 4, >= R), int32 balanced, column-major per node; solved in place */
void panel_solve(const DepGraph &G, i32 *X, int ld, int R, const Zp &F);
/* same solve for a very sparse batch: mask (nnodes x mw words, bit g of a node = its group of right-hand sides
 * 4g..4g+3 may be non-zero) holds the pattern of the right-hand sides on entry and of the solution on exit; only the
 * marked groups are computed (the others are zero and stay untouched) */
void panel_solve_masked(const DepGraph &G, i32 *X, int ld, int R, unsigned *mask, int mw, const Zp &F);

}  // namespace sb
