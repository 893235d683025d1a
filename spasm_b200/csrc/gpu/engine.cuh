/*
 * Device-resident state shared by the drivers (echelonize.cu, api.cu):
 * an echelon form under construction = structural sparse rows U (CSR on the
 * device) + its solve schedule + the dense rows produced by the dense phases.
 */
#pragma once
#include "common.cuh"
#include "dense.cuh"
#include "panel.cuh"
#include "pivots.cuh"
#include "solve.cuh"
#include "zp.cuh"

namespace sb {

/* one batch of rows produced by a dense echelonization: rr RREF rows over the q0 column space */
struct DenseBlock {
	int rr = 0;
	DevBuf<i32> D;                   /* rr x Sm0, row-major, leading dimension ld */
	int ld = 0;
	std::vector<int> pivcol;         /* pivot column (index into q0) of each row, increasing */
	DevBuf<int> d_pivcol;
	DevBuf<unsigned char> d_own;     /* flag per q0 column: pivot of this block */
	DevBuf<int8_t> Dpack;            /* D split into int8 limb planes, tile-packed for the tensor cores (umma_gemm.cu); may be empty */
	/* L mode only (lu.cu): the rows that go to U, unit upper triangular on the pivot columns (Uc * D), and the row of the
	 * block each one was obtained from */
	DevBuf<i32> Dlu;
	std::vector<int> lu_row;
};

struct Engine {
	int m = 0;
	i64 prime = 0;
	Zp F;
	/* structural part */
	DevCsr U;
	DevBuf<int> Uqinv;               /* size m; row of U (structural index) or -1 */
	DepGraph G;                      /* forward dependency graph of U, scheduled */
	bool G_ready = false;
	int lazy_rows = 0;               /* the first lazy_rows rows of U are in column order: assemble() puts them in level order */
	/* L mode (opts->L): row of the input matrix each row of the echelon form was obtained from (fact->p) */
	bool want_L = false;
	int first_round_rows = 0;        /* structural pivots of round 0 (rows 0 .. first_round_rows-1 of U, in any order) */
	std::vector<int> p_struct;       /* per row of U */
	std::vector<int> p_dense;        /* per dense row, block after block */
	/* dense part, over the columns that are non-pivotal after the structural rounds */
	bool dense_ready = false;
	int Sm0 = 0;
	std::vector<int> q0;             /* q0[c] = column of A */
	DevBuf<int> d_q0;
	std::vector<DenseBlock> blocks;
	int dense_rank = 0;
	Panel panel;

	void init(int m_, i64 prime_);
	void rebuild_schedule();                         /* G from U */
	void begin_dense();                              /* freeze q0 */
	int rank() const { return U.n + dense_rank; }

	/* solve the rows `rows` (indices into B) against the structural U: results in panel */
	void solve_rows(const DevCsr &B, const int *d_rows, int R, bool skip_first, bool sparse_batch = false);
	void solve_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w);
	/* dense block (R x Sm0, ld = ldB, room for world * chunk rows) of the rows / combinations reduced by the structural
	 * pivots; with several ranks each one solves a slice and the slices are all-gathered */
	void block_from_rows(const DevCsr &B, const int *d_rows, int R, DevBuf<i32> &out, int &ldB);
	void block_from_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w, DevBuf<i32> &out, int &ldB);
	void account_bytes(int R);
	/* gather the q0 columns of the panel into a dense row-major block (R x Sm0, ld = ldS) */
	void gather_q0(i32 *S, int ldS);
	/* reduce a dense block by the dense rows found so far, echelonize it, keep its pivot rows. Returns rr.
	 * with_lu (L mode): the block also keeps its pivot rows in unit-upper-triangular form and their rows of origin. */
	int absorb_block(i32 *B, int rows, int ldB, bool with_lu = false);
};

int comm_world();
int comm_rank();
void comm_slice(int total, int *chunk, int *begin, int *end);
void comm_allgather_rows(i32 *B, int chunk, int ld);
void comm_allgather_bytes(void *buf, size_t bytes);
void comm_group_begin();
void comm_group_end();
void comm_bcast_bytes(void *buf, size_t bytes, int root);
int comm_result_root();
void comm_send_bytes(const void *buf, size_t bytes, int peer);
void comm_recv_bytes(void *buf, size_t bytes, int peer);

}  // namespace sb
