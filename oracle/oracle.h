/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * Flat C interface of the CPU restatement (oracle.c), meant for ctypes.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

typedef int64_t i64;
typedef int32_t i32;

/* CSR matrix, same conventions as the reference (src/spasm.h:38-51) */
struct ocsr {
	int n, m;
	i64 nzmax;
	i64 *p;
	int *j;
	i32 *x;
	i64 prime;
};

/* mirrors struct echelonize_opts (src/spasm.h:84-108) with plain ints/doubles */
struct oracle_opts {
	int enable_greedy_pivot_search;
	int enable_tall_and_skinny;
	int enable_dense;
	int enable_GPLU;
	double min_pivot_proportion;
	int max_round;
	double sparsity_threshold;
	int dense_block_size;
	double low_rank_ratio;
	double tall_and_skinny_ratio;
	double low_rank_start_weight;
};

#define ORACLE_MAX_ROUNDS 64
#define ORACLE_MAX_BLOCKS 4096

/* what an echelonization returns, plus the trace used for parity checks */
struct oracle_lu {
	int rank;
	struct ocsr *U;
	int *qinv;                 /* size m */

	/* trace: structural pivot search, per round */
	int nrounds;
	int found_FL[ORACLE_MAX_ROUNDS];
	int found_FLcol[ORACLE_MAX_ROUNDS];
	int found_greedy[ORACLE_MAX_ROUNDS];
	int pair_start[ORACLE_MAX_ROUNDS + 1];   /* pairs of round r are [pair_start[r], pair_start[r+1]) */
	int npairs;
	int *pair_row;             /* row index in the ORIGINAL matrix */
	int *pair_col;
	double density[ORACLE_MAX_ROUNDS];

	/* trace: finishing strategy; 0 none (status 1), 1 low-rank, 2 dense, 3 GPLU */
	int finish;
	int nblocks;
	int block_Sn[ORACLE_MAX_BLOCKS], block_Sm[ORACLE_MAX_BLOCKS], block_rr[ORACLE_MAX_BLOCKS], block_w[ORACLE_MAX_BLOCKS];

	/* instrumentation */
	double tsolve_bytes;       /* sum over solved rows of 8*nnz(B[k]) + 8*sum nnz(U[i] reached) + 8*nnz(out) */
	i64 tsolve_rows;
	i64 greedy_edges;          /* entries of pivot rows traversed by the greedy search */
	double dense_fieldops;     /* 2*n*m*r style count of the dense eliminations */
	double seconds_pivots, seconds_schur, seconds_dense, seconds_total;
};

void oracle_default_opts(struct oracle_opts *o);
struct ocsr *oracle_csr_alloc(int n, int m, i64 nzmax, i64 prime);
void oracle_csr_free(struct ocsr *A);
struct ocsr *oracle_compress(int n, int m, i64 nz, const int *Ti, const int *Tj, const i64 *Tx, i64 prime);
struct ocsr *oracle_transpose(const struct ocsr *A);

struct oracle_lu *oracle_echelonize(const struct ocsr *A, const struct oracle_opts *opts);
void oracle_lu_free(struct oracle_lu *f);
struct ocsr *oracle_rref(const struct ocsr *U, const int *qinv, int *Rqinv);
struct ocsr *oracle_kernel(const struct ocsr *U, const int *qinv);

/* single pieces, for unit parity tests */
int oracle_pivots_find(const struct ocsr *A, int greedy, int *pinv, int *qinv, int counts[3], i64 *edges);
int oracle_tsolve(const struct ocsr *U, const struct ocsr *B, int k, int *xj, i32 *x, const int *qinv);
void oracle_schur_dense(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, i32 *S, int *q);
void oracle_schur_dense_randomized(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, i32 *S, int *q, int N, int w);
struct ocsr *oracle_schur(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, double est_density);
double oracle_schur_estimate_density(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, int R);
int oracle_dense_rref_i32(i64 p, int n, int m, i32 *W, int *pivcol);

/* field + prng + sha256 known-answer hooks */
i32 oracle_zp_mul(i64 p, i32 a, i32 b);
i32 oracle_zp_inverse(i64 p, i32 a);
i32 oracle_zp_axpy(i64 p, i32 a, i32 x, i32 y);
void oracle_sha256(const void *data, i64 len, unsigned char out[32]);
void oracle_prng_stream(i64 prime, uint64_t seed, uint32_t seq, int count, i32 *out);

#endif
