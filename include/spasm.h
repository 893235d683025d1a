/*
 * spasm.h -- public C ABI of spasm-b200.
 *
 * This header is the drop-in boundary: it declares the same types and entry
 * points, with the same memory layout, as SpaSM's public header
 * (reference: src/spasm.h).  Programs written against the reference compile
 * against this file and link against libspasm_b200.so unchanged: its
 * tools/rank.c, tools/echelonize.c, tools/kernel.c, and its test programs
 * echelonize, kernel, schur, schur_dense, dense_rref_ffpack, sparse_utsolve,
 * sparse_usolve, dense_usolve, sparse_lsolve, dense_lsolve, sparse_lu_usolve,
 * lu, solve, gesv, rank_cert, dense_lu_ffpack, GFp, prng, sha, spmv,
 * submatrix, transpose, mat_perm, vec_perm (oracle/Makefile: b200_tools,
 * b200_tests; run by tests/test_reference_tests.py).
 *
 * The L side (opts->L / opts->complete, spasm_ffpack_LU, spasm_solve,
 * spasm_gesv, rank certificates, spasm_factorization_verify) is provided
 * since round 2 (csrc/gpu/lu.cu, csrc/host/solve.c).
 *
 * NOT provided -- neither declared nor exported: spasm_dulmage_mendelsohn,
 * spasm_maximum_matching, spasm_strongly_connected_components,
 * spasm_structural_rank, spasm_submatching, spasm_permute_row_matching,
 * spasm_permute_column_matching, spasm_save_pnm (independent graph library
 * and bitmap output, never called by the echelonization) -- tools/dm,
 * tools/bitmap and the tests dm / matching / scc do not link.
 *
 * What differs is behind the boundary: spasm_echelonize(), spasm_rref(),
 * spasm_kernel(), the spasm_schur*() family, spasm_pivots_extract_structural(),
 * spasm_ffpack_rref() and spasm_ffpack_LU() run as CUDA kernels for sm_100a.
 * There is no CPU implementation of these entry points in the library: without a usable
 * CUDA device they abort with errx(1, ...), the reference's error convention
 * (reference: src/spasm_util.c:65-87).
 *
 * Conventions kept from the reference (src/spasm.h:25-51):
 *   - n = #rows, m = #columns
 *   - values are int32 in the balanced range [-(p-1)/2, (p-1)/2]
 *   - CSR rows are NOT sorted by column; row pointers are int64
 *   - in U (and in R = rref) the pivot is the first entry of its row and is 1
 *   - every array reachable from a returned struct is plain malloc() memory
 */
#ifndef _SPASM_H
#define _SPASM_H

#include <stddef.h>
#include <inttypes.h>
#include <stdio.h>
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

/* short integer names used all over the API (reference: src/spasm.h:9-13) */
typedef uint8_t  u8;
typedef uint32_t u32;
typedef uint64_t u64;
typedef int32_t  i32;
typedef int64_t  i64;

#define SPASM_VERSION "1.3-b200"
#define SPASM_BUG_ADDRESS "<spasm-b200>"

#define SPASM_IDENTITY_PERMUTATION NULL
#define SPASM_IGNORE NULL
#define SPASM_IGNORE_VALUES 0

/* ------------------------------------------------------------------ field */

/* element of Z/pZ, balanced representative (reference: src/spasm.h:28) */
typedef i32 spasm_ZZp;

/* reference: src/spasm.h:30-36 -- layout is ABI */
struct spasm_field_struct {
	i64 p;          /* the modulus                      */
	i64 halfp;      /* largest representative  (p/2)    */
	i64 mhalfp;     /* smallest representative          */
	double dinvp;   /* 1.0 / p                          */
};
typedef struct spasm_field_struct spasm_field[1];

/* ------------------------------------------------------------- containers */

/* reference: src/spasm.h:38-51 */
struct spasm_csr {
	i64 nzmax;        /* capacity of j[] and x[]                    */
	int n;            /* rows                                       */
	int m;            /* columns                                    */
	i64 *p;           /* n+1 row pointers; nnz == p[n]              */
	int *j;           /* column index of each entry                 */
	spasm_ZZp *x;     /* value of each entry, or NULL (pattern)     */
	spasm_field field;
};

/* reference: src/spasm.h:53-62 */
struct spasm_triplet {
	i64 nzmax;
	i64 nz;
	int n;
	int m;
	int *i;
	int *j;
	spasm_ZZp *x;     /* or NULL */
	spasm_field field;
};

/* result of an echelonization (reference: src/spasm.h:64-72) */
struct spasm_lu {
	int r;                        /* rank                                      */
	bool complete;                /* when L != NULL: is A == L*U ?             */
	struct spasm_csr *L;
	struct spasm_csr *U;
	int *qinv;                    /* qinv[j] = row of U with its pivot on column j, or -1 */
	int *p;                       /* pivots of L                               */
	struct spasm_triplet *Ltmp;   /* scratch, NULL once the call has returned  */
};

/* reference: src/spasm.h:74-82 (only allocated/freed here) */
struct spasm_dm {
	int *p;
	int *q;
	int *r;
	int *c;
	int nb;
	int rr[5];
	int cc[5];
};

/* knobs of spasm_echelonize (reference: src/spasm.h:84-108, defaults in
 * src/spasm_echelonize.c:9-28).  tools/common.c writes these fields directly,
 * so the layout is ABI. */
struct echelonize_opts {
	bool enable_greedy_pivot_search;

	bool enable_tall_and_skinny;
	bool enable_dense;
	bool enable_GPLU;

	bool L;
	bool complete;
	double min_pivot_proportion;
	int max_round;

	double sparsity_threshold;

	int dense_block_size;
	double low_rank_ratio;
	double tall_and_skinny_ratio;
	double low_rank_start_weight;
};

/* reference: src/spasm.h:110-118 */
struct spasm_rank_certificate {
	int r;
	i64 prime;
	u8 hash[32];
	int *i;
	int *j;
	spasm_ZZp *x;
	spasm_ZZp *y;
};

/* reference: src/spasm.h:120-125 */
typedef struct {
	u32 h[8];
	u32 Nl, Nh;        /* message length in bits, low / high word */
	u32 data[16];      /* pending block                            */
	u32 num, md_len;   /* bytes pending; digest length (32)        */
} spasm_sha256_ctx;

/* reference: src/spasm.h:127-135 */
typedef struct {
	u32 block[11];     /* [0:8] seed, [8] prime, [9] counter, [10] sequence (big endian) */
	u32 hash[8];
	u32 prime;
	u32 mask;
	int counter;
	int i;
	spasm_field field;
} spasm_prng_ctx;

/* element type of the dense blocks handed to spasm_ffpack_* (reference: src/spasm.h:137) */
typedef enum {SPASM_DOUBLE, SPASM_FLOAT, SPASM_I64} spasm_datatype;

/* ------------------------------------------------ field arithmetic (host) */
/* reference: src/spasm_ZZp.c */
void spasm_field_init(i64 p, spasm_field F);
spasm_ZZp spasm_ZZp_init(const spasm_field F, i64 x);
spasm_ZZp spasm_ZZp_add(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_sub(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_mul(const spasm_field F, spasm_ZZp a, spasm_ZZp b);
spasm_ZZp spasm_ZZp_inverse(const spasm_field F, spasm_ZZp a);
spasm_ZZp spasm_ZZp_axpy(const spasm_field F, spasm_ZZp a, spasm_ZZp x, spasm_ZZp y);

/* ----------------------------------------------------------- sha256, prng */
/* reference: src/sha256.c, src/spasm_prng.c */
void spasm_SHA256_init(spasm_sha256_ctx *c);
void spasm_SHA256_update(spasm_sha256_ctx *c, const void *data, size_t len);
void spasm_SHA256_final(u8 *md, spasm_sha256_ctx *c);

void spasm_prng_seed(const u8 *seed, i64 prime, u32 seq, spasm_prng_ctx *ctx);
void spasm_prng_seed_simple(i64 prime, u64 seed, u32 seq, spasm_prng_ctx *ctx);
u32 spasm_prng_u32(spasm_prng_ctx *ctx);
spasm_ZZp spasm_prng_ZZp(spasm_prng_ctx *ctx);

/* -------------------------------------------------------------- utilities */
/* reference: src/spasm_util.c */
double spasm_wtime();
i64 spasm_nnz(const struct spasm_csr *A);
void *spasm_malloc(i64 size);
void *spasm_calloc(i64 count, i64 size);
void *spasm_realloc(void *ptr, i64 size);
struct spasm_csr *spasm_csr_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values);
void spasm_csr_realloc(struct spasm_csr *A, i64 nzmax);
void spasm_csr_resize(struct spasm_csr *A, int n, int m);
void spasm_csr_free(struct spasm_csr *A);
struct spasm_triplet *spasm_triplet_alloc(int m, int n, i64 nzmax, i64 prime, bool with_values);
void spasm_triplet_realloc(struct spasm_triplet *A, i64 nzmax);
void spasm_triplet_free(struct spasm_triplet *A);
struct spasm_dm *spasm_dm_alloc(int n, int m);
void spasm_dm_free(struct spasm_dm *P);
void spasm_lu_free(struct spasm_lu *N);
void spasm_human_format(int64_t n, char *target);
int spasm_get_num_threads();
int spasm_get_thread_num();

static inline i64 spasm_get_prime(const struct spasm_csr *A) { return A->field->p; }
static inline int spasm_max(int a, int b) { return (a > b) ? a : b; }
static inline int spasm_min(int a, int b) { return (a < b) ? a : b; }
static inline int spasm_row_weight(const struct spasm_csr *A, int i) { return (int) (A->p[i + 1] - A->p[i]); }

/* --------------------------------------------------- triplets, I/O, moves */
/* reference: src/spasm_triplet.c, src/spasm_io.c, src/spasm_transpose.c, src/spasm_spmv.c */
void spasm_add_entry(struct spasm_triplet *T, int i, int j, i64 x);
void spasm_triplet_transpose(struct spasm_triplet *T);
struct spasm_csr *spasm_compress(const struct spasm_triplet *T);

struct spasm_triplet *spasm_triplet_load(FILE *f, i64 prime, u8 *hash);
void spasm_triplet_save(const struct spasm_triplet *A, FILE *f);
void spasm_csr_save(const struct spasm_csr *A, FILE *f);

struct spasm_csr *spasm_transpose(const struct spasm_csr *C, int keep_values);

void spasm_xApy(const spasm_ZZp *x, const struct spasm_csr *A, spasm_ZZp *y);
void spasm_Axpy(const struct spasm_csr *A, const spasm_ZZp *x, spasm_ZZp *y);

/* ---------------------- small-grain host verifiers (csrc/host/verify.c) */
/* Re-entrant, one row at a time; the reference's test programs call them inside their own OpenMP regions to CHECK a
 * result (tests/echelonize.c:76-113, tests/kernel.c, tests/schur_dense.c).  The echelonization path never does.
 * reference: src/spasm_scatter.c:7, src/spasm_reach.c:21,:98, src/spasm_triangular.c:21,:65,:109,
 * src/spasm_permutation.c, src/spasm_submatrix.c:7, src/spasm_kernel.c:133 */
void spasm_scatter(const struct spasm_csr *A, int i, spasm_ZZp beta, spasm_ZZp *x);
int spasm_dfs(int i, const struct spasm_csr *G, int top, int *xi, int *pstack, int *marks, const int *pinv);
int spasm_reach(const struct spasm_csr *A, const struct spasm_csr *B, int k, int l, int *xj, const int *qinv);
void spasm_dense_back_solve(const struct spasm_csr *L, spasm_ZZp *b, spasm_ZZp *x, const int *p);
bool spasm_dense_forward_solve(const struct spasm_csr *U, spasm_ZZp *b, spasm_ZZp *x, const int *q);
int spasm_sparse_triangular_solve(const struct spasm_csr *U, const struct spasm_csr *B, int k, int *xj, spasm_ZZp *x, const int *qinv);
void spasm_pvec(const int *p, const spasm_ZZp *b, spasm_ZZp *x, int n);
void spasm_ipvec(const int *p, const spasm_ZZp *b, spasm_ZZp *x, int n);
int *spasm_pinv(int const *p, int n);
struct spasm_csr *spasm_permute(const struct spasm_csr *A, const int *p, const int *qinv, int with_values);
int *spasm_random_permutation(int n);
void spasm_range_pvec(int *x, int a, int b, int *p);
struct spasm_csr *spasm_submatrix(const struct spasm_csr *A, int r_0, int r_1, int c_0, int c_1, int with_values);
struct spasm_csr *spasm_kernel_from_rref(const struct spasm_csr *R, const int *qinv);

/* ------------------------------------------------ the hot path (on B200) */

/* reference: src/spasm_pivots.c:369 */
int spasm_pivots_extract_structural(const struct spasm_csr *A, const int *p_in, struct spasm_lu *fact, int *p, struct echelonize_opts *opts);

/* reference: src/spasm_schur.c:11, :61, :257, :346 */
double spasm_schur_estimate_density(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U, const int *qinv, int R);
struct spasm_csr *spasm_schur(const struct spasm_csr *A, const int *p, int n, const struct spasm_lu *fact,
                   double est_density, struct spasm_triplet *L, const int *p_in, int *p_out);
void spasm_schur_dense(const struct spasm_csr *A, const int *p, int n, const int *p_in,
	struct spasm_lu *fact, void *S, spasm_datatype datatype, int *q, int *p_out);
void spasm_schur_dense_randomized(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U, const int *qinv,
	void *S, spasm_datatype datatype, int *q, int N, int w);

/* dense echelon form; replaces the FFLAS-FFPACK wrapper (reference: src/spasm_ffpack.cpp:78-149) */
int spasm_ffpack_rref(i64 prime, int n, int m, void *A, int ldA, spasm_datatype datatype, size_t *qinv);
int spasm_ffpack_LU(i64 prime, int n, int m, void *A, int ldA, spasm_datatype datatype, size_t *p, size_t *qinv);
spasm_ZZp spasm_datatype_read(const void *A, size_t i, spasm_datatype datatype);
void spasm_datatype_write(void *A, size_t i, spasm_datatype datatype, spasm_ZZp value);
size_t spasm_datatype_size(spasm_datatype datatype);
spasm_datatype spasm_datatype_choose(i64 prime);
const char *spasm_datatype_name(spasm_datatype datatype);

/* reference: src/spasm_echelonize.c:9, :473 */
void spasm_echelonize_init_opts(struct echelonize_opts *opts);
struct spasm_lu *spasm_echelonize(const struct spasm_csr *A, struct echelonize_opts *opts);

/* reference: src/spasm_rref.c:22, src/spasm_kernel.c:9 */
struct spasm_csr *spasm_rref(const struct spasm_lu *fact, int *Rqinv);
struct spasm_csr *spasm_kernel(const struct spasm_lu *fact);

/* ------- entry points of the reference that are outside the B200 hot path.
 * They are exported so that the reference tools link; calling one of them
 * aborts with errx(1, "... not part of spasm-b200").   */
bool spasm_solve(const struct spasm_lu *fact, const spasm_ZZp *b, spasm_ZZp *x);
struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok);
struct spasm_rank_certificate *spasm_certificate_rank_create(const struct spasm_csr *A, const u8 *hash, const struct spasm_lu *fact);
bool spasm_certificate_rank_verify(const struct spasm_csr *A, const u8 *hash, const struct spasm_rank_certificate *proof);
void spasm_rank_certificate_save(const struct spasm_rank_certificate *proof, FILE *f);
bool spasm_rank_certificate_load(FILE *f, struct spasm_rank_certificate *proof);
bool spasm_factorization_verify(const struct spasm_csr *A, const struct spasm_lu *fact, u64 seed);

#ifdef __cplusplus
}
#endif
#endif
