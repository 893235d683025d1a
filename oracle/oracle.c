/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Single-threaded CPU restatement of SpaSM's echelonization path, written to
 * be read next to the reference: every function names the reference lines it
 * follows.  It reproduces the *sequential* (OMP_NUM_THREADS=1) semantics of
 * the reference, including the order-defining details (first eligible entry,
 * DFS order of the reach, glibc rand() call sequence), so that its raw output
 * can be compared array-for-array with oracle/_ref/libspasm_ref.so (the
 * reference's own sources compiled in place) -- that comparison, plus the
 * reference's known-answer vectors (tests/Expected/prng, tests/Expected/hash),
 * is what pins this oracle (tests/test_oracle_pinned.py).
 *
 * The dense echelon form is the restated FFPACK boundary of ffpack_restate.c
 * ("parity unpinned" at the bit level, see its header).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this library; nothing under spasm_b200/ does.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include "oracle.h"

static double wtime(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static void *xmalloc(size_t sz)
{
	void *p = malloc(sz ? sz : 1);
	if (!p) {
		fprintf(stderr, "oracle: out of memory\n");
		abort();
	}
	return p;
}

/* ===================================================================== field
 * reference: src/spasm_ZZp.c.  Exact integer arithmetic; the balanced
 * representative is unique so the values equal the reference's. */
static inline i32 zp_bal(i64 p, i64 x)
{
	x %= p;
	i64 half = p / 2, mhalf = p / 2 - p + 1;      /* ZZp.c:12-13 */
	if (x > half)
		x -= p;
	else if (x < mhalf)
		x += p;
	return (i32) x;
}

i32 oracle_zp_mul(i64 p, i32 a, i32 b) { return zp_bal(p, (i64) a * b); }                      /* ZZp.c:42-46 */
i32 oracle_zp_axpy(i64 p, i32 a, i32 x, i32 y) { return zp_bal(p, (i64) a * x + y); }          /* ZZp.c:77-84 */

i32 oracle_zp_inverse(i64 p, i32 a)                                                             /* ZZp.c:49-74 */
{
	i64 r0 = p, r1 = a < 0 ? a + p : a, t0 = 0, t1 = 1;
	while (r1) {
		i64 q = r0 / r1, r2 = r0 - q * r1, t2 = t0 - q * t1;
		r0 = r1; r1 = r2; t0 = t1; t1 = t2;
	}
	return zp_bal(p, t0);
}

/* ================================================================ containers */
struct ocsr *oracle_csr_alloc(int n, int m, i64 nzmax, i64 prime)              /* util.c:90-103 */
{
	struct ocsr *A = xmalloc(sizeof(*A));
	A->n = n;
	A->m = m;
	A->nzmax = nzmax;
	A->prime = prime;
	A->p = xmalloc((size_t) (n + 1) * sizeof(i64));
	A->j = xmalloc((size_t) nzmax * sizeof(int));
	A->x = xmalloc((size_t) nzmax * sizeof(i32));
	A->p[0] = 0;
	return A;
}

void oracle_csr_free(struct ocsr *A)
{
	if (!A)
		return;
	free(A->p);
	free(A->j);
	free(A->x);
	free(A);
}

static void csr_reserve(struct ocsr *A, i64 nzmax)
{
	if (nzmax <= A->nzmax)
		return;
	A->j = realloc(A->j, (size_t) nzmax * sizeof(int));
	A->x = realloc(A->x, (size_t) nzmax * sizeof(i32));
	A->nzmax = nzmax;
	if (!A->j || !A->x)
		abort();
}

static inline i64 csr_nnz(const struct ocsr *A) { return A->p[A->n]; }

/* triplets (file order) -> CSR.  reference: src/spasm_triplet.c:7-24 (entries reduced mod p, zeros
 * dropped on entry), :99-157 (stable by row), :60-96 (duplicates summed into the first occurrence),
 * :36-57 (cancelled entries removed). */
struct ocsr *oracle_compress(int n, int m, i64 nz, const int *Ti, const int *Tj, const i64 *Tx, i64 prime)
{
	struct ocsr *C = oracle_csr_alloc(n, m, nz, prime);
	i64 *count = calloc((size_t) n + 1, sizeof(i64));
	i32 *val = xmalloc((size_t) nz * sizeof(i32));
	for (i64 k = 0; k < nz; k++) {
		val[k] = zp_bal(prime, Tx[k]);
		if (val[k] != 0)
			count[Ti[k] + 1]++;
	}
	for (int i = 0; i < n; i++)
		count[i + 1] += count[i];
	i64 *where = xmalloc((size_t) (n + 1) * sizeof(i64));
	memcpy(where, count, (size_t) (n + 1) * sizeof(i64));
	int *rawj = xmalloc((size_t) nz * sizeof(int));
	i32 *rawx = xmalloc((size_t) nz * sizeof(i32));
	for (i64 k = 0; k < nz; k++)
		if (val[k] != 0) {
			i64 d = where[Ti[k]]++;
			rawj[d] = Tj[k];
			rawx[d] = val[k];
		}
	i64 *first = xmalloc((size_t) m * sizeof(i64));
	for (int j = 0; j < m; j++)
		first[j] = -1;
	/* pass 1: fold repeated (i,j) into the first occurrence */
	i64 out = 0;
	for (int i = 0; i < n; i++) {
		i64 start = out;
		for (i64 k = count[i]; k < count[i + 1]; k++) {
			int j = rawj[k];
			if (first[j] >= start) {
				C->x[first[j]] = zp_bal(prime, (i64) C->x[first[j]] + rawx[k]);
			} else {
				first[j] = out;
				C->j[out] = j;
				C->x[out] = rawx[k];
				out++;
			}
		}
		C->p[i + 1] = out;
	}
	/* pass 2: drop what cancelled.  QUIRK reproduced on purpose (triplet.c:36-57): the reference
	 * compacts in place and starts the scan of row i at the *already rewritten* p[i] (= the write
	 * cursor), so once an entry has been dropped every later row also re-reads the stale slots
	 * between the cursor and its true start.  This only matters for inputs with repeated (i,j)
	 * whose values cancel mod p; parity with the reference requires the same behaviour. */
	i64 w = 0;
	for (int i = 0; i < n; i++) {
		i64 end = C->p[i + 1];
		for (i64 k = w; k < end; k++)          /* note: k starts at w, not at the old p[i] */
			if (C->x[k] != 0) {
				C->j[w] = C->j[k];
				C->x[w] = C->x[k];
				w++;
			}
		C->p[i + 1] = w;
	}
	free(count); free(val); free(where); free(rawj); free(rawx); free(first);
	return C;
}

/* reference: src/spasm_transpose.c:5-52 (entries of a column appear by increasing row) */
struct ocsr *oracle_transpose(const struct ocsr *A)
{
	struct ocsr *T = oracle_csr_alloc(A->m, A->n, csr_nnz(A), A->prime);
	i64 *w = calloc((size_t) A->m + 1, sizeof(i64));
	for (i64 k = 0; k < csr_nnz(A); k++)
		w[A->j[k] + 1]++;
	for (int j = 0; j < A->m; j++)
		w[j + 1] += w[j];
	memcpy(T->p, w, (size_t) (A->m + 1) * sizeof(i64));
	for (int i = 0; i < A->n; i++)
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++) {
			i64 d = w[A->j[k]]++;
			T->j[d] = i;
			T->x[d] = A->x[k];
		}
	free(w);
	return T;
}

/* ============================================================ sha256 + prng
 * reference: src/sha256.c (standard FIPS 180-4), src/spasm_prng.c */
static const uint32_t SHA_K[64] = {
	0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
	0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
	0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
	0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
	0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
	0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
	0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

#define ROR(x, s) (((x) >> (s)) | ((x) << (32 - (s))))
static void sha_block(uint32_t H[8], const unsigned char *b)
{
	uint32_t w[64], v[8];
	for (int t = 0; t < 16; t++)
		w[t] = (uint32_t) b[4 * t] << 24 | (uint32_t) b[4 * t + 1] << 16 | (uint32_t) b[4 * t + 2] << 8 | b[4 * t + 3];
	for (int t = 16; t < 64; t++)
		w[t] = w[t - 16] + w[t - 7] + (ROR(w[t - 15], 7) ^ ROR(w[t - 15], 18) ^ (w[t - 15] >> 3))
		     + (ROR(w[t - 2], 17) ^ ROR(w[t - 2], 19) ^ (w[t - 2] >> 10));
	memcpy(v, H, sizeof(v));
	for (int t = 0; t < 64; t++) {
		uint32_t t1 = v[7] + (ROR(v[4], 6) ^ ROR(v[4], 11) ^ ROR(v[4], 25)) + ((v[4] & v[5]) ^ (~v[4] & v[6])) + SHA_K[t] + w[t];
		uint32_t t2 = (ROR(v[0], 2) ^ ROR(v[0], 13) ^ ROR(v[0], 22)) + ((v[0] & v[1]) ^ (v[0] & v[2]) ^ (v[1] & v[2]));
		memmove(v + 1, v, 7 * sizeof(uint32_t));
		v[4] += t1;
		v[0] = t1 + t2;
	}
	for (int t = 0; t < 8; t++)
		H[t] += v[t];
}

void oracle_sha256(const void *data, i64 len, unsigned char out[32])
{
	uint32_t H[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
	const unsigned char *in = data;
	i64 left = len;
	for (; left >= 64; left -= 64, in += 64)
		sha_block(H, in);
	unsigned char tail[128] = {0};
	memcpy(tail, in, (size_t) left);
	tail[left] = 0x80;
	int blocks = (left < 56) ? 1 : 2;
	uint64_t bits = (uint64_t) len * 8;
	for (int b = 0; b < 8; b++)
		tail[64 * blocks - 1 - b] = (unsigned char) (bits >> (8 * b));
	for (int b = 0; b < blocks; b++)
		sha_block(H, tail + 64 * b);
	for (int t = 0; t < 8; t++)
		for (int b = 0; b < 4; b++)
			out[4 * t + b] = (unsigned char) (H[t] >> (24 - 8 * b));
}

/* the stream of prng.c: SHA256(seed[32] | prime | counter | seq) in counter mode, big-endian words,
 * masked rejection sampling (prng.c:21-40, :45-73) */
struct prng {
	unsigned char block[44];
	unsigned char digest[32];
	uint32_t counter, prime, mask;
	int pos;
	i64 p;
};

static void put_be32(unsigned char *dst, uint32_t v)
{
	dst[0] = v >> 24; dst[1] = v >> 16; dst[2] = v >> 8; dst[3] = v;
}

static void prng_refill(struct prng *g)
{
	oracle_sha256(g->block, 44, g->digest);
	g->counter += 1;
	put_be32(g->block + 36, g->counter);
	g->pos = 0;
}

static void prng_seed(struct prng *g, i64 prime, uint64_t seed, uint32_t seq)      /* prng.c:66-73 then :45-61 */
{
	memset(g->block, 0, 44);
	put_be32(g->block, (uint32_t) (seed & 0xffffffff));
	put_be32(g->block + 4, (uint32_t) (seed >> 32));
	put_be32(g->block + 32, (uint32_t) prime);
	put_be32(g->block + 40, seq);
	g->counter = 0;
	g->prime = (uint32_t) prime;
	g->p = prime;
	i64 pow2 = 1;
	while (pow2 < prime)
		pow2 <<= 1;
	g->mask = (uint32_t) (pow2 - 1);
	prng_refill(g);
}

static i32 prng_zp(struct prng *g)
{
	for (;;) {
		if (g->pos == 8)
			prng_refill(g);
		const unsigned char *d = g->digest + 4 * g->pos++;
		uint32_t v = ((uint32_t) d[0] << 24 | (uint32_t) d[1] << 16 | (uint32_t) d[2] << 8 | d[3]) & g->mask;
		if (v < g->prime)
			return zp_bal(g->p, v);
	}
}

void oracle_prng_stream(i64 prime, uint64_t seed, uint32_t seq, int count, i32 *out)
{
	struct prng g;
	prng_seed(&g, prime, seed, seq);
	for (int k = 0; k < count; k++)
		out[k] = prng_zp(&g);
}

/* ===================================================== sparse triangular solve
 * Workspace of one solver: x[m] values, order[m] output pattern (filled from the end),
 * stack/next for the DFS, mark[m]. */
struct solver {
	int m;
	i32 *x;
	int *order, *stack, *next;
	char *mark;
};

static struct solver *solver_new(int m)
{
	struct solver *s = xmalloc(sizeof(*s));
	s->m = m;
	s->x = calloc((size_t) m + 1, sizeof(i32));
	s->order = xmalloc((size_t) (m + 1) * sizeof(int));
	s->stack = xmalloc((size_t) (m + 1) * sizeof(int));
	s->next = xmalloc((size_t) (m + 1) * sizeof(int));
	s->mark = calloc((size_t) m + 1, 1);
	return s;
}

static void solver_free(struct solver *s)
{
	free(s->x); free(s->order); free(s->stack); free(s->next); free(s->mark); free(s);
}

/* x += beta * A[i]   (reference: src/spasm_scatter.c:7-15) */
static inline void scatter(const struct ocsr *A, int i, i32 beta, i32 *x)
{
	for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
		x[A->j[k]] = oracle_zp_axpy(A->prime, beta, A->x[k], x[A->j[k]]);
}

/* Depth-first search from column `start` along column -> pivot row -> its columns.
 * A column is emitted (written at order[--top]) when all columns of its pivot row have been
 * emitted, or at once if it is not pivotal: reverse post-order = topological order.
 * reference: src/spasm_reach.c:21-82 (same visiting order: row entries are tried in CSR order). */
static int dfs(const struct ocsr *G, const int *qinv, int start, int top, struct solver *s)
{
	int depth = 0;
	s->stack[0] = start;
	s->mark[start] = 1;
	s->next[0] = 0;
	while (depth >= 0) {
		int j = s->stack[depth];
		int row = qinv[j];
		int pushed = 0;
		if (row >= 0) {
			i64 base = G->p[row];
			int len = (int) (G->p[row + 1] - base);
			for (int k = s->next[depth]; k < len; k++) {
				int c = G->j[base + k];
				if (s->mark[c])
					continue;
				s->next[depth] = k + 1;
				depth++;
				s->stack[depth] = c;
				s->mark[c] = 1;
				s->next[depth] = 0;
				pushed = 1;
				break;
			}
		}
		if (!pushed) {
			s->order[--top] = j;
			depth--;
		}
	}
	return top;
}

/* pattern of the solution of x*U = B[k]: order[top:m], marks cleared on exit.
 * reference: src/spasm_reach.c:98-135 */
static int reach(const struct ocsr *U, const struct ocsr *B, int k, const int *qinv, struct solver *s)
{
	int top = s->m;
	for (i64 e = B->p[k]; e < B->p[k + 1]; e++)
		if (!s->mark[B->j[e]])
			top = dfs(U, qinv, B->j[e], top, s);
	for (int t = top; t < s->m; t++)
		s->mark[s->order[t]] = 0;
	return top;
}

/* reference: src/spasm_triangular.c:109-146.  Returns top; pattern in s->order[top:m], values in s->x.
 * `touched` accumulates the number of U entries scattered (algorithmic-bytes instrumentation). */
static int tsolve(const struct ocsr *U, const struct ocsr *B, int k, const int *qinv, struct solver *s, i64 *touched)
{
	int top = reach(U, B, k, qinv, s);
	for (int t = top; t < s->m; t++)
		s->x[s->order[t]] = 0;
	scatter(B, k, 1, s->x);
	for (int t = top; t < s->m; t++) {
		int j = s->order[t];
		int row = qinv[j];
		if (row < 0)
			continue;
		i32 keep = s->x[j];
		scatter(U, row, -keep, s->x);
		s->x[j] = keep;                      /* the eliminated coordinate keeps its L coefficient */
		if (touched)
			*touched += U->p[row + 1] - U->p[row];
	}
	return top;
}

int oracle_tsolve(const struct ocsr *U, const struct ocsr *B, int k, int *xj, i32 *x, const int *qinv)
{
	struct solver *s = solver_new(U->m);
	int top = tsolve(U, B, k, qinv, s, NULL);
	memcpy(x, s->x, (size_t) U->m * sizeof(i32));
	memcpy(xj, s->order, (size_t) U->m * sizeof(int));
	solver_free(s);
	return top;
}

/* ============================================================= pivot search */

/* Make (i,j) a pivot, evicting whatever row i / column j were matched to.
 * Returns 1 when nothing was evicted.  reference: src/spasm_pivots.c:11-32 */
static int set_pivot(int i, int j, int *pinv, int *qinv)
{
	int fresh = 1;
	if (pinv[i] != -1) {
		qinv[pinv[i]] = -1;
		fresh = 0;
	}
	if (qinv[j] != -1) {
		pinv[qinv[j]] = -1;
		fresh = 0;
	}
	pinv[i] = j;
	qinv[j] = i;
	return fresh;
}

/* Faugere-Lachartre: the leftmost entry of a row is a candidate; per column keep the sparsest row,
 * the earlier one on ties.  reference: src/spasm_pivots.c:41-66 */
static int pivots_FL(const struct ocsr *A, int *pinv, int *qinv)
{
	int found = 0;
	for (int i = 0; i < A->n; i++) {
		if (A->p[i] == A->p[i + 1])
			continue;
		int lead = A->j[A->p[i]];
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
			if (A->j[k] < lead)
				lead = A->j[k];
		int holder = qinv[lead];
		if (holder == -1 || (A->p[i + 1] - A->p[i]) < (A->p[holder + 1] - A->p[holder]))
			found += set_pivot(i, lead, pinv, qinv);
	}
	return found;
}

/* Columns that occur in no pivotal row can be matched greedily, in row order, first open entry
 * first; a matched row closes all its columns.  reference: src/spasm_pivots.c:76-122 */
static int pivots_FL_columns(const struct ocsr *A, int *pinv, int *qinv)
{
	char *open = xmalloc((size_t) A->m + 1);
	memset(open, 1, (size_t) A->m + 1);
	for (int i = 0; i < A->n; i++)
		if (pinv[i] >= 0)
			for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
				open[A->j[k]] = 0;
	int found = 0;
	for (int i = 0; i < A->n; i++) {
		if (pinv[i] >= 0)
			continue;
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++) {
			int j = A->j[k];
			if (!open[j] || qinv[j] >= 0)
				continue;
			found += set_pivot(i, j, pinv, qinv);
			for (i64 e = A->p[i]; e < A->p[i + 1]; e++)
				open[A->j[e]] = 0;
			break;
		}
	}
	free(open);
	return found;
}

/* Greedy search for pivots that create no alternating cycle, rows in increasing order, every row
 * seeing all pivots chosen before it (= the reference with one thread, where every transaction
 * commits at once).  state[j]: 0 untouched, 1 candidate entry of the row, -1 reached.
 * reference: src/spasm_pivots.c:146-294 (BFS :218-225, first survivor in row order :233-237) */
static int pivots_greedy(const struct ocsr *A, int *pinv, int *qinv, i64 *edges)
{
	signed char *state = calloc((size_t) A->m + 1, 1);
	int *queue = xmalloc((size_t) (A->m + 1) * sizeof(int));
	int found = 0;
	for (int i = 0; i < A->n; i++) {
		if (pinv[i] >= 0)
			continue;
		int head = 0, tail = 0, alive = 0;
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++) {
			int j = A->j[k];
			if (qinv[j] < 0) {
				state[j] = 1;
				alive++;
			} else {
				queue[tail++] = j;
				alive -= state[j];        /* state[j] is 0 here (no duplicate columns in a row) */
				state[j] = -1;
			}
		}
		while (head < tail && alive > 0) {
			int row = qinv[queue[head++]];
			if (row < 0)
				continue;
			for (i64 k = A->p[row]; k < A->p[row + 1]; k++) {
				int j = A->j[k];
				if (state[j] >= 0) {
					queue[tail++] = j;
					alive -= state[j];
					state[j] = -1;
				}
			}
			*edges += A->p[row + 1] - A->p[row];
		}
		if (alive > 0) {
			/* first entry still marked 1.  QUIRK kept (pivots.c:232-238): if none is (possible only when
			 * the row holds a repeated column, which inflates `alive`), the LAST entry of the row is
			 * taken, and set_pivot evicts whoever held that column. */
			int j = -1;
			for (i64 k = A->p[i]; k < A->p[i + 1]; k++) {
				j = A->j[k];
				if (state[j] == 1)
					break;
			}
			found += set_pivot(i, j, pinv, qinv);
		}
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
			state[A->j[k]] = 0;
		for (int t = 0; t < tail; t++)
			state[queue[t]] = 0;
	}
	free(state);
	free(queue);
	return found;
}

/* reference: src/spasm_pivots.c:305-319 */
int oracle_pivots_find(const struct ocsr *A, int greedy, int *pinv, int *qinv, int counts[3], i64 *edges)
{
	for (int j = 0; j < A->m; j++)
		qinv[j] = -1;
	for (int i = 0; i < A->n; i++)
		pinv[i] = -1;
	i64 e = 0;
	counts[0] = pivots_FL(A, pinv, qinv);
	counts[1] = pivots_FL_columns(A, pinv, qinv);
	counts[2] = greedy ? pivots_greedy(A, pinv, qinv, &e) : 0;
	if (edges)
		*edges = e;
	return counts[0] + counts[1] + counts[2];
}

/* pivotal rows first, in an order that makes U triangular (DFS from columns 0..m-1), then the others
 * by increasing index.  reference: src/spasm_pivots.c:325-362 */
static void pivots_order_rows(const struct ocsr *A, const int *pinv, const int *qinv, int npiv, int *p)
{
	struct solver *s = solver_new(A->m);
	int top = A->m;
	for (int j = 0; j < A->m; j++)
		if (qinv[j] != -1 && !s->mark[j])
			top = dfs(A, qinv, j, top, s);
	int k = 0;
	for (int t = top; t < A->m; t++)
		if (qinv[s->order[t]] != -1)
			p[k++] = qinv[s->order[t]];
	assert(k == npiv);
	for (int i = 0; i < A->n; i++)
		if (pinv[i] == -1)
			p[k++] = i;
	assert(k == A->n);
	solver_free(s);
}

/* ================================================================ echelonize */
struct state {
	struct oracle_lu *out;
	struct ocsr *U;
	int *qinv;          /* Uqinv */
	i64 prime;
	int m;
};

/* Find structural pivots of A, append the pivotal rows to U with the pivot first and scaled to 1.
 * p receives the row permutation (pivotal rows first).  reference: src/spasm_pivots.c:369-448 */
static int extract_structural(struct state *st, const struct ocsr *A, const int *p_in, int *p, const struct oracle_opts *opts)
{
	struct oracle_lu *out = st->out;
	int *qinv = xmalloc((size_t) (A->m + 1) * sizeof(int));
	int *pinv = xmalloc((size_t) (A->n + 1) * sizeof(int));
	int counts[3];
	i64 edges = 0;
	int npiv = oracle_pivots_find(A, opts->enable_greedy_pivot_search, pinv, qinv, counts, &edges);
	out->greedy_edges += edges;
	pivots_order_rows(A, pinv, qinv, npiv, p);

	int round = out->nrounds;
	if (round < ORACLE_MAX_ROUNDS) {
		out->found_FL[round] = counts[0];
		out->found_FLcol[round] = counts[1];
		out->found_greedy[round] = counts[2];
	}
	out->pair_row = realloc(out->pair_row, (size_t) (out->npairs + npiv + 1) * sizeof(int));
	out->pair_col = realloc(out->pair_col, (size_t) (out->npairs + npiv + 1) * sizeof(int));

	struct ocsr *U = st->U;
	i64 need = csr_nnz(U);
	for (int k = 0; k < npiv; k++)
		need += A->p[p[k] + 1] - A->p[p[k]];
	csr_reserve(U, need);
	i64 unz = csr_nnz(U);
	for (int k = 0; k < npiv; k++) {
		int i = p[k], j = pinv[i];
		out->pair_row[out->npairs] = p_in ? p_in[i] : i;
		out->pair_col[out->npairs] = j;
		out->npairs++;
		st->qinv[j] = U->n;
		i32 pivot = 0;
		for (i64 e = A->p[i]; e < A->p[i + 1]; e++)
			if (A->j[e] == j && A->x[e] != 0) {
				pivot = A->x[e];
				break;
			}
		assert(pivot != 0);
		i32 alpha = oracle_zp_inverse(st->prime, pivot);
		U->j[unz] = j;
		U->x[unz] = 1;
		unz++;
		for (i64 e = A->p[i]; e < A->p[i + 1]; e++)
			if (A->j[e] != j) {
				U->j[unz] = A->j[e];
				U->x[unz] = oracle_zp_mul(st->prime, alpha, A->x[e]);
				unz++;
			}
		U->n++;
		U->p[U->n] = unz;
	}
	free(pinv);
	free(qinv);
	return npiv;
}

/* reference: src/spasm_schur.c:11-44 -- R rows p[rand() % n], mean density on the non-pivotal columns */
double oracle_schur_estimate_density(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, int R)
{
	if (n == 0)
		return 0;
	struct solver *s = solver_new(A->m);
	i64 nnz = 0;
	for (int t = 0; t < R; t++) {
		int row = p[rand() % n];
		int top = tsolve(U, A, row, qinv, s, NULL);
		for (int e = top; e < A->m; e++) {
			int j = s->order[e];
			if (qinv[j] < 0 && s->x[j] != 0)
				nnz++;
		}
	}
	solver_free(s);
	return ((double) nnz) / (A->m - U->n) / R;
}

static struct oracle_lu *g_stats;       /* instrumentation sink of the call in progress */

/* S = rows p[0:n] of A reduced by U, entries on non-pivotal columns, in pattern order.
 * reference: src/spasm_schur.c:61-193 (one thread: rows arrive in p order) */
struct ocsr *oracle_schur(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, double est_density)
{
	int m = A->m;
	if (est_density < 0)
		est_density = oracle_schur_estimate_density(A, p, n, U, qinv, 100);
	i64 cap = (i64) ((est_density * n) * m);
	struct ocsr *S = oracle_csr_alloc(n, m, cap > 0 ? cap : 1, A->prime);
	struct solver *s = solver_new(m);
	i64 snz = 0;
	for (int k = 0; k < n; k++) {
		i64 touched = 0;
		int top = tsolve(U, A, p[k], qinv, s, &touched);
		if (snz + m > S->nzmax)
			csr_reserve(S, 2 * S->nzmax + m);
		i64 before = snz;
		for (int e = top; e < m; e++) {
			int j = s->order[e];
			if (s->x[j] != 0 && qinv[j] < 0) {
				S->j[snz] = j;
				S->x[snz] = s->x[j];
				snz++;
			}
		}
		S->p[k + 1] = snz;
		if (g_stats) {
			g_stats->tsolve_bytes += 8.0 * (A->p[p[k] + 1] - A->p[p[k]]) + 8.0 * touched + 8.0 * (snz - before);
			g_stats->tsolve_rows++;
		}
	}
	solver_free(s);
	return S;
}

/* list of non-pivotal columns, increasing (reference: src/spasm_schur.c:195-203) */
static int nonpivotal_columns(int m, const int *qinv, int *q)
{
	int k = 0;
	for (int j = 0; j < m; j++)
		if (qinv[j] < 0)
			q[k++] = j;
	return k;
}

/* dense block: row k = A[p[k]] reduced by U, gathered through q.  S is n x Sm, row-major int32.
 * reference: src/spasm_schur.c:257-333 */
void oracle_schur_dense(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, i32 *S, int *q)
{
	int m = A->m;
	int Sm = nonpivotal_columns(m, qinv, q);
	struct solver *s = solver_new(m);
	for (int k = 0; k < n; k++) {
		memset(s->x, 0, (size_t) m * sizeof(i32));
		i64 touched = 0;
		tsolve(U, A, p[k], qinv, s, &touched);
		for (int c = 0; c < Sm; c++)
			S[(size_t) k * Sm + c] = s->x[q[c]];
		if (g_stats) {
			g_stats->tsolve_bytes += 8.0 * (A->p[p[k] + 1] - A->p[p[k]]) + 8.0 * touched + 4.0 * Sm;
			g_stats->tsolve_rows++;
		}
	}
	solver_free(s);
}

/* N random combinations of the rows p[0:n] (w rows each, or all of them when w <= 0), reduced by
 * looping over the rows of U in index order (pivot = first entry), gathered through q.
 * rand() is called w times per output row, output rows in increasing order.
 * reference: src/spasm_schur.c:346-413 */
void oracle_schur_dense_randomized(const struct ocsr *A, const int *p, int n, const struct ocsr *U, const int *qinv, i32 *S, int *q, int N, int w)
{
	int m = A->m;
	int Sm = nonpivotal_columns(m, qinv, q);
	i32 *y = xmalloc((size_t) (m + 1) * sizeof(i32));
	for (int k = 0; k < N; k++) {
		struct prng g;
		prng_seed(&g, A->prime, (uint64_t) k, 0);
		memset(y, 0, (size_t) m * sizeof(i32));
		i64 touched = 0, read = 0;
		if (w <= 0) {
			for (int i = 0; i < n; i++) {
				i32 c = prng_zp(&g);
				scatter(A, p[i], c, y);
				read += A->p[p[i] + 1] - A->p[p[i]];
			}
		} else {
			for (int t = 0; t < w; t++) {
				int row = p[rand() % n];
				i32 c = (t == 0) ? 1 : prng_zp(&g);
				scatter(A, row, c, y);
				read += A->p[row + 1] - A->p[row];
			}
		}
		for (int i = 0; i < U->n; i++) {
			int j = U->j[U->p[i]];
			if (y[j] == 0)
				continue;
			scatter(U, i, -y[j], y);
			touched += U->p[i + 1] - U->p[i];
		}
		for (int c = 0; c < Sm; c++)
			S[(size_t) k * Sm + c] = y[q[c]];
		if (g_stats) {
			g_stats->tsolve_bytes += 8.0 * read + 8.0 * touched + 4.0 * Sm;
			g_stats->tsolve_rows++;
		}
	}
	free(y);
}

/* dense RREF of an n x Sm int32 block through the restated FFPACK boundary; pivot columns first in
 * Sqinv, then the others (ffpack_restate.c).  On return rows 0..rr-1 of S hold the RREF rows (all Sm columns). */
static int dense_rref(struct state *st, int n, int Sm, i32 *S, int *Sqinv)
{
	int *pivcol = xmalloc((size_t) (Sm + 1) * sizeof(int));
	int rr = oracle_dense_rref_i32(st->prime, n, Sm, S, pivcol);
	char *isp = calloc((size_t) Sm + 1, 1);
	for (int i = 0; i < rr; i++) {
		Sqinv[i] = pivcol[i];
		isp[pivcol[i]] = 1;
	}
	int k = rr;
	for (int j = 0; j < Sm; j++)
		if (!isp[j])
			Sqinv[k++] = j;
	free(pivcol);
	free(isp);
	st->out->dense_fieldops += 2.0 * n * (double) Sm * rr;
	return rr;
}

/* append the rr echelonized rows to U: pivot (=1) first, then the non-zero entries on the non-pivot
 * columns, mapped back through q.  reference: src/spasm_echelonize.c:192-223 */
static void append_dense_rows(struct state *st, int rr, int Sm, const i32 *S, const int *Sqinv, const int *q)
{
	struct ocsr *U = st->U;
	i64 unz = csr_nnz(U);
	csr_reserve(U, unz + (i64) (1 + Sm - rr) * rr);
	for (int i = 0; i < rr; i++) {
		int pc = q[Sqinv[i]];
		U->j[unz] = pc;
		U->x[unz] = 1;
		unz++;
		st->qinv[pc] = U->n;
		for (int k = rr; k < Sm; k++) {
			i32 v = S[(size_t) i * Sm + Sqinv[k]];
			if (v == 0)
				continue;
			U->j[unz] = q[Sqinv[k]];
			U->x[unz] = v;
			unz++;
		}
		U->n++;
		U->p[U->n] = unz;
	}
}

static void record_block(struct oracle_lu *out, int Sn, int Sm, int rr, int w)
{
	if (out->nblocks < ORACLE_MAX_BLOCKS) {
		out->block_Sn[out->nblocks] = Sn;
		out->block_Sm[out->nblocks] = Sm;
		out->block_rr[out->nblocks] = rr;
		out->block_w[out->nblocks] = w;
	}
	out->nblocks++;
}

/* ceil(128 / log2 p) full random combinations must all reduce to zero.
 * reference: src/spasm_echelonize.c:30-51 */
static int test_completion(struct state *st, const struct ocsr *A, const int *p, int n)
{
	if (n == 0 || csr_nnz(A) == 0)
		return 1;
	int Sm = st->m - st->U->n;
	int Sn = (int) ceil(128 / log2((double) st->prime));
	i32 *S = xmalloc((size_t) Sn * (Sm + 1) * sizeof(i32));
	int *q = xmalloc((size_t) (st->m + 1) * sizeof(int));
	int *Sqinv = xmalloc((size_t) (Sm + 1) * sizeof(int));
	oracle_schur_dense_randomized(A, p, n, st->U, st->qinv, S, q, Sn, 0);
	int rr = dense_rref(st, Sn, Sm, S, Sqinv);
	record_block(st->out, Sn, Sm, rr, 0);
	free(S); free(q); free(Sqinv);
	return rr == 0;
}

/* reference: src/spasm_echelonize.c:315-379 */
static void finish_lowrank(struct state *st, const struct ocsr *A, const int *p, int n, const struct oracle_opts *opts)
{
	int m = st->m;
	int Sm = m - st->U->n;
	i32 *S = xmalloc((size_t) opts->dense_block_size * (Sm + 1) * sizeof(i32));
	int *q = xmalloc((size_t) (m + 1) * sizeof(int));
	int *Sqinv = xmalloc((size_t) (Sm + 1) * sizeof(int));
	int rank_ub = n < Sm ? n : Sm;
	int w = (opts->low_rank_start_weight < 0) ? (int) ceil(-log(0.01) * n / rank_ub) : (int) opts->low_rank_start_weight;
	for (;;) {
		int Sn = rank_ub < opts->dense_block_size ? rank_ub : opts->dense_block_size;
		if (Sn <= 0)
			break;
		oracle_schur_dense_randomized(A, p, n, st->U, st->qinv, S, q, Sn, w);
		int rr = dense_rref(st, Sn, Sm, S, Sqinv);
		record_block(st->out, Sn, Sm, rr, w);
		if (rr == 0) {
			if (test_completion(st, A, p, n))
				break;
			w = 0;
			Sn = 1;           /* the reference stores omp_get_max_threads() here; only the test below reads it */
		}
		if (rr < 0.9 * Sn)
			w *= 2;
		append_dense_rows(st, rr, Sm, S, Sqinv, q);
		n -= rr;              /* sic: the sampling range shrinks (echelonize.c:369) */
		Sm -= rr;
		rank_ub -= rr;
	}
	free(S); free(q); free(Sqinv);
}

/* reference: src/spasm_echelonize.c:385-463 */
static void finish_dense(struct state *st, const struct ocsr *A, const int *p, int n, const struct oracle_opts *opts)
{
	int m = st->m;
	int Sm = m - st->U->n;
	i32 *S = xmalloc((size_t) opts->dense_block_size * (Sm + 1) * sizeof(i32));
	int *q = xmalloc((size_t) (m + 1) * sizeof(int));
	int *Sqinv = xmalloc((size_t) (Sm + 1) * sizeof(int));
	int processed = 0, lowrank = 0;
	int rank_ub = (A->n - st->U->n < A->m - st->U->n) ? A->n - st->U->n : A->m - st->U->n;
	for (;;) {
		int Sn = (opts->dense_block_size < n - processed) ? opts->dense_block_size : n - processed;
		if (Sn <= 0)
			break;
		oracle_schur_dense(A, p, Sn, st->U, st->qinv, S, q);
		int rr = dense_rref(st, Sn, Sm, S, Sqinv);
		record_block(st->out, Sn, Sm, rr, -1);
		append_dense_rows(st, rr, Sm, S, Sqinv, q);
		processed += Sn;
		p += Sn;
		Sm = m - st->U->n;
		rank_ub = (A->n - st->U->n < A->m - st->U->n) ? A->n - st->U->n : A->m - st->U->n;
		if (opts->enable_tall_and_skinny && rr < opts->low_rank_ratio * Sn) {
			lowrank = 1;
			break;
		}
	}
	free(S); free(q); free(Sqinv);
	if (rank_ub > 0 && n - processed > 0 && lowrank)
		finish_lowrank(st, A, p, n - processed, opts);
}

/* row-by-row elimination, leftmost pivot; reference: src/spasm_echelonize.c:54-187 (L == NULL branch) */
static void finish_GPLU(struct state *st, const struct ocsr *A, const int *p, int n)
{
	int m = st->m;
	struct ocsr *U = st->U;
	int r = A->n < m ? A->n : m;
	struct solver *s = solver_new(m);
	int since_pivot = 0, abort_tested = 0;
	for (int i = 0; i < n; i++) {
		if (U->n == r)
			break;
		if (!abort_tested && since_pivot > 10 && since_pivot > n / 100) {
			if (test_completion(st, A, p, n))
				break;
			abort_tested = 1;
		}
		since_pivot++;
		i64 unz = csr_nnz(U);
		if (unz + m > U->nzmax)
			csr_reserve(U, 2 * U->nzmax + m);
		int top = tsolve(U, A, p[i], st->qinv, s, NULL);
		int jpiv = m;
		for (int e = top; e < m; e++) {
			int j = s->order[e];
			if (s->x[j] != 0 && st->qinv[j] < 0 && j < jpiv)
				jpiv = j;
		}
		if (jpiv == m)
			continue;
		st->qinv[jpiv] = U->n;
		U->j[unz] = jpiv;
		U->x[unz] = 1;
		unz++;
		i32 beta = oracle_zp_inverse(st->prime, s->x[jpiv]);
		for (int e = top; e < m; e++) {
			int j = s->order[e];
			if (s->x[j] != 0 && st->qinv[j] < 0) {
				U->j[unz] = j;
				U->x[unz] = oracle_zp_mul(st->prime, beta, s->x[j]);
				unz++;
			}
		}
		U->n++;
		U->p[U->n] = unz;
		since_pivot = 0;
		abort_tested = 0;
	}
	solver_free(s);
}

void oracle_default_opts(struct oracle_opts *o)        /* reference: src/spasm_echelonize.c:9-28 */
{
	o->enable_greedy_pivot_search = 1;
	o->enable_tall_and_skinny = 1;
	o->enable_dense = 1;
	o->enable_GPLU = 1;
	o->min_pivot_proportion = 0.1;
	o->max_round = 3;
	o->sparsity_threshold = 0.05;
	o->tall_and_skinny_ratio = 5;
	o->dense_block_size = 1000;
	o->low_rank_ratio = 0.5;
	o->low_rank_start_weight = -1;
}

/* reference: src/spasm_echelonize.c:473-617 (opts->L == 0) */
struct oracle_lu *oracle_echelonize(const struct ocsr *A0, const struct oracle_opts *opts_in)
{
	struct oracle_opts defaults;
	if (!opts_in) {
		oracle_default_opts(&defaults);
		opts_in = &defaults;
	}
	const struct oracle_opts *opts = opts_in;
	double t_start = wtime();
	int n = A0->n, m = A0->m;
	struct oracle_lu *out = calloc(1, sizeof(*out));
	g_stats = out;
	struct state st;
	st.out = out;
	st.prime = A0->prime;
	st.m = m;
	st.U = oracle_csr_alloc(n, m, csr_nnz(A0) > 0 ? csr_nnz(A0) : 1, A0->prime);
	st.U->n = 0;
	st.qinv = xmalloc((size_t) (m + 1) * sizeof(int));
	for (int j = 0; j < m; j++)
		st.qinv[j] = -1;

	const struct ocsr *A = A0;
	int *p = xmalloc((size_t) (n + 1) * sizeof(int));
	int *p_in = NULL;
	double density = (double) csr_nnz(A) / n / m;
	int npiv = 0, status = 0, round;
	for (round = 0; round < opts->max_round; round++) {
		if (csr_nnz(A) == 0) {
			status = 1;
			break;
		}
		double t0 = wtime();
		out->pair_start[out->nrounds] = out->npairs;
		npiv = extract_structural(&st, A, p_in, p, opts);
		out->nrounds++;
		out->pair_start[out->nrounds] = out->npairs;
		out->seconds_pivots += wtime() - t0;
		int Sm = m - st.U->n;
		if (npiv < opts->min_pivot_proportion * (n < Sm ? n : Sm)) {
			status = 2;
			break;
		}
		t0 = wtime();
		density = oracle_schur_estimate_density(A, p + npiv, n - npiv, st.U, st.qinv, 100);
		if (out->nrounds <= ORACLE_MAX_ROUNDS)
			out->density[out->nrounds - 1] = density;
		if (density > opts->sparsity_threshold) {
			out->seconds_schur += wtime() - t0;
			status = 2;
			break;
		}
		int *p_out = xmalloc((size_t) (n - npiv + 1) * sizeof(int));
		struct ocsr *S = oracle_schur(A, p + npiv, n - npiv, st.U, st.qinv, density);
		for (int k = 0; k < n - npiv; k++) {
			int row = p[npiv + k];
			p_out[k] = p_in ? p_in[row] : row;
		}
		out->seconds_schur += wtime() - t0;
		if (round > 0)
			oracle_csr_free((struct ocsr *) A);
		A = S;
		n -= npiv;
		free(p_in);
		p_in = p_out;
	}
	if (status == 0) {
		npiv = 0;
		for (int i = 0; i < n; i++)
			p[i] = i;
	}
	if (status != 1) {
		double t0 = wtime();
		double aspect = (double) (n - npiv) / (m - st.U->n);
		if (opts->enable_tall_and_skinny && aspect > opts->tall_and_skinny_ratio) {
			out->finish = 1;
			finish_lowrank(&st, A, p + npiv, n - npiv, opts);
		} else if (opts->enable_dense && density > opts->sparsity_threshold) {
			out->finish = 2;
			finish_dense(&st, A, p + npiv, n - npiv, opts);
		} else if (opts->enable_GPLU) {
			out->finish = 3;
			finish_GPLU(&st, A, p + npiv, n - npiv);
		}
		out->seconds_dense += wtime() - t0;
	}
	free(p);
	free(p_in);
	if (round > 0 && A != A0)
		oracle_csr_free((struct ocsr *) A);
	out->U = st.U;
	out->qinv = st.qinv;
	out->rank = st.U->n;
	out->seconds_total = wtime() - t_start;
	g_stats = NULL;
	return out;
}

void oracle_lu_free(struct oracle_lu *f)
{
	if (!f)
		return;
	oracle_csr_free(f->U);
	free(f->qinv);
	free(f->pair_row);
	free(f->pair_col);
	free(f);
}

/* ============================================================ rref and kernel */

/* R = RREF of A*Q from an echelon form: each row of U is solved against U with its own pivot
 * unregistered; entries on non-pivotal columns are kept; pivot moved to the front.
 * reference: src/spasm_rref.c:22-146 */
struct ocsr *oracle_rref(const struct ocsr *U, const int *Uqinv, int *Rqinv)
{
	int n = U->n, m = U->m;
	struct ocsr *R = oracle_csr_alloc(n, m, csr_nnz(U) > 0 ? csr_nnz(U) : 1, U->prime);
	struct solver *s = solver_new(m);
	int *qinv = xmalloc((size_t) (m + 1) * sizeof(int));
	memcpy(qinv, Uqinv, (size_t) m * sizeof(int));
	i64 nnz = 0;
	for (int i = 0; i < n; i++) {
		int pivot = U->j[U->p[i]];
		qinv[pivot] = -1;
		i64 touched = 0;
		int top = tsolve(U, U, i, qinv, s, &touched);
		for (int e = top + 1; e < m; e++)
			if (s->order[e] == pivot) {
				s->order[e] = s->order[top];
				s->order[top] = pivot;
				break;
			}
		if (nnz + m > R->nzmax)
			csr_reserve(R, 2 * R->nzmax + m);
		for (int e = top; e < m; e++) {
			int j = s->order[e];
			if (qinv[j] < 0 && s->x[j] != 0) {
				R->j[nnz] = j;
				R->x[nnz] = s->x[j];
				nnz++;
			}
		}
		R->p[i + 1] = nnz;
		qinv[pivot] = i;
	}
	for (int j = 0; j < m; j++)
		Rqinv[j] = -1;
	for (int i = 0; i < n; i++)
		Rqinv[R->j[R->p[i]]] = i;
	solver_free(s);
	free(qinv);
	return R;
}

/* basis of the right kernel: one vector per non-pivotal column j, (j, -1) first, then the solution of
 * x * Ut = Ut[j] expressed on the pivot columns.  reference: src/spasm_kernel.c:9-127 */
struct ocsr *oracle_kernel(const struct ocsr *U, const int *qinv)
{
	int n = U->n, m = U->m;
	struct ocsr *Ut = oracle_transpose(U);
	struct ocsr *K = oracle_csr_alloc(m - n, m, csr_nnz(U) > 0 ? csr_nnz(U) : 1, U->prime);
	int *Utqinv = xmalloc((size_t) (n + 1) * sizeof(int));
	for (int j = 0; j < m; j++)
		if (qinv[j] >= 0)
			Utqinv[qinv[j]] = j;
	struct solver *s = solver_new(n > 0 ? n : 1);
	i64 nnz = 0;
	int Kn = 0;
	for (int j = 0; j < m; j++) {
		if (qinv[j] >= 0)
			continue;
		int top = tsolve(Ut, Ut, j, Utqinv, s, NULL);
		if (nnz + n + 1 > K->nzmax)
			csr_reserve(K, 2 * K->nzmax + n + 1);
		K->j[nnz] = j;
		K->x[nnz] = -1;
		nnz++;
		for (int e = top; e < n; e++) {
			int jj = s->order[e];
			if (s->x[jj] != 0) {
				K->j[nnz] = Utqinv[jj];
				K->x[nnz] = s->x[jj];
				nnz++;
			}
		}
		Kn++;
		K->p[Kn] = nnz;
	}
	assert(Kn == m - n);
	solver_free(s);
	free(Utqinv);
	oracle_csr_free(Ut);
	return K;
}
