/*
 * Small-grain host entry points of the reference's ABI that its TEST PROGRAMS call to verify a result
 * (SURVEY.md 8b): one-row sparse triangular solves, reachability, scatter, permutations, sub-matrices.
 *
 * These are re-entrant single-row verifiers: the reference's tests call them from inside their own OpenMP
 * regions with per-thread workspaces (tests/echelonize.c:82-111, tests/kernel.c:60-100, tests/schur_dense.c).
 * NOTHING in the echelonization path of this library calls them: spasm_echelonize, spasm_schur*, spasm_rref and
 * spasm_kernel run the batched CUDA solve (csrc/gpu/solve.cu).  They exist so that the reference's unmodified
 * tests link against libspasm_b200.so and check the GPU results with their own row-by-row arithmetic.
 *
 * Written from the documented contracts (workspace sizes, output order, semantics), cited per function.
 */
#include <assert.h>
#include <stdlib.h>
#include "spasm.h"

/* x += beta * A[i]      (reference: src/spasm_scatter.c:7-15) */
void spasm_scatter(const struct spasm_csr *A, int i, spasm_ZZp beta, spasm_ZZp *x)
{
	const i64 lo = A->p[i], hi = A->p[i + 1];
	const int *cols = A->j + lo;
	const spasm_ZZp *vals = A->x + lo;
	for (i64 t = 0; t < hi - lo; t++)
		x[cols[t]] = spasm_ZZp_axpy(A->field, beta, vals[t], x[cols[t]]);
}

/*
 * Depth-first search from column `jstart` along column -> pivot row -> columns of that row
 * (reference: src/spasm_reach.c:21-82).  Contract kept exactly, because callers depend on the ORDER of the output:
 *   - xj[0 .. ) is the recursion stack, xj[top .. m) receives the finished columns (post-order, written downwards),
 *   - pstack[d] = how many entries of the pivot row at depth d were already looked at,
 *   - marks[j] != 0 once column j was entered,
 *   - a row is scanned in storage order and the first unmarked column is entered next.
 */
int spasm_dfs(int jstart, const struct spasm_csr *A, int top, int *xj, int *pstack, int *marks, const int *qinv)
{
	assert(A && xj && pstack && marks && qinv);
	int depth = 0;
	xj[0] = jstart;
	while (depth >= 0) {
		const int col = xj[depth];
		const int owner = qinv[col];
		if (marks[col] == 0) {
			marks[col] = 1;
			pstack[depth] = 0;
		}
		int descend = -1;
		if (owner >= 0) {
			const i64 base = A->p[owner];
			const int len = spasm_row_weight(A, owner);
			int seen = pstack[depth];
			while (seen < len && descend < 0) {
				const int c = A->j[base + seen];
				seen++;
				if (marks[c] == 0)
					descend = c;
			}
			pstack[depth] = seen;
		}
		if (descend >= 0) {
			xj[++depth] = descend;
		} else {
			xj[--top] = col;       /* finished: emit, pop */
			depth--;
		}
	}
	return top;
}

/*
 * Columns reachable from the entries of B[k], in topological order in xj[top:l]
 * (reference: src/spasm_reach.c:98-135).  xj has 3*m ints, zeroed once by the caller; it is left clean.
 */
int spasm_reach(const struct spasm_csr *A, const struct spasm_csr *B, int k, int l, int *xj, const int *qinv)
{
	assert(A && B && xj && qinv);
	const int m = A->m;
	int *pstack = xj + m, *marks = xj + 2 * (size_t) m;
	int top = m;
	for (i64 t = B->p[k]; t < B->p[k + 1]; t++)
		if (marks[B->j[t]] == 0)
			top = spasm_dfs(B->j[t], A, top, xj, pstack, marks, qinv);
	for (int t = top; t < l; t++)
		marks[xj[t]] = 0;
	return top;
}

/*
 * x * U = B[k] with U (permuted) triangular and unit pivots (reference: src/spasm_triangular.c:109-146).
 * On return the pattern is xj[top:m]; entries on pivotal columns hold the multipliers (the value the column had
 * when it was eliminated), entries on non-pivotal columns hold the remainder.
 */
int spasm_sparse_triangular_solve(const struct spasm_csr *U, const struct spasm_csr *B, int k, int *xj, spasm_ZZp *x, const int *qinv)
{
	assert(qinv != NULL);
	const int m = U->m;
	const int top = spasm_reach(U, B, k, m, xj, qinv);
	for (int t = top; t < m; t++)
		x[xj[t]] = 0;
	spasm_scatter(B, k, 1, x);
	for (int t = top; t < m; t++) {
		const int col = xj[t];
		const int row = qinv[col];
		if (row < 0)
			continue;
		const spasm_ZZp keep = x[col];
		spasm_scatter(U, row, -keep, x);
		x[col] = keep;
	}
	return top;
}

/* x.L = b, dense (reference: src/spasm_triangular.c:21-53); b is destroyed */
void spasm_dense_back_solve(const struct spasm_csr *L, spasm_ZZp *b, spasm_ZZp *x, const int *p)
{
	const int n = L->n, r = L->m;
	for (int i = 0; i < n; i++)
		x[i] = 0;
	for (int col = r - 1; col >= 0; col--) {
		const int row = p ? p[col] : col;
		assert(0 <= row && row < n);
		spasm_ZZp diag = 0;
		for (i64 t = L->p[row]; t < L->p[row + 1] && diag == 0; t++)
			if (L->j[t] == col)
				diag = L->x[t];
		assert(diag != 0);
		const spasm_ZZp v = spasm_ZZp_mul(L->field, spasm_ZZp_inverse(L->field, diag), b[col]);
		x[row] = v;
		spasm_scatter(L, row, -v, b);
		x[row] = v;
	}
}

/* x.U = b, dense, U unit upper triangular up to the permutation q (reference: src/spasm_triangular.c:65-87) */
bool spasm_dense_forward_solve(const struct spasm_csr *U, spasm_ZZp *b, spasm_ZZp *x, const int *q)
{
	const int n = U->n, m = U->m;
	assert(n <= m);
	for (int i = 0; i < n; i++)
		x[i] = 0;
	for (int i = 0; i < n; i++) {
		const int col = q ? q[i] : i;
		const spasm_ZZp v = b[col];
		if (v == 0)
			continue;
		x[i] = v;
		spasm_scatter(U, i, -v, b);
	}
	for (int j = 0; j < m; j++)
		if (b[j] != 0)
			return 0;
	return 1;
}

/* ------------------------------------------------------------------ permutations (reference: src/spasm_permutation.c) */

/* x[k] = b[p[k]]   (:17-25) */
void spasm_pvec(const int *p, const spasm_ZZp *b, spasm_ZZp *x, int n)
{
	assert(x && b);
	for (int k = 0; k < n; k++)
		x[k] = b[p ? p[k] : k];
}

/* x[p[k]] = b[k]   (:36-44) */
void spasm_ipvec(const int *p, const spasm_ZZp *b, spasm_ZZp *x, int n)
{
	assert(x && b);
	for (int k = 0; k < n; k++)
		x[p ? p[k] : k] = b[k];
}

/* inverse permutation; NULL (identity) stays NULL   (:47-59) */
int *spasm_pinv(int const *p, int n)
{
	if (p == NULL)
		return NULL;
	int *inv = spasm_malloc((size_t) n * sizeof(*inv));
	for (int k = 0; k < n; k++)
		inv[p[k]] = k;
	return inv;
}

/* C = P.A.Q^-1: row i of C is row p[i] of A, column j of A becomes column qinv[j]   (:65-100) */
struct spasm_csr *spasm_permute(const struct spasm_csr *A, const int *p, const int *qinv, int with_values)
{
	assert(A != NULL);
	const int n = A->n, m = A->m;
	const bool keep = with_values && A->x != NULL;
	struct spasm_csr *C = spasm_csr_alloc(n, m, A->nzmax, spasm_get_prime(A), keep);
	i64 out = 0;
	for (int i = 0; i < n; i++) {
		const int src = p ? p[i] : i;
		C->p[i] = out;
		for (i64 t = A->p[src]; t < A->p[src + 1]; t++, out++) {
			C->j[out] = qinv ? qinv[A->j[t]] : A->j[t];
			if (keep)
				C->x[out] = A->x[t];
		}
	}
	C->p[n] = out;
	return C;
}

/* the reference's shuffle (:102-114): p[i] swapped with p[rand() % i], i = n-1 .. 1 -- same rand() draws */
int *spasm_random_permutation(int n)
{
	int *p = spasm_malloc((size_t) n * sizeof(*p));
	for (int i = 0; i < n; i++)
		p[i] = i;
	for (int i = n - 1; i > 0; i--) {
		const int other = rand() % i;
		const int t = p[i];
		p[i] = p[other];
		p[other] = t;
	}
	return p;
}

/* x[a:b] permuted in place by p, which is destroyed   (:117-125) */
void spasm_range_pvec(int *x, int a, int b, int *p)
{
	const int len = b - a;
	for (int i = 0; i < len; i++)
		p[i] = x[a + p[i]];
	for (int i = 0; i < len; i++)
		x[a + i] = p[i];
}

/* A[r0:r1, c0:c1]   (reference: src/spasm_submatrix.c:7-43) */
struct spasm_csr *spasm_submatrix(const struct spasm_csr *A, int r0, int r1, int c0, int c1, int with_values)
{
	assert(A != NULL);
	const int rows = r1 > r0 ? r1 - r0 : 0, cols = c1 > c0 ? c1 - c0 : 0;
	i64 room = A->p[r1] - A->p[r0];
	if (room < 0)
		room = 0;
	const bool keep = with_values && A->x != NULL;
	struct spasm_csr *B = spasm_csr_alloc(rows, cols, room, spasm_get_prime(A), keep);
	i64 out = 0;
	for (int i = r0; i < r1; i++) {
		B->p[i - r0] = out;
		for (i64 t = A->p[i]; t < A->p[i + 1]; t++) {
			const int c = A->j[t];
			if (c < c0 || c >= c1)
				continue;
			B->j[out] = c - c0;
			if (keep)
				B->x[out] = A->x[t];
			out++;
		}
	}
	B->p[rows] = out;
	spasm_csr_realloc(B, -1);
	return B;
}

/*
 * Right kernel basis read off an RREF (reference: src/spasm_kernel.c:133-178): one vector per non-pivotal column j,
 * with `prime - 1` on column j (the reference's unbalanced -1, kept) and R[i][j] on the pivot column of row i.
 */
struct spasm_csr *spasm_kernel_from_rref(const struct spasm_csr *R, const int *qinv)
{
	assert(qinv != NULL);
	const int n = R->n, m = R->m;
	assert(n <= m);
	const i64 prime = spasm_get_prime(R);
	struct spasm_csr *Rt = spasm_transpose(R, true);
	int *pivcol = spasm_malloc((size_t) (n > 0 ? n : 1) * sizeof(*pivcol));
	for (int i = 0; i < n; i++)
		pivcol[i] = R->j[R->p[i]];
	struct spasm_csr *K = spasm_csr_alloc(m - n, m, spasm_nnz(R) - n + m - n, prime, true);
	i64 out = 0;
	int rows = 0;
	K->p[0] = 0;
	for (int j = 0; j < m; j++) {
		if (qinv[j] >= 0)
			continue;
		K->j[out] = j;
		K->x[out] = (spasm_ZZp) (prime - 1);
		out++;
		for (i64 t = Rt->p[j]; t < Rt->p[j + 1]; t++, out++) {
			K->j[out] = pivcol[Rt->j[t]];
			K->x[out] = Rt->x[t];
		}
		K->p[++rows] = out;
	}
	assert(rows == m - n);
	free(pivcol);
	spasm_csr_free(Rt);
	return K;
}
