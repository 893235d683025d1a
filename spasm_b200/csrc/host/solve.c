/*
 * Consumers of a factorization WITH its L factor: linear solves, Eberly's interactive rank
 * certificate, and the probabilistic check  x*A == (x*L)*U.  Plain re-entrant host C: they run a handful of sparse
 * matrix-vector products and two triangular substitutions on vectors, nothing the GPU would help with; the
 * factorization they consume comes from spasm_echelonize (CUDA) with opts->L / opts->complete.
 * reference: src/spasm_solve.c, src/spasm_certificate.c.
 */
#include <assert.h>
#include <inttypes.h>
#include <stdlib.h>
#include <string.h>
#include "spasm.h"

/*
 * x * A = b  (b of size m, x of size n).  Returns true when a solution exists; x is supported by the pivotal rows.
 * reference: src/spasm_solve.c:13-46 -- z * U = b by substitution over the rows of U in order (they are in
 * topological order), then x * L = z backwards over the pivotal rows fact->p.
 */
bool spasm_solve(const struct spasm_lu *fact, const spasm_ZZp *b, spasm_ZZp *x)
{
	const struct spasm_csr *L = fact->L, *U = fact->U;
	assert(L != NULL);
	const int m = U->m, r = U->n;
	spasm_ZZp *rhs = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*rhs));
	spasm_ZZp *z = spasm_malloc((i64) (r > 0 ? r : 1) * sizeof(*z));
	int *pivot_col = spasm_malloc((i64) (r > 0 ? r : 1) * sizeof(*pivot_col));
	for (int j = 0; j < m; j++) {
		rhs[j] = b[j];
		if (fact->qinv[j] >= 0)
			pivot_col[fact->qinv[j]] = j;
	}
	bool found = spasm_dense_forward_solve(U, rhs, z, pivot_col);
	spasm_dense_back_solve(L, z, x, fact->p);
	free(rhs);
	free(z);
	free(pivot_col);
	return found;
}

/*
 * X * A = B, one solve per row of B; ok[i] (optional) tells whether row i has a solution (X[i] is meaningless
 * otherwise).  reference: src/spasm_solve.c:51-93.
 */
struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok)
{
	assert(fact->L != NULL);
	const i64 prime = spasm_get_prime(B);
	assert(prime == spasm_get_prime(fact->L));
	const int n = B->n, m = B->m, Xm = fact->L->n;
	struct spasm_triplet *X = spasm_triplet_alloc(n, Xm, (i64) n + Xm, prime, true);
	spasm_ZZp *rhs = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*rhs));
	spasm_ZZp *sol = spasm_malloc((i64) (Xm > 0 ? Xm : 1) * sizeof(*sol));
	for (int i = 0; i < n; i++) {
		memset(rhs, 0, (size_t) m * sizeof(*rhs));
		spasm_scatter(B, i, 1, rhs);
		bool found = spasm_solve(fact, rhs, sol);
		if (ok != NULL)
			ok[i] = found;
		for (int j = 0; j < Xm; j++)
			if (sol[j] != 0)
				spasm_add_entry(X, i, j, sol[j]);
	}
	free(rhs);
	free(sol);
	struct spasm_csr *out = spasm_compress(X);
	spasm_triplet_free(X);
	return out;
}

/*
 * Probabilistic check of the factorization: for a random x supported by the pivotal rows (every row when the
 * factorization is complete ... the reference hard-wires "not complete", src/spasm_certificate.c:181),
 * x*A == (x*L)*U.  One PRNG draw per row of A, pivotal or not, in row order.
 * reference: src/spasm_certificate.c:164-217.
 */
bool spasm_factorization_verify(const struct spasm_csr *A, const struct spasm_lu *fact, u64 seed)
{
	assert(fact->L != NULL);
	const struct spasm_csr *U = fact->U, *L = fact->L;
	const int n = A->n, m = A->m, r = U->n;
	bool *is_pivotal = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*is_pivotal));
	spasm_ZZp *x = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*x));
	spasm_ZZp *xL = spasm_malloc((i64) (r > 0 ? r : 1) * sizeof(*xL));
	spasm_ZZp *xLU = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*xLU));
	spasm_ZZp *xA = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*xA));
	memset(is_pivotal, 0, (size_t) n * sizeof(*is_pivotal));
	for (int k = 0; k < r; k++) {
		assert(fact->p[k] >= 0);
		is_pivotal[fact->p[k]] = 1;
	}
	spasm_prng_ctx ctx;
	spasm_prng_seed_simple(spasm_get_prime(A), seed, 0, &ctx);
	for (int i = 0; i < n; i++) {
		spasm_ZZp draw = spasm_prng_ZZp(&ctx);
		x[i] = is_pivotal[i] ? draw : 0;
	}
	memset(xL, 0, (size_t) r * sizeof(*xL));
	memset(xLU, 0, (size_t) m * sizeof(*xLU));
	memset(xA, 0, (size_t) m * sizeof(*xA));
	spasm_xApy(x, A, xA);
	spasm_xApy(x, L, xL);
	spasm_xApy(xL, U, xLU);
	bool same = memcmp(xA, xLU, (size_t) m * sizeof(*xA)) == 0;
	free(is_pivotal);
	free(x);
	free(xL);
	free(xLU);
	free(xA);
	return same;
}

/*
 * Rank certificate (W. Eberly, "A New Interactive Certificate for Matrix Rank", 2015), non-interactive: the
 * challenges come from the PRNG seeded with `hash` (of the matrix and the commitment).
 *   i[], j[]  rows / columns of a non-singular r x r minor (the pivots);
 *   x         supported by i[]: x*A agrees with the challenge c on the columns j[]      (rank >= r);
 *   y         supported by i[]: for the challenge d on the other rows, (y - d)*A == 0   (rank <= r).
 * Draw order (it is part of the format): r values for c (by increasing column), then one value per non-pivotal row
 * by increasing row.  reference: src/spasm_certificate.c:21-97.
 */
struct spasm_rank_certificate *spasm_certificate_rank_create(const struct spasm_csr *A, const u8 *hash, const struct spasm_lu *fact)
{
	assert(fact->L != NULL);
	const int n = fact->L->n, m = fact->U->m, r = fact->U->n;
	const i64 room = r > 0 ? r : 1;
	struct spasm_rank_certificate *proof = spasm_malloc(sizeof(*proof));
	proof->r = r;
	proof->prime = spasm_get_prime(A);
	memcpy(proof->hash, hash, 32);
	proof->i = spasm_malloc(room * sizeof(*proof->i));
	proof->j = spasm_malloc(room * sizeof(*proof->j));
	proof->x = spasm_malloc(room * sizeof(*proof->x));
	proof->y = spasm_malloc(room * sizeof(*proof->y));
	for (int k = 0; k < r; k++)
		proof->i[k] = fact->p[k];
	int count = 0;
	for (int j = 0; j < m; j++)
		if (fact->qinv[j] >= 0)
			proof->j[count++] = j;
	assert(count == r);

	spasm_prng_ctx ctx;
	spasm_prng_seed(hash, proof->prime, 0, &ctx);
	spasm_ZZp *rowvec = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*rowvec));
	spasm_ZZp *colvec = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*colvec));
	bool *is_pivotal = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*is_pivotal));

	/* first response: x*A = c on the pivotal columns */
	memset(colvec, 0, (size_t) m * sizeof(*colvec));
	for (int k = 0; k < r; k++)
		colvec[proof->j[k]] = spasm_prng_ZZp(&ctx);
	spasm_solve(fact, colvec, rowvec);
	for (int k = 0; k < r; k++)
		proof->x[k] = rowvec[proof->i[k]];

	/* second response: the combination -d of the non-pivotal rows, expressed on the pivotal rows */
	memset(is_pivotal, 0, (size_t) n * sizeof(*is_pivotal));
	for (int k = 0; k < r; k++)
		is_pivotal[proof->i[k]] = 1;
	for (int i = 0; i < n; i++)
		rowvec[i] = is_pivotal[i] ? 0 : -spasm_prng_ZZp(&ctx);
	memset(colvec, 0, (size_t) m * sizeof(*colvec));
	spasm_xApy(rowvec, A, colvec);
	spasm_solve(fact, colvec, rowvec);
	for (int k = 0; k < r; k++)
		proof->y[k] = rowvec[proof->i[k]];
	free(rowvec);
	free(colvec);
	free(is_pivotal);
	return proof;
}

/* reference: src/spasm_certificate.c:99-161 */
bool spasm_certificate_rank_verify(const struct spasm_csr *A, const u8 *hash, const struct spasm_rank_certificate *proof)
{
	const int n = A->n, m = A->m, r = proof->r;
	if (memcmp(hash, proof->hash, 32) != 0 || spasm_get_prime(A) != proof->prime)
		return 0;
	for (int k = 0; k < r; k++)
		if (proof->i[k] < 0 || proof->i[k] >= n || proof->j[k] < 0 || proof->j[k] >= m)
			return 0;
	spasm_prng_ctx ctx;
	spasm_prng_seed(proof->hash, proof->prime, 0, &ctx);
	spasm_ZZp *rowvec = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*rowvec));
	spasm_ZZp *colvec = spasm_malloc((i64) (m > 0 ? m : 1) * sizeof(*colvec));
	bool *given = spasm_malloc((i64) (n > 0 ? n : 1) * sizeof(*given));
	bool accept = 1;

	/* x*A must reproduce the first challenge on the columns j[] */
	memset(rowvec, 0, (size_t) n * sizeof(*rowvec));
	for (int k = 0; k < r; k++)
		rowvec[proof->i[k]] = proof->x[k];
	memset(colvec, 0, (size_t) m * sizeof(*colvec));
	spasm_xApy(rowvec, A, colvec);
	for (int k = 0; k < r; k++)
		if (colvec[proof->j[k]] != spasm_prng_ZZp(&ctx))
			accept = 0;

	/* y on the rows i[], the second challenge on the others: the combination must vanish */
	memset(given, 0, (size_t) n * sizeof(*given));
	for (int k = 0; k < r; k++) {
		rowvec[proof->i[k]] = proof->y[k];
		given[proof->i[k]] = 1;
	}
	for (int i = 0; i < n; i++)
		if (!given[i])
			rowvec[i] = spasm_prng_ZZp(&ctx);
	memset(colvec, 0, (size_t) m * sizeof(*colvec));
	spasm_xApy(rowvec, A, colvec);
	for (int j = 0; j < m; j++)
		if (colvec[j] != 0)
			accept = 0;
	free(rowvec);
	free(colvec);
	free(given);
	return accept;
}

/* text format of the reference (src/spasm_certificate.c:219-238): r, prime, hash in hex, then i, j, x, y on one line each */
void spasm_rank_certificate_save(const struct spasm_rank_certificate *proof, FILE *f)
{
	const int r = proof->r;
	fprintf(f, "%d\n%" PRId64 "\n", r, proof->prime);
	for (int t = 0; t < 32; t++)
		fprintf(f, "%02x", proof->hash[t]);
	fprintf(f, "\n");
	const int *lines[4] = {proof->i, proof->j, proof->x, proof->y};
	for (int l = 0; l < 4; l++) {
		for (int k = 0; k < r; k++)
			fprintf(f, "%d ", lines[l][k]);
		fprintf(f, "\n");
	}
}

/* reads what spasm_rank_certificate_save wrote (the reference's loader stores the j line over the i line,
 * src/spasm_certificate.c:262-265; here each line goes to its own array) */
bool spasm_rank_certificate_load(FILE *f, struct spasm_rank_certificate *proof)
{
	int r;
	if (fscanf(f, "%d", &r) != 1 || r < 0)
		return 0;
	proof->r = r;
	const i64 room = r > 0 ? r : 1;
	proof->i = spasm_malloc(room * sizeof(*proof->i));
	proof->j = spasm_malloc(room * sizeof(*proof->j));
	proof->x = spasm_malloc(room * sizeof(*proof->x));
	proof->y = spasm_malloc(room * sizeof(*proof->y));
	if (fscanf(f, "%" SCNd64, &proof->prime) != 1)
		return 0;
	char hex[65];
	if (fscanf(f, "%64s", hex) != 1 || strlen(hex) != 64)
		return 0;
	for (int t = 0; t < 32; t++) {
		unsigned byte;
		if (sscanf(hex + 2 * t, "%2x", &byte) != 1)
			return 0;
		proof->hash[t] = (u8) byte;
	}
	int *lines[4] = {proof->i, proof->j, proof->x, proof->y};
	for (int l = 0; l < 4; l++)
		for (int k = 0; k < r; k++)
			if (fscanf(f, "%d", &lines[l][k]) != 1)
				return 0;
	return 1;
}
