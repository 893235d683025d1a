/*
 * Triplet (coordinate) matrices and their conversion to CSR.
 * reference: src/spasm_triplet.c
 *
 * The order of the entries inside each CSR row is an *input* of the pivot
 * search (first-eligible-entry rules, reference: src/spasm_pivots.c:104-116,
 * :233-237), so spasm_compress() must produce exactly the reference's order:
 * entries of a row in file order, a repeated (i,j) folded into its first
 * occurrence, entries that sum to zero dropped afterwards.
 */
#include <assert.h>
#include <stdlib.h>
#include "spasm.h"

/* reference: src/spasm_triplet.c:7-24 */
void spasm_add_entry(struct spasm_triplet *T, int i, int j, i64 x)
{
	assert(i >= 0 && j >= 0);
	if (T->nz == T->nzmax)
		spasm_triplet_realloc(T, 2 * T->nzmax + 1);
	i64 k = T->nz;
	if (T->x != NULL) {
		spasm_ZZp v = spasm_ZZp_init(T->field, x);
		if (v == 0)
			return;            /* multiples of p never enter the matrix */
		T->x[k] = v;
	}
	T->i[k] = i;
	T->j[k] = j;
	T->nz = k + 1;
	if (i >= T->n)
		T->n = i + 1;
	if (j >= T->m)
		T->m = j + 1;
}

/* O(1): swap the roles of rows and columns (reference: src/spasm_triplet.c:26-34) */
void spasm_triplet_transpose(struct spasm_triplet *T)
{
	int *ti = T->i;
	T->i = T->j;
	T->j = ti;
	int tn = T->n;
	T->n = T->m;
	T->m = tn;
}

/*
 * reference: src/spasm_triplet.c:99-157 (bucket by row, then deduplicate :60-96,
 * then remove_explicit_zeroes :36-57).
 */
struct spasm_csr *spasm_compress(const struct spasm_triplet *T)
{
	const int n = T->n, m = T->m;
	const i64 nz = T->nz;
	const bool valued = (T->x != NULL);
	double start = spasm_wtime();
	fprintf(stderr, "[CSR] Compressing... ");
	fflush(stderr);

	struct spasm_csr *C = spasm_csr_alloc(n, m, nz, T->field->p, valued);
	i64 *Cp = C->p;
	int *Cj = C->j;
	spasm_ZZp *Cx = C->x;

	/* 1. stable bucket sort of the entries by row */
	i64 *fill = spasm_calloc(n + 1, sizeof(i64));
	for (i64 k = 0; k < nz; k++) {
		assert(T->i[k] < n);
		fill[T->i[k] + 1] += 1;
	}
	for (int i = 0; i < n; i++)
		fill[i + 1] += fill[i];
	for (int i = 0; i <= n; i++)
		Cp[i] = fill[i];
	for (i64 k = 0; k < nz; k++) {
		i64 dst = fill[T->i[k]]++;
		Cj[dst] = T->j[k];
		if (valued)
			Cx[dst] = T->x[k];
	}
	free(fill);

	/* 2. fold duplicates into their first occurrence, in place.
	 *    slot[j] = position of column j in the row being rebuilt (valid if >= row start) */
	i64 *slot = spasm_malloc((i64) m * sizeof(i64));
	for (int j = 0; j < m; j++)
		slot[j] = -1;
	i64 out = 0;
	for (int i = 0; i < n; i++) {
		i64 begin = Cp[i], end = Cp[i + 1];
		i64 row_start = out;
		for (i64 k = begin; k < end; k++) {
			int j = Cj[k];
			assert(j < m);
			if (slot[j] >= row_start) {
				if (valued)
					Cx[slot[j]] = spasm_ZZp_add(C->field, Cx[slot[j]], Cx[k]);
			} else {
				slot[j] = out;
				Cj[out] = j;
				if (valued)
					Cx[out] = Cx[k];
				out += 1;
			}
		}
		Cp[i] = row_start;
	}
	Cp[n] = out;
	free(slot);

	/* 3. squeeze out the entries that cancelled.
	 * Reference quirk kept on purpose (src/spasm_triplet.c:36-57): the reference compacts in place and
	 * begins the scan of row i at the already-rewritten row pointer, i.e. at the write cursor, so
	 * after the first dropped entry each later row re-reads the stale slots that precede its true
	 * start.  Only inputs with repeated (i,j) cancelling mod p are affected; being a drop-in means
	 * producing the same CSR as the reference on those too. */
	if (valued) {
		i64 w = 0;
		for (int i = 0; i < n; i++) {
			i64 end = Cp[i + 1];
			for (i64 k = w; k < end; k++)
				if (Cx[k] != 0) {
					Cj[w] = Cj[k];
					Cx[w] = Cx[k];
					w += 1;
				}
			Cp[i + 1] = w;
		}
	}
	spasm_csr_realloc(C, -1);

	char mem[16];
	i64 size = sizeof(int) * (n + nz) + sizeof(spasm_ZZp) * (valued ? nz : 0);
	spasm_human_format(size, mem);
	fprintf(stderr, "%" PRId64 " actual NZ, Mem usage = %sbyte [%.2fs]\n", spasm_nnz(C), mem, spasm_wtime() - start);
	return C;
}
