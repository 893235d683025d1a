/*
 * Host-side CSR transpose and the two sparse matrix-vector products the
 * reference's tests use as verifiers.
 * reference: src/spasm_transpose.c, src/spasm_spmv.c
 * (The transposes needed inside the GPU path are done on the device,
 *  gpu/csr_ops.cu; this one serves callers that hold host matrices.)
 */
#include <stdlib.h>
#include "spasm.h"

/* reference: src/spasm_transpose.c:5-52.  Entries of column j appear in T[j] by increasing row. */
struct spasm_csr *spasm_transpose(const struct spasm_csr *C, int keep_values)
{
	const int n = C->n, m = C->m;
	const i64 nnz = spasm_nnz(C);
	const bool valued = keep_values && (C->x != NULL);
	struct spasm_csr *T = spasm_csr_alloc(m, n, nnz, spasm_get_prime(C), valued);

	i64 *cursor = spasm_calloc(m + 1, sizeof(i64));
	for (i64 k = 0; k < nnz; k++)
		cursor[C->j[k] + 1] += 1;
	for (int j = 0; j < m; j++)
		cursor[j + 1] += cursor[j];
	for (int j = 0; j <= m; j++)
		T->p[j] = cursor[j];
	for (int i = 0; i < n; i++)
		for (i64 k = C->p[i]; k < C->p[i + 1]; k++) {
			i64 dst = cursor[C->j[k]]++;
			T->j[dst] = i;
			if (valued)
				T->x[dst] = C->x[k];
		}
	free(cursor);
	return T;
}

/* y += x * A   (x has n entries, y has m) -- reference: src/spasm_spmv.c:7-20 */
void spasm_xApy(const spasm_ZZp *x, const struct spasm_csr *A, spasm_ZZp *y)
{
	for (int i = 0; i < A->n; i++) {
		spasm_ZZp xi = x[i];
		if (xi == 0)
			continue;
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
			y[A->j[k]] = spasm_ZZp_axpy(A->field, xi, A->x[k], y[A->j[k]]);
	}
}

/* y += A * x   (x has m entries, y has n) -- reference: src/spasm_spmv.c:25-38 */
void spasm_Axpy(const struct spasm_csr *A, const spasm_ZZp *x, spasm_ZZp *y)
{
	for (int i = 0; i < A->n; i++) {
		spasm_ZZp acc = y[i];
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++)
			acc = spasm_ZZp_axpy(A->field, A->x[k], x[A->j[k]], acc);
		y[i] = acc;
	}
}
