/*
 * Containers, allocation, timing: plain host C kept behind the reference's
 * names (reference: src/spasm_util.c).  Everything handed back to a caller is
 * malloc() memory because the reference's tools free()/realloc() it.
 */
#include <stdlib.h>
#include <stdio.h>
#include <sys/time.h>
#include <err.h>
#include "spasm.h"

/* reference: src/spasm_util.c:10-24 -- the library never runs OpenMP teams */
int spasm_get_num_threads() { return 1; }
int spasm_get_thread_num() { return 0; }

/* reference: src/spasm_util.c:27-32 */
double spasm_wtime()
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return (double) tv.tv_sec + 1e-6 * (double) tv.tv_usec;
}

/* reference: src/spasm_util.c:35-38 */
i64 spasm_nnz(const struct spasm_csr *A)
{
	return A->p[A->n];
}

/* "1234", "12.3k", "4.5m", ... in at most 7 characters (reference: src/spasm_util.c:41-63) */
void spasm_human_format(i64 n, char *target)
{
	static const char suffix[] = {'k', 'm', 'g', 't'};
	if (n < 1000) {
		snprintf(target, 8, "%" PRId64, n);
		return;
	}
	i64 unit = 1000;
	for (int s = 0; s < 4; s++, unit *= 1000)
		if (n < 1000 * unit) {
			snprintf(target, 8, "%.1f%c", (double) n / (double) unit, suffix[s]);
			return;
		}
	/* like the reference, leave target untouched beyond 1e15 */
}

/* reference: src/spasm_util.c:65-87 */
void *spasm_malloc(i64 size)
{
	void *x = malloc(size);
	if (x == NULL)
		err(1, "malloc failed (size %" PRId64 ")", size);
	return x;
}

void *spasm_calloc(i64 count, i64 size)
{
	void *x = calloc(count, size);
	if (x == NULL)
		err(1, "calloc failed");
	return x;
}

void *spasm_realloc(void *ptr, i64 size)
{
	void *x = realloc(ptr, size);
	if (x == NULL && ptr != NULL && size != 0)
		err(1, "realloc failed");
	return x;
}

/* reference: src/spasm_util.c:90-103 */
struct spasm_csr *spasm_csr_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values)
{
	struct spasm_csr *A = spasm_malloc(sizeof(*A));
	spasm_field_init(prime, A->field);
	A->n = n;
	A->m = m;
	A->nzmax = nzmax;
	A->p = spasm_malloc((i64) (n + 1) * sizeof(i64));
	A->j = spasm_malloc(nzmax * sizeof(int));
	A->x = with_values ? spasm_malloc(nzmax * sizeof(spasm_ZZp)) : NULL;
	A->p[0] = 0;
	return A;
}

/* reference: src/spasm_util.c:106-118 */
struct spasm_triplet *spasm_triplet_alloc(int n, int m, i64 nzmax, i64 prime, bool with_values)
{
	struct spasm_triplet *T = spasm_malloc(sizeof(*T));
	spasm_field_init(prime, T->field);
	T->n = n;
	T->m = m;
	T->nzmax = nzmax;
	T->nz = 0;
	T->i = spasm_malloc(nzmax * sizeof(int));
	T->j = spasm_malloc(nzmax * sizeof(int));
	T->x = with_values ? spasm_malloc(nzmax * sizeof(spasm_ZZp)) : NULL;
	return T;
}

/* nzmax < 0 means "trim to the current number of entries" (reference: src/spasm_util.c:124-133) */
void spasm_csr_realloc(struct spasm_csr *A, i64 nzmax)
{
	if (nzmax < 0)
		nzmax = spasm_nnz(A);
	A->j = spasm_realloc(A->j, nzmax * sizeof(int));
	if (A->x != NULL)
		A->x = spasm_realloc(A->x, nzmax * sizeof(spasm_ZZp));
	A->nzmax = nzmax;
}

/* reference: src/spasm_util.c:139-151 */
void spasm_triplet_realloc(struct spasm_triplet *T, i64 nzmax)
{
	if (nzmax < 0)
		nzmax = T->nz;
	T->i = spasm_realloc(T->i, nzmax * sizeof(int));
	T->j = spasm_realloc(T->j, nzmax * sizeof(int));
	if (T->x != NULL)
		T->x = spasm_realloc(T->x, nzmax * sizeof(spasm_ZZp));
	T->nzmax = nzmax;
}

/* reference: src/spasm_util.c:154-170 */
void spasm_csr_free(struct spasm_csr *A)
{
	if (A == NULL)
		return;
	free(A->p);
	free(A->j);
	free(A->x);
	free(A);
}

void spasm_triplet_free(struct spasm_triplet *T)
{
	free(T->i);
	free(T->j);
	free(T->x);
	free(T);
}

/* change the number of rows (new rows are empty) and columns (reference: src/spasm_util.c:172-183) */
void spasm_csr_resize(struct spasm_csr *A, int n, int m)
{
	A->m = m;
	A->p = spasm_realloc(A->p, (i64) (n + 1) * sizeof(i64));
	for (int i = A->n + 1; i <= n; i++)
		A->p[i] = A->p[A->n];
	A->n = n;
}

/* reference: src/spasm_util.c:185-206 */
struct spasm_dm *spasm_dm_alloc(int n, int m)
{
	struct spasm_dm *P = spasm_malloc(sizeof(*P));
	P->p = spasm_malloc(n * sizeof(int));
	P->q = spasm_malloc(m * sizeof(int));
	P->r = spasm_malloc((n + 6) * sizeof(int));
	P->c = spasm_malloc((m + 6) * sizeof(int));
	P->nb = 0;
	for (int k = 0; k < 5; k++)
		P->rr[k] = P->cc[k] = 0;
	return P;
}

void spasm_dm_free(struct spasm_dm *P)
{
	free(P->p);
	free(P->q);
	free(P->r);
	free(P->c);
	free(P);
}

/* reference: src/spasm_util.c:208-215 */
void spasm_lu_free(struct spasm_lu *N)
{
	free(N->qinv);
	free(N->p);
	spasm_csr_free(N->U);
	spasm_csr_free(N->L);
	free(N);
}
