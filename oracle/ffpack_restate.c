/*
 * ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's only third-party boundary,
 * src/spasm_ffpack.cpp (148 lines of C++ over FFLAS-FFPACK + Givaro).
 * FFLAS-FFPACK / Givaro are not vendored in /root/reference and not pinned
 * there (pkg_check_modules without version, CMakeLists.txt:33-34; CI uses
 * Debian 12's packages, .gitlab-ci.yml:6, i.e. fflas-ffpack 2.5.x / givaro
 * 4.2.x) and their headers are absent from this image, so that file cannot be
 * compiled.  This file restates what the two wrapped routines are *published*
 * to compute, anchored on the reference's own call sites and tests:
 *
 *   spasm_ffpack_rref  (ffpack.cpp:22-44,78-86  -> FFPACK::pReducedRowEchelonForm)
 *      in-place reduced row echelon form of an n x m row-major matrix over
 *      Z/pZ; returns the rank r; qinv[0:r] = pivot columns (= the column rank
 *      profile, increasing), qinv[r:m] = the other columns; for i < r, k >= r,
 *      A[i*ld + k] is the entry of RREF row i on column qinv[k].
 *      consumers: src/spasm_echelonize.c:192-223 (update_U_after_rref),
 *                 tests/dense_rref_ffpack.c:83-110.
 *   spasm_ffpack_LU    (ffpack.cpp:52-75,88-96  -> FFPACK::pPLUQ, FflasUnit)
 *      P*A*Q = L*U, U unit upper triangular, both packed in place in A;
 *      consumers: src/spasm_echelonize.c:228-313, tests/dense_lu_ffpack.c:145-169.
 *
 * PARITY STATUS: the pivot *set* (column rank profile) and the RREF values are
 * mathematically unique; the order of qinv[r:m] chosen by FFPACK (it comes from
 * LAPACK-style transpositions) is not recoverable without its source, and the
 * reference's tests pin validity and layout only -> at the bit level this
 * boundary is "parity unpinned"; everything downstream is compared on
 * canonical forms (see tests/canonical.py).
 *
 * It is linked (a) into oracle/_ref/libspasm_ref.so next to the reference's
 * own C sources, and (b) into oracle/liboracle.so.  Nothing under spasm_b200/
 * may call it.
 */
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <inttypes.h>
#include <stddef.h>
#include <sys/time.h>

typedef int64_t i64;
typedef int32_t i32;

/* must match the reference's enum order (src/spasm.h:137) */
typedef enum {ORACLE_DOUBLE, ORACLE_FLOAT, ORACLE_I64} oracle_datatype;

static double now(void)
{
	struct timeval tv;
	gettimeofday(&tv, NULL);
	return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ------------------------------------------------------------- Z/pZ helpers */
static inline i64 bal(i64 x, i64 p)          /* any x -> balanced representative */
{
	x %= p;
	i64 half = p / 2, mhalf = p / 2 - p + 1;
	if (x > half) x -= p;
	else if (x < mhalf) x += p;
	return x;
}

static i64 inverse(i64 a, i64 p)
{
	i64 r0 = p, r1 = ((a % p) + p) % p, t0 = 0, t1 = 1;
	while (r1 != 0) {
		i64 q = r0 / r1, r2 = r0 - q * r1, t2 = t0 - q * t1;
		r0 = r1; r1 = r2; t0 = t1; t1 = t2;
	}
	return bal(t0, p);
}

/* ---------------------------------------------------- typed load / store */
static i64 load(const void *A, size_t k, oracle_datatype t)
{
	switch (t) {
	case ORACLE_DOUBLE: return (i64) ((const double *) A)[k];
	case ORACLE_FLOAT:  return (i64) ((const float *) A)[k];
	case ORACLE_I64:    return ((const i64 *) A)[k];
	}
	abort();
}

static void store(void *A, size_t k, oracle_datatype t, i64 v)
{
	switch (t) {
	case ORACLE_DOUBLE: ((double *) A)[k] = (double) v; return;
	case ORACLE_FLOAT:  ((float *) A)[k] = (float) v; return;
	case ORACLE_I64:    ((i64 *) A)[k] = v; return;
	}
	abort();
}

/*
 * Blocked Gauss-Jordan elimination with column-rank-profile pivoting on an
 * n x m int32 matrix W (row-major, leading dimension m, balanced values).
 * On return rows 0..r-1 are the RREF rows ordered by pivot column, pivcol[0:r]
 * the pivot columns (increasing); rows r..n-1 are zero.  Returns r.
 */
#define PANEL 48

/* acc[j] += w * src[j] with 32 x 32 -> 64-bit products (vpmuldq when AVX2 is there; the clone is chosen at load time,
 * so the library built in the container also runs on the GPU box's host) */
__attribute__((target_clones("avx2", "default")))
static void axpy_widening(i64 *restrict acc, i32 w, const i32 *restrict src, int width)
{
	for (int j = 0; j < width; j++)
		acc[j] += (i64) w * (i64) src[j];
}

int oracle_dense_rref_i32(i64 p, int n, int m, i32 *W, int *pivcol)
{
	int r = 0;
	const int fast = (p < (1ll << 26));      /* k*|w|*|v| < 48 * 2^50 fits an i64 accumulator */
	i32 *panel = malloc((size_t) n * PANEL * sizeof(i32));
	int *prow = malloc(PANEL * sizeof(int));
	int *pcol = malloc(PANEL * sizeof(int));
	i64 *Minv = malloc(PANEL * PANEL * sizeof(i64));
	i64 *Mwork = malloc(PANEL * 2 * PANEL * sizeof(i64));
	i32 *Wmul = malloc((size_t) n * PANEL * sizeof(i32));
	i32 *oldrows = malloc((size_t) PANEL * m * sizeof(i32));
	char *ispiv = malloc(n);

	for (int c0 = 0; c0 < m && r < n; c0 += PANEL) {
		int nb = (m - c0 < PANEL) ? m - c0 : PANEL;
		/* 1. discover the pivots of this panel by forward elimination on a copy (rows r..n-1) */
		for (int i = 0; i < n; i++)
			for (int c = 0; c < nb; c++)
				panel[(size_t) i * PANEL + c] = W[(size_t) i * m + c0 + c];
		memset(ispiv, 0, n);
		int k = 0;
		for (int c = 0; c < nb && r + k < n; c++) {
			int piv = -1;
			for (int i = r; i < n; i++)
				if (!ispiv[i] && panel[(size_t) i * PANEL + c] != 0) {
					piv = i;
					break;
				}
			if (piv < 0)
				continue;
			ispiv[piv] = 1;
			prow[k] = piv;
			pcol[k] = c;
			k += 1;
			i64 inv = inverse(panel[(size_t) piv * PANEL + c], p);
			#pragma omp parallel for schedule(static)
			for (int i = r; i < n; i++) {
				if (ispiv[i])
					continue;
				i64 l = panel[(size_t) i * PANEL + c];
				if (l == 0)
					continue;
				l = bal(l * inv, p);
				for (int cc = c; cc < nb; cc++)
					panel[(size_t) i * PANEL + cc] = (i32) bal(panel[(size_t) i * PANEL + cc] - l * panel[(size_t) piv * PANEL + cc], p);
			}
		}
		if (k == 0)
			continue;
		/* 2. M = W[prow, c0 + pcol]; Minv by Gauss-Jordan on [M | I] */
		for (int s = 0; s < k; s++)
			for (int t = 0; t < k; t++) {
				Mwork[s * 2 * PANEL + t] = W[(size_t) prow[s] * m + c0 + pcol[t]];
				Mwork[s * 2 * PANEL + PANEL + t] = (s == t);
			}
		for (int t = 0; t < k; t++) {
			int s = t;
			while (Mwork[s * 2 * PANEL + t] == 0)
				s += 1;                   /* exists: M is invertible */
			assert(s < k);
			if (s != t)
				for (int c = 0; c < 2 * PANEL; c++) {
					i64 tmp = Mwork[s * 2 * PANEL + c];
					Mwork[s * 2 * PANEL + c] = Mwork[t * 2 * PANEL + c];
					Mwork[t * 2 * PANEL + c] = tmp;
				}
			i64 inv = inverse(Mwork[t * 2 * PANEL + t], p);
			for (int c = 0; c < 2 * PANEL; c++)
				Mwork[t * 2 * PANEL + c] = bal(Mwork[t * 2 * PANEL + c] * inv, p);
			for (int s2 = 0; s2 < k; s2++) {
				if (s2 == t)
					continue;
				i64 l = Mwork[s2 * 2 * PANEL + t];
				if (l == 0)
					continue;
				for (int c = 0; c < 2 * PANEL; c++)
					Mwork[s2 * 2 * PANEL + c] = bal(Mwork[s2 * 2 * PANEL + c] - l * Mwork[t * 2 * PANEL + c], p);
			}
		}
		for (int s = 0; s < k; s++)
			for (int t = 0; t < k; t++)
				Minv[s * PANEL + t] = Mwork[s * 2 * PANEL + PANEL + t];
		/* 3. multipliers: Wmul[i] = W[i, pivot columns] * Minv  (pivot rows: I - Minv) */
		int width = m - c0;
		for (int s = 0; s < k; s++)
			memcpy(oldrows + (size_t) s * width, W + (size_t) prow[s] * m + c0, width * sizeof(i32));
		#pragma omp parallel for schedule(static)
		for (int i = 0; i < n; i++) {
			i64 piv_vals[PANEL];
			for (int s = 0; s < k; s++)
				piv_vals[s] = W[(size_t) i * m + c0 + pcol[s]];
			for (int t = 0; t < k; t++) {
				i64 acc = 0;
				if (fast) {                   /* 48 products below 2^50 fit an i64 */
					for (int s = 0; s < k; s++)
						acc += piv_vals[s] * Minv[s * PANEL + t];
				} else {
					for (int s = 0; s < k; s++)
						acc = (acc + (piv_vals[s] * Minv[s * PANEL + t]) % p) % p;
				}
				Wmul[(size_t) i * PANEL + t] = (i32) bal(acc, p);
			}
		}
		for (int s = 0; s < k; s++)
			for (int t = 0; t < k; t++)
				Wmul[(size_t) prow[s] * PANEL + t] = (i32) bal((s == t) - Minv[s * PANEL + t], p);
		/* 4. W[:, c0:] -= Wmul * oldrows   (the "trailing update", a dense product) */
		#pragma omp parallel
		{
			i64 *acc = malloc((size_t) width * sizeof(i64));
			#pragma omp for schedule(dynamic, 8)
			for (int i = 0; i < n; i++) {
				i32 *row = W + (size_t) i * m + c0;
				const i32 *mul = Wmul + (size_t) i * PANEL;
				int any = 0;
				for (int t = 0; t < k; t++)
					any |= (mul[t] != 0);
				if (!any)
					continue;
				if (fast) {
					for (int j = 0; j < width; j++)
						acc[j] = 0;
					for (int t = 0; t < k; t++) {
						if (mul[t] == 0)
							continue;
						axpy_widening(acc, mul[t], oldrows + (size_t) t * width, width);
					}
					for (int j = 0; j < width; j++)
						row[j] = (i32) bal((i64) row[j] - acc[j], p);
				} else {
					for (int j = 0; j < width; j++)
						acc[j] = row[j];
					for (int t = 0; t < k; t++) {
						i64 w = mul[t];
						if (w == 0)
							continue;
						const i32 *src = oldrows + (size_t) t * width;
						for (int j = 0; j < width; j++)
							acc[j] = (acc[j] - (w * (i64) src[j]) % p) % p;
					}
					for (int j = 0; j < width; j++)
						row[j] = (i32) bal(acc[j], p);
				}
			}
			free(acc);
		}
		/* 5. bring the new pivot rows to positions r..r+k-1, in order of pivot column */
		for (int s = 0; s < k; s++) {
			int src = prow[s], dst = r + s;
			if (src != dst) {
				for (int j = 0; j < m; j++) {
					i32 tmp = W[(size_t) src * m + j];
					W[(size_t) src * m + j] = W[(size_t) dst * m + j];
					W[(size_t) dst * m + j] = tmp;
				}
				for (int s2 = s + 1; s2 < k; s2++)     /* a later pivot row may have lived at dst */
					if (prow[s2] == dst)
						prow[s2] = src;
			}
			pivcol[r + s] = c0 + pcol[s];
		}
		r += k;
	}
	free(panel); free(prow); free(pcol); free(Minv); free(Mwork); free(Wmul); free(oldrows); free(ispiv);
	return r;
}

/* ------------------------------------------------------------ the C API */

int spasm_ffpack_rref(i64 prime, int n, int m, void *A, int ldA, oracle_datatype datatype, size_t *qinv)
{
	double start = now();
	fprintf(stderr, "[ffpack/rref (restated)] Matrix of dimension %d x %d mod %" PRId64 "... ", n, m, prime);
	fflush(stderr);
	i32 *W = malloc((size_t) (n > 0 ? n : 1) * (m > 0 ? m : 1) * sizeof(i32));
	int *pivcol = malloc((size_t) (m > 0 ? m : 1) * sizeof(int));
	for (int i = 0; i < n; i++)
		for (int j = 0; j < m; j++)
			W[(size_t) i * m + j] = (i32) bal(load(A, (size_t) i * ldA + j, datatype), prime);
	int r = oracle_dense_rref_i32(prime, n, m, W, pivcol);

	char *is_pivot = calloc(m > 0 ? m : 1, 1);
	for (int i = 0; i < r; i++) {
		qinv[i] = pivcol[i];
		is_pivot[pivcol[i]] = 1;
	}
	int k = r;
	for (int j = 0; j < m; j++)
		if (!is_pivot[j])
			qinv[k++] = j;
	for (int i = 0; i < n; i++)
		for (int kk = 0; kk < m; kk++) {
			i64 v;
			if (i >= r)
				v = 0;
			else if (kk < r)
				v = (i == kk);
			else
				v = W[(size_t) i * m + qinv[kk]];
			store(A, (size_t) i * ldA + kk, datatype, v);
		}
	free(W);
	free(pivcol);
	free(is_pivot);
	fprintf(stderr, "done in %.1fs. Rank %d\n", now() - start, r);
	return r;
}

/*
 * P*A*Q = L*U with U unit upper triangular (row i of U has an implicit 1 on
 * permuted column i), L lower trapezoidal n x r, packed in place:
 * A[i*ld + j] = L[i][j] for j < min(i+1, r), U[i][j] for j > i (i < r).
 * p[i] = original row sitting at position i, qinv[j] = original column at position j.
 * Plain unblocked elimination: this is only exercised on small inputs.
 */
int spasm_ffpack_LU(i64 prime, int n, int m, void *A, int ldA, oracle_datatype datatype, size_t *p, size_t *qinv)
{
	fprintf(stderr, "[ffpack/LU (restated)] Matrix of dimension %d x %d mod %" PRId64 "...\n", n, m, prime);
	size_t nn = n > 0 ? n : 1, mm = m > 0 ? m : 1;
	i64 *W = malloc(nn * mm * sizeof(i64));       /* working copy; becomes U on pivot rows */
	i64 *L = calloc(nn * mm, sizeof(i64));
	int *rowpos = malloc(nn * sizeof(int));       /* rowpos[k] = original row at position k */
	int *pivcol = malloc(mm * sizeof(int));
	for (int i = 0; i < n; i++) {
		rowpos[i] = i;
		for (int j = 0; j < m; j++)
			W[(size_t) i * m + j] = bal(load(A, (size_t) i * ldA + j, datatype), prime);
	}
	int r = 0;
	for (int c = 0; c < m && r < n; c++) {
		int piv = -1;
		for (int k = r; k < n; k++)
			if (W[(size_t) rowpos[k] * m + c] != 0) {
				piv = k;
				break;
			}
		if (piv < 0)
			continue;
		int tmp = rowpos[piv]; rowpos[piv] = rowpos[r]; rowpos[r] = tmp;
		i64 *prow = W + (size_t) rowpos[r] * m;
		i64 d = prow[c];
		i64 dinv = inverse(d, prime);
		L[(size_t) rowpos[r] * m + r] = d;
		for (int j = 0; j < m; j++)
			prow[j] = bal(prow[j] * dinv, prime);
		for (int k = r + 1; k < n; k++) {
			i64 *row = W + (size_t) rowpos[k] * m;
			i64 l = row[c];
			if (l == 0)
				continue;
			L[(size_t) rowpos[k] * m + r] = l;
			for (int j = 0; j < m; j++)
				row[j] = bal(row[j] - l * prow[j], prime);
		}
		pivcol[r] = c;
		r += 1;
	}
	char *is_pivot = calloc(mm, 1);
	for (int i = 0; i < r; i++) {
		qinv[i] = pivcol[i];
		is_pivot[pivcol[i]] = 1;
	}
	int k = r;
	for (int j = 0; j < m; j++)
		if (!is_pivot[j])
			qinv[k++] = j;
	for (int i = 0; i < n; i++) {
		p[i] = rowpos[i];
		for (int j = 0; j < m; j++) {
			i64 v;
			if (j < r && j <= i)
				v = L[(size_t) rowpos[i] * m + j];
			else if (i < r)
				v = W[(size_t) rowpos[i] * m + qinv[j]];
			else
				v = 0;
			store(A, (size_t) i * ldA + j, datatype, v);
		}
	}
	free(W); free(L); free(rowpos); free(pivcol); free(is_pivot);
	return r;
}

/* reference: src/spasm_ffpack.cpp:100-149 */
i32 spasm_datatype_read(const void *A, size_t i, oracle_datatype datatype)
{
	return (i32) load(A, i, datatype);
}

void spasm_datatype_write(void *A, size_t i, oracle_datatype datatype, i32 value)
{
	store(A, i, datatype, value);
}

size_t spasm_datatype_size(oracle_datatype datatype)
{
	switch (datatype) {
	case ORACLE_DOUBLE: return sizeof(double);
	case ORACLE_FLOAT:  return sizeof(float);
	case ORACLE_I64:    return sizeof(i64);
	}
	abort();
}

oracle_datatype spasm_datatype_choose(i64 prime)
{
	if (prime <= 8191)
		return ORACLE_FLOAT;
	if (prime <= 189812531)
		return ORACLE_DOUBLE;
	return ORACLE_I64;
}

const char *spasm_datatype_name(oracle_datatype datatype)
{
	switch (datatype) {
	case ORACLE_DOUBLE: return "double";
	case ORACLE_FLOAT:  return "float";
	case ORACLE_I64:    return "i64";
	}
	abort();
}
