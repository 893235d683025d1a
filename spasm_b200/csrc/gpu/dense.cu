#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include "dense.cuh"
#include "stats.cuh"

namespace sb {

/* ================================================================== C -= A * B (mod p), CUDA cores
 * 64 x 64 tile per CTA, 256 threads, 4 x 4 outputs per thread, K in slabs of 32 staged in shared
 * memory; 64-bit accumulators with delayed reduction (zp.cuh).  This is the portable path; the
 * int8 tensor-core path (umma_gemm.cu) takes over for large products. */
#define GM 64
#define GN 64
#define GK 32

/* SMALL: |a * b| * 4 < 2^31 (p <= 46337, e.g. the default 42013): four products are summed in 32 bits (full-rate IMAD)
 * before they join the 64-bit accumulator -- the rank-32 trailing updates of the panels are bound by the multiply-adds */
template <bool SMALL>
__global__ void __launch_bounds__(256)
k_gemm_sub(i32 *__restrict__ C, int ldc, const i32 *__restrict__ A, int lda, const i32 *__restrict__ B, int ldb,
           int M, int N, int K, Zp F, const int *d_K)
{
	if (d_K)
		K = min(K, *d_K);
	if (K <= 0)
		return;
	__shared__ i32 As[GK][GM + 4];
	__shared__ i32 Bs[GK][GN + 4];
	const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
	const int m0 = blockIdx.y * GM, n0 = blockIdx.x * GN;
	i64 acc[4][4];
#pragma unroll
	for (int a = 0; a < 4; a++)
#pragma unroll
		for (int b = 0; b < 4; b++)
			acc[a][b] = 0;
	int pending = 0;
	for (int k0 = 0; k0 < K; k0 += GK) {
#pragma unroll
		for (int t = 0; t < (GM * GK) / 256; t++) {
			int idx = tid + t * 256;
			int mm = idx / GK, kk = idx % GK;
			int gm = m0 + mm, gk = k0 + kk;
			As[kk][mm] = (gm < M && gk < K) ? A[(size_t) gm * lda + gk] : 0;
		}
#pragma unroll
		for (int t = 0; t < (GK * GN) / 256; t++) {
			int idx = tid + t * 256;
			int kk = idx / GN, nn = idx % GN;
			int gk = k0 + kk, gn = n0 + nn;
			Bs[kk][nn] = (gk < K && gn < N) ? B[(size_t) gk * ldb + gn] : 0;
		}
		__syncthreads();
		if (SMALL) {
#pragma unroll 2
			for (int kk0 = 0; kk0 < GK; kk0 += 4) {
				i32 part[4][4];
#pragma unroll
				for (int u = 0; u < 4; u++)
#pragma unroll
					for (int v = 0; v < 4; v++)
						part[u][v] = 0;
#pragma unroll
				for (int kq = 0; kq < 4; kq++) {
					i32 a[4], b[4];
#pragma unroll
					for (int u = 0; u < 4; u++) {
						a[u] = As[kk0 + kq][ty * 4 + u];
						b[u] = Bs[kk0 + kq][tx * 4 + u];
					}
#pragma unroll
					for (int u = 0; u < 4; u++)
#pragma unroll
						for (int v = 0; v < 4; v++)
							part[u][v] += a[u] * b[v];
				}
#pragma unroll
				for (int u = 0; u < 4; u++)
#pragma unroll
					for (int v = 0; v < 4; v++)
						acc[u][v] += (i64) part[u][v];
			}
			/* delay is astronomically large for such primes (2^63 / p^2 > 2^32 products): no reduction before the end */
		} else
#pragma unroll 4
		for (int kk = 0; kk < GK; kk++) {
			i32 a[4], b[4];
#pragma unroll
			for (int u = 0; u < 4; u++) {
				a[u] = As[kk][ty * 4 + u];
				b[u] = Bs[kk][tx * 4 + u];
			}
#pragma unroll
			for (int u = 0; u < 4; u++)
#pragma unroll
				for (int v = 0; v < 4; v++)
					acc[u][v] += (i64) a[u] * (i64) b[v];
			if (++pending >= F.delay) {
#pragma unroll
				for (int u = 0; u < 4; u++)
#pragma unroll
					for (int v = 0; v < 4; v++)
						acc[u][v] = zp_reduce(acc[u][v], F);
				pending = 0;
			}
		}
		__syncthreads();
	}
#pragma unroll
	for (int u = 0; u < 4; u++) {
		int gm = m0 + ty * 4 + u;
		if (gm >= M)
			continue;
#pragma unroll
		for (int v = 0; v < 4; v++) {
			int gn = n0 + tx * 4 + v;
			if (gn < N) {
				size_t at = (size_t) gm * ldc + gn;
				C[at] = zp_reduce((i64) C[at] - (i64) zp_reduce(acc[u][v], F), F);
			}
		}
	}
}

static void gemm_sub_on(cudaStream_t st, i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F, const int *d_K)
{
	if (M <= 0 || N <= 0 || K <= 0)
		return;
	dim3 grid(cdiv(N, GN), cdiv(M, GM));
	const bool small = F.p <= 46337 && (i64) K < (i64) F.delay;
	if (small)
		k_gemm_sub<true><<<grid, 256, 0, st>>>(C, ldc, A, lda, B, ldb, M, N, K, F, d_K);
	else
		k_gemm_sub<false><<<grid, 256, 0, st>>>(C, ldc, A, lda, B, ldb, M, N, K, F, d_K);
	LAUNCHED(1);
}

void dense_gemm_sub(i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F, const int *d_K)
{
	if (M <= 0 || N <= 0 || K <= 0)
		return;
	/* Large products go to the tensor cores (umma_gemm.cu).  The rank-<=32 trailing updates of dense_rref
	 * (d_K != NULL) stay on the CUDA cores: with K = 32 the tcgen05 kernel is all prologue/epilogue. */
	if (!d_K && K >= 64 && (double) M * N * K >= 4e6 && umma_gemm_available(F)) {
		umma_gemm_sub(C, ldc, A, lda, B, ldb, M, N, K, F);
		stats().pub.gemm_fieldops += 2.0 * M * (double) N * K;
		return;
	}
	dim3 grid(cdiv(N, GN), cdiv(M, GM));
	/* four products of balanced residues fit 32 bits, and K is far below the delayed-reduction bound */
	const bool small = F.p <= 46337 && (i64) K < (i64) F.delay;
	if (small)
		k_gemm_sub<true><<<grid, 256, 0, ctx().stream>>>(C, ldc, A, lda, B, ldb, M, N, K, F, d_K);
	else
		k_gemm_sub<false><<<grid, 256, 0, ctx().stream>>>(C, ldc, A, lda, B, ldb, M, N, K, F, d_K);
	LAUNCHED(1);
	KERNEL_CHECK();
	if (!d_K)      /* the rank-<=32 trailing updates of a block are counted once, as 2 n m rank, at the end of dense_rref */
		stats().pub.gemm_fieldops += 2.0 * M * (double) N * K;
}

/* ================================================================== gathers */

__global__ void k_gather_columns(const i32 *__restrict__ src, int lds, int rows, const int *__restrict__ cols, int ncols, i32 *dst, int ldd)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
	if (t < ncols && r < rows)
		dst[(size_t) r * ldd + t] = src[(size_t) r * lds + cols[t]];
}

void dense_gather_columns(const i32 *src, int lds, int rows, const int *d_cols, int ncols, i32 *dst, int ldd)
{
	if (rows <= 0 || ncols <= 0)
		return;
	dim3 grid(cdiv(ncols, 256), rows);
	k_gather_columns<<<grid, 256, 0, ctx().stream>>>(src, lds, rows, d_cols, ncols, dst, ldd);
	LAUNCHED(1);
	KERNEL_CHECK();
}

__global__ void k_gather_rows(const i32 *__restrict__ src, int lds, const int *__restrict__ rows, int nrows, int width, i32 *dst, int ldd,
                              const int *d_nrows)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
	if (d_nrows)
		nrows = min(nrows, *d_nrows);
	if (c < width && t < nrows)
		dst[(size_t) t * ldd + c] = src[(size_t) rows[t] * lds + c];
}

void dense_gather_rows(const i32 *src, int lds, const int *d_rows, int nrows, int width, i32 *dst, int ldd)
{
	if (nrows <= 0 || width <= 0)
		return;
	dim3 grid(cdiv(width, 256), nrows);
	k_gather_rows<<<grid, 256, 0, ctx().stream>>>(src, lds, d_rows, nrows, width, dst, ldd, nullptr);
	LAUNCHED(1);
	KERNEL_CHECK();
}

__global__ void k_scatter_rows_cols(const i32 *__restrict__ src, int lds, const int *__restrict__ rows, int nrows,
                                    const int *__restrict__ cols, int ncols, i32 *dst, int ldd)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
	if (c < ncols && t < nrows)
		dst[(size_t) t * ldd + cols[c]] = src[(size_t) rows[t] * lds + c];
}

void dense_scatter_rows(const i32 *src, int lds, const int *d_rows, int nrows, const int *d_cols, int ncols, i32 *dst, int ldd)
{
	if (nrows <= 0 || ncols <= 0)
		return;
	dim3 grid(cdiv(ncols, 256), nrows);
	k_scatter_rows_cols<<<grid, 256, 0, ctx().stream>>>(src, lds, d_rows, nrows, d_cols, ncols, dst, ldd);
	LAUNCHED(1);
	KERNEL_CHECK();
}

/* ================================================================== panel factorisation */

#define NB 32
#define PSTRIDE 33

struct PanelInfo {
	int k;
	int prow[NB];
	int pcol[NB];
	i32 Minv[NB][NB];      /* inverse of S[prow, c0 + pcol] */
	int fast_done;         /* the sampled factorisation (k_rref_panel_sample) settled this panel: the full kernels return at once */
};

/* Common tail of the panel kernels (whole CTA): M^-1 for M = S[prow, c0 + pcol] (k x k, original values), then the new
 * pivots are recorded (rowstate, pivcol list, rank counter) and M^-1 is handed to the multiplier kernel. */
__device__ void panel_invert_and_publish(const i32 *__restrict__ S, int ld, int c0, int k, int r0, int *s_prow, int *s_pcol,
                                         i64 (*Mw)[2 * NB + 1], i32 *s_dinv, int *s_sw, int *rowstate, int *rank_dev, int *pivcol_out,
                                         PanelInfo *info, const Zp &F)
{
	const int tid = threadIdx.x, T = blockDim.x;
	/* --- M^-1 by fraction-free Gauss-Jordan on [M | I]: row_s <- d * row_s - l * row_t keeps every step free of
	 * modular inverses (which are ~1 us each on one thread); at the end the left half is diagonal and the k
	 * inverses of the diagonal are computed by k threads at once. */
	for (int idx = tid; idx < k * 2 * NB; idx += T) {
		int s = idx / (2 * NB), t = idx % (2 * NB);
		i64 v = 0;
		if (t < k)
			v = S[(size_t) s_prow[s] * ld + c0 + s_pcol[t]];
		else if (t >= NB)
			v = (t - NB == s);
		Mw[s][t] = v;
	}
	__syncthreads();
	for (int t = 0; t < k; t++) {
		if (tid == 0) {
			int s = t;
			while (Mw[s][t] == 0)
				s++;                         /* M is invertible: a non-zero entry exists below */
			*s_sw = s;
		}
		__syncthreads();
		const int sw = *s_sw;
		if (sw != t && tid < 2 * NB) {
			i64 tmp = Mw[sw][tid];
			Mw[sw][tid] = Mw[t][tid];
			Mw[t][tid] = tmp;
		}
		__syncthreads();
		/* every thread takes its operands (the multiplier column t and the pivot row t are rewritten below) */
		constexpr int PER = 8;                 /* NB * 2NB = 2048 entries, at least 256 threads */
		i64 lv[PER], xv[PER], yv[PER];
		const i64 d = Mw[t][t];
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const int idx = tid + u * T, sa = idx / (2 * NB), ca = idx % (2 * NB);
			lv[u] = 0;
			if (idx < k * 2 * NB && sa != t) {
				lv[u] = Mw[sa][t];
				xv[u] = Mw[sa][ca];
				yv[u] = Mw[t][ca];
			}
		}
		__syncthreads();
#pragma unroll
		for (int u = 0; u < PER; u++) {
			const int idx = tid + u * T, sa = idx / (2 * NB), ca = idx % (2 * NB);
			if (lv[u] != 0)
				Mw[sa][ca] = (ca == t) ? 0 : (i64) zp_reduce(d * xv[u] - lv[u] * yv[u], F);
		}
		__syncthreads();
	}
	/* scale row s by the inverse of its diagonal entry */
	if (tid < k)
		s_dinv[tid] = zp_inverse((i32) Mw[tid][tid], F);
	__syncthreads();
	for (int idx = tid; idx < k * NB; idx += T) {
		int s = idx / NB, c = idx % NB;
		Mw[s][NB + c] = zp_mul((i32) Mw[s][NB + c], s_dinv[s], F);
	}
	__syncthreads();

	/* --- hand M^-1 to the multiplier kernel (the n x k x k product is spread over the whole GPU) */
	for (int idx = tid; idx < NB * NB; idx += T) {
		int s = idx / NB, t = idx % NB;
		info->Minv[s][t] = (s < k && t < k) ? (i32) Mw[s][NB + t] : 0;
	}
	if (tid < k) {
		info->prow[tid] = s_prow[tid];
		info->pcol[tid] = s_pcol[tid];
		rowstate[s_prow[tid]] = r0 + tid;
		pivcol_out[r0 + tid] = c0 + s_pcol[tid];
	}
	__syncthreads();
	if (tid == 0)
		*rank_dev = r0 + k;
}

/*
 * One CTA.  Works on a private copy of the panel S[:, c0:c0+nb]:
 *   1. discovers the pivots of the panel by forward elimination on the rows that are not pivotal yet
 *      (first such row with a non-zero entry, column by column)            -> prow[], pcol[], k
 *   2. inverts M = S[prow, c0 + pcol] (k x k)
 *   3. writes the multipliers W (n x NB): the update of the whole matrix is  S <- S - W * S[prow, :]
 *      (rows outside the panel's pivots: W = S[:, pivot columns] * M^-1; pivot rows: W = I - M^-1)
 * and records the new pivots (rowstate, pivcol list, rank counter).
 */
__global__ void __launch_bounds__(1024)
k_rref_panel(const i32 *__restrict__ S, int ld, int n, int m, int c0, int *rowstate, int *rank_dev, int *pivcol_out,
             i32 *W, PanelInfo *info, i32 *scratch, int use_smem, Zp F)
{
	extern __shared__ unsigned char dyn[];
	__shared__ int s_min, s_k, s_inv;
	__shared__ int s_prow[NB], s_pcol[NB];
	__shared__ i64 Mw[NB][2 * NB + 1];
	const int tid = threadIdx.x, T = blockDim.x;
	if (info->fast_done)
		return;
	const int r0 = *rank_dev;
	if (r0 >= n || c0 >= m) {
		if (tid == 0)
			info->k = 0;
		return;
	}
	const int nbw = min(NB, m - c0);
	i32 *Pw = use_smem ? (i32 *) dyn : scratch;
	unsigned char *taken = use_smem ? (dyn + (size_t) n * PSTRIDE * sizeof(i32)) : (unsigned char *) (scratch + (size_t) n * PSTRIDE);

	for (int idx = tid; idx < n * NB; idx += T) {
		int i = idx / NB, c = idx % NB;
		Pw[i * PSTRIDE + c] = (c < nbw) ? S[(size_t) i * ld + c0 + c] : 0;
	}
	for (int i = tid; i < n; i += T)
		taken[i] = rowstate[i] >= 0;
	if (tid == 0)
		s_k = 0;
	__syncthreads();

	for (int c = 0; c < nbw; c++) {
		if (s_k + r0 >= n)
			break;
		if (tid == 0)
			s_min = 0x7fffffff;
		__syncthreads();
		for (int i = tid; i < n; i += T)
			if (!taken[i] && Pw[i * PSTRIDE + c] != 0) {
				atomicMin(&s_min, i);
				break;
			}
		__syncthreads();
		const int piv = s_min;
		if (piv == 0x7fffffff) {
			__syncthreads();
			continue;
		}
		if (tid == 0) {
			s_prow[s_k] = piv;
			s_pcol[s_k] = c;
			s_k += 1;
			taken[piv] = 1;
		}
		__syncthreads();
		/* Fraction-free elimination: row_i <- d * row_i - l * row_piv.  Only the zero pattern matters here (which
		 * rows and columns are pivots); scaling a row by the unit d does not change it, and no modular inverse sits
		 * on the critical path of the 32 steps.  |d * x - l * y| < 2^63 for balanced residues below 2^31. */
		const i64 d = Pw[piv * PSTRIDE + c];
		for (int i = tid; i < n; i += T) {
			if (taken[i])
				continue;
			const i64 l = Pw[i * PSTRIDE + c];
			if (l == 0)
				continue;
			Pw[i * PSTRIDE + c] = 0;
			for (int cc = c + 1; cc < nbw; cc++)
				Pw[i * PSTRIDE + cc] = zp_reduce(d * Pw[i * PSTRIDE + cc] - l * Pw[piv * PSTRIDE + cc], F);
		}
		__syncthreads();
	}
	__syncthreads();
	const int k = s_k;
	if (tid == 0)
		info->k = k;
	if (k == 0)
		return;

	__shared__ i32 s_dinv[NB];
	panel_invert_and_publish(S, ld, c0, k, r0, s_prow, s_pcol, Mw, s_dinv, &s_min, rowstate, rank_dev, pivcol_out, info, F);
}

/*
 * The same panel factorisation on a thread-block CLUSTER: the rows of the panel are spread over PC_CTAS CTAs (8 SMs),
 * each keeping its slice in shared memory.  Per column: every CTA proposes its first eligible row with a remote
 * atomicMin on a slot in CTA 0's shared memory (DSMEM), ONE cluster barrier, everybody reads the winner, copies the
 * pivot row (32 values) from its owner's shared memory and updates its own rows.  Three rotating slots let CTA 0
 * reset a slot two steps before it is reused, so no second barrier is needed.  The single-CTA kernel spends ~100 us
 * of its ~170 us in these 32 update steps (one SM, 57 % issue-active); here a step is about a microsecond.
 */
#define PC_CTAS 8
#define PC_THREADS 256

__global__ void __cluster_dims__(PC_CTAS, 1, 1) __launch_bounds__(PC_THREADS)
k_rref_panel_cluster(const i32 *__restrict__ S, int ld, int n, int m, int c0, int *rowstate, int *rank_dev, int *pivcol_out,
                     PanelInfo *info, int chunk, Zp F)
{
	namespace cg = cooperative_groups;
	cg::cluster_group cluster = cg::this_cluster();
	const int crank = (int) cluster.block_rank();
	extern __shared__ unsigned char dyn[];
	i32 *Pw = (i32 *) dyn;                                                   /* chunk x PSTRIDE */
	unsigned char *taken = dyn + (size_t) chunk * PSTRIDE * sizeof(i32);    /* chunk */
	__shared__ int s_slot[3];          /* used in CTA 0 only: winner of a step */
	__shared__ int s_k, s_local, s_sw;
	__shared__ int s_prow[NB], s_pcol[NB];
	__shared__ i32 s_piv[NB], s_dinv[NB];
	__shared__ i64 Mw[NB][2 * NB + 1];
	const int tid = threadIdx.x, T = blockDim.x;
	if (info->fast_done)               /* same decision in every CTA of the cluster (written by the previous kernel) */
		return;
	const int r0 = *rank_dev;
	if (r0 >= n || c0 >= m) {          /* same decision in every CTA of the cluster */
		if (crank == 0 && tid == 0)
			info->k = 0;
		return;
	}
	const int nbw = min(NB, m - c0);
	const int row_lo = crank * chunk, nloc = max(0, min(n, row_lo + chunk) - row_lo);
	for (int idx = tid; idx < nloc * NB; idx += T) {
		int i = idx / NB, c = idx % NB;
		Pw[i * PSTRIDE + c] = (c < nbw) ? S[(size_t) (row_lo + i) * ld + c0 + c] : 0;
	}
	for (int i = tid; i < nloc; i += T)
		taken[i] = rowstate[row_lo + i] >= 0;
	if (tid == 0) {
		s_k = 0;
		s_slot[0] = s_slot[1] = s_slot[2] = 0x7fffffff;
	}
	__syncthreads();
	cluster.sync();                    /* every CTA's shared memory is ready before the first remote access */
	int *slot0 = cluster.map_shared_rank(s_slot, 0);

	for (int c = 0; c < nbw; c++) {
		if (s_k + r0 >= n)             /* s_k evolves identically in every CTA */
			break;
		if (tid == 0)
			s_local = 0x7fffffff;
		__syncthreads();
		for (int i = tid; i < nloc; i += T)
			if (!taken[i] && Pw[i * PSTRIDE + c] != 0) {
				atomicMin(&s_local, row_lo + i);
				break;
			}
		__syncthreads();
		if (tid == 0) {
			if (s_local != 0x7fffffff)
				atomicMin(&slot0[c % 3], s_local);
			if (crank == 0)
				s_slot[(c + 1) % 3] = 0x7fffffff;      /* last read during step c - 2: every CTA is past it */
		}
		cluster.sync();
		const int piv = *((volatile int *) &slot0[c % 3]);
		if (piv == 0x7fffffff)
			continue;
		const int owner = piv / chunk;
		const i32 *prow_remote = cluster.map_shared_rank(Pw, owner) + (size_t) (piv - owner * chunk) * PSTRIDE;
		if (tid < NB)
			s_piv[tid] = prow_remote[tid];
		if (tid == 0) {
			s_prow[s_k] = piv;
			s_pcol[s_k] = c;
			s_k += 1;
			if (owner == crank)
				taken[piv - row_lo] = 1;
		}
		__syncthreads();
		/* fraction-free elimination of column c from the local rows, two threads per row (column c itself is never
		 * read again and is left as it is) */
		const i64 d = s_piv[c];
		const int count = nbw - c - 1, h0 = (count + 1) / 2;
		for (int idx = tid; idx < 2 * nloc; idx += T) {
			const int i = idx >> 1, half = idx & 1;
			if (taken[i])
				continue;
			const i64 l = Pw[i * PSTRIDE + c];
			if (l == 0)
				continue;
			const int lo = c + 1 + (half ? h0 : 0), hi = half ? nbw : c + 1 + h0;
			for (int cc = lo; cc < hi; cc++)
				Pw[i * PSTRIDE + cc] = zp_reduce(d * Pw[i * PSTRIDE + cc] - l * s_piv[cc], F);
		}
		__syncthreads();
	}
	__syncthreads();
	const int k = s_k;
	cluster.sync();                    /* nobody reads a neighbour's shared memory after this point */
	if (crank != 0)
		return;
	if (tid == 0)
		info->k = k;
	if (k == 0)
		return;
	panel_invert_and_publish(S, ld, c0, k, r0, s_prow, s_pcol, Mw, s_dinv, &s_sw, rowstate, rank_dev, pivcol_out, info, F);
}

/*
 * Sampled panel factorisation (one CTA, ~1/3 of the time of the cluster kernel).  WHICH row carries a pivot is free: the
 * reduced echelon form of the block and its pivot columns (the column rank profile) do not depend on it.  So the pivots
 * of a panel are looked for among PS_ROWS rows only -- the first rows that are not pivotal yet and hold a non-zero entry
 * in the panel -- by the same forward elimination, in shared memory, without cluster barriers.  If every column of the
 * panel gets a pivot (or the rows run out), each of these columns is independent of the pivot columns before it in the
 * sample, hence in the whole block: the panel is settled exactly as the full kernels would settle it (same pivot columns,
 * a valid choice of pivot rows) and info->fast_done tells them to return at once.  If some column finds no pivot in the
 * sample (it may still have one elsewhere: rank-deficient or very sparse blocks) nothing is published and the full
 * kernel runs.  Same M^-1 tail as the other kernels.
 */
#define PS_ROWS 96
#define PS_THREADS 256
#define PS_THREADS_WIDE 1024      /* SPASM_B200_PANEL_WIDE=1: one thread per (row, column) of the sample; see DESIGN.md 5.3 */

template <int PS_T>
__global__ void __launch_bounds__(PS_T)
k_rref_panel_sample(const i32 *__restrict__ S, int ld, int n, int m, int c0, int *rowstate, int *rank_dev, int *pivcol_out, PanelInfo *info, Zp F)
{
	__shared__ i32 Pw[PS_ROWS * PSTRIDE];
	__shared__ int s_rows[PS_ROWS];
	__shared__ unsigned char used[PS_ROWS];
	__shared__ int s_warp[PS_T / 32];
	__shared__ int s_base, s_min, s_k, s_sw;
	__shared__ int s_prow[NB], s_pcol[NB];
	__shared__ i32 s_piv[NB], s_dinv[NB];
	__shared__ i64 Mw[NB][2 * NB + 1];
	const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5;
	const int r0 = *rank_dev;
	if (r0 >= n || c0 >= m) {
		if (tid == 0) {
			info->k = 0;
			info->fast_done = 1;
		}
		return;
	}
	const int nbw = min(NB, m - c0);
	/* 1. the sample: first PS_ROWS rows that are not pivotal and not zero on the panel, in row order */
	if (tid == 0)
		s_base = 0;
	__syncthreads();
	bool scanned_all = true;
	for (int i0 = 0; i0 < n; i0 += T) {
		const int i = i0 + tid;
		bool cand = false;
		if (i < n && rowstate[i] < 0) {
			const i32 *row = S + (size_t) i * ld + c0;
			i32 any = 0;
			if (nbw == NB && ((((size_t) row) & 15) == 0)) {
				const int4 *r4 = reinterpret_cast<const int4 *>(row);
#pragma unroll
				for (int u = 0; u < NB / 4; u++) {
					const int4 v = r4[u];
					any |= v.x | v.y | v.z | v.w;
				}
			} else {
				for (int c = 0; c < nbw; c++)
					any |= row[c];
			}
			cand = any != 0;
		}
		const unsigned bal = __ballot_sync(0xffffffffu, cand);
		if (lane == 0)
			s_warp[warp] = __popc(bal);
		__syncthreads();
		int before = s_base;
		for (int w = 0; w < warp; w++)
			before += s_warp[w];
		const int pos = before + __popc(bal & ((1u << lane) - 1));
		if (cand && pos < PS_ROWS)
			s_rows[pos] = i;
		__syncthreads();
		if (tid == 0) {
			int tot = 0;
			for (int w = 0; w < T / 32; w++)
				tot += s_warp[w];
			s_base += tot;
		}
		__syncthreads();
		if (s_base >= PS_ROWS && i0 + T < n) {
			scanned_all = false;
			break;
		}
	}
	/* the sample holds EVERY row that is not zero on the panel: a column without a pivot in it has none at all */
	const bool complete = scanned_all && s_base <= PS_ROWS;
	const int ns = min(s_base, PS_ROWS);
	for (int idx = tid; idx < ns * NB; idx += T) {
		const int r = idx / NB, c = idx % NB;
		Pw[r * PSTRIDE + c] = (c < nbw) ? S[(size_t) s_rows[r] * ld + c0 + c] : 0;
	}
	for (int r = tid; r < ns; r += T)
		used[r] = 0;
	if (tid == 0)
		s_k = 0;
	__syncthreads();
	/* 2. forward elimination on the sample, column by column (fraction-free, like the full kernels) */
	bool settled = true;
	for (int c = 0; c < nbw; c++) {
		if (s_k + r0 >= n)
			break;                           /* every row of the block carries a pivot: the other columns have none */
		if (tid == 0)
			s_min = 0x7fffffff;
		__syncthreads();
		if (tid < ns && !used[tid] && Pw[tid * PSTRIDE + c] != 0)
			atomicMin(&s_min, tid);
		__syncthreads();
		const int piv = s_min;
		if (piv == 0x7fffffff) {
			if (complete) {
				__syncthreads();             /* everybody has read s_min before it is reset */
				continue;
			}
			settled = false;                 /* uniform: this column may have its pivot outside the sample */
			break;
		}
		if (tid < NB)
			s_piv[tid] = Pw[piv * PSTRIDE + tid];
		if (tid == 0) {
			s_prow[s_k] = s_rows[piv];
			s_pcol[s_k] = c;
			s_k += 1;
			used[piv] = 1;
		}
		__syncthreads();
		const i64 d = s_piv[c];
		if (PS_T >= PS_THREADS_WIDE) {
			/* one thread per entry: a step is one multiply-subtract-reduce deep instead of a serial loop over half a row
			 * (the 16-iteration loops below, with two threads per row and eight warps to hide their latency, are what
			 * the 1.5 us of a column step are made of); column c itself is only read */
			for (int idx = tid; idx < ns * NB; idx += T) {
				const int r = idx / NB, cc = idx % NB;
				if (cc <= c || cc >= nbw || used[r])
					continue;
				const i64 l = Pw[r * PSTRIDE + c];
				if (l != 0)
					Pw[r * PSTRIDE + cc] = zp_reduce(d * Pw[r * PSTRIDE + cc] - l * s_piv[cc], F);
			}
		} else {
			const int count = nbw - c - 1, h0 = (count + 1) / 2;
			for (int idx = tid; idx < 2 * ns; idx += T) {
				const int r = idx >> 1, half = idx & 1;
				if (used[r])
					continue;
				const i64 l = Pw[r * PSTRIDE + c];
				if (l == 0)
					continue;
				const int lo = c + 1 + (half ? h0 : 0), hi = half ? nbw : c + 1 + h0;
				for (int cc = lo; cc < hi; cc++)
					Pw[r * PSTRIDE + cc] = zp_reduce(d * Pw[r * PSTRIDE + cc] - l * s_piv[cc], F);
			}
		}
		__syncthreads();
	}
	__syncthreads();
	if (!settled) {
		if (tid == 0)
			info->fast_done = 0;
		return;
	}
	const int k = s_k;
	if (tid == 0) {
		info->k = k;
		info->fast_done = 1;
	}
	if (k == 0)
		return;
	panel_invert_and_publish(S, ld, c0, k, r0, s_prow, s_pcol, Mw, s_dinv, &s_sw, rowstate, rank_dev, pivcol_out, info, F);
}

/* W[i][t] = sum_s S[i][c0 + pcol[s]] * Minv[s][t]   (rows of the panel's own pivots: (s == t) - Minv[s][t]); one thread per (i, t) */
__global__ void __launch_bounds__(256)
k_rref_multipliers(const i32 *__restrict__ S, int ld, int n, int c0, const PanelInfo *__restrict__ info, i32 *W, Zp F)
{
	__shared__ i32 sMinv[NB][NB + 1];
	__shared__ int s_prow[NB], s_pcol[NB];
	const int k = info->k;
	if (k == 0)
		return;
	for (int idx = threadIdx.x; idx < NB * NB; idx += blockDim.x)
		sMinv[idx / NB][idx % NB] = info->Minv[idx / NB][idx % NB];
	if (threadIdx.x < k) {             /* the panel kernel wrote k entries */
		s_prow[threadIdx.x] = info->prow[threadIdx.x];
		s_pcol[threadIdx.x] = info->pcol[threadIdx.x];
	}
	__syncthreads();
	const int t = threadIdx.x % NB;
	for (int i = blockIdx.x * (blockDim.x / NB) + threadIdx.x / NB; i < n; i += gridDim.x * (blockDim.x / NB)) {
		i32 w = 0;
		if (t < k) {
			int mine = -1;
			for (int s = 0; s < k; s++)
				if (s_prow[s] == i)
					mine = s;
			if (mine >= 0) {
				w = zp_reduce((i64) (mine == t) - (i64) sMinv[mine][t], F);
			} else {
				i64 acc = 0;
				int pending = 0;
				for (int s = 0; s < k; s++) {
					acc += (i64) S[(size_t) i * ld + c0 + s_pcol[s]] * (i64) sMinv[s][t];
					if (++pending == F.delay) {
						acc = zp_reduce(acc, F);
						pending = 0;
					}
				}
				w = zp_reduce(acc, F);
			}
		}
		W[(size_t) i * NB + t] = w;
	}
}

/* P[t][j] = S[prow[t]][c0 + j] */
__global__ void k_copy_pivot_rows(const i32 *__restrict__ S, int ld, int c0, int width, const PanelInfo *__restrict__ info, i32 *P, int ldp)
{
	int t = blockIdx.y;
	if (t >= info->k)
		return;
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < width)
		P[(size_t) t * ldp + j] = S[(size_t) info->prow[t] * ld + c0 + j];
}

RrefResult dense_rref(i32 *S, int n, int m, int ld, const Zp &F)
{
	RrefResult out;
	if (n <= 0 || m <= 0)
		return out;
	cudaStream_t s = ctx().stream;
	DevBuf<int> rowstate((size_t) n), rank_dev(1), pivcol((size_t) m);
	DevBuf<i32> W((size_t) n * NB), P((size_t) NB * (size_t) m);
	DevBuf<PanelInfo> info(1);
	rowstate.fill_byte(0xff, s);
	rank_dev.zero(s);
	size_t need = (size_t) n * PSTRIDE * sizeof(i32) + (size_t) n + 16;
	int use_smem = need <= 200 * 1024;
	DevBuf<i32> scratch(use_smem ? 1 : ((size_t) n * PSTRIDE + (size_t) n / 4 + 8));
	if (use_smem)      /* static + dynamic shared memory may exceed the 48 KB default: always opt in */
		CUDA_CHECK(cudaFuncSetAttribute(k_rref_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) need));
	/* panels with enough rows are factorised by a cluster of 8 CTAs (SPASM_B200_PANEL_SINGLE=1 keeps the single CTA) */
	static const bool single = getenv("SPASM_B200_PANEL_SINGLE") != NULL;
	const int chunk = (n + PC_CTAS - 1) / PC_CTAS;
	const size_t cluster_smem = (size_t) chunk * PSTRIDE * sizeof(i32) + (size_t) chunk + 16;
	const bool use_cluster = !single && n >= 256 && cluster_smem <= 160 * 1024;
	if (use_cluster)
		CUDA_CHECK(cudaFuncSetAttribute(k_rref_panel_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) cluster_smem));
	/* LOOK-AHEAD.  The factorisation of a panel occupies 8 SMs for ~120 us and only needs its own 32 columns up to
	 * date; the rank-32 update of everything to the right of it (S[:, c0:] -= W * S[pivot rows, c0:]) occupies the rest
	 * of the GPU.  So the update of panel p is split: the columns of panel p and p+1 on the main stream (they gate
	 * panel p+1), the rest on a second stream, where it overlaps the factorisation of panel p+1:
	 *     main: P(p) M(p) | wait G2(p-1) | C1(p) G1(p)            aux: wait M(p) | C2(p) G2(p)
	 * W, the pivot-row copy and the panel record are double-buffered (G2(p) reads them while panel p+1 is written).
	 * Same arithmetic in the same order on every entry (the two updates of an entry stay ordered by the events). */
	static const bool no_ahead = getenv("SPASM_B200_NO_LOOKAHEAD") != NULL;
	static cudaStream_t s2 = nullptr;
	static cudaEvent_t ev_m[2] = {nullptr, nullptr}, ev_g2[2] = {nullptr, nullptr};
	if (!s2) {
		int prio_least = 0, prio_greatest = 0;
		CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
		CUDA_CHECK(cudaStreamCreateWithPriority(&s2, cudaStreamNonBlocking, prio_least));      /* the main stream has the highest */
		for (int k = 0; k < 2; k++) {
			CUDA_CHECK(cudaEventCreateWithFlags(&ev_m[k], cudaEventDisableTiming));
			CUDA_CHECK(cudaEventCreateWithFlags(&ev_g2[k], cudaEventDisableTiming));
		}
	}
	DevBuf<i32> W2((size_t) n * NB), P2((size_t) NB * (size_t) m);
	DevBuf<PanelInfo> info2(1);
	/* sampled factorisation first (SPASM_B200_PANEL_NO_SAMPLE=1: the full kernels only); fast_done = 0 otherwise */
	static const bool use_sample = getenv("SPASM_B200_PANEL_NO_SAMPLE") == NULL;
	static const bool wide_sample = getenv("SPASM_B200_PANEL_WIDE") != NULL;      /* 1024-thread variant (not the default: measured at the very end of the round, the whole GPU suite did not run on it) */
	info.zero(s);
	info2.zero(s);
	i32 *Wb[2] = {W.ptr, W2.ptr}, *Pb[2] = {P.ptr, P2.ptr};
	PanelInfo *ib[2] = {info.ptr, info2.ptr};
	/* everything queued on the main stream so far (the block itself) precedes the first use of the second stream */
	CUDA_CHECK(cudaEventRecord(ev_g2[1], s));
	CUDA_CHECK(cudaStreamWaitEvent(s2, ev_g2[1], 0));
	int panels = 0;
	bool g2_pending = false;
	for (int c0 = 0; c0 < m; c0 += NB) {
		const int width = m - c0;
		const int b = panels & 1;
		if (use_sample) {
			if (wide_sample)
				k_rref_panel_sample<PS_THREADS_WIDE><<<1, PS_THREADS_WIDE, 0, s>>>(S, ld, n, m, c0, rowstate.ptr, rank_dev.ptr, pivcol.ptr, ib[b], F);
			else
				k_rref_panel_sample<PS_THREADS><<<1, PS_THREADS, 0, s>>>(S, ld, n, m, c0, rowstate.ptr, rank_dev.ptr, pivcol.ptr, ib[b], F);
			LAUNCHED(1);
		}
		if (use_cluster)
			k_rref_panel_cluster<<<PC_CTAS, PC_THREADS, cluster_smem, s>>>(S, ld, n, m, c0, rowstate.ptr, rank_dev.ptr, pivcol.ptr, ib[b], chunk, F);
		else
			k_rref_panel<<<1, 1024, use_smem ? need : 0, s>>>(S, ld, n, m, c0, rowstate.ptr, rank_dev.ptr, pivcol.ptr, Wb[b], ib[b], scratch.ptr, use_smem, F);
		k_rref_multipliers<<<std::min(cdiv(n, 8), 148u * 4), 256, 0, s>>>(S, ld, n, c0, ib[b], Wb[b], F);
		LAUNCHED(2);
		const int near = no_ahead ? width : std::min(width, 2 * NB);      /* columns that gate the next panel */
		if (near < width) {
			CUDA_CHECK(cudaEventRecord(ev_m[b], s));
			CUDA_CHECK(cudaStreamWaitEvent(s2, ev_m[b], 0));
			dim3 g2(cdiv(width - near, 256), NB);
			k_copy_pivot_rows<<<g2, 256, 0, s2>>>(S, ld, c0 + near, width - near, ib[b], Pb[b] + near, m);
			LAUNCHED(1);
		}
		if (g2_pending)                       /* the far update of the previous panel touched the near columns too */
			CUDA_CHECK(cudaStreamWaitEvent(s, ev_g2[b ^ 1], 0));
		dim3 g1(cdiv(near, 256), NB);
		k_copy_pivot_rows<<<g1, 256, 0, s>>>(S, ld, c0, near, ib[b], Pb[b], m);
		LAUNCHED(1);
		gemm_sub_on(s, S + c0, ld, Wb[b], NB, Pb[b], m, n, near, NB, F, &ib[b]->k);
		g2_pending = false;
		if (near < width) {
			gemm_sub_on(s2, S + c0 + near, ld, Wb[b], NB, Pb[b] + near, m, n, width - near, NB, F, &ib[b]->k);
			CUDA_CHECK(cudaEventRecord(ev_g2[b], s2));
			g2_pending = true;
		}
		if (++panels % 8 == 0 && fetch(rank_dev.ptr) >= n)
			break;
	}
	/* the far updates rejoin the main stream */
	CUDA_CHECK(cudaEventRecord(ev_g2[0], s2));
	CUDA_CHECK(cudaStreamWaitEvent(s, ev_g2[0], 0));
	KERNEL_CHECK();
	out.rank = fetch(rank_dev.ptr);
	out.pivcol.resize(out.rank);
	out.pivrow.assign(out.rank, -1);
	std::vector<int> hs((size_t) n);
	rowstate.download(hs.data(), (size_t) n, s);
	pivcol.download(out.pivcol.data(), (size_t) out.rank, s);
	sync();
	for (int i = 0; i < n; i++)
		if (hs[i] >= 0)
			out.pivrow[hs[i]] = i;
	stats().pub.gemm_fieldops += 2.0 * n * (double) m * out.rank;
	return out;
}

/* ================================================================== dense rows -> sparse rows of U */

template <bool EMIT>
__global__ void k_rows_to_csr(const i32 *__restrict__ D, int ld, int nrows, int width, const int *__restrict__ pivcol,
                              const unsigned char *__restrict__ skip, const int *__restrict__ colmap, i64 *len, const i64 *__restrict__ Rp, int *Rj, i32 *Rx)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int t = warp; t < nrows; t += nwarps) {
		const i32 *row = D + (size_t) t * ld;
		i64 out = EMIT ? Rp[t] : 0;
		if (EMIT && lane == 0) {
			Rj[out] = colmap[pivcol[t]];
			Rx[out] = 1;
		}
		i64 written = 1;
		const int own = pivcol[t];
		for (int c0 = 0; c0 < width; c0 += 32) {
			int c = c0 + lane;
			/* skip == NULL: rows that are not reduced against each other (L mode), only the row's own pivot is implied */
			i32 v = (c < width && !(skip ? skip[c] : (unsigned char) (c == own))) ? row[c] : 0;
			unsigned mask = __ballot_sync(0xffffffffu, v != 0);
			if (EMIT && v != 0) {
				i64 pos = out + written + __popc(mask & ((1u << lane) - 1));
				Rj[pos] = colmap[c];
				Rx[pos] = v;
			}
			written += __popc(mask);
		}
		if (!EMIT && lane == 0)
			len[t] = written;
	}
}

void dense_rows_to_csr(const i32 *D, int ld, int nrows, int width, const int *d_pivcol, const unsigned char *d_skip,
                       const int *d_colmap, DevBuf<i64> &Rp, DevBuf<int> &Rj, DevBuf<i32> &Rx, i64 &nnz)
{
	cudaStream_t s = ctx().stream;
	Rp.alloc((size_t) nrows + 1);
	nnz = 0;
	if (nrows == 0) {
		CUDA_CHECK(cudaMemsetAsync(Rp.ptr, 0, sizeof(i64), s));
		return;
	}
	DevBuf<i64> len((size_t) nrows + 1);
	len.zero(s);
	unsigned blocks = std::min(cdiv((size_t) nrows * 32, 256), 148u * 8);
	k_rows_to_csr<false><<<blocks, 256, 0, s>>>(D, ld, nrows, width, d_pivcol, d_skip, d_colmap, len.ptr, nullptr, nullptr, nullptr);
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, len.ptr, Rp.ptr, nrows + 1, s);
	tmp.ensure(bytes + 16);
	cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, len.ptr, Rp.ptr, nrows + 1, s);
	LAUNCHED(2);
	nnz = fetch(Rp.ptr + nrows);
	Rj.alloc((size_t) nnz);
	Rx.alloc((size_t) nnz);
	k_rows_to_csr<true><<<blocks, 256, 0, s>>>(D, ld, nrows, width, d_pivcol, d_skip, d_colmap, nullptr, Rp.ptr, Rj.ptr, Rx.ptr);
	LAUNCHED(1);
	KERNEL_CHECK();
}

}  // namespace sb
