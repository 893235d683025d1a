/*
 * The L side of the factorization (opts->L / opts->complete).
 * reference: src/spasm_echelonize.c:228-313 (update_fact_after_LU), src/spasm_ffpack.cpp:52-75 (spasm_ffpack_LU),
 * src/spasm_pivots.c:421-426, src/spasm_schur.c:164-170, :306-318 (L entries of the sparse stages).
 *
 * What the consumers of L need (src/spasm_solve.c, src/spasm_certificate.c, tests/lu.c):
 *     A[i] = sum_k L[i][k] * U[k]          for every pivotal row i (every row when `complete`),
 *     L restricted to the pivotal rows p[0..r) is lower triangular with a non-zero diagonal
 *     (spasm_dense_back_solve walks k = r-1 .. 0 and eliminates with row p[k]).
 * The rows of U are linearly independent, so L is DETERMINED by (A, U): its row i is the solution of x * U = A[i].
 * It is lower triangular exactly when row k of U is "row p[k] of A reduced by the rows 0..k-1 of U, scaled".
 *
 * The echelonization of this library keeps its dense pivots in REDUCED form (block Gauss-Jordan, dense.cu), which
 * does not have that property.  In L mode every dense block B (already reduced by everything found before) with
 * reduced pivot rows R (identity on the pivot columns P) is therefore re-expressed:  B = C * R  with  C = B[:, P]
 * (rows x rank, full column rank);  C = Pi^t * Lc * Uc  (LU with row pivoting, k_lu_fullcol below);  the rows that
 * go to U are  Uc * R  (unit upper triangular on P), row t coming from row Pi[t] of the block.  The structural rows
 * (one scaled row of the current matrix each) have the property by construction.
 * L itself is then read out of ONE batched solve of the rows of A against the final U (the pull-form solve leaves
 * the elimination coefficients on the pivotal columns of the panel): compute_L, echelonize.cu.
 */
#include "dense.cuh"
#include "lu.cuh"
#include "stats.cuh"

namespace sb {

/*
 * In-place LU of C (n x k, leading dimension ld, rank k): step t takes the first unused row with a non-zero entry on
 * column t as pivot row, scales its tail (columns > t) by the inverse of the pivot, and eliminates column t from the
 * other unused rows.  On exit, for a row used at step t: columns <= t hold its row of Lc (pivot = diagonal of Lc at
 * column t), columns > t hold row t of Uc (unit diagonal implied); an unused row holds its row of Lc.
 * One CTA: the matrices are at most dense_block_size (1000) rows, this is not a throughput kernel.
 */
__global__ void __launch_bounds__(1024) k_lu_fullcol(i32 *C, int ld, int n, int k, int *prow, unsigned char *used, int *fail, Zp F)
{
	extern __shared__ i32 urow[];
	__shared__ int s_piv;
	__shared__ i32 s_inv;
	const int tid = threadIdx.x, nthreads = blockDim.x;
	const int warp = tid >> 5, lane = tid & 31, nwarps = nthreads >> 5;
	for (int i = tid; i < n; i += nthreads)
		used[i] = 0;
	if (tid == 0)
		s_piv = INT_MAX;
	__syncthreads();
	for (int t = 0; t < k; t++) {
		int best = INT_MAX;
		for (int i = tid; i < n; i += nthreads)
			if (!used[i] && C[(size_t) i * ld + t] != 0) {
				best = i;
				break;
			}
		if (best != INT_MAX)
			atomicMin(&s_piv, best);
		__syncthreads();
		const int pr = s_piv;
		if (pr == INT_MAX) {          /* uniform: C does not have full column rank (internal error, reported by the host) */
			if (tid == 0)
				*fail = t + 1;
			return;
		}
		if (tid == 0) {
			s_inv = zp_inverse(C[(size_t) pr * ld + t], F);
			prow[t] = pr;
			used[pr] = 1;
		}
		__syncthreads();
		const i32 inv = s_inv;
		for (int c = t + 1 + tid; c < k; c += nthreads) {
			i32 v = zp_mul(C[(size_t) pr * ld + c], inv, F);
			urow[c] = v;
			C[(size_t) pr * ld + c] = v;
		}
		if (tid == 0)
			s_piv = INT_MAX;
		__syncthreads();
		for (int i = warp; i < n; i += nwarps) {
			if (used[i])
				continue;
			i32 *row = C + (size_t) i * ld;
			const i32 l = row[t];
			if (l == 0)
				continue;
			for (int c = t + 1 + lane; c < k; c += 32)
				row[c] = zp_reduce((i64) row[c] - (i64) l * urow[c], F);
		}
		__syncthreads();
	}
}

/* N (k x ldn) = -Uc, read from the factored C: row t = (0 .. 0, -1, -C[prow[t]][t+1 ..]) */
__global__ void k_lu_neg_upper(const i32 *__restrict__ C, int ld, int k, const int *__restrict__ prow, i32 *N, int ldn)
{
	int t = blockIdx.y;
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= ldn)
		return;
	i32 v = 0;
	if (c == t)
		v = -1;
	else if (c > t && c < k)
		v = -C[(size_t) prow[t] * ld + c];
	N[(size_t) t * ldn + c] = v;
}

void dense_lu_fullcol(i32 *C, int n, int k, int ld, const Zp &F, std::vector<int> &prow)
{
	prow.assign((size_t) k, -1);
	if (k == 0)
		return;
	cudaStream_t s = ctx().stream;
	DevBuf<int> d_prow((size_t) k), fail(1);
	DevBuf<unsigned char> used((size_t) std::max(n, 1));
	fail.zero(s);
	size_t smem = (size_t) k * sizeof(i32);
	if (smem > 200 * 1024)
		errx(1, "[spasm-b200] dense LU: %d pivots in one block do not fit the shared memory of the LU kernel (lower dense_block_size)", k);
	CUDA_CHECK(cudaFuncSetAttribute(k_lu_fullcol, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	k_lu_fullcol<<<1, 1024, smem, s>>>(C, ld, n, k, d_prow.ptr, used.ptr, fail.ptr, F);
	LAUNCHED(1);
	KERNEL_CHECK();
	d_prow.download(prow.data(), (size_t) k, s);
	int f = fetch(fail.ptr);
	if (f != 0)
		errx(1, "[spasm-b200] internal: dense LU found no pivot at step %d of %d", f - 1, k);
}

/* out (k x width, ldo) = Uc * R   where Uc is read from the factored C (dense_lu_fullcol) and R (k x width, ldr) are the reduced rows */
void dense_lu_rows(const i32 *C, int ldc, int k, const std::vector<int> &prow, const i32 *R, int ldr, int width, i32 *out, int ldo, const Zp &F)
{
	if (k == 0)
		return;
	cudaStream_t s = ctx().stream;
	DevBuf<int> d_prow;
	d_prow.upload(prow.data(), prow.size(), s);
	const int ldn = std::max((k + 3) & ~3, 4);
	DevBuf<i32> N((size_t) k * ldn);
	dim3 grid(cdiv(ldn, 256), k);
	k_lu_neg_upper<<<grid, 256, 0, s>>>(C, ldc, k, d_prow.ptr, N.ptr, ldn);
	LAUNCHED(1);
	CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t) k * ldo * sizeof(i32), s));
	dense_gemm_sub(out, ldo, N.ptr, ldn, R, ldr, k, width, k, F);
	sync();              /* prow's device copy and N are locals */
}

}  // namespace sb
