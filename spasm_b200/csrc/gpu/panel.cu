#include <cub/cub.cuh>
#include "panel.cuh"
#include "stats.cuh"

namespace sb {

void Panel::shape(int nnodes_, int R_, bool masked_)
{
	nnodes = nnodes_;
	R = R_;
	ld = (R_ + 3) & ~3;
	if (ld == 0)
		ld = 4;
	X.ensure((size_t) std::max(nnodes, 1) * ld);
	CUDA_CHECK(cudaMemsetAsync(X.ptr, 0, (size_t) std::max(nnodes, 1) * ld * sizeof(i32), ctx().stream));
	masked = masked_;
	mw = masked ? (ld / 4 + 31) / 32 : 0;
	if (masked) {
		mask.ensure((size_t) std::max(nnodes, 1) * mw);
		CUDA_CHECK(cudaMemsetAsync(mask.ptr, 0, (size_t) std::max(nnodes, 1) * mw * sizeof(unsigned), ctx().stream));
	}
}

int panel_capacity(int nnodes, double gb)
{
	const char *s = getenv("SPASM_B200_PANEL_GB");
	if (s)
		gb = atof(s);
	double cap = gb * 1e9 / (4.0 * std::max(nnodes, 1));
	int R = cap > 32768 ? 32768 : (int) cap;
	R &= ~3;
	return R < 4 ? 4 : R;
}

/* ------------------------------------------------------------------ scatter */

__global__ void k_scatter_rows(int R, const int *__restrict__ rows, const i64 *__restrict__ Bp, const int *__restrict__ Bj,
                               const i32 *__restrict__ Bx, i32 *X, int ld, Zp F, int skip_first, unsigned *mask, int mw)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int r = warp; r < R; r += nwarps) {
		int i = rows[r];
		for (i64 k = Bp[i] + skip_first + lane; k < Bp[i + 1]; k += 32) {
			zp_atomic_add(&X[(size_t) Bj[k] * ld + r], Bx[k], F);
			if (mask)
				atomicOr(&mask[(size_t) Bj[k] * mw + (r >> 7)], 1u << ((r >> 2) & 31));
		}
	}
}

void panel_scatter_rows(const DevCsr &B, const int *d_rows, int R, Panel &P, const Zp &F, bool skip_first, int r_off)
{
	if (R == 0)
		return;
	k_scatter_rows<<<std::min(cdiv((size_t) R * 32, 256), 148u * 16), 256, 0, ctx().stream>>>(R, d_rows, B.p, B.j, B.x, P.X + r_off, P.ld, F, skip_first ? 1 : 0,
	                                                                                             P.masked ? P.mask.ptr : nullptr, P.mw);
	LAUNCHED(1);
	KERNEL_CHECK();
}

__global__ void k_scatter_combos(i64 total, int w, const int *__restrict__ rows, const i32 *__restrict__ coef,
                                 const i64 *__restrict__ Ap, const int *__restrict__ Aj, const i32 *__restrict__ Ax,
                                 i32 *X, int ld, Zp F)
{
	for (i64 t = (i64) blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (i64) gridDim.x * blockDim.x) {
		int k = (int) (t / w);
		i32 c = coef[t];
		if (c == 0)
			continue;
		int i = rows[t];
		for (i64 e = Ap[i]; e < Ap[i + 1]; e++)
			zp_atomic_add(&X[(size_t) Aj[e] * ld + k], zp_mul(c, Ax[e], F), F);
	}
}

void panel_scatter_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w, Panel &P, const Zp &F, int r_off)
{
	i64 total = (i64) N * w;
	if (total == 0)
		return;
	k_scatter_combos<<<std::min(cdiv((size_t) total, 256), 148u * 32), 256, 0, ctx().stream>>>(total, w, d_rows, d_coef, A.p, A.j, A.x, P.X + r_off, P.ld, F);
	LAUNCHED(1);
	KERNEL_CHECK();
}

__global__ void k_scatter_columns(int n, const i64 *__restrict__ Up, const int *__restrict__ Uj, const i32 *__restrict__ Ux,
                                  const int *__restrict__ colslot, i32 *X, int ld, Zp F)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int i = warp; i < n; i += nwarps)
		for (i64 k = Up[i] + lane; k < Up[i + 1]; k += 32) {
			int slot = colslot[Uj[k]];
			if (slot >= 0)
				zp_atomic_add(&X[(size_t) i * ld + slot], Ux[k], F);
		}
}

void panel_scatter_columns(const DevCsr &U, const int *d_colslot, Panel &P)
{
	if (U.n == 0)
		return;
	Zp F = make_zp(U.prime);
	k_scatter_columns<<<std::min(cdiv((size_t) U.n * 32, 256), 148u * 16), 256, 0, ctx().stream>>>(U.n, U.p, U.j, U.x, d_colslot, P.X, P.ld, F);
	LAUNCHED(1);
	KERNEL_CHECK();
}

/* ------------------------------------------------------------------ gather */

/* 32 x 32 tile transpose through shared memory: reads are contiguous in r, writes contiguous in c */
__global__ void k_gather_dense(int R, int Sm, const int *__restrict__ q, const i32 *__restrict__ X, int ld, i32 *S, int ldS)
{
	__shared__ i32 tile[32][33];
	int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		int c = c0 + dy, r = r0 + threadIdx.x;
		tile[dy][threadIdx.x] = (c < Sm && r < R) ? X[(size_t) q[c] * ld + r] : 0;
	}
	__syncthreads();
	for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
		int r = r0 + dy, c = c0 + threadIdx.x;
		if (r < R && c < Sm)
			S[(size_t) r * ldS + c] = tile[threadIdx.x][dy];
	}
}

void panel_gather_dense(const Panel &P, const int *d_q, int Sm, i32 *S, int ldS, int r_off, int count)
{
	const int R = count < 0 ? P.R - r_off : count;
	if (R <= 0 || Sm == 0)
		return;
	dim3 grid(cdiv(Sm, 32), cdiv(R, 32)), block(32, 8);
	k_gather_dense<<<grid, block, 0, ctx().stream>>>(R, Sm, d_q, P.X + r_off, P.ld, S, ldS);
	LAUNCHED(1);
	KERNEL_CHECK();
}

#define CHUNK 64

/* blockIdx.y = chunk of CHUNK nodes, threadIdx.x -> right-hand side (coalesced); writes the number of kept
 * entries of (chunk, r) when part != NULL, or emits them when Sj != NULL */
template <bool EMIT>
__global__ void k_panel_sparse(int nnodes, int R, const i32 *__restrict__ X, int ld, const int *__restrict__ flag,
                               int *part, const i64 *__restrict__ Sp, const int *__restrict__ node_to_col, int *Sj, i32 *Sx, int has_first,
                               const unsigned *__restrict__ mask, int mw)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= R)
		return;
	int chunk = blockIdx.y;
	int c0 = chunk * CHUNK, c1 = min(nnodes, c0 + CHUNK);
	int count = 0;
	i64 base = EMIT ? Sp[r] + has_first + part[(size_t) chunk * R + r] : 0;
	for (int c = c0; c < c1; c++) {
		if (flag && flag[c] >= 0)
			continue;
		/* masked panels: one word per 128 right-hand sides says which groups of 4 can be non-zero (a warp reads one word) */
		if (mask && !((mask[(size_t) c * mw + (r >> 7)] >> ((r >> 2) & 31)) & 1u))
			continue;
		i32 v = X[(size_t) c * ld + r];
		if (v == 0)
			continue;
		if (EMIT) {
			Sj[base + count] = node_to_col ? node_to_col[c] : c;
			Sx[base + count] = v;
		}
		count++;
	}
	if (!EMIT)
		part[(size_t) chunk * R + r] = count;
}

/* per right-hand side: exclusive prefix over the chunks, total in rowlen */
__global__ void k_chunk_prefix(int nchunks, int R, int *part, i64 *rowlen, int has_first)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= R)
		return;
	int run = 0;
	for (int ch = 0; ch < nchunks; ch++) {
		int c = part[(size_t) ch * R + r];
		part[(size_t) ch * R + r] = run;
		run += c;
	}
	rowlen[r] = run + has_first;
}

__global__ void k_emit_first(int R, const i64 *__restrict__ Sp, const int *__restrict__ first_col, i32 first_val, int *Sj, i32 *Sx)
{
	int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r < R) {
		Sj[Sp[r]] = first_col[r];
		Sx[Sp[r]] = first_val;
	}
}

__global__ void k_sum_i64(int n, const i64 *__restrict__ v, unsigned long long *out)
{
	unsigned long long acc = 0;
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x)
		acc += (unsigned long long) v[t];
	for (int o = 16; o > 0; o >>= 1)
		acc += __shfl_down_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc)
		atomicAdd(out, acc);
}

i64 panel_count_nonzero(const Panel &P, const int *d_flag, int R_limit)
{
	const int R = R_limit < 0 ? P.R : std::min(P.R, R_limit);
	if (R == 0 || P.nnodes == 0)
		return 0;
	cudaStream_t s = ctx().stream;
	int nchunks = cdiv(P.nnodes, CHUNK);
	DevBuf<int> part((size_t) nchunks * R);
	DevBuf<i64> rowlen((size_t) R);
	DevBuf<unsigned long long> total(1);
	total.zero(s);
	dim3 grid(cdiv(R, 128), nchunks);
	k_panel_sparse<false><<<grid, 128, 0, s>>>(P.nnodes, R, P.X, P.ld, d_flag, part.ptr, nullptr, nullptr, nullptr, nullptr, 0,
	                                           P.masked ? P.mask.ptr : nullptr, P.mw);
	k_chunk_prefix<<<cdiv(R, 128), 128, 0, s>>>(nchunks, R, part.ptr, rowlen.ptr, 0);
	k_sum_i64<<<std::min(cdiv(R, 256), 64u), 256, 0, s>>>(R, rowlen.ptr, total.ptr);
	LAUNCHED(3);
	KERNEL_CHECK();
	return (i64) fetch(total.ptr);
}

void panel_to_csr(const Panel &P, const int *d_flag, const int *d_first_col, i32 first_val, const int *d_node_to_col,
                  DevBuf<i64> &Sp, DevBuf<int> &Sj, DevBuf<i32> &Sx, i64 &nnz, bool count_only)
{
	cudaStream_t s = ctx().stream;
	int R = P.R;
	Sp.alloc((size_t) R + 1);
	nnz = 0;
	if (R == 0) {
		CUDA_CHECK(cudaMemsetAsync(Sp.ptr, 0, sizeof(i64), s));
		return;
	}
	int nchunks = std::max(1u, cdiv(P.nnodes, CHUNK));
	int has_first = d_first_col ? 1 : 0;
	DevBuf<int> part((size_t) nchunks * R);
	DevBuf<i64> rowlen((size_t) R + 1);
	rowlen.zero(s);
	dim3 grid(cdiv(R, 128), nchunks);
	k_panel_sparse<false><<<grid, 128, 0, s>>>(P.nnodes, R, P.X, P.ld, d_flag, part.ptr, nullptr, nullptr, nullptr, nullptr, 0,
	                                           P.masked ? P.mask.ptr : nullptr, P.mw);
	k_chunk_prefix<<<cdiv(R, 128), 128, 0, s>>>(nchunks, R, part.ptr, rowlen.ptr, has_first);
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, rowlen.ptr, Sp.ptr, R + 1, s);
	tmp.ensure(bytes + 16);
	cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, rowlen.ptr, Sp.ptr, R + 1, s);
	LAUNCHED(3);
	nnz = fetch(Sp.ptr + R);
	Sj.alloc((size_t) std::max<i64>(nnz, 1));
	Sx.alloc((size_t) std::max<i64>(nnz, 1));
	if (nnz > 0 && !count_only) {
		k_panel_sparse<true><<<grid, 128, 0, s>>>(P.nnodes, R, P.X, P.ld, d_flag, part.ptr, Sp.ptr, d_node_to_col, Sj.ptr, Sx.ptr, has_first,
		                                          P.masked ? P.mask.ptr : nullptr, P.mw);
		LAUNCHED(1);
		if (has_first) {
			k_emit_first<<<cdiv(R, 128), 128, 0, s>>>(R, Sp.ptr, d_first_col, first_val, Sj.ptr, Sx.ptr);
			LAUNCHED(1);
		}
	}
	KERNEL_CHECK();
}

}  // namespace sb

namespace sb {

/* Algorithmic bytes of the solved batch in the reference's per-row accounting:
 * 8 bytes per entry of every pivotal row of U that a right-hand side reaches.  A pivot is counted as reached when its
 * elimination coefficient X[pivot column][r] is non-zero (the structural reach can only be larger). */
__global__ void k_reached_bytes(int m, int R, const i32 *__restrict__ X, int ld, const int *__restrict__ qinv,
                                const i64 *__restrict__ Up, unsigned long long *total)
{
	unsigned long long acc = 0;
	for (int c = blockIdx.x; c < m; c += gridDim.x) {
		int row = qinv[c];
		if (row < 0)
			continue;
		unsigned long long w = (unsigned long long) (Up[row + 1] - Up[row]);
		int cnt = 0;
		for (int r = threadIdx.x; r < R; r += blockDim.x)
			cnt += X[(size_t) c * ld + r] != 0;
		acc += w * (unsigned long long) cnt;
	}
	for (int o = 16; o > 0; o >>= 1)
		acc += __shfl_down_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc)
		atomicAdd(total, 8ull * acc);
}

double panel_reached_bytes(const Panel &P, const int *d_qinv, const i64 *d_Up)
{
	if (P.R == 0 || P.nnodes == 0)
		return 0;
	DevBuf<unsigned long long> total(1);
	total.zero(ctx().stream);
	k_reached_bytes<<<std::min(P.nnodes, 148 * 16), 128, 0, ctx().stream>>>(P.nnodes, P.R, P.X, P.ld, d_qinv, d_Up, total.ptr);
	LAUNCHED(1);
	return (double) fetch(total.ptr);
}

}  // namespace sb

namespace sb {

/*
 * Entries of the sparse Schur rows in the REFERENCE's order.  The reference stores a row of S in the order of the
 * pattern xj[top:m] produced by its depth-first reach (src/spasm_reach.c:21-135, src/spasm_schur.c:156-171), and that
 * order decides which entry the next pivot round sees first (src/spasm_pivots.c:104-116, :233-237).  A non-pivotal
 * column is a leaf of the DFS and is finished the moment it is first visited, and xj is filled from the end, so the
 * kept entries of a row are its non-zero leaves in REVERSE order of first visit.  One thread replays the DFS of one
 * row (same traversal: start columns in row order, row entries in CSR order) with a private mark bitmap and an
 * explicit stack whose depth is bounded by the depth of the pivot DAG; the values come from the solved panel.
 */
__global__ void k_schur_emit_dfs(int R, const int *__restrict__ rows, const i64 *__restrict__ Bp, const int *__restrict__ Bj,
                                 const i64 *__restrict__ Up, const int *__restrict__ Uj, const int *__restrict__ qinv,
                                 const i32 *__restrict__ X, int ld, const i64 *__restrict__ Sp, int *Sj, i32 *Sx,
                                 int words, unsigned *marks_pool, int *stack_pool, int maxdepth)
{
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	const int nslots = gridDim.x * blockDim.x;
	unsigned *mark = marks_pool + (size_t) slot * words;
	int *st_col = stack_pool + (size_t) slot * 2 * maxdepth;
	int *st_next = st_col + maxdepth;
	for (int r = slot; r < R; r += nslots) {
		for (int w = 0; w < words; w++)
			mark[w] = 0;
		i64 pos = Sp[r + 1] - 1;                     /* filled from the end, like xj */
		const int brow = rows[r];
		for (i64 e = Bp[brow]; e < Bp[brow + 1]; e++) {
			const int jstart = Bj[e];
			if (mark[jstart >> 5] & (1u << (jstart & 31)))
				continue;
			int head = 0;
			st_col[0] = jstart;
			while (head >= 0) {
				const int j = st_col[head];
				const int i = qinv[j];
				if (!(mark[j >> 5] & (1u << (j & 31)))) {
					mark[j >> 5] |= 1u << (j & 31);
					st_next[head] = 0;
				}
				if (i < 0) {
					const i32 v = X[(size_t) j * ld + r];
					if (v != 0) {
						Sj[pos] = j;
						Sx[pos] = v;
						pos--;
					}
					head--;
					continue;
				}
				const i64 base = Up[i];
				const int len = (int) (Up[i + 1] - base);
				bool pushed = false;
				for (int k = st_next[head]; k < len; k++) {
					const int c = Uj[base + k];
					if (mark[c >> 5] & (1u << (c & 31)))
						continue;
					st_next[head] = k + 1;
					if (head + 1 < maxdepth) {
						head++;
						st_col[head] = c;
						pushed = true;
					}
					break;
				}
				if (!pushed)
					head--;
			}
		}
	}
}

void panel_emit_reference_order(const Panel &P, const DevCsr &B, const int *d_rows, const DevCsr &U, const int *d_qinv,
                                int dag_depth, const i64 *d_Sp, int *d_Sj, i32 *d_Sx)
{
	if (P.R == 0)
		return;
	cudaStream_t s = ctx().stream;
	int words = (P.nnodes + 31) / 32;
	int maxdepth = dag_depth + 8;
	int threads = 64;
	int blocks = std::min<int>(cdiv(P.R, threads), 148 * 8);
	size_t nslots = (size_t) blocks * threads;
	DevBuf<unsigned> marks(nslots * words);
	DevBuf<int> stacks(nslots * 2 * maxdepth);
	k_schur_emit_dfs<<<blocks, threads, 0, s>>>(P.R, d_rows, B.p, B.j, U.p, U.j, d_qinv, P.X, P.ld, d_Sp, d_Sj, d_Sx, words, marks.ptr,
	                                           stacks.ptr, maxdepth);
	LAUNCHED(1);
	KERNEL_CHECK();
}

}  // namespace sb
