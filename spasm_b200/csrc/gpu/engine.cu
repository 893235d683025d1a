#include "engine.cuh"
#include "lu.cuh"
#include "stats.cuh"

namespace sb {

void Engine::init(int m_, i64 prime_)
{
	m = m_;
	prime = prime_;
	F = make_zp(prime_);
	U = DevCsr();
	U.m = m_;
	U.prime = prime_;
	U.p.alloc(1);
	CUDA_CHECK(cudaMemsetAsync(U.p.ptr, 0, sizeof(i64), ctx().stream));
	Uqinv.alloc((size_t) std::max(m_, 1));
	Uqinv.fill_byte(0xff, ctx().stream);
	G = DepGraph();
	G_ready = false;
	lazy_rows = 0;
	dense_ready = false;
	blocks.clear();
	dense_rank = 0;
	Sm0 = 0;
	want_L = false;
	first_round_rows = 0;
	p_struct.clear();
	p_dense.clear();
}

void Engine::rebuild_schedule()
{
	G = DepGraph();
	depgraph_forward(U, G);
	depgraph_schedule(G);
	G_ready = true;
	stats().pub.dag_depth = G.nlevels;
}

void Engine::begin_dense()
{
	if (dense_ready)
		return;
	std::vector<int> hq((size_t) std::max(m, 1));
	Uqinv.download(hq.data(), (size_t) m, ctx().stream);
	sync();
	q0.clear();
	for (int j = 0; j < m; j++)
		if (hq[j] < 0)
			q0.push_back(j);
	Sm0 = (int) q0.size();
	d_q0.upload(q0.data(), q0.size(), ctx().stream);
	dense_ready = true;
}

void Engine::solve_rows(const DevCsr &B, const int *d_rows, int R, bool skip_first, bool sparse_batch)
{
	if (!G_ready)
		rebuild_schedule();
	GpuTimer t;
	t.start();
	/* Opt-in (SPASM_B200_MASKED_SOLVE=1): on the 500k x 500k configuration the eliminations of a row of the RREF
	 * reach a large share of the pivots, the masks are dense and the plain solve is 5 times faster. */
	const bool use_mask = getenv("SPASM_B200_MASKED_SOLVE") != NULL;       /* read at every call: the tests toggle it */
	const bool masked = sparse_batch && use_mask;
	panel.shape(m, R, masked);
	panel_scatter_rows(B, d_rows, R, panel, F, skip_first);
	if (masked)
		panel_solve_masked(G, panel.X, panel.ld, R, panel.mask.ptr, panel.mw, F);
	else
		panel_solve(G, panel.X, panel.ld, R, F);
	stats().pub.ms_solve += t.stop_ms();
	if (!masked)
		account_bytes(R);
}

void Engine::solve_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w)
{
	if (!G_ready)
		rebuild_schedule();
	GpuTimer t;
	t.start();
	panel.shape(m, N);
	panel_scatter_combos(A, d_rows, d_coef, N, w, panel, F);
	panel_solve(G, panel.X, panel.ld, N, F);
	stats().pub.ms_solve += t.stop_ms();
	account_bytes(N);
}

void Engine::block_from_rows(const DevCsr &B, const int *d_rows, int R, DevBuf<i32> &out, int &ldB)
{
	ldB = std::max((Sm0 + 3) & ~3, 4);
	int chunk, begin, end;
	comm_slice(R, &chunk, &begin, &end);
	out.ensure((size_t) chunk * comm_world() * ldB);
	const int cap = panel_capacity(m);
	for (int at = begin; at < end; at += cap) {
		int R2 = std::min(cap, end - at);
		solve_rows(B, d_rows + at, R2, false);
		gather_q0(out.ptr + (size_t) at * ldB, ldB);
	}
	comm_allgather_rows(out.ptr, chunk, ldB);
}

void Engine::block_from_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w, DevBuf<i32> &out, int &ldB)
{
	ldB = std::max((Sm0 + 3) & ~3, 4);
	int chunk, begin, end;
	comm_slice(N, &chunk, &begin, &end);
	out.ensure((size_t) chunk * comm_world() * ldB);
	const int cap = panel_capacity(m);
	for (int at = begin; at < end; at += cap) {
		int N2 = std::min(cap, end - at);
		solve_combos(A, d_rows + (size_t) at * w, d_coef + (size_t) at * w, N2, w);
		gather_q0(out.ptr + (size_t) at * ldB, ldB);
	}
	comm_allgather_rows(out.ptr, chunk, ldB);
}

/* instrumentation (outside the timed solve): algorithmic bytes of the batch, SURVEY 8d accounting */
void Engine::account_bytes(int R)
{
	static const bool off = getenv("SPASM_B200_NO_BYTE_COUNT") != NULL;
	if (off || U.n == 0)
		return;
	/* reached pivot rows (8 B per entry) + the dense output row (4 B per non-pivotal column); the right-hand sides
	 * themselves (8 B per entry) are small and left out */
	stats().pub.solve_bytes += panel_reached_bytes(panel, Uqinv.ptr, U.p.ptr) + 4.0 * (double) (m - U.n) * R;
}

void Engine::gather_q0(i32 *S, int ldS)
{
	panel_gather_dense(panel, d_q0.ptr, Sm0, S, ldS);
}

int Engine::absorb_block(i32 *B, int rows, int ldB, bool with_lu)
{
	if (rows <= 0 || Sm0 <= 0)
		return 0;
	GpuTimer t, tg;
	t.start();
	cudaStream_t s = ctx().stream;
	/* eliminate the pivots of the earlier dense blocks, block after block (block-triangular solve):
	 *   B <- B - B[:, P_b] * D_b        D_b = [ I on P_b | ... ] */
	double gemm_ms = 0;
	DevBuf<i32> Ac;
	DevBuf<int8_t> Apack;
	for (DenseBlock &blk : blocks) {
		const int lda = (blk.rr + 3) & ~3;
		Ac.ensure((size_t) rows * lda);
		tg.start();
		dense_gather_columns(B, ldB, rows, blk.d_pivcol.ptr, blk.rr, Ac.ptr, lda);
		if (blk.Dpack.ptr && (double) rows * Sm0 * blk.rr >= 4e6) {
			/* tensor cores: the block's rows were split into limb planes when it was created */
			Apack.ensure(umma_packed_bytes(rows, blk.rr, umma_limbs(F)));
			umma_pack(Ac.ptr, lda, rows, blk.rr, true, Apack.ptr, F);
			umma_gemm_sub_packed(B, ldB, Apack.ptr, blk.Dpack.ptr, rows, Sm0, blk.rr, F);
			stats().pub.gemm_fieldops += 2.0 * rows * (double) Sm0 * blk.rr;
		} else {
			dense_gemm_sub(B, ldB, Ac.ptr, lda, blk.D.ptr, blk.ld, rows, Sm0, blk.rr, F);
		}
		gemm_ms += tg.stop_ms();
	}
	/* Echelonize on the columns that are not pivotal yet only: after the reduction the block is zero on the pivot
	 * columns of the earlier blocks, and panels made of such columns would be factorised for nothing.  The block is
	 * compacted (gather of the remaining columns), echelonized, and its pivot rows are expanded back to the q0 space. */
	int remaining = Sm0 - dense_rank;
	DevBuf<i32> Bc;
	const i32 *src = B;
	int lds = ldB;
	std::vector<int> cols;            /* compact column -> q0 column */
	DevBuf<int> d_cols;
	bool compact = !blocks.empty();
	if (compact) {
		std::vector<char> taken((size_t) Sm0, 0);
		for (DenseBlock &blk : blocks)
			for (int cpiv : blk.pivcol)
				taken[cpiv] = 1;
		cols.reserve(remaining);
		for (int cidx = 0; cidx < Sm0; cidx++)
			if (!taken[cidx])
				cols.push_back(cidx);
		d_cols.upload(cols.data(), cols.size(), s);
		lds = std::max((remaining + 3) & ~3, 4);
		Bc.alloc((size_t) rows * lds);
		dense_gather_columns(B, ldB, rows, d_cols.ptr, remaining, Bc.ptr, lds);
		src = Bc.ptr;
	}
	const int width = compact ? remaining : Sm0;
	DevBuf<i32> before;               /* L mode: the block as it is before the echelonization (lu.cu: B = C * R) */
	if (with_lu) {
		before.alloc((size_t) rows * lds);
		CUDA_CHECK(cudaMemcpyAsync(before.ptr, src, (size_t) rows * lds * sizeof(i32), cudaMemcpyDeviceToDevice, s));
	}
	RrefResult res = dense_rref(const_cast<i32 *>(src), rows, width, lds, F);
	if (res.rank > 0) {
		DenseBlock blk;
		blk.rr = res.rank;
		blk.ld = (Sm0 + 3) & ~3;
		blk.D.alloc((size_t) blk.rr * blk.ld);
		DevBuf<int> d_rows;
		d_rows.upload(res.pivrow.data(), res.pivrow.size(), s);
		if (compact) {
			blk.D.zero(s);
			dense_scatter_rows(src, lds, d_rows.ptr, blk.rr, d_cols.ptr, remaining, blk.D.ptr, blk.ld);
			blk.pivcol.resize(res.rank);
			for (int t2 = 0; t2 < res.rank; t2++)
				blk.pivcol[t2] = cols[res.pivcol[t2]];
		} else {
			dense_gather_rows(src, lds, d_rows.ptr, blk.rr, Sm0, blk.D.ptr, blk.ld);
			blk.pivcol = res.pivcol;
		}
		blk.d_pivcol.upload(blk.pivcol.data(), blk.pivcol.size(), s);
		std::vector<unsigned char> own((size_t) Sm0, 0);
		for (int cpiv : blk.pivcol)
			own[cpiv] = 1;
		blk.d_own.upload(own.data(), own.size(), s);
		if (blk.rr >= 64 && umma_gemm_available(F)) {
			blk.Dpack.alloc(umma_packed_bytes(Sm0, blk.rr, umma_limbs(F)));
			umma_pack(blk.D.ptr, blk.ld, Sm0, blk.rr, false, blk.Dpack.ptr, F);
		}
		if (with_lu) {
			/* C = (block before)[:, pivot columns];  C = Pi^t Lc Uc;  rows for U = Uc * D */
			DevBuf<int> d_pc;
			d_pc.upload(res.pivcol.data(), res.pivcol.size(), s);
			const int ldc = std::max((blk.rr + 3) & ~3, 4);
			DevBuf<i32> C((size_t) rows * ldc);
			dense_gather_columns(before.ptr, lds, rows, d_pc.ptr, blk.rr, C.ptr, ldc);
			dense_lu_fullcol(C.ptr, rows, blk.rr, ldc, F, blk.lu_row);
			blk.Dlu.alloc((size_t) blk.rr * blk.ld);
			dense_lu_rows(C.ptr, ldc, blk.rr, blk.lu_row, blk.D.ptr, blk.ld, Sm0, blk.Dlu.ptr, blk.ld, F);
		}
		sync();
		dense_rank += blk.rr;
		blocks.push_back(std::move(blk));
	}
	stats().pub.ms_dense += t.stop_ms();
	stats().pub.ms_dense_gemm += gemm_ms;
	return res.rank;
}

}  // namespace sb
