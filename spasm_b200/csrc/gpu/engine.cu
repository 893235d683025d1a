#include "engine.cuh"
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
	dense_ready = false;
	blocks.clear();
	dense_rank = 0;
	Sm0 = 0;
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

void Engine::solve_rows(const DevCsr &B, const int *d_rows, int R, bool skip_first)
{
	if (!G_ready)
		rebuild_schedule();
	GpuTimer t;
	t.start();
	panel.shape(m, R);
	panel_scatter_rows(B, d_rows, R, panel, F, skip_first);
	panel_solve(G, panel.X, panel.ld, R, F);
	stats().pub.ms_solve += t.stop_ms();
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
}

void Engine::block_from_rows(const DevCsr &B, const int *d_rows, int R, DevBuf<i32> &out, int &ldB)
{
	ldB = std::max((Sm0 + 3) & ~3, 4);
	int chunk, begin, end;
	comm_slice(R, &chunk, &begin, &end);
	out.ensure((size_t) chunk * comm_world() * ldB);
	if (end > begin) {
		solve_rows(B, d_rows + begin, end - begin, false);
		gather_q0(out.ptr + (size_t) begin * ldB, ldB);
	}
	comm_allgather_rows(out.ptr, chunk, ldB);
}

void Engine::block_from_combos(const DevCsr &A, const int *d_rows, const i32 *d_coef, int N, int w, DevBuf<i32> &out, int &ldB)
{
	ldB = std::max((Sm0 + 3) & ~3, 4);
	int chunk, begin, end;
	comm_slice(N, &chunk, &begin, &end);
	out.ensure((size_t) chunk * comm_world() * ldB);
	if (end > begin) {
		solve_combos(A, d_rows + (size_t) begin * w, d_coef + (size_t) begin * w, end - begin, w);
		gather_q0(out.ptr + (size_t) begin * ldB, ldB);
	}
	comm_allgather_rows(out.ptr, chunk, ldB);
}

void Engine::gather_q0(i32 *S, int ldS)
{
	panel_gather_dense(panel, d_q0.ptr, Sm0, S, ldS);
}

int Engine::absorb_block(i32 *B, int rows, int ldB)
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
	for (DenseBlock &blk : blocks) {
		const int lda = (blk.rr + 3) & ~3;      /* 16-byte aligned rows for the vector loads of the tensor-core staging */
		Ac.ensure((size_t) rows * lda);
		tg.start();
		dense_gather_columns(B, ldB, rows, blk.d_pivcol.ptr, blk.rr, Ac.ptr, lda);
		dense_gemm_sub(B, ldB, Ac.ptr, lda, blk.D.ptr, blk.ld, rows, Sm0, blk.rr, F);
		gemm_ms += tg.stop_ms();
	}
	RrefResult res = dense_rref(B, rows, Sm0, ldB, F);
	if (res.rank > 0) {
		DenseBlock blk;
		blk.rr = res.rank;
		blk.ld = (Sm0 + 3) & ~3;
		blk.D.alloc((size_t) blk.rr * blk.ld);
		DevBuf<int> d_rows;
		d_rows.upload(res.pivrow.data(), res.pivrow.size(), s);
		dense_gather_rows(B, ldB, d_rows.ptr, blk.rr, Sm0, blk.D.ptr, blk.ld);
		blk.pivcol = res.pivcol;
		blk.d_pivcol.upload(blk.pivcol.data(), blk.pivcol.size(), s);
		std::vector<unsigned char> own((size_t) Sm0, 0);
		for (int c : blk.pivcol)
			own[c] = 1;
		blk.d_own.upload(own.data(), own.size(), s);
		sync();
		dense_rank += blk.rr;
		blocks.push_back(std::move(blk));
	}
	stats().pub.ms_dense += t.stop_ms();
	stats().pub.ms_dense_gemm += gemm_ms;
	return res.rank;
}

}  // namespace sb
