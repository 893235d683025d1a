/*
 * The remaining GPU entry points of include/spasm.h, as thin host wrappers:
 * they move the caller's host matrices to HBM, run the same kernel families as
 * spasm_echelonize, and hand back malloc'ed host results.
 */
#include <math.h>
#include <functional>
#include "engine.cuh"
#include "lu.cuh"
#include "stats.cuh"

namespace sb {
int extract_structural(Engine &E, const DevCsr &A, const int *p_in, int *p, bool greedy, int round, bool allow_lazy);
double estimate_density(Engine &E, const DevCsr &A, const int *p, int n, int R);
void schur_sparse(Engine &E, const DevCsr &A, const int *p, int n, DevCsr &S, const std::function<void(int, int)> &after_batch);

#define LOG(...) do { if (ctx().verbose) { fprintf(stderr, __VA_ARGS__); fflush(stderr); } } while (0)

/* engine over a caller-provided echelon form (host U + qinv) */
static void engine_from_host(Engine &E, const struct spasm_csr *U, const int *qinv)
{
	E.init(U->m, spasm_get_prime(U));
	if (U->n > 0) {
		struct spasm_csr view = *U;
		E.U.upload(&view);
	}
	E.U.m = U->m;
	E.U.prime = spasm_get_prime(U);
	E.Uqinv.upload(qinv, (size_t) U->m, ctx().stream);
	E.rebuild_schedule();
}

/* device CSR -> malloc'ed host CSR with n rows */
static struct spasm_csr *csr_to_host(const DevCsr &S)
{
	struct spasm_csr *H = spasm_csr_alloc(S.n, S.m, std::max<i64>(S.nnz, 1), S.prime, true);
	cudaStream_t s = ctx().stream;
	S.p.download(H->p, (size_t) S.n + 1, s);
	S.j.download(H->j, (size_t) S.nnz, s);
	S.x.download(H->x, (size_t) S.nnz, s);
	sync();
	stats().pub.d2h_bytes += S.nnz * 8 + (i64) (S.n + 1) * 8;
	spasm_csr_realloc(H, -1);
	return H;
}

/* concatenate per-batch CSR pieces (device) into one host CSR */
/* a batch of result rows, kept in HBM until the whole result is known: the host matrix is then allocated once and
 * every batch is downloaded straight to its place (no intermediate host copy: a 3 GB result used to be touched
 * three times by one host thread) */
struct HostPiece {
	DevBuf<i64> p;
	DevBuf<int> j;
	DevBuf<i32> x;
	int rows = 0;
	i64 nnz = 0;
};

static void piece_download(DevBuf<i64> &Sp, DevBuf<int> &Sj, DevBuf<i32> &Sx, int rows, i64 nnz, HostPiece &out)
{
	out.p = std::move(Sp);
	out.j = std::move(Sj);
	out.x = std::move(Sx);
	out.rows = rows;
	out.nnz = nnz;
}

static struct spasm_csr *pieces_to_host(const std::vector<HostPiece> &pieces, int n, int m, i64 prime)
{
	cudaStream_t s = ctx().stream;
	i64 total = 0;
	for (const HostPiece &pc : pieces)
		total += pc.nnz;
	struct spasm_csr *H = spasm_csr_alloc(n, m, std::max<i64>(total, 1), prime, true);
	i64 off = 0;
	int row = 0;
	std::vector<i64> hp;
	for (const HostPiece &pc : pieces) {
		hp.resize((size_t) pc.rows + 1);
		pc.p.download(hp.data(), hp.size(), s);
		pc.j.download(H->j + off, (size_t) pc.nnz, s);
		pc.x.download(H->x + off, (size_t) pc.nnz, s);
		sync();
		for (int t = 0; t < pc.rows; t++)
			H->p[row + t + 1] = off + hp[t + 1];
		off += pc.nnz;
		row += pc.rows;
		stats().pub.d2h_bytes += pc.nnz * 8 + (i64) (pc.rows + 1) * 8;
	}
	for (; row < n; row++)
		H->p[row + 1] = off;
	return H;
}

/*
 * Multi-GPU exchange of result rows (spasm_rref shards the rows of U, spasm_kernel the non-pivotal
 * columns; reference loops: src/spasm_rref.c:44-56, src/spasm_kernel.c:42-51 are independent-iteration).
 * Every rank computed the pieces of ITS contiguous slice; on return `pieces` holds the pieces of every rank in rank
 * order (= row order of the result), in HBM.  One all-gather of the piece sizes, then one NCCL group of broadcasts,
 * each piece from its owner (variable lengths, no padding).
 */
#define MAX_PIECES_PER_RANK 4095
static void exchange_pieces(std::vector<HostPiece> &pieces)
{
	const int world = comm_world(), me = comm_rank();
	if (world == 1)
		return;
	cudaStream_t s = ctx().stream;
	if (pieces.size() > MAX_PIECES_PER_RANK)
		errx(1, "[spasm-b200] internal: too many result batches on one rank");
	const size_t per = 2 * MAX_PIECES_PER_RANK + 2;
	std::vector<i64> meta(per * world, 0);
	i64 *mine = meta.data() + per * me;
	mine[0] = (i64) pieces.size();
	for (size_t k = 0; k < pieces.size(); k++) {
		mine[1 + 2 * k] = pieces[k].rows;
		mine[2 + 2 * k] = pieces[k].nnz;
	}
	DevBuf<i64> d_meta;
	d_meta.upload(meta.data(), meta.size(), s);
	comm_allgather_bytes(d_meta.ptr, per * sizeof(i64));
	d_meta.download(meta.data(), meta.size(), s);
	sync();
	const int root = comm_result_root();
	if (root >= 0 && me != root) {
		/* gather to one rank (spasm_b200_comm_result_root): this rank only ships its pieces and returns an empty result */
		comm_group_begin();
		for (HostPiece &pc : pieces) {
			comm_send_bytes(pc.p.ptr, ((size_t) pc.rows + 1) * sizeof(i64), root);
			comm_send_bytes(pc.j.ptr, (size_t) pc.nnz * sizeof(int), root);
			comm_send_bytes(pc.x.ptr, (size_t) pc.nnz * sizeof(i32), root);
		}
		comm_group_end();
		sync();
		pieces.clear();
		return;
	}
	std::vector<HostPiece> all;
	for (int r = 0; r < world; r++) {
		const i64 *mr = meta.data() + per * r;
		for (i64 k = 0; k < mr[0]; k++) {
			if (r == me) {
				all.push_back(std::move(pieces[k]));
			} else {
				all.emplace_back();
				HostPiece &pc = all.back();
				pc.rows = (int) mr[1 + 2 * k];
				pc.nnz = mr[2 + 2 * k];
				pc.p.alloc((size_t) pc.rows + 1);
				pc.j.alloc((size_t) std::max<i64>(pc.nnz, 1));
				pc.x.alloc((size_t) std::max<i64>(pc.nnz, 1));
			}
		}
	}
	comm_group_begin();
	size_t at = 0;
	for (int r = 0; r < world; r++) {
		const i64 *mr = meta.data() + per * r;
		for (i64 k = 0; k < mr[0]; k++, at++) {
			HostPiece &pc = all[at];
			if (root >= 0) {
				if (r != me) {
					comm_recv_bytes(pc.p.ptr, ((size_t) pc.rows + 1) * sizeof(i64), r);
					comm_recv_bytes(pc.j.ptr, (size_t) pc.nnz * sizeof(int), r);
					comm_recv_bytes(pc.x.ptr, (size_t) pc.nnz * sizeof(i32), r);
				}
				continue;
			}
			comm_bcast_bytes(pc.p.ptr, ((size_t) pc.rows + 1) * sizeof(i64), r);
			comm_bcast_bytes(pc.j.ptr, (size_t) pc.nnz * sizeof(int), r);
			comm_bcast_bytes(pc.x.ptr, (size_t) pc.nnz * sizeof(i32), r);
		}
	}
	comm_group_end();
	pieces = std::move(all);
}

/* flag[c] = -1 on pivotal columns (the selection convention of panel_to_csr) */
__global__ void k_flag_pivotal(int m, const int *__restrict__ qinv, int *flag)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < m)
		flag[c] = qinv[c] >= 0 ? -1 : 0;
}

/* The L coefficients of a solved batch (reference: src/spasm_schur.c:306-318): for right-hand side r and every pivotal
 * column j whose multiplier x[j] = X[j][r] is non-zero, the triplet (row_out[r], Uqinv[j], x[j]).  The pull-form solve
 * leaves exactly these multipliers on the pivotal columns of the panel. */
static void append_L_from_panel(Engine &E, struct spasm_triplet *L, const int *row_out, int R)
{
	cudaStream_t s = ctx().stream;
	DevBuf<int> flag((size_t) E.m);
	k_flag_pivotal<<<cdiv(E.m, 256), 256, 0, s>>>(E.m, E.Uqinv.ptr, flag.ptr);
	LAUNCHED(1);
	DevBuf<i64> Lp;
	DevBuf<int> Lj;
	DevBuf<i32> Lx;
	i64 nnz = 0;
	panel_to_csr(E.panel, flag.ptr, nullptr, 0, E.Uqinv.ptr, Lp, Lj, Lx, nnz);
	if (nnz == 0)
		return;
	std::vector<i64> hp((size_t) R + 1);
	std::vector<int> hj((size_t) nnz);
	std::vector<i32> hx((size_t) nnz);
	Lp.download(hp.data(), hp.size(), s);
	Lj.download(hj.data(), hj.size(), s);
	Lx.download(hx.data(), hx.size(), s);
	sync();
	for (int r = 0; r < R; r++)
		for (i64 e = hp[r]; e < hp[r + 1]; e++)
			spasm_add_entry(L, row_out[r], hj[e], hx[e]);
}

/*
 * L of the factorization, read out of the final echelon form (lu.cu explains why this is THE L of the reference's
 * definition: the rows of U are independent, so A[i] = sum_k L[i][k] U[k] has one solution).  Row i of L = the
 * elimination coefficients the batched solve of x * U = A[i] leaves on the pivotal columns; what it leaves on the
 * other columns must be zero (U spans the rows of A), which is checked.  Rows: all of them when `complete`, the
 * pivotal ones (fact->p) otherwise -- the reference also keeps the pivotal rows only (echelonize.c:252-270).
 * reference: src/spasm_echelonize.c:606-614 (L->m = rank, spasm_compress(L), fact->complete).
 */
void compute_L(struct spasm_lu *fact, const DevCsr &dA, const struct spasm_csr *A, bool complete, int n_first_round)
{
	cudaStream_t s = ctx().stream;
	const int n = dA.n, r = fact->U->n;
	const struct spasm_csr *U = fact->U;
	struct spasm_triplet *L = spasm_triplet_alloc(n, r, std::max<i64>(spasm_nnz(fact->U) + n, 1), dA.prime, true);
	if (r > 0) {
		/* The rows behind the structural pivots of the FIRST round need no solve: row k of U is row p[k] of A divided by
		 * its pivot, so row p[k] of L is the single entry (k, pivot) (pivots.c:421-426).  On the 500k x 500k configuration
		 * that is 494 k of the 499 k pivotal rows. */
		std::vector<char> settled((size_t) std::max(n, 1), 0);
		n_first_round = std::min(n_first_round, r);
		for (int k = 0; k < n_first_round; k++) {
			const int i = fact->p[k], j = U->j[U->p[k]];
			spasm_ZZp pivot = 0;
			for (i64 px = A->p[i]; px < A->p[i + 1] && pivot == 0; px++)
				if (A->j[px] == j)
					pivot = A->x[px];
			if (pivot == 0)
				errx(1, "[spasm-b200] internal: pivot (%d, %d) not found in the input (L mode)", i, j);
			spasm_add_entry(L, i, k, pivot);
			settled[i] = 1;
		}
		std::vector<int> rows;
		if (complete) {
			for (int i = 0; i < n; i++)
				if (!settled[i])
					rows.push_back(i);
		} else {
			for (int k = n_first_round; k < r; k++)
				rows.push_back(fact->p[k]);
		}
		if (!rows.empty()) {
			Engine E;
			engine_from_host(E, fact->U, fact->qinv);
			const int cap = panel_capacity(E.m);
			for (size_t done = 0; done < rows.size(); done += cap) {
				const int R = (int) std::min<size_t>(cap, rows.size() - done);
				DevBuf<int> d_rows;
				d_rows.upload(rows.data() + done, (size_t) R, s);
				E.solve_rows(dA, d_rows.ptr, R, false);
				if (panel_count_nonzero(E.panel, E.Uqinv.ptr) != 0)
					errx(1, "[spasm-b200] internal: a row of the input is not in the span of the echelon form (L mode)");
				append_L_from_panel(E, L, rows.data() + done, R);
			}
		}
	}
	LOG("[echelonize] L : %" PRId64 " entries\n", L->nz);
	fact->L = spasm_compress(L);
	spasm_triplet_free(L);
	fact->Ltmp = NULL;
	fact->complete = complete;
}

static void store_dense(void *S, spasm_datatype datatype, const std::vector<i32> &host, int rows, int ld, int Sm)
{
	for (int r = 0; r < rows; r++)
		for (int c = 0; c < Sm; c++)
			spasm_datatype_write(S, (size_t) r * Sm + c, datatype, host[(size_t) r * ld + c]);
}

}  // namespace sb

using namespace sb;

extern "C" {

/* reference: src/spasm_pivots.c:369-448 */
int spasm_pivots_extract_structural(const struct spasm_csr *A, const int *p_in, struct spasm_lu *fact, int *p, struct echelonize_opts *opts)
{
	struct echelonize_opts defaults;
	if (opts == NULL) {
		spasm_echelonize_init_opts(&defaults);
		opts = &defaults;
	}
	ctx();
	struct spasm_csr *U = fact->U;
	Engine E;
	E.init(A->m, spasm_get_prime(A));
	if (U->n > 0)
		E.U.upload(U);
	E.Uqinv.upload(fact->qinv, (size_t) A->m, ctx().stream);
	int un0 = U->n;
	i64 unz0 = spasm_nnz(U);
	DevCsr dA;
	dA.upload(A);
	stats().pair_row.clear();
	stats().pair_col.clear();
	int npiv = extract_structural(E, dA, p_in, p, opts->enable_greedy_pivot_search, 0, false);      /* this entry point returns U and p at once: eager levels */
	/* append the new rows to the caller's U */
	i64 extra = E.U.nnz - unz0;
	if (spasm_nnz(U) + extra > U->nzmax)
		spasm_csr_realloc(U, spasm_nnz(U) + extra);
	cudaStream_t s = ctx().stream;
	std::vector<i64> hp((size_t) npiv + 1);
	if (npiv > 0) {
		CUDA_CHECK(cudaMemcpyAsync(hp.data(), E.U.p.ptr + un0, ((size_t) npiv + 1) * sizeof(i64), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(U->j + unz0, E.U.j.ptr + unz0, (size_t) extra * sizeof(int), cudaMemcpyDeviceToHost, s));
		CUDA_CHECK(cudaMemcpyAsync(U->x + unz0, E.U.x.ptr + unz0, (size_t) extra * sizeof(i32), cudaMemcpyDeviceToHost, s));
		sync();
		for (int k = 0; k < npiv; k++)
			U->p[un0 + k + 1] = hp[k + 1];
		U->n = un0 + npiv;
		if (fact->Ltmp != NULL) {
			/* L[row of A, new row of U] = the pivot before it was scaled to 1 (pivots.c:409-426); row p[k] of A became
			 * row un0 + k of U, its pivot column is the first entry of that row */
			for (int k = 0; k < npiv; k++) {
				const int i = p[k], j = U->j[U->p[un0 + k]];
				spasm_ZZp pivot = 0;
				for (i64 px = A->p[i]; px < A->p[i + 1] && pivot == 0; px++)
					if (A->j[px] == j)
						pivot = A->x[px];
				const int i_out = (p_in != NULL) ? p_in[i] : i;
				spasm_add_entry(fact->Ltmp, i_out, un0 + k, pivot);
				if (fact->p != NULL)
					fact->p[un0 + k] = i_out;
			}
		}
	}
	E.Uqinv.download(fact->qinv, (size_t) A->m, s);
	sync();
	return npiv;
}

/* reference: src/spasm_schur.c:11-44 */
double spasm_schur_estimate_density(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U, const int *qinv, int R)
{
	if (n == 0)
		return 0;
	ctx();
	Engine E;
	engine_from_host(E, U, qinv);
	DevCsr dA;
	dA.upload(A);
	return estimate_density(E, dA, p, n, R);
}

/* reference: src/spasm_schur.c:61-193 */
struct spasm_csr *spasm_schur(const struct spasm_csr *A, const int *p, int n, const struct spasm_lu *fact,
                              double est_density, struct spasm_triplet *L, const int *p_in, int *p_out)
{
	ctx();
	Engine E;
	engine_from_host(E, fact->U, fact->qinv);
	DevCsr dA;
	dA.upload(A);
	if (est_density < 0)
		est_density = estimate_density(E, dA, p, n, 100);     /* keeps the reference's rand() consumption */
	DevCsr S;
	std::function<void(int, int)> keep_L;
	if (L != NULL)
		keep_L = [&](int done, int R) {
			std::vector<int> row_out((size_t) R);
			for (int r = 0; r < R; r++)
				row_out[r] = (p_in != NULL) ? p_in[p[done + r]] : p[done + r];
			append_L_from_panel(E, L, row_out.data(), R);
		};
	schur_sparse(E, dA, p, n, S, keep_L);
	if (p_out != NULL)
		for (int k = 0; k < n; k++)
			p_out[k] = (p_in != NULL) ? p_in[p[k]] : p[k];
	LOG("Schur complement: %d * %d [%" PRId64 " nz / density= %.3f]\n", n, A->m, S.nnz, 1.0 * S.nnz / (1.0 * A->m * n));
	return csr_to_host(S);
}

/* reference: src/spasm_schur.c:257-333 */
void spasm_schur_dense(const struct spasm_csr *A, const int *p, int n, const int *p_in,
                       struct spasm_lu *fact, void *S, spasm_datatype datatype, int *q, int *p_out)
{
	ctx();
	Engine E;
	engine_from_host(E, fact->U, fact->qinv);
	E.begin_dense();
	int Sm = E.Sm0;
	for (int c = 0; c < Sm; c++)
		q[c] = E.q0[c];
	LOG("[schur/dense] dimension %d x %d...\n", n, Sm);
	DevCsr dA;
	dA.upload(A);
	cudaStream_t s = ctx().stream;
	const int cap = panel_capacity(E.m);
	int ld = (Sm + 3) & ~3;
	DevBuf<i32> B;
	std::vector<i32> host;
	for (int done = 0; done < n; done += cap) {
		int R = std::min(cap, n - done);
		DevBuf<int> d_rows;
		d_rows.upload(p + done, (size_t) R, s);
		E.solve_rows(dA, d_rows.ptr, R, false);
		if (fact->Ltmp != NULL) {
			std::vector<int> row_out((size_t) R);
			for (int r = 0; r < R; r++)
				row_out[r] = (p_in != NULL) ? p_in[p[done + r]] : p[done + r];
			append_L_from_panel(E, fact->Ltmp, row_out.data(), R);
		}
		B.ensure((size_t) R * std::max(ld, 4));
		E.gather_q0(B.ptr, ld);
		host.resize((size_t) R * std::max(ld, 4));
		B.download(host.data(), (size_t) R * ld, s);
		sync();
		stats().pub.d2h_bytes += (i64) R * ld * 4;
		char *dst = (char *) S + (size_t) done * Sm * spasm_datatype_size(datatype);
		store_dense(dst, datatype, host, R, ld, Sm);
	}
	for (int k = 0; k < n; k++)
		p_out[k] = (p_in != NULL) ? p_in[p[k]] : p[k];
}

/* reference: src/spasm_schur.c:346-413 */
void spasm_schur_dense_randomized(const struct spasm_csr *A, const int *p, int n, const struct spasm_csr *U, const int *qinv,
                                  void *S, spasm_datatype datatype, int *q, int N, int w)
{
	ctx();
	Engine E;
	engine_from_host(E, U, qinv);
	E.begin_dense();
	int Sm = E.Sm0;
	for (int c = 0; c < Sm; c++)
		q[c] = E.q0[c];
	LOG("[schur/dense/random] dimension %d x %d, weight %d...\n", N, Sm, w);
	DevCsr dA;
	dA.upload(A);
	cudaStream_t s = ctx().stream;
	int ww = (w <= 0) ? n : w;
	std::vector<int> rows((size_t) N * ww);
	std::vector<i32> coef((size_t) N * ww);
	i64 prime = spasm_get_prime(A);
	for (int k = 0; k < N; k++) {
		spasm_prng_ctx prng;
		spasm_prng_seed_simple(prime, (u64) k, 0, &prng);
		for (int t = 0; t < ww; t++) {
			if (w <= 0) {
				rows[(size_t) k * ww + t] = p[t];
				coef[(size_t) k * ww + t] = spasm_prng_ZZp(&prng);
			} else {
				rows[(size_t) k * ww + t] = p[rand() % n];
				coef[(size_t) k * ww + t] = (t == 0) ? 1 : spasm_prng_ZZp(&prng);
			}
		}
	}
	DevBuf<int> d_rows;
	DevBuf<i32> d_coef;
	d_rows.upload(rows.data(), rows.size(), s);
	d_coef.upload(coef.data(), coef.size(), s);
	E.solve_combos(dA, d_rows.ptr, d_coef.ptr, N, ww);
	int ld = (Sm + 3) & ~3;
	DevBuf<i32> B((size_t) N * std::max(ld, 4));
	E.gather_q0(B.ptr, ld);
	std::vector<i32> host((size_t) N * std::max(ld, 4));
	B.download(host.data(), (size_t) N * ld, s);
	sync();
	store_dense(S, datatype, host, N, ld, Sm);
}

/* reference: src/spasm_ffpack.cpp:78-86 (boundary) -- computed by dense.cu */
int spasm_ffpack_rref(i64 prime, int n, int m, void *A, int ldA, spasm_datatype datatype, size_t *qinv)
{
	ctx();
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	int ld = (m + 3) & ~3;
	std::vector<i32> host((size_t) std::max(n, 1) * std::max(ld, 4), 0);
	spasm_field field;
	spasm_field_init(prime, field);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < m; j++)
			host[(size_t) i * ld + j] = spasm_ZZp_init(field, (i64) spasm_datatype_read(A, (size_t) i * ldA + j, datatype));
	DevBuf<i32> D;
	D.upload(host.data(), host.size(), s);
	RrefResult res = dense_rref(D.ptr, n, m, ld, F);
	D.download(host.data(), host.size(), s);
	sync();
	std::vector<char> is_pivot((size_t) std::max(m, 1), 0);
	for (int t = 0; t < res.rank; t++) {
		qinv[t] = res.pivcol[t];
		is_pivot[res.pivcol[t]] = 1;
	}
	int k = res.rank;
	for (int j = 0; j < m; j++)
		if (!is_pivot[j])
			qinv[k++] = j;
	/* packed layout of the reference's consumers: row i, position k >= rank  <->  column qinv[k] */
	std::vector<i32> packed((size_t) std::max(n, 1) * std::max(m, 1), 0);
	for (int t = 0; t < res.rank; t++) {
		const i32 *row = host.data() + (size_t) res.pivrow[t] * ld;
		for (int kk = 0; kk < m; kk++)
			packed[(size_t) t * m + kk] = (kk < res.rank) ? (kk == t) : row[qinv[kk]];
	}
	for (int i = 0; i < n; i++)
		for (int kk = 0; kk < m; kk++)
			spasm_datatype_write(A, (size_t) i * ldA + kk, datatype, packed[(size_t) i * m + kk]);
	return res.rank;
}

/*
 * reference: src/spasm_ffpack.cpp:52-75, :88-96 (FFPACK::pPLUQ behind the same signature).  Output layout, as read by
 * update_fact_after_LU (src/spasm_echelonize.c:276-312) and tests/dense_lu_ffpack.c:88-121: position (i, j) of A is row
 * p[i], column qinv[j] of the input; j <= min(i, r-1) holds L (its diagonal is not 1), i < r and j > i holds U (unit
 * diagonal implied).  Computed as in lu.cu: reduced echelon form R (dense.cu) for the pivot columns, C = A[:, pivots],
 * C = Pi^t Lc Uc, U = Uc * R.  The pivot columns are the column rank profile, in increasing order.
 */
int spasm_ffpack_LU(i64 prime, int n, int m, void *A, int ldA, spasm_datatype datatype, size_t *p, size_t *qinv)
{
	ctx();
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	const int ld = std::max((m + 3) & ~3, 4);
	std::vector<i32> host((size_t) std::max(n, 1) * ld, 0);
	spasm_field field;
	spasm_field_init(prime, field);
	for (int i = 0; i < n; i++)
		for (int j = 0; j < m; j++)
			host[(size_t) i * ld + j] = spasm_ZZp_init(field, (i64) spasm_datatype_read(A, (size_t) i * ldA + j, datatype));
	DevBuf<i32> D, before;
	D.upload(host.data(), host.size(), s);
	before.upload(host.data(), host.size(), s);
	RrefResult res = dense_rref(D.ptr, n, m, ld, F);
	const int r = res.rank;
	std::vector<char> is_pivot((size_t) std::max(m, 1), 0);
	for (int t = 0; t < r; t++) {
		qinv[t] = res.pivcol[t];
		is_pivot[res.pivcol[t]] = 1;
	}
	int k = r;
	for (int j = 0; j < m; j++)
		if (!is_pivot[j])
			qinv[k++] = j;
	std::vector<int> lu_row;
	const int ldc = std::max((r + 3) & ~3, 4);
	std::vector<i32> hC((size_t) std::max(n, 1) * ldc, 0), hU((size_t) std::max(r, 1) * ld, 0);
	if (r > 0) {
		DevBuf<int> d_pc, d_pr;
		d_pc.upload(res.pivcol.data(), res.pivcol.size(), s);
		d_pr.upload(res.pivrow.data(), res.pivrow.size(), s);
		DevBuf<i32> C((size_t) n * ldc), Dr((size_t) r * ld), Ulu((size_t) r * ld);
		dense_gather_columns(before.ptr, ld, n, d_pc.ptr, r, C.ptr, ldc);
		dense_lu_fullcol(C.ptr, n, r, ldc, F, lu_row);
		dense_gather_rows(D.ptr, ld, d_pr.ptr, r, m, Dr.ptr, ld);
		dense_lu_rows(C.ptr, ldc, r, lu_row, Dr.ptr, ld, m, Ulu.ptr, ld, F);
		C.download(hC.data(), (size_t) n * ldc, s);
		Ulu.download(hU.data(), (size_t) r * ld, s);
		sync();
	}
	std::vector<char> used((size_t) std::max(n, 1), 0);
	for (int t = 0; t < r; t++) {
		p[t] = lu_row[t];
		used[lu_row[t]] = 1;
	}
	k = r;
	for (int i = 0; i < n; i++)
		if (!used[i])
			p[k++] = i;
	for (int i = 0; i < n; i++) {
		const i32 *Lrow = hC.data() + (size_t) p[i] * ldc;
		for (int j = 0; j < m; j++) {
			i32 v = 0;
			if (j < r && j <= i)
				v = Lrow[j];
			else if (i < r)
				v = hU[(size_t) i * ld + qinv[j]];
			spasm_datatype_write(A, (size_t) i * ldA + j, datatype, v);
		}
	}
	return r;
}

/* reference: src/spasm_rref.c:22-146 */
struct spasm_csr *spasm_rref(const struct spasm_lu *fact, int *Rqinv)
{
	const struct spasm_csr *U = fact->U;
	int n = U->n, m = U->m;
	char hnnz[8];
	spasm_human_format(spasm_nnz(U), hnnz);
	LOG("[rref] start. U is %d x %d (%s nnz)\n", n, m, hnnz);
	ctx();
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sb::sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace] rref/%-22s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	Engine E;
	engine_from_host(E, U, fact->qinv);
	lap("upload + schedule");
	cudaStream_t s = ctx().stream;
	std::vector<int> pivcol((size_t) std::max(n, 1));
	for (int i = 0; i < n; i++)
		pivcol[i] = U->j[U->p[i]];
	const int cap = panel_capacity(m, 16.0);
	std::vector<HostPiece> pieces;
	/* several ranks: each one reduces a contiguous slice of the rows (they are independent given U, rref.c:44-56) */
	int chunk, slice_begin, slice_end;
	comm_slice(n, &chunk, &slice_begin, &slice_end);
	for (int done = slice_begin; done < slice_end; done += cap) {
		int R = std::min(cap, slice_end - done);
		std::vector<int> rows(R);
		for (int r = 0; r < R; r++)
			rows[r] = done + r;
		DevBuf<int> d_rows, d_first;
		d_rows.upload(rows.data(), (size_t) R, s);
		d_first.upload(pivcol.data() + done, (size_t) R, s);
		/* row i is solved against U with its own pivot unregistered (rref.c:56-60): its pivot entry is not
		 * scattered, so nothing propagates from it, and it is emitted first, as 1 */
		E.solve_rows(E.U, d_rows.ptr, R, true, true);      /* rows of the RREF are very sparse: masked solve */
		lap("solve");
		DevBuf<i64> Sp;
		DevBuf<int> Sj;
		DevBuf<i32> Sx;
		i64 nnz;
		panel_to_csr(E.panel, E.Uqinv.ptr, d_first.ptr, 1, nullptr, Sp, Sj, Sx, nnz);
		pieces.emplace_back();
		piece_download(Sp, Sj, Sx, R, nnz, pieces.back());
		lap("panel -> sparse rows");
	}
	exchange_pieces(pieces);
	lap("exchange");
	struct spasm_csr *Rm = pieces_to_host(pieces, n, m, spasm_get_prime(U));
	lap("download");
	for (int j = 0; j < m; j++)
		Rqinv[j] = -1;
	for (int i = 0; i < n; i++)
		Rqinv[pivcol[i]] = i;          /* the pivot of row i of R is the pivot of row i of U, emitted first (valid on the ranks that hold no rows too) */
	spasm_human_format(spasm_nnz(Rm), hnnz);
	LOG("[rref] done. NNZ(R) = %s\n", hnnz);
	return Rm;
}

/* reference: src/spasm_kernel.c:9-127 */
struct spasm_csr *spasm_kernel(const struct spasm_lu *fact)
{
	const struct spasm_csr *U = fact->U;
	const int *qinv = fact->qinv;
	int n = U->n, m = U->m;
	char hnnz[8];
	spasm_human_format(spasm_nnz(U), hnnz);
	LOG("[kernel] start. U is %d x %d (%s nnz)\n", n, m, hnnz);
	ctx();
	cudaStream_t s = ctx().stream;
	Engine E;
	E.init(m, spasm_get_prime(U));
	if (n > 0)
		E.U.upload(U);
	E.Uqinv.upload(qinv, (size_t) m, s);
	DepGraph Gt;
	depgraph_transposed(E.U, E.Uqinv.ptr, Gt);
	depgraph_schedule(Gt);
	std::vector<int> pivcol((size_t) std::max(n, 1));
	for (int i = 0; i < n; i++)
		pivcol[i] = U->j[U->p[i]];
	DevBuf<int> d_pivcol;
	d_pivcol.upload(pivcol.data(), pivcol.size(), s);
	std::vector<int> freecols;
	for (int j = 0; j < m; j++)
		if (qinv[j] < 0)
			freecols.push_back(j);
	const int cap = panel_capacity(std::max(n, 1), 16.0);
	std::vector<HostPiece> pieces;
	std::vector<int> colslot((size_t) std::max(m, 1));
	Panel P;
	/* several ranks: each one takes a contiguous slice of the non-pivotal columns (kernel.c:42-51) */
	int chunk, slice_begin = 0, slice_end = (int) freecols.size();
	/* a kernel that fits ONE batch is a single depth-bound pass: slicing it would only add the exchange (measured at 8
	 * ranks on config 4, 998 vectors: 0.80 s sharded against 0.27 s replicated) */
	const bool shard_kernel = comm_world() > 1 && (int) freecols.size() > cap;
	if (shard_kernel)
		comm_slice((int) freecols.size(), &chunk, &slice_begin, &slice_end);
	for (size_t done = (size_t) slice_begin; done < (size_t) slice_end; done += cap) {
		int R = (int) std::min<size_t>(cap, (size_t) slice_end - done);
		std::fill(colslot.begin(), colslot.end(), -1);
		for (int r = 0; r < R; r++)
			colslot[freecols[done + r]] = r;
		DevBuf<int> d_slot, d_first;
		d_slot.upload(colslot.data(), colslot.size(), s);
		d_first.upload(freecols.data() + done, (size_t) R, s);
		GpuTimer t;
		t.start();
		P.shape(n, R);
		panel_scatter_columns(E.U, d_slot.ptr, P);
		panel_solve(Gt, P.X, P.ld, R, E.F);
		stats().pub.ms_solve += t.stop_ms();
		DevBuf<i64> Sp;
		DevBuf<int> Sj;
		DevBuf<i32> Sx;
		i64 nnz;
		panel_to_csr(P, nullptr, d_first.ptr, -1, d_pivcol.ptr, Sp, Sj, Sx, nnz);
		pieces.emplace_back();
		piece_download(Sp, Sj, Sx, R, nnz, pieces.back());
	}
	if (shard_kernel)
		exchange_pieces(pieces);
	struct spasm_csr *K = pieces_to_host(pieces, m - n, m, spasm_get_prime(U));
	spasm_human_format(spasm_nnz(K), hnnz);
	LOG("[kernel] done. NNZ(K) = %s\n", hnnz);
	return K;
}

}  /* extern "C" */
