/*
 * spasm_echelonize on the GPU: host orchestration (same decision logic as the
 * reference, because the branch decisions are part of the result) driving the
 * CUDA kernel families of pivots.cu / solve.cu / panel.cu / dense.cu.
 * reference: src/spasm_echelonize.c
 *
 * What is restructured:
 *  - all data stays in HBM between the steps: A, the structural rows of U, the
 *    solve schedule, the dense rows; the host only sees counters;
 *  - a dense block is reduced by the structural pivots with ONE batched sparse
 *    solve and by the rows of the earlier dense blocks with dense products
 *    (the reference scatters those dense rows one entry at a time,
 *    spasm_schur.c:291-304 / :389-394 -- its dominant cost on config 1);
 *  - U is assembled on the host once, at the end.
 * What is kept bit for bit: option handling, round structure, thresholds,
 * glibc rand() call sequence, PRNG coefficients, block sizes, weight doubling,
 * the "n -= rr" quirk of the low-rank loop (echelonize.c:369).
 */
#include <math.h>
#include <functional>
#include <cub/cub.cuh>
#include "engine.cuh"
#include "lu.cuh"
#include "stats.cuh"

namespace sb {

#define LOG(...) do { if (ctx().verbose) { fprintf(stderr, __VA_ARGS__); fflush(stderr); } } while (0)

/* ------------------------------------------------------------------ small device helpers */

struct IsPivotal {
	__device__ bool operator()(const int &v) const { return v >= 0; }
};

/* rows_out = qinv[j] for the columns j (in the order of `cols`, or 0..m-1 when cols == NULL) with qinv[j] >= base */
__global__ void k_select_rows(int count, const int *__restrict__ cols, const int *__restrict__ qinv, int *flags, int *rows)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count)
		return;
	int j = cols ? cols[t] : t;
	int i = qinv[j];
	flags[t] = i >= 0;
	rows[t] = i;
}

/* same for the rows below `limit` only (the rows of the first, lazily ordered round) */
__global__ void k_select_rows_below(int count, const int *__restrict__ cols, const int *__restrict__ qinv, int limit, int *flags, int *rows)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count)
		return;
	int i = qinv[cols[t]];
	flags[t] = i >= 0 && i < limit;
	rows[t] = i;
}

/* perm[t] for t >= L is the identity; len[t] = length of row perm[t] */
__global__ void k_perm_lengths(int n, int L, int *perm, const i64 *__restrict__ Up, i64 *len)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t > n)
		return;
	if (t == n) {
		len[t] = 0;
		return;
	}
	if (t >= L)
		perm[t] = t;
	const int r = t >= L ? t : perm[t];
	len[t] = Up[r + 1] - Up[r];
}

/* row t of the output = row perm[t] of U; qinv of its pivot (first entry) = t */
__global__ void k_perm_rows(int n, const int *__restrict__ perm, const i64 *__restrict__ Up, const int *__restrict__ Uj, const i32 *__restrict__ Ux,
                            const i64 *__restrict__ newp, int *newj, i32 *newx, int *qinv)
{
	const int lane = threadIdx.x & 31, nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n; t += nwarps) {
		const int r = perm[t];
		const i64 src = Up[r], dst = newp[t], cnt = Up[r + 1] - src;
		for (i64 e = lane; e < cnt; e += 32) {
			newj[dst + e] = Uj[src + e];
			newx[dst + e] = Ux[src + e];
		}
		if (lane == 0 && cnt > 0)
			qinv[Uj[src]] = t;
	}
}

static int compact_flagged(const int *d_in, const int *d_flags, int count, int *d_out)
{
	static DevBuf<char> tmp;
	DevBuf<int> nsel(1);
	size_t bytes = 0;
	cudaStream_t s = ctx().stream;
	cub::DeviceSelect::Flagged(nullptr, bytes, d_in, d_flags, d_out, nsel.ptr, count, s);
	tmp.ensure(bytes + 16);
	cub::DeviceSelect::Flagged(tmp.ptr, bytes, d_in, d_flags, d_out, nsel.ptr, count, s);
	LAUNCHED(1);
	return fetch(nsel.ptr);
}

__global__ void k_add_offset(i64 *p, int count, i64 off)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < count)
		p[t] += off;
}

/* ------------------------------------------------------------------ one round of structural pivots */

/*
 * reference: spasm_pivots_extract_structural (src/spasm_pivots.c:369-448).
 * p (host, size A.n) receives the row permutation: pivotal rows first (in the order they get in U),
 * then the other rows by increasing index.
 */
int extract_structural(Engine &E, const DevCsr &A, const int *p_in, int *p, bool greedy, int round, bool allow_lazy)
{
	GpuTimer timer;
	timer.start();
	cudaStream_t s = ctx().stream;
	const int n = A.n, m = A.m;
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace] %-28s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	DevBuf<int> d_pinv((size_t) std::max(n, 1)), d_qinv((size_t) std::max(m, 1));
	PivotCounts cnt = pivots_find(A, d_pinv.ptr, d_qinv.ptr, greedy);
	lap("pivots_find");
	int npiv = cnt.fl + cnt.flcol + cnt.greedy;
	LOG("[pivots] Faugère-Lachartre: %d pivots found\n[pivots] ``Faugère-Lachartre on columns'': %d pivots found\n"
	    "[pivots] greedy alternating cycle-free search: %d pivots found\n[pivots] %d pivots found\n",
	    cnt.fl, cnt.flcol, cnt.greedy, npiv);
	Stats &st = stats();
	if (round < 64) {
		st.pub.found_FL[round] = cnt.fl;
		st.pub.found_FLcol[round] = cnt.flcol;
		st.pub.found_greedy[round] = cnt.greedy;
	}
	std::vector<int> h_pinv((size_t) std::max(n, 1));
	d_pinv.download(h_pinv.data(), (size_t) n, s);

	std::vector<int> h_rows;       /* pivotal rows in their final order */
	if (npiv > 0) {
		/* 1. the new pivotal rows alone, in column order, give the dependency DAG among new pivots */
		DevBuf<int> flags((size_t) m), rows_all((size_t) m), rows_tmp((size_t) npiv);
		k_select_rows<<<cdiv(m, 256), 256, 0, s>>>(m, nullptr, d_qinv.ptr, flags.ptr, rows_all.ptr);
		LAUNCHED(1);
		int got = compact_flagged(rows_all.ptr, flags.ptr, m, rows_tmp.ptr);
		if (got != npiv)
			errx(1, "[spasm-b200] internal: pivot count mismatch (%d vs %d)", got, npiv);
		lap("select rows");
		if (allow_lazy && E.U.n == 0) {
			/* First round of an echelonization: the rows go to U in column order and only what the dataflow solve
			 * needs is scheduled.  The levels of the pivot DAG (a full traversal: 14 ms on config 2) come out of
			 * the first solve pass, and assemble() puts the rows in level order at the end. */
			append_pivotal_rows(A, rows_tmp.ptr, npiv, d_pinv.ptr, E.U, E.Uqinv);
			lap("U (column order)");
			E.G = DepGraph();
			depgraph_forward(E.U, E.G);
			lap("depgraph_forward");
			depgraph_schedule(E.G, true);
			lap("depgraph_schedule (lazy)");
			E.G_ready = true;
			E.lazy_rows = E.G.levels_known ? 0 : npiv;
			st.pub.dag_depth = E.G.nlevels;
			h_rows.resize(npiv);
			rows_tmp.download(h_rows.data(), (size_t) npiv, s);
			sync();
		} else {
		DevCsr Unew;
		Unew.m = m;
		Unew.prime = A.prime;
		DevBuf<int> scratch_qinv((size_t) m);
		append_pivotal_rows(A, rows_tmp.ptr, npiv, d_pinv.ptr, Unew, scratch_qinv);
		lap("U (column order)");
		DepGraph Gnew;
		depgraph_forward(Unew, Gnew);
		lap("depgraph_forward");
		depgraph_schedule(Gnew);
		lap("depgraph_schedule");
		/* 2. pivot columns by (level, column) -> final order of the new rows */
		DevBuf<int> flags2((size_t) m), rows_ord((size_t) m), rows_sorted((size_t) npiv);
		k_select_rows<<<cdiv(m, 256), 256, 0, s>>>(m, Gnew.order.ptr, d_qinv.ptr, flags2.ptr, rows_ord.ptr);
		LAUNCHED(1);
		got = compact_flagged(rows_ord.ptr, flags2.ptr, m, rows_sorted.ptr);
		if (got != npiv)
			errx(1, "[spasm-b200] internal: pivot order mismatch");
		bool first_rows = (E.U.n == 0);
		append_pivotal_rows(A, rows_sorted.ptr, npiv, d_pinv.ptr, E.U, E.Uqinv);
		lap("U (level order)");
		h_rows.resize(npiv);
		rows_sorted.download(h_rows.data(), (size_t) npiv, s);
		sync();
		if (first_rows) {
			E.G = std::move(Gnew);           /* same columns, same dependencies: the schedule carries over */
			E.G_ready = true;
			st.pub.dag_depth = E.G.nlevels;
		} else {
			E.G_ready = false;
		}
		}
	}
	sync();
	int k = 0;
	for (int t = 0; t < npiv; t++) {
		int i = h_rows[t];
		p[k++] = i;
		st.pair_row.push_back(p_in ? p_in[i] : i);
		st.pair_col.push_back(h_pinv[i]);
		E.p_struct.push_back(p_in ? p_in[i] : i);      /* fact->p of the L mode: row t of U comes from this row of the input */
	}
	for (int i = 0; i < n; i++)
		if (h_pinv[i] < 0)
			p[k++] = i;
	lap("host permutation");
	st.pub.ms_pivots += timer.stop_ms();
	return npiv;
}

/* ------------------------------------------------------------------ Schur complements */

/* reference: src/spasm_schur.c:11-44.  rand() is called exactly R times, like the reference. */
double estimate_density(Engine &E, const DevCsr &A, const int *p, int n, int R)
{
	if (n == 0)
		return 0;
	std::vector<int> rows(R);
	for (int t = 0; t < R; t++)
		rows[t] = p[rand() % n];
	DevBuf<int> d_rows;
	d_rows.upload(rows.data(), rows.size(), ctx().stream);
	E.solve_rows(A, d_rows.ptr, R, false);
	i64 nnz = panel_count_nonzero(E.panel, E.Uqinv.ptr);
	return ((double) nnz) / (E.m - E.U.n) / R;
}

/* reference: src/spasm_schur.c:61-193.  Rows come out in p order (the reference with one thread) and the entries of
 * a row in the order of the reference's reach pattern (panel.cu: k_schur_emit_dfs). */
void schur_sparse(Engine &E, const DevCsr &A, const int *p, int n, DevCsr &S, const std::function<void(int, int)> &after_batch)
{
	cudaStream_t s = ctx().stream;
	const int cap = panel_capacity(E.m);
	struct Piece { DevBuf<i64> p; DevBuf<int> j; DevBuf<i32> x; i64 nnz; int rows; };
	std::vector<Piece> pieces;
	i64 total = 0;
	for (int done = 0; done < n; done += cap) {
		int R = std::min(cap, n - done);
		DevBuf<int> d_rows;
		d_rows.upload(p + done, (size_t) R, s);
		E.solve_rows(A, d_rows.ptr, R, false);
		if (after_batch)
			after_batch(done, R);          /* the elimination coefficients are on the pivotal columns of the panel (L, schur.c:164-170) */
		Piece pc;
		pc.rows = R;
		/* count, then emit every row in the order of the reference's DFS pattern (it feeds the next pivot round) */
		panel_to_csr(E.panel, E.Uqinv.ptr, nullptr, 0, nullptr, pc.p, pc.j, pc.x, pc.nnz, true);
		if (pc.nnz > 0)
			panel_emit_reference_order(E.panel, A, d_rows.ptr, E.U, E.Uqinv.ptr, E.G.nlevels, pc.p.ptr, pc.j.ptr, pc.x.ptr);
		total += pc.nnz;
		pieces.push_back(std::move(pc));
	}
	S.n = n;
	S.m = E.m;
	S.prime = E.prime;
	S.nnz = total;
	S.p.alloc((size_t) n + 1);
	S.j.alloc((size_t) std::max<i64>(total, 1));
	S.x.alloc((size_t) std::max<i64>(total, 1));
	CUDA_CHECK(cudaMemsetAsync(S.p.ptr, 0, sizeof(i64), s));
	i64 off = 0;
	int row = 0;
	for (Piece &pc : pieces) {
		if (off)
			k_add_offset<<<cdiv(pc.rows + 1, 256), 256, 0, s>>>(pc.p.ptr, pc.rows + 1, off);
		CUDA_CHECK(cudaMemcpyAsync(S.p.ptr + row, pc.p.ptr, ((size_t) pc.rows + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
		if (pc.nnz) {
			CUDA_CHECK(cudaMemcpyAsync(S.j.ptr + off, pc.j.ptr, (size_t) pc.nnz * sizeof(int), cudaMemcpyDeviceToDevice, s));
			CUDA_CHECK(cudaMemcpyAsync(S.x.ptr + off, pc.x.ptr, (size_t) pc.nnz * sizeof(i32), cudaMemcpyDeviceToDevice, s));
		}
		off += pc.nnz;
		row += pc.rows;
	}
	sync();
}

void prng_combo_coefficients(i64 prime, int N, int w, i32 *d_coef);     /* prng.cu */
void compute_L(struct spasm_lu *fact, const DevCsr &dA, const struct spasm_csr *A, bool complete, int n_first_round);    /* api.cu */

/* ------------------------------------------------------------------ finishing strategies */

static void record_block(int Sn, int Sm, int rr, int w)
{
	Stats &st = stats();
	int k = st.pub.nblocks;
	if (k < 4096) {
		st.pub.block_Sn[k] = Sn;
		st.pub.block_Sm[k] = Sm;
		st.pub.block_rr[k] = rr;
		st.pub.block_w[k] = w;
	}
	st.pub.nblocks = k + 1;
}

/* N random combinations of the rows p[0:n] of A, reduced by the structural rows, as a dense N x Sm0 block.
 * Host side: the row choices (glibc rand(), k-major) and the coefficients (SHA-256 PRNG seeded with
 * (prime, k, 0)) of the reference, spasm_schur.c:367-386. */
static void randomized_block(Engine &E, const DevCsr &A, const int *p, int n, int N, int w, DevBuf<i32> &B, int &ldB)
{
	cudaStream_t s = ctx().stream;
	int ww = (w <= 0) ? n : w;
	std::vector<int> rows((size_t) N * ww);
	DevBuf<int> d_rows;
	DevBuf<i32> d_coef((size_t) N * ww);
	if (w <= 0) {
		/* full combinations (completion test only): one coefficient per row, drawn on the host */
		std::vector<i32> coef((size_t) N * ww);
		for (int k = 0; k < N; k++) {
			spasm_prng_ctx prng;
			spasm_prng_seed_simple(E.prime, (u64) k, 0, &prng);
			for (int i = 0; i < n; i++) {
				rows[(size_t) k * ww + i] = p[i];
				coef[(size_t) k * ww + i] = spasm_prng_ZZp(&prng);
			}
		}
		CUDA_CHECK(cudaMemcpyAsync(d_coef.ptr, coef.data(), coef.size() * sizeof(i32), cudaMemcpyHostToDevice, s));
		sync();
	} else {
		/* the row choices are the reference's glibc rand() sequence (host, k-major); the coefficients come
		 * from the SHA-256 streams, generated on the device (prng.cu) */
		for (int k = 0; k < N; k++)
			for (int t = 0; t < w; t++)
				rows[(size_t) k * ww + t] = p[rand() % n];
		prng_combo_coefficients(E.prime, N, w, d_coef.ptr);
	}
	d_rows.upload(rows.data(), rows.size(), s);
	stats().pub.h2d_bytes += (i64) rows.size() * 4;
	E.block_from_combos(A, d_rows.ptr, d_coef.ptr, N, ww, B, ldB);
}

/* reference: src/spasm_echelonize.c:30-51 */
static bool test_completion(Engine &E, const DevCsr &A, const int *p, int n)
{
	if (n == 0 || A.nnz == 0)
		return true;
	int Sm = E.m - E.rank();
	int Sn = (int) ceil(128 / log2((double) E.prime));
	LOG("[echelonize/completion] Testing completion with %d random linear combinations (rank %d)\n", Sn, E.rank());
	DevBuf<i32> B;
	int ldB;
	randomized_block(E, A, p, n, Sn, 0, B, ldB);
	/* rank of the block after reduction by everything found so far: absorb on a scratch engine state */
	size_t before = E.blocks.size();
	int rank_before = E.dense_rank;
	int rr = E.absorb_block(B.ptr, Sn, ldB);
	/* the reference discards these rows (it only looks at rr) */
	while (E.blocks.size() > before)
		E.blocks.pop_back();
	E.dense_rank = rank_before;
	record_block(Sn, Sm, rr, 0);
	return rr == 0;
}

/* ------------------------------------------------------------------ solve-ahead
 *
 * The elimination of a block by the STRUCTURAL pivots does not depend on the dense rows found by the earlier blocks
 * (those are applied afterwards, as dense products).  So the structural solves of several consecutive blocks can
 * share one pass over the pivot DAG, whose depth -- not its size -- is what a pass costs.
 *   - finish_dense: the rows of all blocks are known in advance: they are solved in large batches.
 *   - finish_lowrank: the rows of the next blocks depend on rand(), on the block sizes and on the weight, which the
 *     reference only knows after each block.  We PREDICT them (every block finds as many pivots as it has rows, the
 *     weight does not change: the normal course), peek at the corresponding glibc rand() draws, and solve the
 *     predicted blocks together.  rand() is then rewound; each block consumes its draws for real when it is
 *     processed, exactly like the reference, and if the prediction fails at some block the remaining predicted
 *     blocks are discarded and recomputed.  The result is identical to the block-by-block course.
 */
struct RandSnapshot {
	char *where = nullptr;
	char saved[256];
	size_t bytes = 0;
};

/* glibc keeps the state of rand()/random() in a table whose first word encodes the generator type and the position
 * of its pointers whenever another table is installed (initstate/setstate).  Installing a temporary table therefore
 * makes the live state a plain, copyable array. */
static bool rand_snapshot(RandSnapshot &snap)
{
	alignas(8) static char scratch[256];
	char *old = initstate(1, scratch, sizeof(scratch));
	if (old == NULL)
		return false;
	static const size_t size_of_type[5] = {8, 32, 64, 128, 256};
	int type = (int) (((const int32_t *) old)[0] % 5);
	if (type < 0 || type > 4) {
		setstate(old);
		return false;
	}
	snap.where = old;
	snap.bytes = size_of_type[type];
	memcpy(snap.saved, old, snap.bytes);
	setstate(old);
	return true;
}

static void rand_restore(const RandSnapshot &snap)
{
	alignas(8) static char scratch[256];
	initstate(1, scratch, sizeof(scratch));       /* park the generator elsewhere while its table is rewritten */
	memcpy(snap.where, snap.saved, snap.bytes);
	setstate(snap.where);
}

/* does snapshot/restore really rewind rand() on this libc?  (checked once; without it blocks are solved one by one) */
static bool rand_rewind_works()
{
	static int known = -1;
	if (known >= 0)
		return known;
	RandSnapshot s0;
	known = 0;
	if (!rand_snapshot(s0))
		return false;
	int a1 = rand(), a2 = rand(), a3 = rand();
	rand_restore(s0);
	int b1 = rand(), b2 = rand(), b3 = rand();
	rand_restore(s0);
	known = (a1 == b1 && a2 == b2 && a3 == b3);
	return known;
}

/*
 * Fast access to the reference's random row choices.  The low-rank finisher draws Sn x w values of glibc's rand() per
 * block (635 k for config 2) and, with solve-ahead, every value is drawn twice (peeked, then consumed): at ~15 ns per
 * locked call that was 15-20 ms of a 135 ms step.  glibc's default generator (TYPE_3: r[i] = r[i-31] + r[i-3], output
 * r[i] >> 1) lives in the table that initstate()/setstate() hand over, so the draws are replayed here on that table
 * directly and the table is handed back with its position updated.  Checked once against rand() itself; if the libc
 * behaves differently the code falls back to calling rand().
 */
struct FastRand {
	int32_t *table = nullptr;      /* table[0] = position * 5 + type, table[1..31] = the 31 state words */
	int rear = 0, front = 0;
	bool active = false;
	char saved[256];

	static bool usable()
	{
		static int known = -1;
		if (known >= 0)
			return known;
		known = 0;
		if (getenv("SPASM_B200_NO_FAST_RAND") != NULL || !rand_rewind_works())
			return false;
		RandSnapshot s0;
		if (!rand_snapshot(s0))
			return false;
		int want[64];
		for (int t = 0; t < 64; t++)
			want[t] = rand();
		rand_restore(s0);
		FastRand g;
		bool same = g.begin();
		for (int t = 0; same && t < 64; t++)
			same = (g.next() == want[t]);
		if (g.active)
			g.end(false);
		/* and the hand-over: after a committed run rand() must continue the sequence */
		if (same) {
			FastRand h;
			h.begin();
			for (int t = 0; t < 10; t++)
				(void) h.next();
			h.end(true);
			for (int t = 10; same && t < 20; t++)
				same = (rand() == want[t]);
			rand_restore(s0);
		}
		known = same;
		return known;
	}

	bool begin()
	{
		alignas(8) static char scratch[256];
		char *old = initstate(1, scratch, sizeof(scratch));       /* park glibc's generator: its table is ours now */
		if (old == NULL)
			return false;
		table = (int32_t *) old;
		int type = table[0] % 5;
		if (type != 3) {
			setstate(old);
			return false;
		}
		rear = table[0] / 5;
		front = (rear + 3) % 31;
		memcpy(saved, old, 128);
		active = true;
		return true;
	}

	inline int next()
	{
		uint32_t *st = (uint32_t *) (table + 1);
		uint32_t val = (st[front] += st[rear]);
		if (++front >= 31) {
			front = 0;
			++rear;
		} else if (++rear >= 31) {
			rear = 0;
		}
		return (int) (val >> 1);
	}

	void end(bool commit)
	{
		if (!active)
			return;
		if (commit)
			table[0] = rear * 5 + 3;
		else
			memcpy(table, saved, 128);
		setstate((char *) table);
		active = false;
	}
};

/* `count` draws of rand(), thrown away (the blocks solved ahead consume their draws when they are processed) */
static void rand_skip(i64 count)
{
	FastRand g;
	if (FastRand::usable() && g.begin()) {
		for (i64 t = 0; t < count; t++)
			(void) g.next();
		g.end(true);
	} else {
		for (i64 t = 0; t < count; t++)
			(void) rand();
	}
}

struct AheadBlock { int Sn, n, w; };          /* predicted shape of a low-rank block */

/* Row choices and coefficients of several predicted low-rank blocks, as the reference will draw them; rand() is left
 * untouched (peek: state snapshot, then rewound).  Block b starts at row offset[b] of the stacked list. */
static size_t peek_combinations(Engine &E, const int *p, const std::vector<AheadBlock> &plan, DevBuf<int> &d_rows, DevBuf<i32> &d_coef,
                                std::vector<int> &offset)
{
	cudaStream_t s = ctx().stream;
	int w = plan[0].w;
	size_t total = 0;
	offset.clear();
	for (const AheadBlock &b : plan) {
		offset.push_back((int) total);
		total += b.Sn;
	}
	std::vector<int> rows(total * w);
	d_coef.alloc(total * w);
	size_t at = 0;
	for (const AheadBlock &b : plan) {
		prng_combo_coefficients(E.prime, b.Sn, w, d_coef.ptr + at * w);     /* streams are seeded per row OF ITS BLOCK */
		at += b.Sn;
	}
	FastRand g;
	if (FastRand::usable() && g.begin()) {
		at = 0;
		for (const AheadBlock &b : plan) {
			/* x % n without a division (Lemire): n is the same for the whole block */
			const uint64_t M = UINT64_MAX / (uint32_t) b.n + 1;
			for (size_t e = at * w; e < (at + b.Sn) * w; e++) {
				const uint64_t low = M * (uint32_t) g.next();
				rows[e] = p[(uint32_t) (((unsigned __int128) low * (uint32_t) b.n) >> 64)];
			}
			at += b.Sn;
		}
		g.end(false);                    /* peek: the table goes back as it was */
	} else {
		RandSnapshot snap;
		rand_snapshot(snap);
		at = 0;
		for (const AheadBlock &b : plan) {
			for (size_t e = at * w; e < (at + b.Sn) * w; e++)
				rows[e] = p[rand() % b.n];
			at += b.Sn;
		}
		rand_restore(snap);
	}
	d_rows.upload(rows.data(), rows.size(), s);
	sync();                              /* `rows` is a local */
	stats().pub.h2d_bytes += (i64) rows.size() * 4;
	return total;
}

/* The dense blocks (reduced by the structural pivots) of several predicted low-rank blocks, stacked in `out`. */
static void randomized_blocks_ahead(Engine &E, const DevCsr &A, const int *p, const std::vector<AheadBlock> &plan,
                                    DevBuf<i32> &out, int &ldB, std::vector<int> &offset)
{
	DevBuf<int> d_rows;
	DevBuf<i32> d_coef;
	size_t total = peek_combinations(E, p, plan, d_rows, d_coef, offset);
	E.block_from_combos(A, d_rows.ptr, d_coef.ptr, (int) total, plan[0].w, out, ldB);
}

/* blocks the low-rank finisher will most likely process next (every block has full rank and keeps the weight) */
static std::vector<AheadBlock> plan_lowrank(int n, int rank_ub, int w, const struct echelonize_opts *opts)
{
	std::vector<AheadBlock> plan;
	int n2 = n, ub2 = rank_ub;
	size_t rows_ahead = 0;
	const size_t row_budget = (size_t) 8 * opts->dense_block_size;
	while (ub2 > 0 && rows_ahead < row_budget && (size_t) (plan.size() + 1) * w * opts->dense_block_size < ((size_t) 1 << 28)) {
		int sn2 = spasm_min(ub2, opts->dense_block_size);
		plan.push_back({sn2, n2, w});
		rows_ahead += sn2;
		n2 -= sn2;
		ub2 -= sn2;
	}
	return plan;
}

/* rows of the Schur complement solved in one batch by the dense finisher (bounded by the memory of the stacked blocks) */
static int dense_rows_ahead(int Sm0, int remaining, int Sn, const struct echelonize_opts *opts, bool no_ahead)
{
	size_t per_row = (size_t) std::max((Sm0 + 3) & ~3, 4) * sizeof(i32);
	size_t max_rows = std::max<size_t>((size_t) Sn, ((size_t) 12 << 30) / per_row);
	int take = no_ahead ? Sn : (int) std::min<size_t>((size_t) remaining, max_rows);
	return std::max(Sn, take - take % opts->dense_block_size);
}

/*
 * Speculation across the density estimate.  A pass over the pivot DAG costs its depth, not its width, and the density
 * estimate (100 rows) is a full pass of its own.  When the finisher that will run if the Schur complement turns out
 * dense is known beforehand (it only depends on the aspect ratio and on the options), its first batch of right-hand
 * sides is solved IN THE SAME PASS as the 100 sampled rows.  If the estimate says "sparse" the extra columns are
 * thrown away.  The rand() draws happen in the reference's order: 100 for the estimate, then the blocks' draws are
 * peeked (and consumed for real when the blocks are processed).
 */
struct Speculation {
	bool valid = false;
	int kind = 0;                      /* 1 = low-rank blocks, 2 = dense blocks */
	int n = 0, w = 0;                  /* what the finisher must start with for the batch to be usable */
	std::vector<AheadBlock> plan;      /* kind 1 */
	std::vector<int> offset;
	int take = 0;                      /* kind 2: rows p[0:take] */
	DevBuf<i32> B;
	int ldB = 0;
	bool froze_columns = false;        /* begin_dense() was called for this batch */
	void discard(Engine &E)
	{
		if (froze_columns)
			E.dense_ready = false;
		valid = false;
		froze_columns = false;
		B.release();
	}
};

/* reference: src/spasm_schur.c:11-44 (density) + the first batch of the predicted finisher */
static double estimate_density_speculative(Engine &E, const DevCsr &A, const int *p, int n, int R, const struct echelonize_opts *opts, Speculation &spec)
{
	static const bool off = getenv("SPASM_B200_NO_SPECULATION") != NULL || getenv("SPASM_B200_NO_SOLVE_AHEAD") != NULL;
	spec.valid = false;
	const int Sm = E.m - E.U.n;
	/* with several ranks the merged pass is replicated (it is bound by the depth of the DAG, not by its width: slicing it
	 * would only add a collective), every rank ends up with the same block */
	if (off || n == 0 || Sm <= 0 || E.dense_ready)
		return estimate_density(E, A, p, n, R);
	cudaStream_t s = ctx().stream;
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace]   density+batch/%-14s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	/* 1. the estimate's own draws */
	std::vector<int> rows(R);
	for (int t = 0; t < R; t++)
		rows[t] = p[rand() % n];
	DevBuf<int> d_rows;
	d_rows.upload(rows.data(), rows.size(), s);
	/* 2. which finisher, with which first batch */
	const double aspect_ratio = (double) n / Sm;
	DevBuf<int> d_rows2;
	DevBuf<i32> d_coef;
	int N = 0;
	if (opts->enable_tall_and_skinny && aspect_ratio > opts->tall_and_skinny_ratio) {
		const int rank_ub = spasm_min(n, Sm);
		const int w = (opts->low_rank_start_weight < 0) ? (int) ceil(-log(0.01) * n / rank_ub) : (int) opts->low_rank_start_weight;
		if (w > 0 && rand_rewind_works()) {
			spec.plan = plan_lowrank(n, rank_ub, w, opts);
			if (spec.plan.size() >= 1) {
				spec.kind = 1;
				spec.w = w;
				N = (int) peek_combinations(E, p, spec.plan, d_rows2, d_coef, spec.offset);
			}
		}
	} else if (opts->enable_dense || opts->enable_GPLU) {
		spec.kind = 2;
		spec.take = N = dense_rows_ahead(Sm, n, spasm_min(opts->dense_block_size, n), opts, false);
		d_rows2.upload(p, (size_t) N, s);
	}
	if (N <= 0 || R + N > panel_capacity(E.m)) {
		spec.kind = 0;
		sync();
		E.solve_rows(A, d_rows.ptr, R, false);
		i64 nnz0 = panel_count_nonzero(E.panel, E.Uqinv.ptr);
		return ((double) nnz0) / Sm / R;
	}
	lap("draws");
	/* 3. one pass for both */
	if (!E.G_ready)
		E.rebuild_schedule();
	GpuTimer t;
	t.start();
	const int R_off = (R + 3) & ~3;
	E.panel.shape(E.m, R_off + N);
	panel_scatter_rows(A, d_rows.ptr, R, E.panel, E.F, false);
	if (spec.kind == 1)
		panel_scatter_combos(A, d_rows2.ptr, d_coef.ptr, N, spec.w, E.panel, E.F, R_off);
	else
		panel_scatter_rows(A, d_rows2.ptr, N, E.panel, E.F, false, R_off);
	lap("scatter");
	panel_solve(E.G, E.panel.X, E.panel.ld, R_off + N, E.F);
	stats().pub.ms_solve += t.stop_ms();
	lap("solve");
	E.account_bytes(R_off + N);
	lap("byte accounting");
	i64 nnz = panel_count_nonzero(E.panel, E.Uqinv.ptr, R);
	lap("count");
	/* 4. the finisher's batch as a dense block over the columns that are not pivotal now */
	E.begin_dense();
	spec.froze_columns = true;
	spec.ldB = std::max((E.Sm0 + 3) & ~3, 4);
	spec.B.ensure((size_t) N * spec.ldB);
	panel_gather_dense(E.panel, E.d_q0.ptr, E.Sm0, spec.B.ptr, spec.ldB, R_off, N);
	spec.n = n;
	spec.valid = true;
	sync();
	lap("gather");
	return ((double) nnz) / Sm / R;
}

/* reference: src/spasm_echelonize.c:315-379 */
static void finish_lowrank(Engine &E, const DevCsr &A, const int *p, int n, const struct echelonize_opts *opts, Speculation *spec = nullptr)
{
	E.begin_dense();
	int Sm = E.m - E.rank();
	LOG("[echelonize/dense/low-rank] processing dense schur complement of dimension %d x %d; block size=%d\n", n, Sm, opts->dense_block_size);
	int rank_ub = spasm_min(n, Sm);
	int w = (opts->low_rank_start_weight < 0) ? (int) ceil(-log(0.01) * n / rank_ub) : (int) opts->low_rank_start_weight;
	DevBuf<i32> B, ahead;
	std::vector<AheadBlock> plan;     /* predicted blocks already solved, plan[next_ahead] is the next one */
	std::vector<int> ahead_offset;
	size_t next_ahead = 0;
	int ahead_ld = 0;
	static const bool no_ahead = getenv("SPASM_B200_NO_SOLVE_AHEAD") != NULL;
	if (spec && spec->valid && spec->kind == 1 && spec->n == n && spec->w == w) {
		/* the first blocks were solved together with the density estimate */
		plan = spec->plan;
		ahead_offset = spec->offset;
		ahead = std::move(spec->B);
		ahead_ld = spec->ldB;
		spec->valid = false;
	}
	int round = 0;
	for (;;) {
		int Sn = spasm_min(rank_ub, opts->dense_block_size);
		if (Sn <= 0)
			break;
		LOG("[echelonize/dense/low-rank] Round %d. Weight %d. Processing chunk (%d x %d)\n", round, w, Sn, Sm);
		int ldB;
		i32 *block = nullptr;
		if (next_ahead < plan.size() && plan[next_ahead].Sn == Sn && plan[next_ahead].n == n && plan[next_ahead].w == w) {
			/* the prediction holds: consume this block's rand() draws like the reference, use the solved rows */
			rand_skip((i64) Sn * w);
			block = ahead.ptr + (size_t) ahead_offset[next_ahead] * ahead_ld;
			ldB = ahead_ld;
			next_ahead++;
		} else {
			plan.clear();
			next_ahead = 0;
			if (w > 0 && !no_ahead && rand_rewind_works())
				plan = plan_lowrank(n, rank_ub, w, opts);      /* predict: every block has full rank and keeps the weight */
			if (plan.size() > 1) {
				randomized_blocks_ahead(E, A, p, plan, ahead, ahead_ld, ahead_offset);
				rand_skip((i64) Sn * w);
				block = ahead.ptr;
				ldB = ahead_ld;
				next_ahead = 1;
			} else {
				plan.clear();
				randomized_block(E, A, p, n, Sn, w, B, ldB);
				block = B.ptr;
			}
		}
		int rr = E.absorb_block(block, Sn, ldB);
		record_block(Sn, Sm, rr, w);
		if (rr == 0) {
			if (test_completion(E, A, p, n))
				break;
			LOG("[echelonize/dense/low-rank] Failed termination test; switching to full linear combinations\n");
			w = 0;
			Sn = 1;          /* the reference stores omp_get_max_threads(); only the comparison below reads it */
		}
		if (rr < 0.9 * Sn) {
			w *= 2;
			LOG("[echelonize/dense/low-rank] Not enough pivots, increasing weight to %d\n", w);
		}
		n -= rr;             /* sic (echelonize.c:369): the sampling range shrinks */
		Sm -= rr;
		rank_ub -= rr;
		round += 1;
		LOG("[echelonize/dense/low-rank] found %d new pivots\n", rr);
	}
}

/* reference: src/spasm_echelonize.c:385-463 */
static void finish_dense(Engine &E, const DevCsr &A, const int *p, int n, const struct echelonize_opts *opts, Speculation *spec = nullptr,
                         const int *p_in = nullptr)
{
	E.begin_dense();
	cudaStream_t s = ctx().stream;
	int Sm = E.m - E.rank();
	LOG("[echelonize/dense] processing dense schur complement of dimension %d x %d; block size=%d\n", n, Sm, opts->dense_block_size);
	int processed = 0, round = 0;
	bool lowrank_mode = false;
	int rank_ub = spasm_min(A.n - E.rank(), A.m - E.rank());
	DevBuf<i32> B;
	int ahead_begin = 0, ahead_end = 0, ldB = 0;          /* rows [ahead_begin, ahead_end) of p are solved and sit in B */
	const int base = processed;
	const int *p0 = p;
	static const bool no_ahead = getenv("SPASM_B200_NO_SOLVE_AHEAD") != NULL;
	if (spec && spec->valid && spec->kind == 2 && spec->n == n && spec->take > 0) {
		/* the first batch was solved together with the density estimate */
		B = std::move(spec->B);
		ldB = spec->ldB;
		ahead_begin = 0;
		ahead_end = spec->take;
		spec->valid = false;
	}
	for (;;) {
		int Sn = spasm_min(opts->dense_block_size, n - processed);
		if (Sn <= 0)
			break;
		LOG("[echelonize/dense] Round %d. processing S[%d:%d] (%d x %d)\n", round, processed, processed + Sn, Sn, Sm);
		if (processed >= ahead_end) {
			/* structural solves of the next blocks in one batch */
			int take = dense_rows_ahead(E.Sm0, n - processed, Sn, opts, no_ahead);
			DevBuf<int> d_rows;
			d_rows.upload(p0 + (processed - base), (size_t) take, s);
			E.block_from_rows(A, d_rows.ptr, take, B, ldB);
			ahead_begin = processed;
			ahead_end = processed + take;
		}
		int rr = E.absorb_block(B.ptr + (size_t) (processed - ahead_begin) * ldB, Sn, ldB, E.want_L);
		if (E.want_L && rr > 0)
			for (int local : E.blocks.back().lu_row) {       /* update_fact_after_LU, echelonize.c:252-287: Lp[U->n + i] = iorig */
				const int row = p[local];
				E.p_dense.push_back(p_in ? p_in[row] : row);
			}
		record_block(Sn, Sm, rr, -1);
		round += 1;
		processed += Sn;
		p += Sn;
		Sm = E.m - E.rank();
		rank_ub = spasm_min(A.n - E.rank(), A.m - E.rank());
		LOG("[echelonize/dense] found %d new pivots\n", rr);
		if (opts->enable_tall_and_skinny && (rr < opts->low_rank_ratio * Sn)) {
			lowrank_mode = true;
			break;
		}
	}
	if (rank_ub > 0 && n - processed > 0 && lowrank_mode) {
		LOG("[echelonize/dense] Too few pivots; switching to low-rank mode\n");
		finish_lowrank(E, A, p, n - processed, opts);
	}
}

/* ------------------------------------------------------------------ sparse finisher (GPLU) */

__global__ void k_gplu_mark_cols(i64 nnz, const int *__restrict__ Sj, int *flag)
{
	i64 e = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (e < nnz)
		flag[Sj[e]] = 1;
}

__global__ void k_gplu_list_cols(int m, const int *__restrict__ flag, const int *__restrict__ off, int *cols, int *colmap)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m)
		return;
	colmap[j] = flag[j] ? off[j] : -1;
	if (flag[j])
		cols[off[j]] = j;
}

/* D[r][colmap[j]] = x for every entry (j, x) of row r */
__global__ void k_gplu_csr_to_dense(int R, const i64 *__restrict__ Sp, const int *__restrict__ Sj, const i32 *__restrict__ Sx,
                                    const int *__restrict__ colmap, i32 *D, int ld)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (warp >= R)
		return;
	for (i64 e = Sp[warp] + lane; e < Sp[warp + 1]; e += 32)
		D[(size_t) warp * ld + colmap[Sj[e]]] = Sx[e];
}

__global__ void k_gplu_register(int nnew, const i64 *__restrict__ Rp, const int *__restrict__ Rj, int un0, int *Uqinv)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < nnew)
		Uqinv[Rj[Rp[t]]] = un0 + t;
}

/* append nnew rows (device CSR, pivot first and equal to 1) to the structural U */
static void append_device_rows(Engine &E, DevBuf<i64> &Rp, DevBuf<int> &Rj, DevBuf<i32> &Rx, int nnew, i64 nnz)
{
	cudaStream_t s = ctx().stream;
	DevCsr &U = E.U;
	const int un0 = U.n;
	const i64 unz0 = U.nnz;
	DevBuf<i64> np((size_t) un0 + nnew + 1);
	DevBuf<int> nj((size_t) std::max<i64>(unz0 + nnz, 1));
	DevBuf<i32> nx((size_t) std::max<i64>(unz0 + nnz, 1));
	if (un0 > 0) {
		CUDA_CHECK(cudaMemcpyAsync(np.ptr, U.p.ptr, ((size_t) un0 + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(nj.ptr, U.j.ptr, (size_t) unz0 * sizeof(int), cudaMemcpyDeviceToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(nx.ptr, U.x.ptr, (size_t) unz0 * sizeof(i32), cudaMemcpyDeviceToDevice, s));
	} else {
		CUDA_CHECK(cudaMemsetAsync(np.ptr, 0, sizeof(i64), s));
	}
	k_gplu_register<<<cdiv(nnew, 256), 256, 0, s>>>(nnew, Rp.ptr, Rj.ptr, un0, E.Uqinv.ptr);
	if (unz0)
		k_add_offset<<<cdiv(nnew + 1, 256), 256, 0, s>>>(Rp.ptr, nnew + 1, unz0);
	LAUNCHED(2);
	CUDA_CHECK(cudaMemcpyAsync(np.ptr + un0, Rp.ptr, ((size_t) nnew + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
	CUDA_CHECK(cudaMemcpyAsync(nj.ptr + unz0, Rj.ptr, (size_t) nnz * sizeof(int), cudaMemcpyDeviceToDevice, s));
	CUDA_CHECK(cudaMemcpyAsync(nx.ptr + unz0, Rx.ptr, (size_t) nnz * sizeof(i32), cudaMemcpyDeviceToDevice, s));
	sync();
	U.p = std::move(np);
	U.j = std::move(nj);
	U.x = std::move(nx);
	U.n = un0 + nnew;
	U.nnz = unz0 + nnz;
	E.G_ready = false;                 /* the solve schedule is rebuilt before the next batch */
}

/* reference: src/spasm_echelonize.c:30-51, against a U that is entirely sparse: ceil(128 / log2 p) random combinations of
 * all the rows must reduce to zero */
static bool test_completion_sparse(Engine &E, const DevCsr &A, const int *p, int n)
{
	if (n == 0 || A.nnz == 0)
		return true;
	cudaStream_t s = ctx().stream;
	const int Sn = (int) ceil(128 / log2((double) E.prime));
	LOG("[echelonize/completion] Testing completion with %d random linear combinations (rank %d)\n", Sn, E.rank());
	std::vector<int> rows((size_t) Sn * n);
	std::vector<i32> coef((size_t) Sn * n);
	for (int k = 0; k < Sn; k++) {
		spasm_prng_ctx prng;
		spasm_prng_seed_simple(E.prime, (u64) k, 0, &prng);
		for (int i = 0; i < n; i++) {
			rows[(size_t) k * n + i] = p[i];
			coef[(size_t) k * n + i] = spasm_prng_ZZp(&prng);
		}
	}
	DevBuf<int> d_rows;
	DevBuf<i32> d_coef;
	d_rows.upload(rows.data(), rows.size(), s);
	d_coef.upload(coef.data(), coef.size(), s);
	E.solve_combos(A, d_rows.ptr, d_coef.ptr, Sn, n);
	const i64 left = panel_count_nonzero(E.panel, E.Uqinv.ptr);
	record_block(Sn, E.m - E.rank(), left == 0 ? 0 : 1, 0);
	return left == 0;
}

/*
 * The reference's sparse finisher (echelonize_GPLU, src/spasm_echelonize.c:54-187) processes the rows one at a time:
 * solve against the current U, leftmost entry on a non-pivotal column = new pivot, append the scaled row to U (which
 * stays SPARSE), early abort when the rank is reached or when no pivot was found for a while and random combinations of
 * the rows all reduce to zero.  Here the rows are processed in batches: one batched solve against the current U, the
 * reduced rows (sparse, on non-pivotal columns) are compacted to the columns they touch and echelonized among
 * themselves (dense_rref on a batch x touched-columns block), the resulting rows go back to U as sparse rows.  The pivot
 * columns are the leading positions of the row space either way (an invariant of the row space), so rank, pivot columns,
 * RREF and kernel are the reference's; U never holds a dense row over all remaining columns (the dense block
 * path needs rank x Sm x 4 bytes).  Chosen when the Schur complement is sparse (density <= sparsity_threshold).
 */
static void finish_gplu(Engine &E, const DevCsr &A, const int *p, int n, const struct echelonize_opts *opts, const int *p_in = nullptr)
{
	(void) opts;
	cudaStream_t s = ctx().stream;
	const int m = E.m;
	const int r_ub = spasm_min(A.n, m);
	const int batch = getenv("SPASM_B200_GPLU_BATCH") ? std::max(1, atoi(getenv("SPASM_B200_GPLU_BATCH"))) : 1024;
	LOG("[echelonize/GPLU] processing matrix of dimension %d x %d\n", n, m);
	int rows_since_last_pivot = 0;
	bool early_abort_done = false;
	DevBuf<int> flag((size_t) m + 1), off((size_t) m + 1), cols((size_t) m), colmap((size_t) m);
	static DevBuf<char> tmp;
	for (int done = 0; done < n;) {
		if (!E.want_L && E.U.n == r_ub) {  /* the reference's test, literally (:88): A is the CURRENT matrix, U->n the total rank */
			LOG("\n[echelonize/GPLU] full rank reached\n");
			break;
		}
		if (!E.want_L && !early_abort_done && rows_since_last_pivot > 10 && rows_since_last_pivot > n / 100) {
			LOG("\n[echelonize/GPLU] testing for early abort...\n");
			if (test_completion_sparse(E, A, p, n))
				break;
			early_abort_done = true;
		}
		const int R = std::min(batch, n - done);
		DevBuf<int> d_rows;
		d_rows.upload(p + done, (size_t) R, s);
		E.solve_rows(A, d_rows.ptr, R, false);
		const int batch_begin = done;
		done += R;
		DevBuf<i64> Sp;
		DevBuf<int> Sj;
		DevBuf<i32> Sx;
		i64 nnz = 0;
		panel_to_csr(E.panel, E.Uqinv.ptr, nullptr, 0, nullptr, Sp, Sj, Sx, nnz);
		if (nnz == 0) {
			rows_since_last_pivot += R;
			continue;
		}
		/* the columns the batch touches */
		flag.zero(s);
		k_gplu_mark_cols<<<cdiv((size_t) nnz, 256), 256, 0, s>>>(nnz, Sj.ptr, flag.ptr);
		size_t bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.ptr, off.ptr, m + 1, s);
		tmp.ensure(bytes + 16);
		cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, flag.ptr, off.ptr, m + 1, s);
		k_gplu_list_cols<<<cdiv(m, 256), 256, 0, s>>>(m, flag.ptr, off.ptr, cols.ptr, colmap.ptr);
		LAUNCHED(3);
		const int C = fetch(off.ptr + m);
		const int ld = std::max((C + 3) & ~3, 4);
		DevBuf<i32> D((size_t) R * ld);
		D.zero(s);
		k_gplu_csr_to_dense<<<cdiv((size_t) R * 32, 256), 256, 0, s>>>(R, Sp.ptr, Sj.ptr, Sx.ptr, colmap.ptr, D.ptr, ld);
		LAUNCHED(1);
		DevBuf<i32> before;              /* L mode (lu.cu): the batch before it is echelonized */
		if (E.want_L) {
			before.alloc((size_t) R * ld);
			CUDA_CHECK(cudaMemcpyAsync(before.ptr, D.ptr, (size_t) R * ld * sizeof(i32), cudaMemcpyDeviceToDevice, s));
		}
		GpuTimer t;
		t.start();
		RrefResult res = dense_rref(D.ptr, R, C, ld, E.F);
		stats().pub.ms_dense += t.stop_ms();
		record_block(R, C, res.rank, -2);
		if (res.rank == 0) {
			rows_since_last_pivot += R;
			continue;
		}
		/* the pivot rows of the block, as sparse rows of U (pivot first) */
		DevBuf<int> d_prow, d_pcol;
		d_prow.upload(res.pivrow.data(), res.pivrow.size(), s);
		d_pcol.upload(res.pivcol.data(), res.pivcol.size(), s);
		DevBuf<i32> Dr((size_t) res.rank * ld);
		dense_gather_rows(D.ptr, ld, d_prow.ptr, res.rank, C, Dr.ptr, ld);
		std::vector<unsigned char> own((size_t) std::max(C, 1), 0);
		for (int c : res.pivcol)
			own[c] = 1;
		DevBuf<unsigned char> d_own;
		d_own.upload(own.data(), own.size(), s);
		DevBuf<i64> Rp;
		DevBuf<int> Rj;
		DevBuf<i32> Rx;
		i64 rnz = 0;
		if (E.want_L) {
			/* rows of U in elimination order (row t = a row of the batch reduced by the rows before it): batch = Cm * Dr,
			 * Cm = Pi^t Lc Uc, rows = Uc * Dr; L is read out of the final U (compute_L) */
			const int ldc = std::max((res.rank + 3) & ~3, 4);
			DevBuf<i32> Cm((size_t) R * ldc), Dlu((size_t) res.rank * ld);
			dense_gather_columns(before.ptr, ld, R, d_pcol.ptr, res.rank, Cm.ptr, ldc);
			std::vector<int> lu_row;
			dense_lu_fullcol(Cm.ptr, R, res.rank, ldc, E.F, lu_row);
			dense_lu_rows(Cm.ptr, ldc, res.rank, lu_row, Dr.ptr, ld, C, Dlu.ptr, ld, E.F);
			dense_rows_to_csr(Dlu.ptr, ld, res.rank, C, d_pcol.ptr, nullptr, cols.ptr, Rp, Rj, Rx, rnz);
			for (int local : lu_row) {
				const int row = p[batch_begin + local];
				E.p_struct.push_back(p_in ? p_in[row] : row);
			}
		} else {
			dense_rows_to_csr(Dr.ptr, ld, res.rank, C, d_pcol.ptr, d_own.ptr, cols.ptr, Rp, Rj, Rx, rnz);
		}
		append_device_rows(E, Rp, Rj, Rx, res.rank, rnz);
		rows_since_last_pivot = 0;
		early_abort_done = false;
		LOG("\r[echelonize/GPLU] %d / %d [|U| = %" PRId64 "] --- rank >= %d", done, n, E.U.nnz, E.U.n);
	}
	LOG("\n");
}

/* ------------------------------------------------------------------ result assembly */

/* copy the echelon form to malloc'ed host memory: structural rows, then the dense rows in the
 * order they were found (reference: update_U_after_rref, echelonize.c:192-223) */
static struct spasm_lu *assemble(Engine &E, int n_rows_alloc)
{
	cudaStream_t s = ctx().stream;
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace] assemble/%-19s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	struct Piece { DevBuf<i64> p; DevBuf<int> j; DevBuf<i32> x; i64 nnz; int rows; };
	std::vector<Piece> pieces;
	i64 total = E.U.nnz;
	for (DenseBlock &blk : E.blocks) {
		Piece pc;
		pc.rows = blk.rr;
		if (blk.Dlu.ptr)       /* L mode: unit upper triangular rows, only the row's own pivot is implied */
			dense_rows_to_csr(blk.Dlu.ptr, blk.ld, blk.rr, E.Sm0, blk.d_pivcol.ptr, nullptr, E.d_q0.ptr, pc.p, pc.j, pc.x, pc.nnz);
		else
			dense_rows_to_csr(blk.D.ptr, blk.ld, blk.rr, E.Sm0, blk.d_pivcol.ptr, blk.d_own.ptr, E.d_q0.ptr, pc.p, pc.j, pc.x, pc.nnz);
		total += pc.nnz;
		pieces.push_back(std::move(pc));
	}
	lap("dense rows -> CSR");
	int rank = E.rank();
	struct spasm_csr *U = spasm_csr_alloc(std::max(rank, n_rows_alloc), E.m, std::max<i64>(total, 1), E.prime, true);
	int *qinv = (int *) spasm_malloc((i64) std::max(E.m, 1) * sizeof(int));
	if (E.lazy_rows == 0)
		E.Uqinv.download(qinv, (size_t) E.m, s);
	int *Lp = NULL;                   /* L mode: fact->p, row of the input behind every row of U */
	if (E.want_L) {
		if ((int) (E.p_struct.size() + E.p_dense.size()) != rank || (int) E.p_struct.size() != E.U.n)
			errx(1, "[spasm-b200] internal: %zu + %zu rows of origin for %d + %d pivots", E.p_struct.size(), E.p_dense.size(), E.U.n, E.dense_rank);
		Lp = (int *) spasm_malloc((i64) std::max(rank, 1) * sizeof(int));
		for (int t = 0; t < E.U.n; t++)
			Lp[t] = E.p_struct[t];
		for (int t = 0; t < E.dense_rank; t++)
			Lp[E.U.n + t] = E.p_dense[t];
	}
	if (E.lazy_rows > 0) {
		/* The first lazy_rows rows of U are in column order (lazy schedule, extract_structural).  Put them in
		 * (level, column) order like the eager path does: the levels came out of the first solve pass; if no solve
		 * ever ran, or if later rounds invalidated the schedule, they are computed now. */
		if (!E.G_ready || !E.G.levels_known) {
			E.G = DepGraph();
			depgraph_forward(E.U, E.G);
			depgraph_schedule(E.G);
			E.G_ready = true;
			stats().pub.dag_depth = E.G.nlevels;
		}
		/* the permutation and the permuted CSR are built on the device and downloaded in place (the host loop over
		 * 135 k rows and three downloads of unpermuted arrays was 6 ms of a 100 ms call on config 2, 17 ms on config 4) */
		const int L = E.lazy_rows, un = E.U.n;
		DevBuf<int> flags((size_t) E.m), rows_ord((size_t) E.m), perm((size_t) std::max(un, 1));
		k_select_rows_below<<<cdiv(E.m, 256), 256, 0, s>>>(E.m, E.G.order.ptr, E.Uqinv.ptr, L, flags.ptr, rows_ord.ptr);
		LAUNCHED(1);
		const int got = compact_flagged(rows_ord.ptr, flags.ptr, E.m, perm.ptr);
		if (got != L)
			errx(1, "[spasm-b200] internal: level order covers %d of %d structural rows", got, L);
		DevBuf<i64> len((size_t) un + 1), newp((size_t) un + 1);
		DevBuf<int> newj((size_t) std::max<i64>(E.U.nnz, 1)), d_qinv((size_t) std::max(E.m, 1));
		DevBuf<i32> newx((size_t) std::max<i64>(E.U.nnz, 1));
		k_perm_lengths<<<cdiv((size_t) un + 1, 256), 256, 0, s>>>(un, L, perm.ptr, E.U.p.ptr, len.ptr);
		{
			static DevBuf<char> tmp;
			size_t bytes = 0;
			cub::DeviceScan::ExclusiveSum(nullptr, bytes, len.ptr, newp.ptr, un + 1, s);
			tmp.ensure(bytes + 16);
			cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, len.ptr, newp.ptr, un + 1, s);
		}
		CUDA_CHECK(cudaMemcpyAsync(d_qinv.ptr, E.Uqinv.ptr, (size_t) E.m * sizeof(int), cudaMemcpyDeviceToDevice, s));
		k_perm_rows<<<std::min(cdiv((size_t) un * 32, 256), 148u * 16), 256, 0, s>>>(un, perm.ptr, E.U.p.ptr, E.U.j.ptr, E.U.x.ptr, newp.ptr, newj.ptr, newx.ptr, d_qinv.ptr);
		LAUNCHED(4);
		newp.download(U->p, (size_t) un + 1, s);
		newj.download(U->j, (size_t) E.U.nnz, s);
		newx.download(U->x, (size_t) E.U.nnz, s);
		d_qinv.download(qinv, (size_t) E.m, s);
		if (Lp) {
			std::vector<int> hperm((size_t) std::max(un, 1));
			perm.download(hperm.data(), (size_t) un, s);
			sync();
			for (int t = 0; t < un; t++)
				Lp[t] = E.p_struct[hperm[t]];
		}
	} else {
		E.U.p.download(U->p, (size_t) E.U.n + 1, s);
		E.U.j.download(U->j, (size_t) E.U.nnz, s);
		E.U.x.download(U->x, (size_t) E.U.nnz, s);
	}
	sync();
	lap("structural rows");
	i64 off = E.U.nnz;
	int row = E.U.n;
	for (size_t b = 0; b < pieces.size(); b++) {
		Piece &pc = pieces[b];
		std::vector<i64> hp((size_t) pc.rows + 1);
		pc.p.download(hp.data(), hp.size(), s);
		pc.j.download(U->j + off, (size_t) pc.nnz, s);
		pc.x.download(U->x + off, (size_t) pc.nnz, s);
		sync();
		for (int t = 0; t < pc.rows; t++) {
			U->p[row + t + 1] = off + hp[t + 1];
			qinv[E.q0[E.blocks[b].pivcol[t]]] = row + t;
		}
		off += pc.nnz;
		row += pc.rows;
	}
	lap("download dense rows");
	stats().pub.d2h_bytes += total * 8 + (i64) (rank + 1) * 8 + (i64) E.m * 4;
	U->n = rank;
	/* trim like the reference does (echelonize.c:603-604) */
	spasm_csr_resize(U, rank, E.m);
	spasm_csr_realloc(U, -1);
	struct spasm_lu *fact = (struct spasm_lu *) spasm_malloc(sizeof(*fact));
	fact->r = rank;
	fact->complete = false;
	fact->L = NULL;
	fact->U = U;
	fact->qinv = qinv;
	fact->p = Lp;
	fact->Ltmp = NULL;
	lap("trim");
	return fact;
}

/* reference: src/spasm_echelonize.c:473-617.  A lives in HBM already; the result stays in E. */
static void echelonize_core(Engine &E, const DevCsr &dA0, struct echelonize_opts *opts)
{
	Stats &st = stats();
	st.pair_row.clear();
	st.pair_col.clear();
	st.pair_start.assign(1, 0);
	st.pub.nrounds = 0;
	st.pub.nblocks = 0;
	st.pub.finish = 0;
	int n = dA0.n, m = dA0.m;
	LOG("[echelonize] Start on %d x %d matrix with %" PRId64 " nnz\n", n, m, dA0.nnz);
	if (opts->complete)
		opts->L = 1;
	if (opts->L)
		opts->enable_tall_and_skinny = 0;       /* echelonize.c:489-490 ("for now") */
	if (opts->dense_block_size <= 0)
		errx(1, "[spasm-b200] dense_block_size must be positive");

	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace] core/%-23s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	E.init(m, dA0.prime);
	E.want_L = opts->L != 0;
	lap("init");
	DevCsr dS;                       /* current Schur complement once a round has run */
	const DevCsr *cur = &dA0;

	Speculation spec;
	std::vector<int> p((size_t) std::max(n, 1));
	std::vector<int> p_in;           /* empty = identity */
	double density = (double) dA0.nnz / n / m;
	int npiv = 0, status = 0, round;
	for (round = 0; round < opts->max_round; round++) {
		if (cur->nnz == 0) {
			LOG("[echelonize] empty matrix\n");
			status = 1;
			break;
		}
		LOG("[echelonize] round %d\n", round);
		npiv = extract_structural(E, *cur, p_in.empty() ? NULL : p_in.data(), p.data(), opts->enable_greedy_pivot_search, round, true);
		if (round == 0)
			E.first_round_rows = npiv;       /* rows 0 .. npiv-1 of U: one scaled row of the INPUT each (compute_L) */
		lap("extract_structural");
		st.pub.nrounds = round + 1;
		st.pair_start.push_back((int) st.pair_row.size());
		if (npiv < opts->min_pivot_proportion * spasm_min(n, m - E.U.n)) {
			LOG("[echelonize] not enough pivots found; stopping\n");
			status = 2;
			break;
		}
		density = estimate_density_speculative(E, *cur, p.data() + npiv, n - npiv, 100, opts, spec);
		lap("estimate_density");
		if (round < 64)
			st.pub.density[round] = density;
		if (density > opts->sparsity_threshold) {
			LOG("[echelonize] Schur complement is dense (estimated %.2f%%)\n", 100 * density);
			status = 2;
			break;
		}
		spec.discard(E);                 /* sparse after all: the finisher's batch is not needed */
		LOG("Schur complement is %d x %d, estimated density : %.2f\n", n - npiv, m - E.U.n, density);
		DevCsr S;
		schur_sparse(E, *cur, p.data() + npiv, n - npiv, S, nullptr);
		lap("schur_sparse");
		std::vector<int> p_out((size_t) (n - npiv));
		for (int k = 0; k < n - npiv; k++) {
			int row = p[npiv + k];
			p_out[k] = p_in.empty() ? row : p_in[row];
		}
		LOG("Schur complement: %d * %d [%" PRId64 " nz / density= %.3f]\n", n - npiv, m, S.nnz, 1.0 * S.nnz / (1.0 * m * (n - npiv)));
		dS = std::move(S);
		cur = &dS;
		n = n - npiv;
		p_in = std::move(p_out);
		if (p_in.empty())
			p_in.push_back(0);           /* keep "non-empty == not the identity" even for zero rows */
	}
	if (status == 0) {
		npiv = 0;
		for (int i = 0; i < n; i++)
			p[i] = i;
	}
	if (status != 1) {
		if (!opts->enable_tall_and_skinny)
			LOG("[echelonize] dense low-rank mode disabled\n");
		if (!opts->enable_dense)
			LOG("[echelonize] regular dense mode disabled\n");
		if (!opts->enable_GPLU)
			LOG("[echelonize] GPLU mode disabled\n");
		double aspect_ratio = (double) (n - npiv) / (m - E.U.n);
		LOG("[echelonize] finishing; density = %.3f; aspect ratio = %.1f\n", density, aspect_ratio);
		if (opts->enable_tall_and_skinny && aspect_ratio > opts->tall_and_skinny_ratio) {
			st.pub.finish = 1;
			finish_lowrank(E, *cur, p.data() + npiv, n - npiv, opts, &spec);
		} else if (opts->enable_dense && density > opts->sparsity_threshold) {
			st.pub.finish = 2;
			finish_dense(E, *cur, p.data() + npiv, n - npiv, opts, &spec, p_in.empty() ? NULL : p_in.data());
		} else if (opts->enable_GPLU) {
			/* The reference finishes row by row (echelonize_GPLU, echelonize.c:54-187): leftmost pivot of each
			 * reduced row.  That rule selects the column rank profile of the Schur complement, which is what
			 * the dense echelon form selects too, so the canonical result (rank, pivot columns, RREF, kernel)
			 * is the same; here it is computed block-wise on the GPU.  (a sequential GPU GPLU
			 * is a later row.)  No low-rank switch: GPLU processes every row. */
			st.pub.finish = 3;
			static const bool gplu_dense = getenv("SPASM_B200_GPLU_DENSE") != NULL;
			if (!gplu_dense && comm_world() == 1) {
				finish_gplu(E, *cur, p.data() + npiv, n - npiv, opts, p_in.empty() ? NULL : p_in.data());
				lap("finish");
				return;
			}
			{
				/* the block path keeps every new pivot row as a dense row over the columns that are non-pivotal now:
				 * rank_ub x Sm x 4 bytes (+ the packed copy).  The reference keeps U sparse here.  Refuse early, with
				 * the numbers, instead of failing inside an allocation half way through. */
				const double Sm_now = (double) (m - E.U.n);
				const double rank_ub = (double) std::min(n - npiv, m - E.U.n);
				const double need = rank_ub * Sm_now * 5.0;
				size_t free_b = 0, total_b = 0;
				CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
				if (need > 0.8 * (double) total_b)
					errx(1, "[spasm-b200] the sparse (GPLU) finisher is executed block-wise with dense pivot rows: %.0f x %.0f remaining "
					        "rows x columns need about %.1f GB of HBM (%.1f GB on this device). Pre-process the matrix (tools/vertical_swap, "
					        "--dense-threshold) so that fewer columns remain, or raise --max-iterations", rank_ub, Sm_now, need / 1e9,
					     (double) total_b / 1e9);
			}
			struct echelonize_opts o2 = *opts;
			o2.enable_tall_and_skinny = 0;
			finish_dense(E, *cur, p.data() + npiv, n - npiv, &o2, &spec, p_in.empty() ? NULL : p_in.data());
		} else {
			LOG("[echelonize] Cannot finish (no valid method enabled). Incomplete echelonization returned\n");
		}
		lap("finish");
	}
}

}  // namespace sb

using namespace sb;

/* ================================================================== C ABI */

extern "C" {

/* reference: src/spasm_echelonize.c:9-28 */
void spasm_echelonize_init_opts(struct echelonize_opts *opts)
{
	opts->enable_greedy_pivot_search = 1;
	opts->enable_tall_and_skinny = 1;
	opts->enable_dense = 1;
	opts->enable_GPLU = 1;
	opts->L = 0;
	opts->complete = 0;
	opts->min_pivot_proportion = 0.1;
	opts->max_round = 3;
	opts->sparsity_threshold = 0.05;
	opts->tall_and_skinny_ratio = 5;
	opts->dense_block_size = 1000;
	opts->low_rank_ratio = 0.5;
	opts->low_rank_start_weight = -1;
}

/* reference: src/spasm_echelonize.c:473-617 -- host matrix in, malloc'ed host echelon form out */
struct spasm_lu *spasm_echelonize(const struct spasm_csr *A, struct echelonize_opts *opts)
{
	struct echelonize_opts default_opts;
	if (opts == NULL) {
		LOG("[echelonize] using default settings\n");
		opts = &default_opts;
		spasm_echelonize_init_opts(opts);
	}
	double start = spasm_wtime();
	ctx();
	GpuTimer timer;
	timer.start();
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	DevCsr dA0;
	dA0.upload(A);
	Engine E;
	echelonize_core(E, dA0, opts);
	if (trace) {
		sb::sync();
		fprintf(stderr, "[trace] upload + echelonize_core %8.3f ms\n", 1e3 * (spasm_wtime() - start));
	}
	struct spasm_lu *fact = assemble(E, 0);
	if (opts->L)
		compute_L(fact, dA0, A, opts->complete != 0, E.first_round_rows);
	stats().pub.ms_device_echelonize = timer.stop_ms();
	stats().pub.ms_total_echelonize = 1e3 * (spasm_wtime() - start);
	LOG("[echelonize] Done in %.3fs. Rank %d, %" PRId64 " nz in basis\n", spasm_wtime() - start, fact->U->n, spasm_nnz(fact->U));
	return fact;
}

/* ---- device-resident variant (include/spasm_b200.h): the input stays in HBM between calls */
void *spasm_b200_upload_csr(const struct spasm_csr *A)
{
	ctx();
	DevCsr *d = new DevCsr();
	d->upload(A);
	sb::sync();
	return d;
}

void spasm_b200_free_csr(void *handle) { delete (DevCsr *) handle; }

int spasm_b200_echelonize_resident(void *handle, struct echelonize_opts *opts, double *ms_device)
{
	struct echelonize_opts default_opts;
	if (opts == NULL) {
		opts = &default_opts;
		spasm_echelonize_init_opts(opts);
	}
	double start = spasm_wtime();
	GpuTimer timer;
	timer.start();
	Engine E;
	echelonize_core(E, *(DevCsr *) handle, opts);
	double ms = timer.stop_ms();
	stats().pub.ms_device_echelonize = ms;
	stats().pub.ms_total_echelonize = 1e3 * (spasm_wtime() - start);
	if (ms_device)
		*ms_device = ms;
	return E.rank();
}

/* write a buffer larger than the 126 MB L2 so that the next timed step starts cold */
void spasm_b200_flush_l2(void)
{
	static DevBuf<char> scrub;
	size_t bytes = (size_t) 512 << 20;
	scrub.ensure(bytes);
	static int v = 0;
	CUDA_CHECK(cudaMemsetAsync(scrub.ptr, ++v & 0xff, bytes, ctx().stream));
	sb::sync();
}

}  /* extern "C" */
