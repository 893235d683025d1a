/*
 * Dependency graphs of U, their level schedule, and the batched pull-form
 * triangular solve.  See solve.cuh for the formulation.
 */
#include <cooperative_groups.h>
#include <cub/cub.cuh>
#include "solve.cuh"
#include "stats.cuh"

namespace cg = cooperative_groups;

namespace sb {

/* ------------------------------------------------------------------ small helpers */

static void exclusive_scan_i64(const i64 *d_in, i64 *d_out, size_t n)
{
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, d_in, d_out, n, ctx().stream);
	tmp.ensure(bytes + 16);
	cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, d_in, d_out, n, ctx().stream);
	LAUNCHED(1);
}

template <typename K> static int coop_blocks(K kernel, int threads, size_t smem = 0)
{
	int occ = 0;
	CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem));
	if (occ < 1)
		errx(1, "[spasm-b200] cooperative kernel does not fit on an SM");
	return occ * ctx().sm_count;
}

/* ------------------------------------------------------------------ forward dependency graph */

/* one warp per row of U; pass 0 counts, pass 1 fills */
template <int PASS>
__global__ void k_fwd_deps(int n, const i64 *__restrict__ Up, const int *__restrict__ Uj, const i32 *__restrict__ Ux,
                           unsigned long long *cnt_or_cursor, int *src, i32 *val)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int i = warp; i < n; i += nwarps) {
		i64 b = Up[i], e = Up[i + 1];
		if (b == e)
			continue;
		int pc = Uj[b];
		for (i64 k = b + 1 + lane; k < e; k += 32) {
			int c = Uj[k];
			if (c == pc)
				continue;          /* a repeated pivot column carries no dependency (reference: pivots.c:435) */
			if (PASS == 0) {
				atomicAdd(&cnt_or_cursor[c], 1ull);
			} else {
				unsigned long long pos = atomicAdd(&cnt_or_cursor[c], 1ull);
				src[pos] = pc;
				val[pos] = Ux[k];
			}
		}
	}
}

void depgraph_forward(const DevCsr &U, DepGraph &G)
{
	cudaStream_t s = ctx().stream;
	G.nnodes = U.m;
	DevBuf<i64> cnt((size_t) U.m + 1);
	cnt.zero(s);
	int blocks = std::max(1u, std::min(cdiv((size_t) U.n * 32, 256), 148u * 16));
	if (U.n > 0) {
		k_fwd_deps<0><<<blocks, 256, 0, s>>>(U.n, U.p, U.j, U.x, (unsigned long long *) cnt.ptr, nullptr, nullptr);
		LAUNCHED(1);
	}
	G.ptr.alloc((size_t) U.m + 1);
	exclusive_scan_i64(cnt.ptr, G.ptr.ptr, (size_t) U.m + 1);
	G.ndeps = fetch(G.ptr.ptr + U.m);
	G.src.alloc((size_t) std::max<i64>(G.ndeps, 1));
	G.val.alloc((size_t) std::max<i64>(G.ndeps, 1));
	if (U.n > 0 && G.ndeps > 0) {
		CUDA_CHECK(cudaMemcpyAsync(cnt.ptr, G.ptr.ptr, ((size_t) U.m + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
		k_fwd_deps<1><<<blocks, 256, 0, s>>>(U.n, U.p, U.j, U.x, (unsigned long long *) cnt.ptr, G.src, G.val);
		LAUNCHED(1);
	}
	KERNEL_CHECK();
	sync();
}

/* ------------------------------------------------------------------ transposed dependency graph */

template <int PASS>
__global__ void k_tr_deps(int n, const i64 *__restrict__ Up, const int *__restrict__ Uj, const i32 *__restrict__ Ux,
                          const int *__restrict__ qinv, i64 *cnt, const i64 *__restrict__ ptr, int *src, i32 *val)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int i = warp; i < n; i += nwarps) {
		i64 b = Up[i], e = Up[i + 1];
		if (b == e)
			continue;
		int pc = Uj[b];
		/* entries of row i on pivotal columns other than its own pivot; the warp compacts them in order */
		i64 base = (PASS == 1) ? ptr[i] : 0;
		int total = 0;
		for (i64 k0 = b + 1; k0 < e; k0 += 32) {
			i64 k = k0 + lane;
			int c = (k < e) ? Uj[k] : -1;
			int owner = (c >= 0 && c != pc) ? qinv[c] : -1;
			unsigned mask = __ballot_sync(0xffffffffu, owner >= 0);
			if (PASS == 1 && owner >= 0) {
				int off = __popc(mask & ((1u << lane) - 1));
				src[base + total + off] = owner;
				val[base + total + off] = Ux[k];
			}
			total += __popc(mask);
		}
		if (PASS == 0 && lane == 0)
			cnt[i] = total;
	}
}

void depgraph_transposed(const DevCsr &U, const int *d_qinv, DepGraph &G)
{
	cudaStream_t s = ctx().stream;
	G.nnodes = U.n;
	DevBuf<i64> cnt((size_t) U.n + 1);
	cnt.zero(s);
	int blocks = std::max(1u, std::min(cdiv((size_t) U.n * 32, 256), 148u * 16));
	if (U.n > 0) {
		k_tr_deps<0><<<blocks, 256, 0, s>>>(U.n, U.p, U.j, U.x, d_qinv, cnt.ptr, nullptr, nullptr, nullptr);
		LAUNCHED(1);
	}
	G.ptr.alloc((size_t) U.n + 1);
	exclusive_scan_i64(cnt.ptr, G.ptr.ptr, (size_t) U.n + 1);
	G.ndeps = fetch(G.ptr.ptr + U.n);
	G.src.alloc((size_t) std::max<i64>(G.ndeps, 1));
	G.val.alloc((size_t) std::max<i64>(G.ndeps, 1));
	if (U.n > 0 && G.ndeps > 0) {
		k_tr_deps<1><<<blocks, 256, 0, s>>>(U.n, U.p, U.j, U.x, d_qinv, nullptr, G.ptr, G.src, G.val);
		LAUNCHED(1);
	}
	KERNEL_CHECK();
	sync();
}

/* ------------------------------------------------------------------ level schedule (Kahn) */

__global__ void k_rev_count(int nnodes, const i64 *__restrict__ ptr, const int *__restrict__ src, unsigned long long *rcnt, int *indeg)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nnodes)
		return;
	i64 b = ptr[c], e = ptr[c + 1];
	indeg[c] = (int) (e - b);
	for (i64 k = b; k < e; k++)
		atomicAdd(&rcnt[src[k]], 1ull);
}

__global__ void k_rev_fill(int nnodes, const i64 *__restrict__ ptr, const int *__restrict__ src, unsigned long long *cursor, int *rdst)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nnodes)
		return;
	for (i64 k = ptr[c]; k < ptr[c + 1]; k++)
		rdst[atomicAdd(&cursor[src[k]], 1ull)] = c;
}

__global__ void k_kahn_seed(int nnodes, const int *__restrict__ indeg, int *order, unsigned long long *state, int *tail)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= nnodes)
		return;
	state[c] = (unsigned long long) (unsigned) indeg[c];        /* level 0, all dependencies pending */
	if (indeg[c] == 0)
		order[atomicAdd(tail, 1)] = c;
}

/*
 * Levels of the dependency DAG without a barrier per level ("asynchronous Kahn").
 * order[] doubles as a work queue: slot t is claimed by exactly one thread (atomic ticket) which spins until the
 * slot is filled, then releases the dependents of that node: level[c] = max(level[c], level[s] + 1) and, when the
 * last dependency of c is released, c is appended to the queue.  Every node is appended exactly once (the graph
 * is acyclic), so all n tickets terminate; the critical path is depth x (one atomic round trip) instead of
 * depth x (two grid barriers).  A cycle (invalid input) would leave slots empty for ever: a spin budget turns
 * that into an error flag.  The grid is sized to be co-resident (the spinning threads must all run).
 */
__global__ void k_kahn_async(int n, const i64 *__restrict__ rptr, const int *__restrict__ rdst, unsigned long long *state, int *order,
                             int *level, int *tail, int *ticket, int *error, int *done)
{
	/* One WARP per node (idle lanes that spin in the same warp as a working lane would steal its issue slots).
	 * state[c] = (level so far << 32) | dependencies not yet released: one 64-bit CAS releases a dependency and
	 * raises the level at once, so the lane that releases the last one knows the final level.  The dependents of
	 * a node are released by the lanes in parallel.  The warp keeps one released node for itself (no queue round
	 * trip on the critical path: a chain of the DAG is walked by one warp); the others go to the shared queue. */
	const int lane = threadIdx.x & 31;
	int next = -1;
	for (;;) {
		int s = next;
		next = -1;
		if (s < 0) {
			if (lane == 0) {
				const int t = atomicAdd(ticket, 1);
				/* watchdog on PROGRESS, not on elapsed spins: the budget restarts whenever another node completes, so a
				 * long valid pass is never mistaken for a cycle; only "nothing completes any more" raises the flag */
				int seen_done = -1;
				for (long spin = 0;; spin++) {
					if (t < n) {
						s = *((volatile int *) &order[t]);
						if (s >= 0)
							break;
					}
					const int d_now = *((volatile int *) done);
					if (d_now >= n || *((volatile int *) error)) {
						s = -2;
						break;
					}
					if (d_now != seen_done) {
						seen_done = d_now;
						spin = 0;
					}
					if (spin > (1L << 24)) {
						*error = 1;
						s = -2;
						break;
					}
					__nanosleep(200);
				}
			}
			s = __shfl_sync(0xffffffffu, s, 0);
			if (s < 0)
				return;
		}
		const unsigned ls = (unsigned) (*((volatile unsigned long long *) &state[s]) >> 32);
		const i64 b = rptr[s], e = rptr[s + 1];
		if (lane == 0)
			level[s] = (int) ls;
		for (i64 k0 = b; k0 < e; k0 += 32) {
			const i64 k = k0 + lane;
			int released = -1;
			if (k < e) {
				const int c = rdst[k];
				unsigned long long old = *((volatile unsigned long long *) &state[c]), assumed, upd;
				do {
					assumed = old;
					unsigned lv = (unsigned) (assumed >> 32), cnt = (unsigned) assumed;
					lv = lv > ls + 1 ? lv : ls + 1;
					upd = ((unsigned long long) lv << 32) | (cnt - 1);
					old = atomicCAS(&state[c], assumed, upd);
				} while (old != assumed);
				if ((unsigned) upd == 0)
					released = c;
			}
			unsigned mask = __ballot_sync(0xffffffffu, released >= 0);
			if (mask) {
				int keep_lane = -1;
				if (next < 0) {
					keep_lane = __ffs(mask) - 1;
					next = __shfl_sync(0xffffffffu, released, keep_lane);
				}
				if (released >= 0 && lane != keep_lane) {
					const int pos = atomicAdd(tail, 1);
					__threadfence();
					*((volatile int *) &order[pos]) = released;
				}
			}
		}
		if (lane == 0)
			atomicAdd(done, 1);
	}
}

__global__ void k_make_keys(int n, const int *__restrict__ order, const int *__restrict__ level, unsigned long long *keys)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < n)
		keys[c] = ((unsigned long long) (unsigned) level[c] << 32) | (unsigned) c;
}

__global__ void k_keys_to_order(int n, const unsigned long long *__restrict__ keys, int *order)
{
	int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n)
		order[t] = (int) (keys[t] & 0xffffffffull);
}

/* level_ptr[L] = first position in the sorted keys whose level is >= L (L = 0 .. nlevels) */
__global__ void k_level_bounds(int n, const unsigned long long *__restrict__ keys, int nlevels, int *level_ptr)
{
	int L = blockIdx.x * blockDim.x + threadIdx.x;
	if (L > nlevels)
		return;
	int lo = 0, hi = n;
	while (lo < hi) {
		int mid = (lo + hi) >> 1;
		if ((int) (keys[mid] >> 32) < L)
			lo = mid + 1;
		else
			hi = mid;
	}
	level_ptr[L] = lo;
}

/* pending0[c] = dependencies of c whose source is not a level-0 node; seeds = level >= 1 nodes with pending0 == 0 */
__global__ void k_flow_prepare(int n, const i64 *__restrict__ ptr, const int *__restrict__ src, const int *__restrict__ level,
                               int *pending0, int *seeds, int *nseeds)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n)
		return;
	int cnt = 0;
	for (i64 e = ptr[c]; e < ptr[c + 1]; e++)
		cnt += level[src[e]] > 0;
	pending0[c] = cnt;
	if (level[c] >= 1 && cnt == 0)
		seeds[atomicAdd(nseeds, 1)] = c;
}

void depgraph_finish_levels(DepGraph &G)
{
	cudaStream_t s = ctx().stream;
	const int n = G.nnodes;
	/* deterministic order inside each level: sort by (level, node); level boundaries by binary search on the keys */
	DevBuf<unsigned long long> keys((size_t) n), keys2((size_t) n);
	k_make_keys<<<cdiv(n, 256), 256, 0, s>>>(n, G.order, G.level, keys.ptr);
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys.ptr, keys2.ptr, n, 0, 64, s);
	tmp.ensure(bytes + 16);
	cub::DeviceRadixSort::SortKeys(tmp.ptr, bytes, keys.ptr, keys2.ptr, n, 0, 64, s);
	k_keys_to_order<<<cdiv(n, 256), 256, 0, s>>>(n, keys2.ptr, G.order);
	unsigned long long last_key = 0;
	CUDA_CHECK(cudaMemcpyAsync(&last_key, keys2.ptr + (n - 1), sizeof(last_key), cudaMemcpyDeviceToHost, s));
	sync();
	G.nlevels = (int) (last_key >> 32) + 1;
	k_level_bounds<<<cdiv(G.nlevels + 1, 256), 256, 0, s>>>(n, keys2.ptr, G.nlevels, G.level_ptr.ptr);
	LAUNCHED(4);
	KERNEL_CHECK();
	G.level_ptr_h.resize((size_t) G.nlevels + 1);
	CUDA_CHECK(cudaMemcpyAsync(G.level_ptr_h.data(), G.level_ptr.ptr, ((size_t) G.nlevels + 1) * sizeof(int), cudaMemcpyDeviceToHost, s));
	sync();
	G.levels_known = true;
}

/* lazy schedule: pending0[c] = dependencies of c that have dependencies themselves; seeds = nodes with dependencies all
 * of whose dependencies are sources; *nsched = nodes with dependencies */
__global__ void k_flow_prepare_lazy(int n, const i64 *__restrict__ ptr, const int *__restrict__ src, int *pending0, int *seeds,
                                    int *nseeds, int *nsched)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n)
		return;
	const i64 b = ptr[c], e = ptr[c + 1];
	int cnt = 0;
	for (i64 k = b; k < e; k++) {
		const int sc = src[k];
		cnt += ptr[sc + 1] > ptr[sc];
	}
	pending0[c] = cnt;
	if (e > b) {
		atomicAdd(nsched, 1);
		if (cnt == 0)
			seeds[atomicAdd(nseeds, 1)] = c;
	}
}

void depgraph_schedule(DepGraph &G, bool lazy)
{
	cudaStream_t s = ctx().stream;
	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t_prev = spasm_wtime();
	auto lap = [&](const char *what) {
		if (trace) {
			sync();
			double now = spasm_wtime();
			fprintf(stderr, "[trace]     schedule/%-18s %8.3f ms\n", what, 1e3 * (now - t_prev));
			t_prev = now;
		}
	};
	int n = G.nnodes;
	G.level.alloc((size_t) n + 1);
	G.order.alloc((size_t) n + 1);
	G.level_ptr.alloc((size_t) n + 2);
	G.level_ptr_h.assign(1, 0);
	G.nlevels = 0;
	if (n == 0)
		return;
	DevBuf<i64> rcnt((size_t) n + 1);
	G.rptr.alloc((size_t) n + 1);
	G.rdst.alloc((size_t) std::max<i64>(G.ndeps, 1));
	DevBuf<i64> &rptr = G.rptr;
	DevBuf<int> &rdst = G.rdst;
	DevBuf<int> indeg((size_t) n), counters(4);
	rcnt.zero(s);
	counters.zero(s);
	G.order.fill_byte(0xff, s);        /* empty queue slots */
	k_rev_count<<<cdiv(n, 256), 256, 0, s>>>(n, G.ptr, G.src, (unsigned long long *) rcnt.ptr, indeg);
	exclusive_scan_i64(rcnt.ptr, rptr.ptr, (size_t) n + 1);
	CUDA_CHECK(cudaMemcpyAsync(rcnt.ptr, rptr.ptr, ((size_t) n + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
	k_rev_fill<<<cdiv(n, 256), 256, 0, s>>>(n, G.ptr, G.src, (unsigned long long *) rcnt.ptr, rdst);
	static const bool force_eager = getenv("SPASM_B200_EAGER_LEVELS") != NULL || getenv("SPASM_B200_SOLVE_LEVELS") != NULL;
	if (lazy && !force_eager) {
		G.level.zero(s);                      /* sources stay at level 0; the solve fills the rest */
		G.pending0.alloc((size_t) n);
		G.seeds.alloc((size_t) n);
		k_flow_prepare_lazy<<<cdiv(n, 256), 256, 0, s>>>(n, G.ptr, G.src, G.pending0.ptr, G.seeds.ptr, counters.ptr, counters.ptr + 1);
		LAUNCHED(3);
		KERNEL_CHECK();
		int hc[2];
		CUDA_CHECK(cudaMemcpyAsync(hc, counters.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
		sync();
		G.nseeds = hc[0];
		G.nscheduled = hc[1];
		G.nlevels = G.nscheduled > 0 ? 2 : 1;         /* placeholder until the first solve */
		G.level_ptr_h.assign({0, n - G.nscheduled, n});
		G.scheduled_deps = G.ndeps;
		G.levels_known = false;
		lap("lazy schedule");
		if (G.nscheduled == 0)
			depgraph_finish_levels(G);        /* no dependency at all: every node is at level 0 */
		return;
	}
	DevBuf<unsigned long long> state((size_t) n);
	k_kahn_seed<<<cdiv(n, 256), 256, 0, s>>>(n, indeg, G.order, state.ptr, counters.ptr);
	LAUNCHED(3);
	KERNEL_CHECK();
	lap("reverse graph");

	{
		int threads = 128;
		int occ = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_kahn_async, threads, 0));
		int blocks = std::max(1, std::min(occ, 4)) * ctx().sm_count;          /* co-resident: idle warps spin on the queue */
		k_kahn_async<<<blocks, threads, 0, s>>>(n, rptr.ptr, rdst.ptr, state.ptr, G.order.ptr, G.level.ptr, counters.ptr, counters.ptr + 1, counters.ptr + 2, counters.ptr + 3);
		LAUNCHED(1);
		KERNEL_CHECK();
	}
	int h[4];
	CUDA_CHECK(cudaMemcpyAsync(h, counters.ptr, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
	sync();
	lap("kahn");
	if (h[3] != n || h[2] != 0) {
		/* say which side is wrong: redo the topological sort on the host */
		std::vector<i64> hp((size_t) n + 1);
		std::vector<int> hs((size_t) std::max<i64>(G.ndeps, 1));
		CUDA_CHECK(cudaMemcpy(hp.data(), G.ptr.ptr, hp.size() * sizeof(i64), cudaMemcpyDeviceToHost));
		CUDA_CHECK(cudaMemcpy(hs.data(), G.src.ptr, (size_t) G.ndeps * sizeof(int), cudaMemcpyDeviceToHost));
		std::vector<int> deg((size_t) n, 0), stack;
		std::vector<std::vector<int>> out((size_t) n);
		for (int c = 0; c < n; c++)
			for (i64 k = hp[c]; k < hp[c + 1]; k++) {
				deg[c]++;
				out[hs[k]].push_back(c);
			}
		for (int c = 0; c < n; c++)
			if (deg[c] == 0)
				stack.push_back(c);
		int seen = 0;
		while (!stack.empty()) {
			int u = stack.back();
			stack.pop_back();
			seen++;
			for (int c : out[u])
				if (--deg[c] == 0)
					stack.push_back(c);
		}
		errx(1, "[spasm-b200] the pivots do not form a triangular system (%d of %d nodes scheduled on the device, %d on the host, "
		        "error flag %d): %s", h[3], n, seen, h[2], seen == n ? "internal error of the device scheduler" : "invalid U / qinv");
	}

	depgraph_finish_levels(G);
	/* dataflow schedule of the solve */
	G.pending0.alloc((size_t) n);
	G.seeds.alloc((size_t) n);
	CUDA_CHECK(cudaMemsetAsync(counters.ptr, 0, sizeof(int), s));
	k_flow_prepare<<<cdiv(n, 256), 256, 0, s>>>(n, G.ptr, G.src, G.level, G.pending0.ptr, G.seeds.ptr, counters.ptr);
	LAUNCHED(1);
	G.nseeds = fetch(counters.ptr);
	G.nscheduled = n - G.level_ptr_h[1];
	lap("sort + bounds");
	G.scheduled_deps = G.ndeps;
}

/* ------------------------------------------------------------------ the solve */

/* one work item = 4 consecutive right-hand sides (one 16-byte vector) of one scheduled column */
__device__ __forceinline__ void solve_item(const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
                                           int c, int r, int4 *X, int ld4, const Zp &F)
{
	const i64 e0 = ptr[c], e1 = ptr[c + 1];
	int4 *Xc = X + (size_t) c * ld4;
	int4 b = Xc[r];
	i64 a0 = b.x, a1 = b.y, a2 = b.z, a3 = b.w;
	int pending = 0;
	for (i64 e = e0; e < e1; e++) {
		const i64 v = val[e];
		const int4 xs = X[(size_t) src[e] * ld4 + r];
		a0 -= v * xs.x;
		a1 -= v * xs.y;
		a2 -= v * xs.z;
		a3 -= v * xs.w;
		if (++pending == F.delay) {
			a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
			pending = 0;
		}
	}
	b.x = zp_reduce(a0, F); b.y = zp_reduce(a1, F); b.z = zp_reduce(a2, F); b.w = zp_reduce(a3, F);
	Xc[r] = b;
}

/*
 * Persistent cooperative kernel, one CTA of 1024 threads per SM.
 * The schedule is a list of segments.  A WIDE segment is one level: its (column, vector) items are spread over
 * the whole grid and a grid barrier follows.  A NARROW segment is a run of consecutive levels that each hold
 * few columns (the long tail of the pivot DAG: thousands of levels of ~10 columns): CTA 0 walks the run alone,
 * separated by block barriers only, while the other CTAs wait at the single grid barrier that ends the run.
 */
struct SolveSegment { int level_begin, level_end, narrow; };

__global__ void __launch_bounds__(1024, 1)
k_panel_solve(const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
              const int *__restrict__ order, const int *__restrict__ level_ptr, const SolveSegment *__restrict__ segs, int nsegs,
              int4 *X, int ld4, int R4, Zp F)
{
	cg::grid_group grid = cg::this_grid();
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsize = gridDim.x * blockDim.x;
	for (int sgi = 0; sgi < nsegs; sgi++) {
		const SolveSegment sg = segs[sgi];
		if (!sg.narrow) {
			const int begin = level_ptr[sg.level_begin];
			const i64 items = (i64) (level_ptr[sg.level_begin + 1] - begin) * R4;
			for (i64 it = gtid; it < items; it += gsize)
				solve_item(ptr, src, val, order[begin + (int) (it / R4)], (int) (it % R4), X, ld4, F);
		} else if (blockIdx.x == 0) {
			for (int L = sg.level_begin; L < sg.level_end; L++) {
				const int begin = level_ptr[L];
				const int items = (level_ptr[L + 1] - begin) * R4;
				for (int it = threadIdx.x; it < items; it += blockDim.x)
					solve_item(ptr, src, val, order[begin + it / R4], it % R4, X, ld4, F);
				__syncthreads();
			}
		}
		grid.sync();
	}
}

/*
 * Dataflow variant: no barrier between levels.  One CTA owns one column at a time: it computes the column for
 * every right-hand side, publishes it (__threadfence), then releases the columns that depend on it by
 * decrementing their pending counters; a CTA that releases a column keeps it as its next job (a chain of the DAG
 * is walked by one CTA without a queue round trip), further released columns go to a shared queue that idle CTAs
 * poll.  Sources (level-0 columns) are final from the start and are never scheduled.  The values do not depend
 * on the schedule (each column is still written once, from final columns).
 */
#define FLOW_MAXD 8      /* dependents of a column whose metadata is prefetched */
#define FLOW_MAXE 8      /* dependencies per column kept in shared memory */

struct FlowMeta {
	int node;
	int cnt;             /* number of dependencies (may exceed FLOW_MAXE: the rest is read from global memory) */
	i64 e0;
	i64 rb, re;          /* its dependents: rdst[rb:re] */
	int src[FLOW_MAXE];
	i32 val[FLOW_MAXE];
};

__global__ void __launch_bounds__(1024)
k_panel_solve_flow(const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
                   const i64 *__restrict__ rptr, const int *__restrict__ rdst, const int *__restrict__ level,
                   int *pending, int *queue, int nseeds, int nscheduled, int *tail, int *ticket, int *done, int *error,
                   int4 *X, int ld4, int R4, Zp F, int *level_out)
{
	/* The critical path of a pass is a chain of the DAG walked by one CTA.  To keep a hop short, the metadata of the
	 * columns that depend on the current one (their dependency lists) is prefetched into shared memory while the
	 * current column is being computed: when one of them is released and becomes the next job, only the panel
	 * vectors themselves remain to be loaded. */
	__shared__ int s_node, s_next;
	__shared__ FlowMeta cur, dep[FLOW_MAXD];
	const int tid = threadIdx.x;
	int next_slot = -1;          /* >= 0: next job is dep[next_slot] */
	for (;;) {
		/* ---- 1. the job and its metadata */
		if (next_slot >= 0) {
			if (tid < FLOW_MAXE) {
				cur.src[tid] = dep[next_slot].src[tid];
				cur.val[tid] = dep[next_slot].val[tid];
			}
			if (tid == 0) {
				cur.node = dep[next_slot].node;
				cur.cnt = dep[next_slot].cnt;
				cur.e0 = dep[next_slot].e0;
				cur.rb = dep[next_slot].rb;
				cur.re = dep[next_slot].re;
			}
		} else {
			if (tid == 0) {
				const int t = atomicAdd(ticket, 1);
				int got = -2;
				int seen_done = -1;      /* progress-based watchdog: the budget restarts whenever a column completes */
				for (long spin = 0;; spin++) {
					if (t < nscheduled) {
						got = *((volatile int *) &queue[t]);
						if (got >= 0)
							break;
					}
					const int d_now = *((volatile int *) done);
					if (d_now >= nscheduled || *((volatile int *) error)) {
						got = -2;
						break;
					}
					if (d_now != seen_done) {
						seen_done = d_now;
						spin = 0;
					}
					if (spin > (1L << 24)) {
						*error = 1;
						got = -2;
						break;
					}
					__nanosleep(100);
				}
				s_node = got;
			}
			__syncthreads();
			const int c0 = s_node;
			if (c0 < 0)
				return;
			const i64 e0 = ptr[c0];
			const int cnt = (int) (ptr[c0 + 1] - e0);
			if (tid < FLOW_MAXE && tid < cnt) {
				cur.src[tid] = src[e0 + tid];
				cur.val[tid] = val[e0 + tid];
			}
			if (tid == 0) {
				cur.node = c0;
				cur.cnt = cnt;
				cur.e0 = e0;
				cur.rb = rptr[c0];
				cur.re = rptr[c0 + 1];
			}
		}
		__syncthreads();
		const int c = cur.node;
		const i64 rb = cur.rb, re = cur.re;
		/* ---- 2. the LAST WARP prefetches the metadata of the dependents (their dependency lists and their own lists
		 * of dependents): a chain of four dependent loads that used to be issued by threads that also compute the
		 * column, which put it in front of the column's own loads on the critical path of a hop */
		/* (when the batch is wider than the CTA every thread computes and the first 64 threads prefetch first) */
		const bool dedicated = R4 <= (int) blockDim.x - 32;
		const int T3 = dedicated ? blockDim.x - 32 : blockDim.x;          /* threads that compute the column */
		if (dedicated ? tid >= T3 : tid < FLOW_MAXD * FLOW_MAXE) {
			for (int item = dedicated ? tid - T3 : tid; item < FLOW_MAXD * FLOW_MAXE; item += dedicated ? 32 : FLOW_MAXD * FLOW_MAXE) {
				const int di = item / FLOW_MAXE, ei = item % FLOW_MAXE;
				if (rb + di < re) {
					const int d = rdst[rb + di];
					const i64 de0 = ptr[d];
					const int dcnt = (int) (ptr[d + 1] - de0);
					if (ei < dcnt) {
						dep[di].src[ei] = src[de0 + ei];
						dep[di].val[ei] = val[de0 + ei];
					}
					if (ei == 0) {
						dep[di].node = d;
						dep[di].cnt = dcnt;
						dep[di].e0 = de0;
						dep[di].rb = rptr[d];
						dep[di].re = rptr[d + 1];
					}
				}
			}
		}
		/* ---- 3. the column, for all right-hand sides; dependencies are read around L1 (written by other SMs) */
		if (tid < T3) {
			const int cnt = cur.cnt, cached = min(cnt, FLOW_MAXE);
			const i64 e0 = cur.e0;
			int4 *Xc = X + (size_t) c * ld4;
			for (int r = tid; r < R4; r += T3) {
				int4 b = Xc[r];
				i64 a0 = b.x, a1 = b.y, a2 = b.z, a3 = b.w;
				int pendingred = 0;
				int e = 0;
				/* four dependencies at a time: their 16-byte loads are independent and in flight together (wide
				 * batches are bound by the bytes in flight per SM, not by the arithmetic) */
				for (; F.delay >= 4 && e + 4 <= cnt; e += 4) {
					i64 v[4];
					int4 xs[4];
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const int sc = (e + u < cached) ? cur.src[e + u] : src[e0 + e + u];
						v[u] = (e + u < cached) ? cur.val[e + u] : val[e0 + e + u];
						xs[u] = __ldcg(&X[(size_t) sc * ld4 + r]);
					}
					if (pendingred + 4 > F.delay) {
						a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
						pendingred = 0;
					}
#pragma unroll
					for (int u = 0; u < 4; u++) {
						a0 -= v[u] * xs[u].x;
						a1 -= v[u] * xs[u].y;
						a2 -= v[u] * xs[u].z;
						a3 -= v[u] * xs[u].w;
					}
					pendingred += 4;
				}
				for (; e < cnt; e++) {
					const i64 v = (e < cached) ? cur.val[e] : val[e0 + e];
					const int sc = (e < cached) ? cur.src[e] : src[e0 + e];
					const int4 xs = __ldcg(&X[(size_t) sc * ld4 + r]);
					if (pendingred + 1 > F.delay) {
						a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
						pendingred = 0;
					}
					a0 -= v * xs.x;
					a1 -= v * xs.y;
					a2 -= v * xs.z;
					a3 -= v * xs.w;
					pendingred++;
				}
				b.x = zp_reduce(a0, F); b.y = zp_reduce(a1, F); b.z = zp_reduce(a2, F); b.w = zp_reduce(a3, F);
				Xc[r] = b;
			}
		}
		/* lazy schedule: the level of the column, 1 + max over its dependencies (all final: they released this
		 * column), is a by-product of the pass; published by the fence below with the column itself */
		if (level_out && tid == blockDim.x - 1) {
			const int cnt = cur.cnt;
			const i64 e0 = cur.e0;
			int lv = 0;
			for (int e = 0; e < cnt; e++)
				lv = max(lv, __ldcg(&level_out[(e < FLOW_MAXE) ? cur.src[e] : src[e0 + e]]));
			level_out[c] = lv + 1;
		}
		__threadfence();
		if (tid == 0)
			s_next = -1;
		__syncthreads();
		/* ---- 4. release the dependents; keep one of the prefetched ones as the next job */
		for (i64 k = rb + tid; k < re; k += blockDim.x) {
			const int di = (int) (k - rb);
			const int d = (di < FLOW_MAXD) ? dep[di].node : rdst[k];
			if (atomicSub(&pending[d], 1) == 1) {
				if (di >= FLOW_MAXD || atomicCAS(&s_next, -1, di) != -1) {
					const int pos = atomicAdd(tail, 1);
					__threadfence();
					*((volatile int *) &queue[pos]) = d;
				}
			}
		}
		__syncthreads();
		next_slot = s_next;
		if (tid == 0)
			atomicAdd(done, 1);
		__syncthreads();
	}
}

/*
 * Dataflow solve, second version: the hop of a chain without a global-memory round trip on its critical path.
 *
 * In k_panel_solve_flow a CTA that walks a chain c -> d pays, per hop: store X[c], __threadfence, atomicSub on the
 * pending counter of d (an L2 round trip), then the reload of X[c] it has just written.  Here:
 *   - REGISTER FORWARDING: thread r keeps X[c][r] in a register; the next column reads its dependency on c from there.
 *   - CERTAIN CONTINUATION: every column gets a `doneflag`, set by its CTA AFTER it has decremented the counters of all
 *     its dependents.  While c is being computed, a prefetch warp loads the dependency lists of c's dependents and
 *     checks the flags of their OTHER dependencies: if they are all set, every other decrement of pending[d] has
 *     happened, c's own would be the last one, so this CTA owns d -- it goes on with d at once, without the atomic.
 *   - DEFERRED PUBLICATION: the fence, the decrements of the other dependents, the flag and the `done` count of c are
 *     issued by a publisher warp during the computation of d (two service warps per CTA; the metadata buffers are
 *     double-buffered so the prefetch of d's dependents does not overwrite what the publication of c still reads).
 * When no dependent is certain the CTA publishes at once and continues with a dependent its decrements released
 * (as before), or polls the queue.  Each column is still written once, from final columns: same values.
 * Needs at most FLOW2_G right-hand-side groups per compute thread (R4 <= 2 (blockDim.x - 64)); wider batches take the first version.
 */
struct FlowMeta2 {             /* 24 ints = 96 bytes: one dependent of a column in the per-column blob (see k_flow2_blob) */
	int node, cnt, ready, pad;
	i64 e0, rb, re;
	int src[FLOW_MAXE];
	i32 val[FLOW_MAXE];
};
static_assert(sizeof(FlowMeta2) == 88 + 8 || sizeof(FlowMeta2) == 104, "FlowMeta2 layout");
#define FLOW2_WORDS ((int) (sizeof(FlowMeta2) / sizeof(int)))
#define FLOW2_G 2               /* groups of four right-hand sides per compute thread: batches of up to 7680 */

/* blob[c] = the FlowMeta2 records of the first FLOW_MAXD dependents of column c (their dependency lists and their own
 * ranges of dependents), contiguous: the prefetch of a hop is ONE coalesced copy instead of a chain of four dependent
 * loads (rdst -> ptr/rptr -> src/val) */
__global__ void k_flow2_blob(int n, const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
                             const i64 *__restrict__ rptr, const int *__restrict__ rdst, FlowMeta2 *blob)
{
	const i64 item = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	const int c = (int) (item / FLOW_MAXD), di = (int) (item % FLOW_MAXD);
	if (c >= n)
		return;
	FlowMeta2 m;
	m.node = -1; m.cnt = 0; m.ready = 0; m.pad = 0; m.e0 = 0; m.rb = 0; m.re = 0;
#pragma unroll
	for (int e = 0; e < FLOW_MAXE; e++) {
		m.src[e] = -1;
		m.val[e] = 0;
	}
	const i64 rb = rptr[c], re = rptr[c + 1];
	if (rb + di < re) {
		const int d = rdst[rb + di];
		m.node = d;
		m.e0 = ptr[d];
		m.cnt = (int) (ptr[d + 1] - m.e0);
		m.rb = rptr[d];
		m.re = rptr[d + 1];
		for (int e = 0; e < FLOW_MAXE && e < m.cnt; e++) {
			m.src[e] = src[m.e0 + e];
			m.val[e] = val[m.e0 + e];
		}
	}
	blob[(size_t) c * FLOW_MAXD + di] = m;
	if (di == 0) {
		/* the column's OWN record (what a CTA needs when it takes the column from the queue), after the dependents' blobs */
		FlowMeta2 o;
		o.node = c; o.ready = 0; o.pad = 0;
		o.e0 = ptr[c];
		o.cnt = (int) (ptr[c + 1] - o.e0);
		o.rb = rb;
		o.re = re;
#pragma unroll
		for (int e = 0; e < FLOW_MAXE; e++) {
			o.src[e] = (e < o.cnt) ? src[o.e0 + e] : -1;
			o.val[e] = (e < o.cnt) ? val[o.e0 + e] : 0;
		}
		blob[(size_t) n * FLOW_MAXD + c] = o;
	}
}

struct FlowPub {
	int valid, node, buf, skip;
	i64 rb, re;
};

__device__ __forceinline__ void flow2_publish(const FlowPub &pb, const FlowMeta2 *depbuf, const int *__restrict__ rdst, int *pending,
                                              int *queue, int *tail, int *done, int *doneflag, int *s_next, bool capture, int lane)
{
	__threadfence();                 /* the column was stored by the compute threads before the CTA barrier that precedes this */
	for (i64 k = pb.rb + lane; k < pb.re; k += 32) {
		const int di = (int) (k - pb.rb);
		if (di == pb.skip)
			continue;                /* the dependent this CTA went on with: its counter is left alone */
		const int d = (di < FLOW_MAXD) ? depbuf[di].node : rdst[k];
		if (atomicSub(&pending[d], 1) == 1) {
			if (!(capture && di < FLOW_MAXD && atomicCAS(s_next, -1, di) == -1)) {
				const int pos = atomicAdd(tail, 1);
				__threadfence();
				*((volatile int *) &queue[pos]) = d;
			}
		}
	}
	__syncwarp();
	__threadfence();                 /* the decrements before the flag: a set flag means "my decrements are done" */
	if (lane == 0) {
		*((volatile int *) &doneflag[pb.node]) = 1;
		atomicAdd(done, 1);
	}
}

/* 1-D bulk copy (the TMA engine without a tensor map) of a column's blob of dependents' records into shared memory:
 * one request of 832 bytes instead of 208 word loads by the prefetch warp, completion on an mbarrier */
__device__ __forceinline__ uint32_t flow_smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ bool flow_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
	uint32_t ok;
	asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
	             : "=r"(ok) : "r"(flow_smem_u32(bar)), "r"(parity) : "memory");
	return ok != 0;
}

__global__ void __launch_bounds__(1024)
k_panel_solve_flow2(const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
                    const i64 *__restrict__ rptr, const int *__restrict__ rdst,
                    int *pending, int *queue, int nscheduled, int *tail, int *ticket, int *done, int *error, int *doneflag,
                    int4 *X, int ld4, int R4, Zp F, int *level_out, unsigned long long *hopstats, const FlowMeta2 *__restrict__ blob, int nnodes,
                    int bulk_blob)
{
	__shared__ int s_node, s_next, s_capt;
	__shared__ FlowMeta2 cur;
	__shared__ __align__(16) FlowMeta2 dep[2][FLOW_MAXD];
	__shared__ __align__(8) uint64_t dep_bar;
	uint32_t dep_phase = 0;               /* prefetch warp: parity of the next completion of dep_bar */
	unsigned n_certain = 0, n_released = 0, n_polled = 0;      /* how the CTA came by its columns (thread 0; development) */
	__shared__ FlowPub pub;
	const int tid = threadIdx.x, lane = tid & 31;
	const int T3 = blockDim.x - 64;                       /* compute threads */
	const bool is_prefetch = tid >= T3 && tid < T3 + 32;
	const bool is_publish = tid >= T3 + 32;
	int buf = 0;
	bool have_cur = false;
	int fwd_col = -1;
	int4 fwd[FLOW2_G];
#pragma unroll
	for (int g = 0; g < FLOW2_G; g++)
		fwd[g] = make_int4(0, 0, 0, 0);
	if (tid == 0) {
		pub.valid = 0;
		asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(flow_smem_u32(&dep_bar)), "r"(1));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	for (;;) {
		/* ---- 1. the job */
		if (!have_cur) {
			if (tid == 0) {
				const int t = atomicAdd(ticket, 1);
				int got = -2;
				int seen_done = -1;      /* progress-based watchdog */
				for (long spin = 0;; spin++) {
					if (t < nscheduled) {
						got = *((volatile int *) &queue[t]);
						if (got >= 0)
							break;
					}
					const int d_now = *((volatile int *) done);
					if (d_now >= nscheduled || *((volatile int *) error)) {
						got = -2;
						break;
					}
					if (d_now != seen_done) {
						seen_done = d_now;
						spin = 0;
					}
					if (spin > (1L << 24)) {
						*error = 1;
						got = -2;
						break;
					}
					__nanosleep(100);
				}
				__threadfence();
				s_node = got;
			}
			__syncthreads();
			const int c0 = s_node;
			if (c0 < 0) {
				if (tid == 0 && hopstats) {
					atomicAdd(&hopstats[0], (unsigned long long) n_certain);
					atomicAdd(&hopstats[1], (unsigned long long) n_released);
					atomicAdd(&hopstats[2], (unsigned long long) n_polled);
				}
				return;
			}
			n_polled++;
			{
				/* one coalesced load of the column's own record instead of ptr -> src/val -> rptr */
				const int *gsrc = reinterpret_cast<const int *>(blob + (size_t) nnodes * FLOW_MAXD + c0);
				int *gdst = reinterpret_cast<int *>(&cur);
				if (tid < FLOW2_WORDS)
					gdst[tid] = gsrc[tid];
			}
			fwd_col = -1;
		}
		__syncthreads();                                  /* [A] cur is ready */
		const int c = cur.node;
		const i64 rb = cur.rb, re = cur.re;
		if (tid < T3) {
			/* ---- 2a. the column: up to FLOW2_G groups of four right-hand sides per thread, each kept in a register for
			 * the next hop */
#pragma unroll
			for (int g = 0; g < FLOW2_G; g++) {
				const int r = tid + g * T3;
				if (r >= R4)
					break;
				const int cnt = cur.cnt, cached = min(cnt, FLOW_MAXE);
				const i64 e0 = cur.e0;
				int4 *Xc = X + (size_t) c * ld4;
				int4 b = Xc[r];
				i64 a0 = b.x, a1 = b.y, a2 = b.z, a3 = b.w;
				int pendingred = 0;
				int e = 0;
				for (; F.delay >= 4 && e + 4 <= cnt; e += 4) {
					i64 v[4];
					int4 xs[4];
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const int sc = (e + u < cached) ? cur.src[e + u] : src[e0 + e + u];
						v[u] = (e + u < cached) ? cur.val[e + u] : val[e0 + e + u];
						xs[u] = (sc == fwd_col) ? fwd[g] : __ldcg(&X[(size_t) sc * ld4 + r]);
					}
					if (pendingred + 4 > F.delay) {
						a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
						pendingred = 0;
					}
#pragma unroll
					for (int u = 0; u < 4; u++) {
						a0 -= v[u] * xs[u].x;
						a1 -= v[u] * xs[u].y;
						a2 -= v[u] * xs[u].z;
						a3 -= v[u] * xs[u].w;
					}
					pendingred += 4;
				}
				for (; e < cnt; e++) {
					const i64 v = (e < cached) ? cur.val[e] : val[e0 + e];
					const int sc = (e < cached) ? cur.src[e] : src[e0 + e];
					const int4 xs = (sc == fwd_col) ? fwd[g] : __ldcg(&X[(size_t) sc * ld4 + r]);
					if (pendingred + 1 > F.delay) {
						a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
						pendingred = 0;
					}
					a0 -= v * xs.x;
					a1 -= v * xs.y;
					a2 -= v * xs.z;
					a3 -= v * xs.w;
					pendingred++;
				}
				b.x = zp_reduce(a0, F); b.y = zp_reduce(a1, F); b.z = zp_reduce(a2, F); b.w = zp_reduce(a3, F);
				Xc[r] = b;
				fwd[g] = b;
			}
			fwd_col = c;
			/* lazy schedule: the level of the column is a by-product of the pass (its dependencies are final) */
			if (level_out && tid == T3 - 1) {
				const int cnt = cur.cnt;
				const i64 e0 = cur.e0;
				int lv = 0;
				for (int e = 0; e < cnt; e++)
					lv = max(lv, __ldcg(&level_out[(e < FLOW_MAXE) ? cur.src[e] : src[e0 + e]]));
				level_out[c] = lv + 1;
			}
		} else if (is_prefetch) {
			/* ---- 2b. metadata of the dependents of c (one coalesced copy of the column's blob), and whether c is the
			 * last dependency they wait for (the flags of their other dependencies) */
			FlowMeta2 *dp = dep[buf];
			if (bulk_blob) {
				/* U's entries for the next hop, staged by the TMA engine: cp.async.bulk global -> shared, mbarrier expect_tx.
				 * The buffer was last read (generic proxy) two hops ago, before several CTA barriers; the proxy fence orders
				 * those reads before the asynchronous write. */
				if (lane == 0) {
					const uint32_t bytes = (uint32_t) (FLOW_MAXD * sizeof(FlowMeta2));
					asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
					asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(flow_smem_u32(&dep_bar)), "r"(bytes) : "memory");
					asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					             ::"r"(flow_smem_u32(dp)), "l"(blob + (size_t) c * FLOW_MAXD), "r"(bytes), "r"(flow_smem_u32(&dep_bar)) : "memory");
				}
				__syncwarp();
				bool arrived = false;
				for (int spin = 0; spin < (1 << 24) && !(arrived = flow_mbar_try_wait(&dep_bar, dep_phase)); spin++)
					;
				if (!arrived && lane == 0)
					*error = 1;                   /* never seen; the host reports an internal error instead of hanging */
				dep_phase ^= 1;
			} else {
				const int *gsrc = reinterpret_cast<const int *>(blob + (size_t) c * FLOW_MAXD);
				int *gdst = reinterpret_cast<int *>(dp);
#pragma unroll
				for (int w = lane; w < FLOW_MAXD * FLOW2_WORDS; w += 32)
					gdst[w] = gsrc[w];
			}
			__syncwarp();
			/* warm the L2 for the next hop: the right-hand sides and the blob of every dependent (the panel and the blob
			 * array are far larger than the L2, a cold column costs a DRAM round trip on the critical path) */
			for (int di = 0; di < FLOW_MAXD; di++) {
				const int d = dp[di].node;
				if (d < 0)
					break;
				const char *xrow = reinterpret_cast<const char *>(X + (size_t) d * ld4);
				for (int off = lane * 128; off < R4 * 16; off += 32 * 128)
					asm volatile("prefetch.global.L2 [%0];" ::"l"(xrow + off));
				const char *brow = reinterpret_cast<const char *>(blob + (size_t) d * FLOW_MAXD);
				if (lane * 128 < (int) (FLOW_MAXD * sizeof(FlowMeta2)))
					asm volatile("prefetch.global.L2 [%0];" ::"l"(brow + lane * 128));
			}
			for (int pass = 0; pass < (FLOW_MAXD * FLOW_MAXE) / 32; pass++) {
				const int item = pass * 32 + lane;
				const int di = item / FLOW_MAXE, ei = item % FLOW_MAXE;
				bool ok = true;                       /* this dependency does not stand in the way */
				const int dcnt = dp[di].cnt;
				if (dp[di].node >= 0) {
					if (ei < dcnt) {
						const int sc = dp[di].src[ei];
						ok = (sc == c) || (*((volatile int *) &doneflag[sc]) != 0);
					}
				} else {
					ok = false;
				}
				const unsigned okmask = __ballot_sync(0xffffffffu, ok);
				if (ei == 0) {
					const unsigned mine = (okmask >> (lane & ~(FLOW_MAXE - 1))) & ((1u << FLOW_MAXE) - 1);
					dp[di].ready = dp[di].node >= 0 && dcnt <= FLOW_MAXE && mine == ((1u << FLOW_MAXE) - 1);
				}
			}
			__threadfence();                          /* flags read before the panel vectors of those columns are */
		} else if (is_publish) {
			/* ---- 2c. publication of the previous column of the chain (deferred by one hop) */
			if (pub.valid)
				flow2_publish(pub, dep[pub.buf], rdst, pending, queue, tail, done, doneflag, &s_capt, false, lane);
		}
		__syncthreads();                                  /* [B] */
		/* ---- 3. go on with a dependent that is certain to be released by c, if there is one */
		if (tid == 0) {
			int pick = -1;
			const int nd = (int) min((i64) FLOW_MAXD, re - rb);
			for (int di = 0; di < nd && pick < 0; di++)
				if (dep[buf][di].ready)
					pick = di;
			s_next = pick;
			s_capt = -1;                     /* written by the capture of an immediate publication only, read after its barrier */
			pub.valid = pick >= 0;
			pub.node = c;
			pub.buf = buf;
			pub.skip = pick;
			pub.rb = rb;
			pub.re = re;
		}
		__syncthreads();                                  /* [C] */
		int nxt = s_next;
		if (nxt >= 0)
			n_certain++;
		if (nxt < 0) {
			/* nobody is certain: publish now, keep a dependent the decrements release (if any) */
			if (is_publish) {
				FlowPub now;
				now.valid = 1; now.node = c; now.buf = buf; now.skip = -1; now.rb = rb; now.re = re;
				flow2_publish(now, dep[buf], rdst, pending, queue, tail, done, doneflag, &s_capt, true, lane);
			}
			__syncthreads();
			nxt = s_capt;
			if (nxt >= 0)
				n_released++;
		}
		if (nxt >= 0) {
			const FlowMeta2 *nx = &dep[buf][nxt];
			if (tid < FLOW_MAXE) {
				cur.src[tid] = nx->src[tid];
				cur.val[tid] = nx->val[tid];
			}
			if (tid == 0) {
				cur.node = nx->node;
				cur.cnt = nx->cnt;
				cur.e0 = nx->e0;
				cur.rb = nx->rb;
				cur.re = nx->re;
			}
			have_cur = true;
		} else {
			have_cur = false;
		}
		buf ^= 1;
	}
}

/* Dataflow solve of a very sparse batch (spasm_rref: a row of the RREF touches a fraction of a percent of the
 * columns).  Same schedule as k_panel_solve_flow; per column the CTA first ORs the occupancy masks of the column and
 * of its dependencies (mw words: 128 right-hand sides per word), lists the marked groups in shared memory and computes
 * those only.  The traffic is the masks (1/128 of the panel) plus the marked groups. */
__global__ void __launch_bounds__(1024)
k_panel_solve_flow_masked(const i64 *__restrict__ ptr, const int *__restrict__ src, const i32 *__restrict__ val,
                          const i64 *__restrict__ rptr, const int *__restrict__ rdst, int *pending, int *queue, int nscheduled,
                          int *tail, int *ticket, int *done, int *error, int4 *X, int ld4, int R4, unsigned *mask, int mw, Zp F)
{
	extern __shared__ int s_list[];      /* marked groups of the current column (at most R4) */
	__shared__ int s_node, s_next, s_count;
	const int tid = threadIdx.x, T = blockDim.x;
	int next = -1;
	for (;;) {
		if (tid == 0) {
			int got = next;
			if (got < 0) {
				const int t = atomicAdd(ticket, 1);
				got = -2;
				int seen_done = -1;      /* progress-based watchdog: the budget restarts whenever a column completes */
				for (long spin = 0;; spin++) {
					if (t < nscheduled) {
						got = *((volatile int *) &queue[t]);
						if (got >= 0)
							break;
					}
					const int d_now = *((volatile int *) done);
					if (d_now >= nscheduled || *((volatile int *) error)) {
						got = -2;
						break;
					}
					if (d_now != seen_done) {
						seen_done = d_now;
						spin = 0;
					}
					if (spin > (1L << 24)) {
						*error = 1;
						got = -2;
						break;
					}
					__nanosleep(100);
				}
			}
			s_node = got;
			s_count = 0;
			s_next = -1;
		}
		__syncthreads();
		const int c = s_node;
		if (c < 0)
			return;
		const i64 e0 = ptr[c];
		const int cnt = (int) (ptr[c + 1] - e0);
		/* ---- occupancy of the column: its right-hand sides or any dependency */
		for (int w = tid; w < mw; w += T) {
			unsigned m = mask[(size_t) c * mw + w];
			for (int e = 0; e < cnt; e++)
				m |= __ldcg(&mask[(size_t) src[e0 + e] * mw + w]);
			mask[(size_t) c * mw + w] = m;
			while (m) {
				const int b = __ffs(m) - 1;
				m &= m - 1;
				s_list[atomicAdd(&s_count, 1)] = w * 32 + b;
			}
		}
		__syncthreads();
		const int ngroups = s_count;
		/* ---- the marked groups; dependencies are read around L1 (written by other SMs) */
		int4 *Xc = X + (size_t) c * ld4;
		for (int t = tid; t < ngroups; t += T) {
			const int g = s_list[t];
			int4 b = Xc[g];
			i64 a0 = b.x, a1 = b.y, a2 = b.z, a3 = b.w;
			int pendingred = 0;
			for (int e = 0; e < cnt; e++) {
				const i64 v = val[e0 + e];
				const int4 xs = __ldcg(&X[(size_t) src[e0 + e] * ld4 + g]);
				a0 -= v * xs.x;
				a1 -= v * xs.y;
				a2 -= v * xs.z;
				a3 -= v * xs.w;
				if (++pendingred == F.delay) {
					a0 = zp_reduce(a0, F); a1 = zp_reduce(a1, F); a2 = zp_reduce(a2, F); a3 = zp_reduce(a3, F);
					pendingred = 0;
				}
			}
			b.x = zp_reduce(a0, F); b.y = zp_reduce(a1, F); b.z = zp_reduce(a2, F); b.w = zp_reduce(a3, F);
			Xc[g] = b;
		}
		__threadfence();
		__syncthreads();
		/* ---- release the dependents; keep one as the next job of this CTA */
		const i64 rb = rptr[c], re = rptr[c + 1];
		for (i64 k = rb + tid; k < re; k += T) {
			const int d = rdst[k];
			if (atomicSub(&pending[d], 1) == 1) {
				if (atomicCAS(&s_next, -1, d) != -1) {
					const int pos = atomicAdd(tail, 1);
					__threadfence();
					*((volatile int *) &queue[pos]) = d;
				}
			}
		}
		__syncthreads();
		next = s_next;
		if (tid == 0)
			atomicAdd(done, 1);
		__syncthreads();
	}
}

void panel_solve_masked(const DepGraph &G, i32 *X, int ld, int R, unsigned *mask, int mw, const Zp &F)
{
	if (G.nlevels <= 1 || R <= 0 || G.nscheduled <= 0)
		return;
	if (!G.levels_known)
		errx(1, "[spasm-b200] internal: the masked solve needs an eagerly scheduled graph");
	if (ld % 4 != 0)
		errx(1, "[spasm-b200] internal: panel leading dimension must be a multiple of 4");
	int R4 = (R + 3) / 4, ld4 = ld / 4;
	cudaStream_t s = ctx().stream;
	int n = G.nnodes;
	DevBuf<int> pending((size_t) n), queue((size_t) n + 1), counters(4);
	CUDA_CHECK(cudaMemcpyAsync(pending.ptr, G.pending0.ptr, (size_t) n * sizeof(int), cudaMemcpyDeviceToDevice, s));
	queue.fill_byte(0xff, s);
	CUDA_CHECK(cudaMemcpyAsync(queue.ptr, G.seeds.ptr, (size_t) G.nseeds * sizeof(int), cudaMemcpyDeviceToDevice, s));
	int h_init[4] = {G.nseeds, 0, 0, 0};      /* tail, ticket, done, error */
	CUDA_CHECK(cudaMemcpyAsync(counters.ptr, h_init, sizeof(h_init), cudaMemcpyHostToDevice, s));
	const int threads = 256;
	size_t smem = (size_t) mw * 32 * sizeof(int);
	CUDA_CHECK(cudaFuncSetAttribute(k_panel_solve_flow_masked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int occ = 0;
	CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_panel_solve_flow_masked, threads, smem));
	if (occ < 1)
		errx(1, "[spasm-b200] internal: masked solve kernel does not fit (%zu bytes of shared memory)", smem);
	int blocks = std::min(occ, 8) * ctx().sm_count;      /* co-resident: idle CTAs poll the queue */
	GpuTimer tk;
	tk.start();
	k_panel_solve_flow_masked<<<blocks, threads, smem, s>>>(G.ptr.ptr, G.src.ptr, G.val.ptr, G.rptr.ptr, G.rdst.ptr, pending.ptr, queue.ptr,
	                                                       G.nscheduled, counters.ptr, counters.ptr + 1, counters.ptr + 2, counters.ptr + 3,
	                                                       (int4 *) X, ld4, R4, mask, mw, F);
	LAUNCHED(1);
	KERNEL_CHECK();
	stats().pub.ms_k_panel_solve += tk.stop_ms();
	int h[4];
	CUDA_CHECK(cudaMemcpyAsync(h, counters.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
	sync();
	if (h[3] != 0 || h[2] != G.nscheduled)
		errx(1, "[spasm-b200] internal: masked dataflow solve did not complete (%d of %d columns)", h[2], G.nscheduled);
	Stats &st = stats();
	st.pub.solve_batches += 1;
	st.pub.solve_rows += R;
}

/* columns without dependency are final from the start */
__global__ void k_flow2_flags(int n, const i64 *__restrict__ ptr, int *doneflag)
{
	int c = blockIdx.x * blockDim.x + threadIdx.x;
	if (c < n)
		doneflag[c] = (ptr[c + 1] == ptr[c]);
}

void panel_solve(const DepGraph &G, i32 *X, int ld, int R, const Zp &F)
{
	if (G.nlevels <= 1 || R <= 0)
		return;
	if (ld % 4 != 0)
		errx(1, "[spasm-b200] internal: panel leading dimension must be a multiple of 4");
	int R4 = (R + 3) / 4, ld4 = ld / 4;
	cudaStream_t s = ctx().stream;
	static const bool use_levels = getenv("SPASM_B200_SOLVE_LEVELS") != NULL;
	if ((!use_levels || !G.levels_known) && G.nscheduled > 0) {
		int n = G.nnodes;
		DevBuf<int> pending((size_t) n), queue((size_t) n + 1), counters(4);
		CUDA_CHECK(cudaMemcpyAsync(pending.ptr, G.pending0.ptr, (size_t) n * sizeof(int), cudaMemcpyDeviceToDevice, s));
		queue.fill_byte(0xff, s);
		CUDA_CHECK(cudaMemcpyAsync(queue.ptr, G.seeds.ptr, (size_t) G.nseeds * sizeof(int), cudaMemcpyDeviceToDevice, s));
		int h_init[4] = {G.nseeds, 0, 0, 0};      /* tail, ticket, done, error */
		CUDA_CHECK(cudaMemcpyAsync(counters.ptr, h_init, sizeof(h_init), cudaMemcpyHostToDevice, s));
		/* one pass over the right-hand sides per column when possible: the column is the unit of the critical path */
		int threads = getenv("SPASM_B200_FLOW_THREADS") ? atoi(getenv("SPASM_B200_FLOW_THREADS")) : (R4 > 480 ? 1024 : R4 > 224 ? 512 : 256);      /* one warp of the CTA is the prefetcher when the batch fits */
		threads = std::max(64, std::min(1024, threads & ~31));
		static const bool flow1 = getenv("SPASM_B200_FLOW1") != NULL;
		const bool flow2 = !flow1 && R4 <= FLOW2_G * (1024 - 64);      /* the groups of a thread stay in registers: register forwarding */
		GpuTimer tk;
		if (flow2) {
			threads = R4 + 64 > 512 ? 1024 : R4 + 64 > 256 ? 512 : 256;
			DevBuf<int> doneflag((size_t) n);
			static const bool hoptrace = getenv("SPASM_B200_TRACE") != NULL;
			static const bool no_bulk = getenv("SPASM_B200_FLOW2_NO_BULK") != NULL;      /* blob copied by the prefetch warp's own loads */
			DevBuf<unsigned long long> hopstats(3);
			hopstats.zero(s);
			DevBuf<char> blob(((size_t) n * FLOW_MAXD + (size_t) n) * sizeof(FlowMeta2));
			k_flow2_blob<<<cdiv((size_t) n * FLOW_MAXD, 256), 256, 0, s>>>(n, G.ptr.ptr, G.src.ptr, G.val.ptr, G.rptr.ptr, G.rdst.ptr, (FlowMeta2 *) blob.ptr);
			LAUNCHED(1);
			k_flow2_flags<<<cdiv(n, 256), 256, 0, s>>>(n, G.ptr.ptr, doneflag.ptr);
			int occ = 0;
			CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_panel_solve_flow2, threads, 0));
			int blocks = std::max(1, std::min(occ, 8)) * ctx().sm_count;      /* co-resident: idle CTAs poll the queue */
			tk.start();
			k_panel_solve_flow2<<<blocks, threads, 0, s>>>(G.ptr.ptr, G.src.ptr, G.val.ptr, G.rptr.ptr, G.rdst.ptr, pending.ptr, queue.ptr,
			                                          G.nscheduled, counters.ptr, counters.ptr + 1, counters.ptr + 2, counters.ptr + 3, doneflag.ptr,
			                                          (int4 *) X, ld4, R4, F, G.levels_known ? nullptr : G.level.ptr, hoptrace ? hopstats.ptr : nullptr,
			                                          (const FlowMeta2 *) blob.ptr, n, no_bulk ? 0 : 1);
			LAUNCHED(2);
			KERNEL_CHECK();
			stats().pub.ms_k_panel_solve += tk.stop_ms();      /* before doneflag goes out of scope: the stop synchronises */
			if (hoptrace) {
				unsigned long long h3[3];
				CUDA_CHECK(cudaMemcpyAsync(h3, hopstats.ptr, sizeof(h3), cudaMemcpyDeviceToHost, s));
				sync();
				fprintf(stderr, "[trace]     dataflow solve: %d columns, R = %d; continued with a certain dependent %llu times, with a released one %llu, polled the queue %llu\n",
				        G.nscheduled, R, h3[0], h3[1], h3[2]);
			}
		} else {
		int occ = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_panel_solve_flow, threads, 0));
		int blocks = std::max(1, std::min(occ, 8)) * ctx().sm_count;      /* co-resident: idle CTAs poll the queue */
		tk.start();
		k_panel_solve_flow<<<blocks, threads, 0, s>>>(G.ptr.ptr, G.src.ptr, G.val.ptr, G.rptr.ptr, G.rdst.ptr, G.level.ptr, pending.ptr, queue.ptr,
		                                         G.nseeds, G.nscheduled, counters.ptr, counters.ptr + 1, counters.ptr + 2, counters.ptr + 3,
		                                         (int4 *) X, ld4, R4, F, G.levels_known ? nullptr : G.level.ptr);
		LAUNCHED(1);
		KERNEL_CHECK();
		stats().pub.ms_k_panel_solve += tk.stop_ms();
		}
		int h[4];
		CUDA_CHECK(cudaMemcpyAsync(h, counters.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
		sync();
		if (h[3] != 0 || h[2] != G.nscheduled) {
			if (!G.levels_known)      /* lazy schedule: this pass is also the cycle check */
				errx(1, "[spasm-b200] the pivots do not form a triangular system (%d of %d dependent columns solved): invalid U / qinv",
				     h[2], G.nscheduled);
			errx(1, "[spasm-b200] internal: dataflow solve did not complete (%d of %d columns)", h[2], G.nscheduled);
		}
		if (!G.levels_known) {
			/* the graph is logically const for a solve; completing its lazily built schedule is the one exception */
			DepGraph &Gm = const_cast<DepGraph &>(G);
			depgraph_finish_levels(Gm);
			stats().pub.dag_depth = Gm.nlevels;
		}
		Stats &st = stats();
		st.pub.solve_batches += 1;
		st.pub.solve_rows += R;
		st.pub.solve_traffic_model += ((double) G.ndeps + 2.0 * G.nscheduled) * 4.0 * R;
		return;
	}
	/* segments: a level is narrow when one CTA covers it in at most two passes */
	std::vector<SolveSegment> segs;
	const int narrow_items = getenv("SPASM_B200_NARROW_ITEMS") ? atoi(getenv("SPASM_B200_NARROW_ITEMS")) : 2048;
	for (int L = 1; L < G.nlevels;) {
		i64 items = (i64) (G.level_ptr_h[L + 1] - G.level_ptr_h[L]) * R4;
		if (items > narrow_items) {
			segs.push_back({L, L + 1, 0});
			L++;
		} else {
			int E = L;
			while (E < G.nlevels && (i64) (G.level_ptr_h[E + 1] - G.level_ptr_h[E]) * R4 <= narrow_items)
				E++;
			segs.push_back({L, E, 1});
			L = E;
		}
	}
	DevBuf<SolveSegment> d_segs;
	d_segs.upload(segs.data(), segs.size(), s);
	int blocks = coop_blocks(k_panel_solve, 1024);
	blocks = std::min(blocks, ctx().sm_count);
	const i64 *a_ptr = G.ptr.ptr;
	const int *a_src = G.src.ptr;
	const i32 *a_val = G.val.ptr;
	const int *a_order = G.order.ptr, *a_lp = G.level_ptr.ptr;
	const SolveSegment *a_segs = d_segs.ptr;
	int nsegs = (int) segs.size();
	int4 *a_X = (int4 *) X;
	Zp Fc = F;
	void *args[] = {&a_ptr, &a_src, &a_val, &a_order, &a_lp, &a_segs, &nsegs, &a_X, &ld4, &R4, &Fc};
	GpuTimer tk;
	tk.start();
	CUDA_CHECK(cudaLaunchCooperativeKernel((void *) k_panel_solve, dim3(blocks), dim3(1024), args, 0, s));
	LAUNCHED(1);
	stats().pub.ms_k_panel_solve += tk.stop_ms();
	Stats &st = stats();
	st.pub.solve_batches += 1;
	st.pub.solve_rows += R;
	/* what this formulation must move: one panel vector read per dependency, one read+write per scheduled column */
	i64 scheduled = (i64) G.nnodes - (G.level_ptr_h.size() > 1 ? G.level_ptr_h[1] : 0);
	st.pub.solve_traffic_model += ((double) G.ndeps + 2.0 * scheduled) * 4.0 * R;
}

}  // namespace sb
