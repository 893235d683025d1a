/*
 * Structural pivot search on the GPU (reference: src/spasm_pivots.c).
 *
 * The three searches of the reference are sequential scans whose result depends
 * on the order in which rows are visited.  Each kernel family below computes
 * exactly the result of the reference run with ONE thread (rows in increasing
 * order), with a parallel schedule:
 *
 *  - Faugere-Lachartre (pivots.c:41-66): per column, the row whose leftmost entry
 *    is that column, of minimum weight, lowest index on ties  ==  a 64-bit
 *    atomicMin on (weight, row) keys.
 *  - FL on columns (pivots.c:76-122): a lexicographically-first greedy.  Rounds of
 *    "deterministic reservations": every undecided row reserves its open columns
 *    with atomicMin(row index); a row whose open columns are all reserved by
 *    itself can decide now (no lower row can close any of them); it takes its
 *    first open entry and closes its columns.  The lowest undecided row always
 *    decides, so the rounds terminate, and the outcome is the sequential one.
 *  - greedy alternating-cycle-free search (pivots.c:146-294): the reference's own
 *    transactional scheme (journal of new pivots, replay on commit failure), with
 *    the commit section taken in increasing row order.  Each CTA owns one
 *    candidate row at a time, runs the BFS with a visited bitmap in shared memory
 *    against the live pivot set, and when it holds a survivor waits until every
 *    lower row is resolved, replays the journal entries it has not seen
 *    (pivots.c:260-274) and commits.  Rows that exhaust their survivors are
 *    resolved at once: reachability only grows with the pivot set.
 *
 * Pivotal rows are then ordered by the levels of the pivot DAG (a valid
 * triangular order; the reference uses a DFS order, pivots.c:325-362 -- the row
 * order inside U is not part of the canonical result) and copied
 * to U with the pivot first and scaled to 1 (pivots.c:405-443).
 */
#include <cub/cub.cuh>
#include "pivots.cuh"
#include "stats.cuh"

namespace sb {

/* ============================================================ Faugere-Lachartre */

__global__ void k_fl_keys(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, unsigned long long *key)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	i64 b = Ap[i], e = Ap[i + 1];
	if (b == e)
		return;
	int lead = Aj[b];
	for (i64 k = b + 1; k < e; k++)
		lead = min(lead, Aj[k]);
	unsigned long long mine = ((unsigned long long) (e - b) << 32) | (unsigned) i;
	atomicMin(&key[lead], mine);
}

__global__ void k_fl_assign(int m, const unsigned long long *__restrict__ key, int *pinv, int *qinv, int *count)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m)
		return;
	unsigned long long k = key[j];
	if (k == ~0ull)
		return;
	int i = (int) (k & 0xffffffffull);
	qinv[j] = i;
	pinv[i] = j;
	atomicAdd(count, 1);
}

/* ============================================================ FL on columns */

__global__ void k_flc_close_pivotal(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, const int *__restrict__ pinv,
                                    unsigned char *open, unsigned char *decided)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	if (pinv[i] >= 0) {
		decided[i] = 1;
		for (i64 k = Ap[i]; k < Ap[i + 1]; k++)
			open[Aj[k]] = 0;
	}
}

/* undecided rows reserve their open columns; rows without any are decided (no pivot) */
__global__ void k_flc_reserve(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, const int *__restrict__ qinv,
                              const unsigned char *__restrict__ open, unsigned char *decided, int *minrow, int *undecided)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || decided[i])
		return;
	bool any = false;
	for (i64 k = Ap[i]; k < Ap[i + 1]; k++) {
		int j = Aj[k];
		if (open[j]) {              /* an open column is never pivotal: pivot columns lie in a pivotal row, all of which are closed */
			any = true;
			atomicMin(&minrow[j], i);
		}
	}
	if (!any)
		decided[i] = 1;
	else
		atomicAdd(undecided, 1);
}

__global__ void k_flc_commit(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, int *pinv, int *qinv,
                             const unsigned char *__restrict__ open, unsigned char *decided, const int *__restrict__ minrow,
                             unsigned char *closing, int *count)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || decided[i])
		return;
	int first = -1;
	for (i64 k = Ap[i]; k < Ap[i + 1]; k++) {
		int j = Aj[k];
		if (open[j]) {              /* open[] is not written by this kernel: every row sees the same snapshot */
			if (minrow[j] != i)
				return;            /* a lower undecided row may still close this column */
			if (first < 0)
				first = j;
		}
	}
	/* safe: the open columns of this row are disjoint from those of every other safe row, so the
	 * writes below cannot conflict (qinv[first] is read by no other committing row) */
	qinv[first] = i;
	pinv[i] = first;
	decided[i] = 1;
	closing[i] = 1;
	atomicAdd(count, 1);
}

__global__ void k_flc_close(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, unsigned char *closing, unsigned char *open)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !closing[i])
		return;
	closing[i] = 0;
	for (i64 k = Ap[i]; k < Ap[i + 1]; k++)
		open[Aj[k]] = 0;
}

/* ============================================================ greedy cycle-free search */

struct GreedyArgs {
	int n, m, words, queue_cap, use_smem;
	const i64 *Ap;
	const int *Aj;
	int *qinv, *pinv;
	int *journal;
	int *npiv;       /* committed pivots (journal length) */
	int *found;      /* sum of register_pivot results */
	int *ticket;
	int *status;     /* per row: 1 once resolved */
	int *hint;       /* every row below *hint is resolved */
	unsigned *bitmaps;
	int *queues;
	unsigned long long *edges;
};

__device__ __forceinline__ int ld_volatile(const int *p) { return *((const volatile int *) p); }

/* Expansion of a pivot row: its column indices are loaded 8 at a time BEFORE any of them is marked -- the marks are
 * shared-memory atomics the compiler will not move loads across, and one L2 round trip per entry, serialised, was
 * most of the latency of a BFS step (which is the critical path of a chain of dependent commits). */
#define EXPAND_ROW(MARK, b, e)                                                         \
	for (i64 k0_ = (b); k0_ < (e); k0_ += 8) {                                        \
		int c_[8];                                                                     \
		_Pragma("unroll") for (int u_ = 0; u_ < 8; u_++)                               \
			c_[u_] = (k0_ + u_ < (e)) ? __ldg(&a.Aj[k0_ + u_]) : -1;                   \
		_Pragma("unroll") for (int u_ = 0; u_ < 8; u_++)                               \
			if (c_[u_] >= 0)                                                           \
				MARK(c_[u_], vis, srv, queue, &sh);                                    \
	}

struct GreedyShared {
	int row, head, tail, alive, npiv_local, scan, replayed, np_now, turn;
};

__device__ __forceinline__ void greedy_mark(int c, unsigned *vis, unsigned *srv, int *queue, GreedyShared *sh)
{
	unsigned bit = 1u << (c & 31);
	int word = c >> 5;
	unsigned old = atomicOr(&vis[word], bit);
	if (old & bit)
		return;
	unsigned olds = atomicAnd(&srv[word], ~bit);
	if (olds & bit)
		atomicSub(&sh->alive, 1);
	queue[atomicAdd(&sh->tail, 1)] = c;
}

__global__ void __launch_bounds__(256, 8) k_greedy(GreedyArgs a)
{
	extern __shared__ unsigned dyn_smem[];
	__shared__ GreedyShared sh;
	const int tid = threadIdx.x, T = blockDim.x;
	unsigned *vis = a.use_smem ? dyn_smem : a.bitmaps + (size_t) blockIdx.x * 2 * a.words;
	unsigned *srv = vis + a.words;
	int *queue = a.queues + (size_t) blockIdx.x * a.queue_cap;
	unsigned long long my_edges = 0;

	for (int w = tid; w < 2 * a.words; w += T)
		vis[w] = 0;
	if (tid == 0)
		sh.scan = 0;
	__syncthreads();

	for (;;) {
		if (tid == 0)
			sh.row = atomicAdd(a.ticket, 1);
		__syncthreads();
		const int i = sh.row;
		if (i >= a.n)
			break;
		if (ld_volatile(&a.pinv[i]) >= 0) {        /* already pivotal (pivots.c:192) */
			if (tid == 0) {
				__threadfence();
				*((volatile int *) &a.status[i]) = 1;
			}
			__syncthreads();
			continue;
		}
		const i64 rb = a.Ap[i], re = a.Ap[i + 1];

		/* ---- begin the transaction: read the journal length, then scatter the row (pivots.c:198-215).
		 * `alive` counts DISTINCT candidate columns.  (On a row that repeats a column the reference counts the
		 * repetitions too, which can make it select a reachable entry and close an alternating cycle,
		 * pivots.c:232-238; that failure mode is deliberately not reproduced -- DESIGN.md, "quirks".) */
		if (tid == 0) {
			int np = ld_volatile(a.npiv);
			__threadfence();
			sh.npiv_local = np;
			int tail = 0, alive = 0;
			for (i64 k = rb; k < re; k++) {
				int j = a.Aj[k];
				unsigned bit = 1u << (j & 31);
				int word = j >> 5;
				if ((vis[word] | srv[word]) & bit)
					continue;                   /* repeated column */
				if (ld_volatile(&a.qinv[j]) < 0) {
					srv[word] |= bit;
					alive += 1;
				} else {
					queue[tail++] = j;
					vis[word] |= bit;
				}
			}
			sh.head = 0;
			sh.tail = tail;
			sh.alive = alive;
			sh.replayed = 0;
		}
		__syncthreads();

		/* ---- search / replay / wait loop.
		 * The BFS is kept closed under the pivots published so far WHILE the row waits for its turn: every poll
		 * replays the journal entries it has not seen (pivots.c:260-274).  When the turn comes (every lower row is
		 * resolved, so nobody else can commit) at most the last few entries remain to be replayed: the serial
		 * section of a commit stays short.  A row whose survivors are exhausted is resolved at once, out of
		 * order: reachability only grows with the pivot set. */
		bool final_pass = false;
		for (;;) {
			/* BFS along alternating paths (pivots.c:218-225), one slice of the queue per iteration */
			for (;;) {
				const int head = sh.head, tail = sh.tail, alive = sh.alive;
				__syncthreads();
				if (head >= tail || alive <= 0)
					break;
				const int cnt = min(T, tail - head);
				if (tid < cnt) {
					int j = queue[head + tid];
					int I = ld_volatile(&a.qinv[j]);
					if (I >= 0) {
						i64 b = a.Ap[I], e = a.Ap[I + 1];
						EXPAND_ROW(greedy_mark, b, e)
						my_edges += (unsigned long long) (e - b);
					}
				}
				if (tid == 0)
					sh.head = head + cnt;
				__syncthreads();
			}
			if (sh.alive <= 0 || final_pass)
				break;                          /* no survivor: final, whatever happens later; or exact state at our turn */
			/* poll: new pivots?  our turn? */
			if (tid == 0) {
				int k = max(sh.scan, ld_volatile(a.hint));
				while (k < i && ld_volatile(&a.status[k]))
					k++;
				sh.scan = k;
				__threadfence();
				sh.np_now = ld_volatile(a.npiv);     /* read AFTER the status scan: includes every commit of the rows below k */
				sh.turn = (k >= i);
			}
			__syncthreads();
			const int np_now = sh.np_now, np_old = sh.npiv_local;
			const bool turn = sh.turn;
			__syncthreads();
			if (np_now == np_old) {
				if (turn)
					break;                      /* closed under every pivot that can precede this row: commit */
				__nanosleep(40);
				continue;
			}
			if (turn)
				final_pass = true;              /* nobody else can commit any more: after this replay the state is exact */
			for (int t = np_old + tid; t < np_now; t += T) {
				int j = ld_volatile(&a.journal[t]);
				unsigned bit = 1u << (j & 31);
				int word = j >> 5;
				if (srv[word] & bit) {          /* a survivor became pivotal */
					unsigned olds = atomicAnd(&srv[word], ~bit);
					if (olds & bit) {
						atomicSub(&sh.alive, 1);
						atomicOr(&vis[word], bit);
						queue[atomicAdd(&sh.tail, 1)] = j;
					}
				} else if (vis[word] & bit) {   /* a reached column became pivotal: expand its row */
					int I = ld_volatile(&a.qinv[j]);
					i64 b = a.Ap[I], e = a.Ap[I + 1];
					EXPAND_ROW(greedy_mark, b, e)
					my_edges += (unsigned long long) (e - b);
				}
			}
			if (tid == 0)
				sh.npiv_local = np_now;
			__syncthreads();
		}
		const bool my_turn = true;              /* (a commit only happens at the row's turn; a failure needs none) */

		/* ---- commit (pivots.c:227-255) */
		if (tid == 0) {
			if (sh.alive > 0) {
				int j = -1;
				for (i64 k = rb; k < re; k++) {     /* first survivor in row order (pivots.c:233-237) */
					int c = a.Aj[k];
					if (srv[c >> 5] & (1u << (c & 31))) {
						j = c;
						break;
					}
				}
				a.pinv[i] = j;
				*((volatile int *) &a.qinv[j]) = i;
				int np = ld_volatile(a.npiv);
				a.journal[np] = j;
				__threadfence();
				*((volatile int *) a.npiv) = np + 1;
				atomicAdd(a.found, 1);
			}
			__threadfence();
			*((volatile int *) &a.status[i]) = 1;
			if (sh.alive > 0 && my_turn)
				atomicMax(a.hint, i + 1);      /* a commit happens at the row's turn: every row up to i is resolved */
		}
		__syncthreads();

		/* ---- reset the bitmaps (pivots.c:276-285) */
		for (i64 k = rb + tid; k < re; k += T) {
			int j = a.Aj[k];
			atomicAnd(&vis[j >> 5], ~(1u << (j & 31)));
			atomicAnd(&srv[j >> 5], ~(1u << (j & 31)));
		}
		const int tail = sh.tail;
		for (int t = tid; t < tail; t += T) {
			int j = queue[t];
			atomicAnd(&vis[j >> 5], ~(1u << (j & 31)));
		}
		__syncthreads();
	}
	if (my_edges)
		atomicAdd(a.edges, my_edges);
}

/* ------------------------------------------------------------------ out-of-order commits
 *
 * The ordered kernel above serialises every commit behind all lower rows.  The sequential semantics only needs less:
 *   (1) VISIBILITY: a greedy pivot (I, j) matters to row i only if I < i.  Greedy pivots are stored in qinv with a
 *       tag bit; a row ignores tagged pivots whose owner is above it, so rows may commit in any order without
 *       disturbing the rows below them.
 *   (2) SAFETY: row i may commit as soon as its search is closed under the published pivots of lower rows and no
 *       UNRESOLVED lower row can still produce a pivot that it would see: every unresolved lower row r publishes its
 *       surviving candidate columns (a superset of whatever it will finally pick, only ever shrinking); if none of
 *       them is marked (visited or survivor) in row i's bitmaps, a future pivot of r adds edges out of a column the
 *       search of i never reaches, so i's result is final.  Otherwise i waits for those rows only.
 * The lowest unresolved row never waits, so the scheme terminates, and the result is the sequential one.
 */
#define GREEDY_TAG 0x40000000
#define SURV_SLOTS 8
#define PEND_CAP 1024

struct GreedyOooArgs {
	int n, m, words, queue_cap, use_smem;
	const i64 *Ap;
	const int *Aj;
	int *qinv, *pinv;
	int *journal;
	int *reserved;   /* journal slots handed out */
	int *npiv;       /* journal slots published (every slot below is written) */
	int *ticket;
	int *status;     /* per row: 0 not initialised, 1 candidates published, 2 resolved */
	int *surv;       /* per row: SURV_SLOTS candidate columns (-1 = none), or surv[0] == -2: too many to list */
	int *hint;       /* every row below *hint is resolved */
	unsigned *bitmaps;
	int *queues;
	unsigned long long *edges;
};

#define ROWC_MAX 128
struct GreedyOooShared {
	int row, head, tail, alive, npiv_local, np_now, scan, blocked, npend, changed;
	int pend[PEND_CAP];
	int rowc[ROWC_MAX];      /* column indices of the current row when it has at most ROWC_MAX entries */
};

/* row owning the pivot of column j as seen by row i, or -1 */
__device__ __forceinline__ int visible_owner(const int *qinv, int j, int i)
{
	int q = ld_volatile(&qinv[j]);
	if (q < 0)
		return -1;
	if (q & GREEDY_TAG) {
		q &= ~GREEDY_TAG;
		return q < i ? q : -1;
	}
	return q;
}

__device__ __forceinline__ void ooo_mark(int c, unsigned *vis, unsigned *srv, int *queue, GreedyOooShared *sh)
{
	unsigned bit = 1u << (c & 31);
	int word = c >> 5;
	unsigned old = atomicOr(&vis[word], bit);
	if (old & bit)
		return;
	unsigned olds = atomicAnd(&srv[word], ~bit);
	if (olds & bit) {
		atomicSub(&sh->alive, 1);
		sh->changed = 1;
	}
	queue[atomicAdd(&sh->tail, 1)] = c;
}

__global__ void __launch_bounds__(256, 6) k_greedy_ooo(GreedyOooArgs a)
{
	extern __shared__ unsigned dyn_smem[];
	__shared__ GreedyOooShared sh;
	const int tid = threadIdx.x, T = blockDim.x;
	unsigned *vis = a.use_smem ? dyn_smem : a.bitmaps + (size_t) blockIdx.x * 2 * a.words;
	unsigned *srv = vis + a.words;
	int *queue = a.queues + (size_t) blockIdx.x * a.queue_cap;
	unsigned long long my_edges = 0;

	for (int w = tid; w < 2 * a.words; w += T)
		vis[w] = 0;
	if (tid == 0)
		sh.scan = 0;
	__syncthreads();

	for (;;) {
		if (tid == 0)
			sh.row = atomicAdd(a.ticket, 1);
		__syncthreads();
		const int i = sh.row;
		if (i >= a.n)
			break;
		if (ld_volatile(&a.pinv[i]) >= 0) {        /* pivotal since FL / FL-columns (pivots.c:192) */
			if (tid == 0) {
				__threadfence();
				*((volatile int *) &a.status[i]) = 2;
			}
			__syncthreads();
			continue;
		}
		const i64 rb = a.Ap[i], re = a.Ap[i + 1];
		int *mysurv = a.surv + (size_t) i * SURV_SLOTS;

		/* ---- scatter the row (pivots.c:198-215) and publish its candidate columns.  Warp 0: the lanes fetch the
		 * entries and their owners in parallel (two round trips for the whole row instead of two per entry), lane 0
		 * consumes them in row order.  Short rows are also kept in shared memory for the commit and the reset. */
		if (tid < 32) {
			int np = 0;
			if (tid == 0) {
				np = ld_volatile(a.npiv);
				__threadfence();
				sh.npiv_local = np;
			}
			np = __shfl_sync(0xffffffffu, np, 0);      /* the owners below are read after the journal length */
			__threadfence();
			const bool cache_row = (re - rb) <= ROWC_MAX;
			int tail = 0, alive = 0;
			for (i64 k0 = rb; k0 < re; k0 += 32) {
				const i64 k = k0 + tid;
				const int jmine = (k < re) ? a.Aj[k] : -1;
				const int omine = (jmine >= 0) ? visible_owner(a.qinv, jmine, i) : -1;
				if (cache_row && k < re)
					sh.rowc[k - rb] = jmine;
				const int cnt = (int) min((i64) 32, re - k0);
				for (int u = 0; u < cnt; u++) {
					const int j = __shfl_sync(0xffffffffu, jmine, u);
					const int own = __shfl_sync(0xffffffffu, omine, u);
					if (tid != 0)
						continue;
					unsigned bit = 1u << (j & 31);
					int word = j >> 5;
					if ((vis[word] | srv[word]) & bit)
						continue;               /* repeated column */
					if (own < 0) {
						srv[word] |= bit;
						if (alive < SURV_SLOTS)
							mysurv[alive] = j;
						alive += 1;
					} else {
						queue[tail++] = j;
						vis[word] |= bit;
					}
				}
			}
			if (tid == 0) {
				for (int t = alive; t < SURV_SLOTS; t++)
					mysurv[t] = -1;
				if (alive > SURV_SLOTS)
					mysurv[0] = -2;             /* too many to list: blocks the rows above until resolved */
				__threadfence();
				*((volatile int *) &a.status[i]) = 1;
				sh.head = 0;
				sh.tail = tail;
				sh.alive = alive;
				sh.npend = -1;                  /* pending list not built yet */
				sh.changed = 0;
			}
		}
		__syncthreads();

		for (;;) {
			/* ---- BFS along alternating paths (pivots.c:218-225) */
			for (;;) {
				const int head = sh.head, tail = sh.tail, alive = sh.alive;
				__syncthreads();
				if (head >= tail || alive <= 0)
					break;
				const int cnt = min(T, tail - head);
				if (tid < cnt) {
					int j = queue[head + tid];
					int I = visible_owner(a.qinv, j, i);
					if (I >= 0) {
						i64 b = a.Ap[I], e = a.Ap[I + 1];
						EXPAND_ROW(ooo_mark, b, e)
						my_edges += (unsigned long long) (e - b);
					}
				}
				if (tid == 0)
					sh.head = head + cnt;
				__syncthreads();
			}
			if (sh.alive <= 0)
				break;                          /* no survivor: final */
			/* ---- survivors that were reached are no candidates any more: shrink the published list */
			if (sh.changed) {
				if (tid < SURV_SLOTS) {
					int c = mysurv[tid];
					if (c >= 0 && !(srv[c >> 5] & (1u << (c & 31))))
						*((volatile int *) &mysurv[tid]) = -1;
				}
				__syncthreads();
				if (tid == 0)
					sh.changed = 0;
			}
			/* ---- replay the pivots of lower rows published since the last look (pivots.c:260-274) */
			if (tid == 0)
				sh.np_now = ld_volatile(a.npiv);
			__syncthreads();
			const int np_now = sh.np_now, np_old = sh.npiv_local;
			/* read here, between two barriers: thread 0 resets sh.npend as soon as it enters the collection below,
			 * and a warp that evaluated the condition after that would skip the block and its barriers */
			const bool collect_pending = sh.npend < 0;
			__syncthreads();
			if (np_now != np_old) {
				for (int t = np_old + tid; t < np_now; t += T) {
					int j = ld_volatile(&a.journal[t]);
					int owner = ld_volatile(&a.qinv[j]) & ~GREEDY_TAG;
					if (owner >= i)
						continue;               /* a pivot of a higher row: invisible to this row */
					unsigned bit = 1u << (j & 31);
					int word = j >> 5;
					if (srv[word] & bit) {      /* a candidate became pivotal */
						unsigned olds = atomicAnd(&srv[word], ~bit);
						if (olds & bit) {
							atomicSub(&sh.alive, 1);
							sh.changed = 1;
							atomicOr(&vis[word], bit);
							queue[atomicAdd(&sh.tail, 1)] = j;
						}
					} else if (vis[word] & bit) {   /* a reached column became pivotal: expand its row */
						i64 b = a.Ap[owner], e = a.Ap[owner + 1];
						EXPAND_ROW(ooo_mark, b, e)
						my_edges += (unsigned long long) (e - b);
					}
				}
				if (tid == 0)
					sh.npiv_local = np_now;
				__syncthreads();
				continue;                       /* close the search again */
			}
			/* ---- can any unresolved lower row still matter? */
			if (collect_pending) {              /* first time: collect the unresolved rows below i */
				if (tid == 0) {
					sh.npend = 0;
					sh.blocked = 0;
					int k = max(sh.scan, ld_volatile(a.hint));
					while (k < i && ld_volatile(&a.status[k]) == 2)
						k++;
					sh.scan = k;
				}
				__syncthreads();
				for (int r = sh.scan + tid; r < i; r += T)
					if (ld_volatile(&a.status[r]) != 2) {
						int pos = atomicAdd(&sh.npend, 1);
						if (pos < PEND_CAP)
							sh.pend[pos] = r;
					}
				__syncthreads();
				const bool overflow = sh.npend > PEND_CAP;
				__syncthreads();                /* every thread has read the count before thread 0 resets it */
				if (overflow) {                 /* too many to track: look again later */
					if (tid == 0)
						sh.npend = -1;
					__syncthreads();
					__nanosleep(500);
					continue;
				}
			}
			if (tid == 0)
				sh.blocked = 0;
			__syncthreads();
			const int npend = sh.npend;
			for (int t = tid; t < npend; t += T) {
				int r = sh.pend[t];
				if (r < 0)
					continue;
				int st = ld_volatile(&a.status[r]);
				if (st == 2) {
					sh.pend[t] = -1;            /* resolved: its pivot, if any, is in the journal */
					continue;
				}
				bool conflict = (st == 0);      /* candidates not published yet: unknown */
				if (!conflict) {
					/* the 8 slots in two 16-byte loads (one round trip; the slots are independent words, a list
					 * that shrinks between the two halves is still a superset of the final one) */
					const int *rs = a.surv + (size_t) r * SURV_SLOTS;
					int cand[SURV_SLOTS];
					asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];"
					             : "=r"(cand[0]), "=r"(cand[1]), "=r"(cand[2]), "=r"(cand[3]) : "l"(rs) : "memory");
					asm volatile("ld.volatile.global.v4.s32 {%0, %1, %2, %3}, [%4];"
					             : "=r"(cand[4]), "=r"(cand[5]), "=r"(cand[6]), "=r"(cand[7]) : "l"(rs + 4) : "memory");
#pragma unroll
					for (int q = 0; q < SURV_SLOTS; q++) {
						const int c = cand[q];
						if (c == -2)
							conflict = true;
						else if (c >= 0 && ((vis[c >> 5] | srv[c >> 5]) & (1u << (c & 31))))
							conflict = true;
					}
				}
				if (conflict)
					sh.blocked = 1;
			}
			__syncthreads();
			/* a row seen as resolved may have published its pivot after our replay: look at the journal once more */
			if (tid == 0)
				sh.np_now = ld_volatile(a.npiv);
			__syncthreads();
			const bool fresh = (sh.np_now == sh.npiv_local);
			const bool blocked = sh.blocked != 0;
			__syncthreads();
			if (!fresh)
				continue;
			if (!blocked)
				break;                          /* commit */
			__nanosleep(100);
		}

		/* ---- commit (pivots.c:227-255) */
		if (tid == 0) {
			if (sh.alive > 0) {
				int j = -1;
				const bool cached = (re - rb) <= ROWC_MAX;
				for (i64 k = rb; k < re; k++) {     /* first survivor in row order (pivots.c:233-237) */
					int c = cached ? sh.rowc[k - rb] : a.Aj[k];
					if (srv[c >> 5] & (1u << (c & 31))) {
						j = c;
						break;
					}
				}
				a.pinv[i] = j;
				*((volatile int *) &a.qinv[j]) = i | GREEDY_TAG;
				__threadfence();
				int pos = atomicAdd(a.reserved, 1);
				*((volatile int *) &a.journal[pos]) = j;
				__threadfence();
				while (ld_volatile(a.npiv) != pos)
					;                           /* publish in slot order: readers assume every slot below npiv is written */
				*((volatile int *) a.npiv) = pos + 1;
			}
			__threadfence();
			*((volatile int *) &a.status[i]) = 2;
			if (sh.npend >= 0 && sh.scan >= 0) {
				/* if everything below was resolved, move the hint */
				bool all = true;
				for (int t = 0; t < sh.npend && all; t++)
					all = sh.pend[t] < 0;
				if (all && sh.alive > 0)
					atomicMax(a.hint, i + 1);
			}
		}
		__syncthreads();

		/* ---- reset the bitmaps (pivots.c:276-285) */
		for (i64 k = rb + tid; k < re; k += T) {
			int j = a.Aj[k];
			atomicAnd(&vis[j >> 5], ~(1u << (j & 31)));
			atomicAnd(&srv[j >> 5], ~(1u << (j & 31)));
		}
		const int tail = sh.tail;
		for (int t = tid; t < tail; t += T) {
			int j = queue[t];
			atomicAnd(&vis[j >> 5], ~(1u << (j & 31)));
		}
		__syncthreads();
	}
	if (my_edges)
		atomicAdd(a.edges, my_edges);
}

__global__ void k_strip_tags(int m, int *qinv)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j < m && qinv[j] >= 0)
		qinv[j] &= ~GREEDY_TAG;
}

/* ============================================================ driver */

__global__ void k_longest_row(int n, const i64 *__restrict__ Ap, i64 *out)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	i64 len = (i < n) ? Ap[i + 1] - Ap[i] : 0;
	for (int o = 16; o > 0; o >>= 1)
		len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
	if ((threadIdx.x & 31) == 0 && len > 0)
		atomicMax((unsigned long long *) out, (unsigned long long) len);
}

bool greedy_windowed(const DevCsr &A, int *d_pinv, int *d_qinv, i64 longest_row, int *d_found, unsigned long long *d_edges);

PivotCounts pivots_find(const DevCsr &A, int *d_pinv, int *d_qinv, bool greedy)
{
	cudaStream_t s = ctx().stream;
	const int n = A.n, m = A.m;
	PivotCounts out = {0, 0, 0};
	CUDA_CHECK(cudaMemsetAsync(d_pinv, 0xff, (size_t) n * sizeof(int), s));
	CUDA_CHECK(cudaMemsetAsync(d_qinv, 0xff, (size_t) m * sizeof(int), s));
	if (n == 0 || m == 0 || A.nnz == 0)
		return out;
	DevBuf<int> counters(8);
	counters.zero(s);

	/* --- Faugere-Lachartre */
	{
		DevBuf<unsigned long long> key((size_t) m);
		key.fill_byte(0xff, s);
		k_fl_keys<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, key.ptr);
		k_fl_assign<<<cdiv(m, 256), 256, 0, s>>>(m, key.ptr, d_pinv, d_qinv, counters.ptr + 0);
		LAUNCHED(2);
		KERNEL_CHECK();
		out.fl = fetch(counters.ptr + 0);
	}

	static const bool trace = getenv("SPASM_B200_TRACE") != NULL;
	double t0 = spasm_wtime();
	if (trace) { sync(); fprintf(stderr, "[trace]   FL %.3f ms\n", 1e3 * (spasm_wtime() - t0)); t0 = spasm_wtime(); }
	/* --- FL on columns */
	{
		DevBuf<unsigned char> open((size_t) m), decided((size_t) n), closing((size_t) n);
		DevBuf<int> minrow((size_t) m);
		open.fill_byte(1, s);
		decided.zero(s);
		closing.zero(s);
		k_flc_close_pivotal<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, d_pinv, open.ptr, decided.ptr);
		LAUNCHED(1);
		for (int round = 0;; round++) {
			minrow.fill_byte(0x7f, s);
			CUDA_CHECK(cudaMemsetAsync(counters.ptr + 1, 0, sizeof(int), s));
			k_flc_reserve<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, d_qinv, open.ptr, decided.ptr, minrow.ptr, counters.ptr + 1);
			LAUNCHED(1);
			if (fetch(counters.ptr + 1) == 0)
				break;
			k_flc_commit<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, d_pinv, d_qinv, open.ptr, decided.ptr, minrow.ptr, closing.ptr, counters.ptr + 2);
			k_flc_close<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, closing.ptr, open.ptr);
			LAUNCHED(2);
		}
		KERNEL_CHECK();
		out.flcol = fetch(counters.ptr + 2);
	}

	if (trace) { sync(); fprintf(stderr, "[trace]   FL-columns %.3f ms\n", 1e3 * (spasm_wtime() - t0)); t0 = spasm_wtime(); }
	/* --- greedy */
	if (greedy) {
		GpuTimer timer;
		timer.start();
		GreedyArgs a;
		a.n = n;
		a.m = m;
		a.words = (m + 31) / 32;
		a.Ap = A.p;
		a.Aj = A.j;
		a.qinv = d_qinv;
		a.pinv = d_pinv;
		size_t bitmap_bytes = (size_t) 2 * a.words * sizeof(unsigned);
		/* tuning knobs (development): threads per CTA, bitmaps in shared or global memory, CTAs per SM */
		int threads = getenv("SPASM_B200_GREEDY_THREADS") ? atoi(getenv("SPASM_B200_GREEDY_THREADS")) : 256;
		int want_smem = getenv("SPASM_B200_GREEDY_SMEM") ? atoi(getenv("SPASM_B200_GREEDY_SMEM")) : 1;
		int per_sm = getenv("SPASM_B200_GREEDY_PER_SM") ? atoi(getenv("SPASM_B200_GREEDY_PER_SM")) : 8;
		a.use_smem = want_smem && bitmap_bytes <= 200 * 1024;
		size_t smem = a.use_smem ? bitmap_bytes : 0;
		if (smem > 0)
			CUDA_CHECK(cudaFuncSetAttribute(k_greedy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
		int occ = 0;
		CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_greedy, threads, smem));
		if (occ < 1)
			errx(1, "[spasm-b200] greedy pivot search kernel does not fit");
		int blocks = std::min(occ, per_sm) * ctx().sm_count;
		blocks = std::min(blocks, std::max(1, n));
		/* longest row bounds the extra queue slots used by the initial scatter */
		a.queue_cap = m + 64;
		i64 longest = 0;
		{
			/* rows longer than 64 entries: enlarge the queues by the longest row */
			DevBuf<i64> d_longest(1);
			d_longest.zero(s);
			k_longest_row<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, d_longest.ptr);
			LAUNCHED(1);
			longest = fetch(d_longest.ptr);
			a.queue_cap = m + (int) longest + 64;
		}
		static const bool ordered = getenv("SPASM_B200_GREEDY_ORDERED") != NULL;
		/* the windowed search (greedy_win.cu) serves short rows; the journal kernels below serve the rest */
		const bool try_windowed = !ordered && n < GREEDY_TAG;
		DevBuf<int> journal, status, queues;
		DevBuf<unsigned> bitmaps;
		DevBuf<unsigned long long> edges(1);
		edges.zero(s);
		CUDA_CHECK(cudaMemsetAsync(counters.ptr + 3, 0, 5 * sizeof(int), s));
		auto journal_buffers = [&]() {
			journal.alloc((size_t) n + 1);
			status.alloc((size_t) n + 1);
			queues.alloc((size_t) blocks * a.queue_cap);
			bitmaps.alloc(a.use_smem ? 1 : (size_t) blocks * 2 * a.words);
			status.zero(s);
			journal.zero(s);       /* every slot is written before it is published; zeroed so that initcheck sees it that way too */
		};
		/* development: run the ordered kernel as well, from the same starting point, and report the differences */
		static const bool shadow = getenv("SPASM_B200_GREEDY_SHADOW") != NULL;
		DevBuf<int> pinv0, qinv0;
		if (shadow) {
			pinv0.alloc((size_t) n);
			qinv0.alloc((size_t) m);
			CUDA_CHECK(cudaMemcpyAsync(pinv0.ptr, d_pinv, (size_t) n * sizeof(int), cudaMemcpyDeviceToDevice, s));
			CUDA_CHECK(cudaMemcpyAsync(qinv0.ptr, d_qinv, (size_t) m * sizeof(int), cudaMemcpyDeviceToDevice, s));
		}
		GpuTimer tk;
		bool windowed = false;
		if (try_windowed) {
			tk.start();
			windowed = greedy_windowed(A, d_pinv, d_qinv, longest, counters.ptr + 4, edges.ptr);
		}
		if (windowed) {
			/* done */
		} else if (ordered || n >= GREEDY_TAG) {
			journal_buffers();
			a.journal = journal.ptr;
			a.npiv = counters.ptr + 3;
			a.found = counters.ptr + 4;
			a.ticket = counters.ptr + 5;
			a.hint = counters.ptr + 6;
			a.status = status.ptr;
			a.bitmaps = bitmaps.ptr;
			a.queues = queues.ptr;
			a.edges = edges.ptr;
			tk.start();
			k_greedy<<<blocks, threads, smem, s>>>(a);
			LAUNCHED(1);
		} else {
			journal_buffers();
			GreedyOooArgs o;
			o.n = n; o.m = m; o.words = a.words; o.queue_cap = a.queue_cap; o.use_smem = a.use_smem;
			o.Ap = A.p; o.Aj = A.j; o.qinv = d_qinv; o.pinv = d_pinv;
			DevBuf<int> surv((size_t) n * SURV_SLOTS);
			surv.fill_byte(0xff, s);
			o.journal = journal.ptr;
			o.npiv = counters.ptr + 3;
			o.reserved = counters.ptr + 4;
			o.ticket = counters.ptr + 5;
			o.hint = counters.ptr + 6;
			o.status = status.ptr;
			o.surv = surv.ptr;
			o.bitmaps = bitmaps.ptr;
			o.queues = queues.ptr;
			o.edges = edges.ptr;
			if (smem > 0)
				CUDA_CHECK(cudaFuncSetAttribute(k_greedy_ooo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			int occ2 = 0;
			CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k_greedy_ooo, threads, smem));
			if (occ2 < 1)
				errx(1, "[spasm-b200] greedy pivot search kernel does not fit");
			int blocks2 = std::min(blocks, std::min(occ2, per_sm) * ctx().sm_count);
			/* an unresolved row is held by a CTA, so a row never has more pending lower rows than there are CTAs:
			 * keep that below the capacity of the pending list (the overflow path stays as a safety net) */
			blocks2 = std::min(blocks2, PEND_CAP);
			tk.start();
			k_greedy_ooo<<<blocks2, threads, smem, s>>>(o);
			k_strip_tags<<<cdiv(m, 256), 256, 0, s>>>(m, d_qinv);
			LAUNCHED(2);
		}
		KERNEL_CHECK();
		stats().pub.ms_k_greedy += tk.stop_ms();
		if (shadow && !(ordered || n >= GREEDY_TAG)) {
			int got = fetch(counters.ptr + 4);
			DevBuf<int> counters2(8), journal2((size_t) n + 1), status2((size_t) n + 1);
			if (windowed)
				journal_buffers();
			counters2.zero(s);
			status2.zero(s);
			a.qinv = qinv0.ptr;
			a.pinv = pinv0.ptr;
			a.journal = journal2.ptr;
			a.npiv = counters2.ptr + 3;
			a.found = counters2.ptr + 4;
			a.ticket = counters2.ptr + 5;
			a.hint = counters2.ptr + 6;
			a.status = status2.ptr;
			a.bitmaps = bitmaps.ptr;
			a.queues = queues.ptr;
			a.edges = edges.ptr;
			k_greedy<<<blocks, threads, smem, s>>>(a);
			KERNEL_CHECK();
			std::vector<int> p1((size_t) n), p2((size_t) n);
			CUDA_CHECK(cudaMemcpyAsync(p1.data(), d_pinv, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, s));
			CUDA_CHECK(cudaMemcpyAsync(p2.data(), pinv0.ptr, (size_t) n * sizeof(int), cudaMemcpyDeviceToHost, s));
			int want = fetch(counters2.ptr + 4);
			sync();
			int ndiff = 0, first = -1;
			for (int i = 0; i < n; i++)
				if (p1[i] != p2[i]) {
					if (first < 0)
						first = i;
					ndiff++;
				}
			if (ndiff)
				errx(1, "[spasm-b200] greedy pivot search: the %s kernel (%d pivots) and the ordered kernel (%d pivots) differ on %d rows, "
				        "first on row %d (column %d vs %d)", windowed ? "windowed" : "out-of-order", got, want, ndiff, first, p1[first], p2[first]);
		}
		out.greedy = fetch(counters.ptr + 4);
		stats().pub.greedy_edges += (i64) fetch(edges.ptr);
		stats().pub.ms_pivots_greedy += timer.stop_ms();
	}
	return out;
}

/* ============================================================ copy pivotal rows to U */

__global__ void k_urow_len(int npiv, const int *__restrict__ rows, const i64 *__restrict__ Ap, const int *__restrict__ Aj,
                           const int *__restrict__ pinv, i64 *len)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= npiv)
		return;
	int i = rows[k], j = pinv[i];
	i64 c = 1;
	for (i64 e = Ap[i]; e < Ap[i + 1]; e++)
		c += (Aj[e] != j);
	len[k] = c;
}

/* one warp per pivotal row: pivot first and equal to 1, the rest scaled by its inverse (pivots.c:405-443) */
__global__ void k_urow_fill(int npiv, const int *__restrict__ rows, const i64 *__restrict__ Ap, const int *__restrict__ Aj,
                            const i32 *__restrict__ Ax, const int *__restrict__ pinv, const i64 *__restrict__ Up_new, i64 unz0,
                            int un0, i64 *Up, int *Uj, i32 *Ux, int *Uqinv, Zp F)
{
	int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	int nwarps = (gridDim.x * blockDim.x) >> 5;
	for (int k = warp; k < npiv; k += nwarps) {
		int i = rows[k], j = pinv[i];
		i64 b = Ap[i], e = Ap[i + 1];
		/* the pivot value: first entry on column j (pivots.c:413-419) */
		i64 where = e;
		for (i64 k0 = b; k0 < e && where == e; k0 += 32) {
			i64 q = k0 + lane;
			bool hit = (q < e) && (Aj[q] == j) && (Ax[q] != 0);
			unsigned mask = __ballot_sync(0xffffffffu, hit);
			if (mask)
				where = k0 + (__ffs(mask) - 1);
		}
		i32 alpha = zp_inverse(Ax[where], F);
		i64 out = unz0 + Up_new[k];
		if (lane == 0) {
			Uj[out] = j;
			Ux[out] = 1;
			Uqinv[j] = un0 + k;
			Up[un0 + k + 1] = unz0 + Up_new[k + 1];
		}
		i64 written = 1;
		for (i64 k0 = b; k0 < e; k0 += 32) {
			i64 q = k0 + lane;
			bool keep = (q < e) && (Aj[q] != j);
			unsigned mask = __ballot_sync(0xffffffffu, keep);
			if (keep) {
				i64 pos = out + written + __popc(mask & ((1u << lane) - 1));
				Uj[pos] = Aj[q];
				Ux[pos] = zp_mul(alpha, Ax[q], F);
			}
			written += __popc(mask);
		}
	}
}

void append_pivotal_rows(const DevCsr &A, const int *d_rows, int npiv, const int *d_pinv, DevCsr &U, DevBuf<int> &Uqinv)
{
	if (npiv == 0)
		return;
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(A.prime);
	DevBuf<i64> len((size_t) npiv + 1), off((size_t) npiv + 1);
	len.zero(s);
	k_urow_len<<<cdiv(npiv, 256), 256, 0, s>>>(npiv, d_rows, A.p, A.j, d_pinv, len.ptr);
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, len.ptr, off.ptr, npiv + 1, s);
	tmp.ensure(bytes + 16);
	cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, len.ptr, off.ptr, npiv + 1, s);
	LAUNCHED(2);
	i64 extra = fetch(off.ptr + npiv);
	/* grow U (old rows are kept) */
	i64 unz0 = U.nnz;
	int un0 = U.n;
	DevBuf<i64> np((size_t) un0 + npiv + 1);
	DevBuf<int> nj((size_t) (unz0 + extra));
	DevBuf<i32> nx((size_t) (unz0 + extra));
	if (un0 > 0) {
		CUDA_CHECK(cudaMemcpyAsync(np.ptr, U.p.ptr, ((size_t) un0 + 1) * sizeof(i64), cudaMemcpyDeviceToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(nj.ptr, U.j.ptr, (size_t) unz0 * sizeof(int), cudaMemcpyDeviceToDevice, s));
		CUDA_CHECK(cudaMemcpyAsync(nx.ptr, U.x.ptr, (size_t) unz0 * sizeof(i32), cudaMemcpyDeviceToDevice, s));
	} else {
		CUDA_CHECK(cudaMemsetAsync(np.ptr, 0, sizeof(i64), s));
	}
	k_urow_fill<<<std::min(cdiv((size_t) npiv * 32, 256), 148u * 16), 256, 0, s>>>(npiv, d_rows, A.p, A.j, A.x, d_pinv, off.ptr, unz0, un0,
	                                                                                 np.ptr, nj.ptr, nx.ptr, Uqinv.ptr, F);
	LAUNCHED(1);
	KERNEL_CHECK();
	sync();
	U.p = std::move(np);
	U.j = std::move(nj);
	U.x = std::move(nx);
	U.n = un0 + npiv;
	U.nnz = unz0 + extra;
	U.m = A.m;
	U.prime = A.prime;
}

}  // namespace sb
