/*
 * Greedy alternating-cycle-free pivot search (reference: src/spasm_pivots.c:146-294), windowed formulation.
 *
 * The reference visits the non-pivotal rows in increasing order; row i searches the columns reachable from its
 * pivotal entries through the pivots committed so far, and takes its first unreached entry on a non-pivotal column.
 * The result depends on the commit order, so the one-thread reference is the parity target.  What makes it slow on a
 * GPU is not the searches but the chain of dependent commits (three out of four searching rows of BASELINE config 2
 * commit, and every commit is seen by a tenth of the later rows).  This file separates the two:
 *
 *  A. SEARCH, embarrassingly parallel.  A window of up to Wn candidate rows (rows that still own an entry on a
 *     non-pivotal column) is searched against the SAME snapshot of the pivots: one CTA per row, visited bitmap in
 *     shared memory, the pivot row of a column read with one 16..64-byte load from a per-column adjacency record
 *     (`padj`, rebuilt incrementally) instead of the qinv -> Ap -> Aj chain.  A row whose candidates are all reached
 *     has failed for good (reachability only grows with the pivot set).  A row with survivors runs its search to
 *     exhaustion: its visited set is then CLOSED under the snapshot, and it publishes, for every candidate column c
 *     of the window, the bit "this row reaches c" (R[c], one bit per window row).
 *
 *  B. RESOLUTION, sequential but tiny: one CTA walks the surviving rows in increasing order with, per row t, the set
 *     E[t] of window rows whose searches t INHERITS.  When row t0 commits the pivot (t0, c0), every later row t' that
 *     reaches c0 -- (R[c0] & E[t']) != 0, or c0 is one of its own candidate entries -- now also reaches whatever t0
 *     reaches (c0's new out-edges lead to the other entries of row t0, all of them inside t0's closed search, plus
 *     t0's other candidates):  E[t'] |= E[t0].  Row t's surviving candidates are those c with
 *     ((R[c] | (C[c] & committed)) & E[t]) == 0, where C[c] = window rows holding c as a candidate; it takes the first
 *     one in row order (pivots.c:233-237).  This is exactly the reference's journal replay (pivots.c:260-274) with the
 *     replayed sub-searches replaced by unions of searches that were already run, so the pivots are those of the
 *     one-thread reference, bit for bit.  A commit costs a few hundred shared-memory cycles instead of a round of
 *     polling, replaying and re-closing in global memory.
 *
 * Rows longer than WIN_MAXC entries, or column counts whose bitmap does not fit shared memory, take the journal
 * kernels of pivots.cu instead (later rounds on sparse Schur complements).
 */
#include <cub/cub.cuh>
#include "pivots.cuh"
#include "stats.cuh"

namespace sb {

#define WIN_MAXC 16
#define WIN_ECAP 256           /* candidate entries whose vectors are staged in shared memory at a time (phase B) */
#define WIN_BFS_THREADS 256
#define WIN_BFS_PER 4            /* frontier entries per thread and slice (phase A) */
#define WIN_CHASE 0              /* hops a thread follows on its own before the next slice (measured: 6 makes wide slices 7x longer, config 2 57 -> 88 ms) */
#define WIN_RING 4096          /* frontier ring in shared memory (phase A); the global queue holds everything */

struct WinArgs {
	int n, m, words;
	int Wn, W, K;
	const i64 *Ap;
	const int *Aj;
	int *qinv, *pinv;
	int *padj;                 /* m * K: the other entries of the pivot row of column j; [0] == -2: not pivotal; a last
	                            * slot <= -16 links to an overflow record of 16 more entries (rows longer than K + 1) */
	int *ovf, *novf;           /* overflow records, bump counter */
	const int *list;           /* candidate rows after FL / FL on columns, increasing */
	const int *lrow;           /* nlist * WIN_MAXC: their entries, padded with -1 */
	int nlist;
	int *pos;                  /* next position in `list` */
	int *nslots;               /* rows in the current window */
	int *slotrow, *slotlidx, *slotnc;   /* Wn: row, its index in `list`, number of candidate entries */
	int *slotcol, *slotvec;    /* Wn * WIN_MAXC: candidate columns in row order, and the vector each of them uses */
	int *tent;                 /* Wn: row has survivors against the snapshot */
	int *colmark;              /* m, 0x7f7f7f7f when unused: first window entry holding the column */
	int *replist, *nrep;       /* entries that own a vector */
	unsigned *Rvec, *Cvec;     /* (Wn * WIN_MAXC) * W */
	int *queues;
	int queue_cap;
	int *found;
	unsigned long long *edges;
};

/* ------------------------------------------------------------------ adjacency records */

#define WIN_OVF 16             /* entries of an overflow record */

/* adjacency record of pivot column j from the entries of its row (ent[0:n), -1 = hole): the other columns, the first
 * K - 1 (or K when they all fit) inline, the rest in an overflow record linked from the last slot */
template <int K>
__device__ __forceinline__ void win_build_record(const int *ent, int n, int j, int *rec, int *ovf, int *novf)
{
	int others = 0;
	for (int k = 0; k < n; k++)
		others += (ent[k] >= 0 && ent[k] != j);
	const bool spill = others > K;
	int cnt = 0, o = -1, ocnt = 0;
	if (spill)
		o = atomicAdd(novf, 1);
	for (int k = 0; k < n; k++) {
		const int c = ent[k];
		if (c < 0 || c == j)
			continue;
		if (!spill || cnt < K - 1)
			rec[cnt++] = c;
		else if (ocnt < WIN_OVF)
			ovf[(size_t) o * WIN_OVF + ocnt++] = c;
	}
	if (spill) {
		for (; ocnt < WIN_OVF; ocnt++)
			ovf[(size_t) o * WIN_OVF + ocnt] = -1;
		rec[K - 1] = -(16 + o);
	} else {
		for (; cnt < K; cnt++)
			rec[cnt] = -1;
	}
}

template <int K>
__global__ void k_win_padj_init(int m, const i64 *__restrict__ Ap, const int *__restrict__ Aj, const int *__restrict__ qinv, int *padj,
                                int *ovf, int *novf)
{
	int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= m)
		return;
	int *rec = padj + (size_t) j * K;
	const int I = qinv[j];
	if (I < 0) {
		rec[0] = -2;
		for (int k = 1; k < K; k++)
			rec[k] = -1;
		return;
	}
	int ent[WIN_MAXC];
	int n = 0;
	for (i64 e = Ap[I]; e < Ap[I + 1] && n < WIN_MAXC; e++)
		ent[n++] = Aj[e];
	win_build_record<K>(ent, n, j, rec, ovf, novf);
}

/* number of rows too long for an inline record */
__global__ void k_win_count_long(int n, const i64 *__restrict__ Ap, int limit, int *count)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n && Ap[i + 1] - Ap[i] > limit)
		atomicAdd(count, 1);
}

/* marks the columns of one adjacency record (and of its overflow record); NEW(c) is what happens to a column seen for
 * the first time */
#define WIN_EXPAND(ADJ, NEWCOL)                                                              \
	_Pragma("unroll") for (int k_ = 0; k_ < K; k_++) {                                       \
		const int c_ = (ADJ)[k_];                                                            \
		if (c_ >= 0) {                                                                       \
			my_edges += 1;                                                                   \
			const unsigned bit_ = 1u << (c_ & 31);                                           \
			if (!(atomicOr(&vis[c_ >> 5], bit_) & bit_)) { NEWCOL(c_) }                      \
		} else if (c_ <= -16) {                                                              \
			const int *o_ = a.ovf + (size_t) (-c_ - 16) * WIN_OVF;                           \
			for (int q_ = 0; q_ < WIN_OVF; q_++) {                                           \
				const int d_ = o_[q_];                                                       \
				if (d_ < 0)                                                                  \
					break;                                                                   \
				my_edges += 1;                                                               \
				const unsigned bit_ = 1u << (d_ & 31);                                       \
				if (!(atomicOr(&vis[d_ >> 5], bit_) & bit_)) { NEWCOL(d_) }                  \
			}                                                                                \
		}                                                                                    \
	}

/* rows that are not pivotal and hold an entry on a non-pivotal column */
__global__ void k_win_flag_rows(int n, const i64 *__restrict__ Ap, const int *__restrict__ Aj, const int *__restrict__ pinv,
                                const int *__restrict__ qinv, int *flag)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	int f = 0;
	if (pinv[i] < 0)
		for (i64 e = Ap[i]; e < Ap[i + 1] && !f; e++)
			f = qinv[Aj[e]] < 0;
	flag[i] = f;
}

/* the list of candidate rows and a padded copy of their entries (distinct columns, row order): the window kernels
 * read a row with four 16-byte loads instead of walking Ap -> Aj */
__global__ void k_win_list_rows(int n, const int *__restrict__ flag, const int *__restrict__ off, const i64 *__restrict__ Ap,
                                const int *__restrict__ Aj, int *list, int *lrow)
{
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !flag[i])
		return;
	const int t = off[i];
	list[t] = i;
	int ent[WIN_MAXC];
#pragma unroll
	for (int k = 0; k < WIN_MAXC; k++)
		ent[k] = -1;
	int cnt = 0;
	for (i64 e = Ap[i]; e < Ap[i + 1]; e++) {
		const int c = Aj[e];
		bool dup = false;
#pragma unroll
		for (int k = 0; k < WIN_MAXC; k++)
			dup |= (ent[k] == c);
		if (!dup) {                        /* a repeated column counts once */
#pragma unroll
			for (int k = 0; k < WIN_MAXC; k++)
				if (k == cnt)
					ent[k] = c;
			cnt++;
		}
	}
#pragma unroll
	for (int k = 0; k < WIN_MAXC; k += 4)
		*reinterpret_cast<int4 *>(lrow + (size_t) t * WIN_MAXC + k) = make_int4(ent[k], ent[k + 1], ent[k + 2], ent[k + 3]);
}

__device__ __forceinline__ void load_row16(const int *lrow, int lidx, int (&ent)[WIN_MAXC])
{
#pragma unroll
	for (int k = 0; k < WIN_MAXC; k += 4) {
		const int4 v = *reinterpret_cast<const int4 *>(lrow + (size_t) lidx * WIN_MAXC + k);
		ent[k] = v.x; ent[k + 1] = v.y; ent[k + 2] = v.z; ent[k + 3] = v.w;
	}
}

/* ------------------------------------------------------------------ window formation (one CTA of 1024 threads) */

__global__ void __launch_bounds__(1024) k_win_form(WinArgs a)
{
	typedef cub::BlockScan<int, 1024> Scan;
	__shared__ typename Scan::TempStorage tmp;
	__shared__ int s_newpos, s_total;
	const int tid = threadIdx.x;
	const int pos = *a.pos;
	if (pos >= a.nlist) {
		if (tid == 0)
			*a.nslots = 0;
		return;
	}
	const int per = (4 * a.Wn + 1023) / 1024;          /* list rows per thread: the chunk holds 4 Wn rows */
	const int chunk = min(per * 1024, a.nlist - pos);
	if (tid == 0)
		s_newpos = pos + chunk;
	/* which rows of the chunk still own a candidate entry */
	unsigned mine = 0;
	int cnt = 0;
	for (int u = 0; u < per; u++) {
		const int idx = tid * per + u;
		if (idx < chunk) {
			int ent[WIN_MAXC];
			load_row16(a.lrow, pos + idx, ent);
			int q[WIN_MAXC];
#pragma unroll
			for (int k = 0; k < WIN_MAXC; k++)
				q[k] = ent[k] >= 0 ? a.qinv[ent[k]] : 0;
			bool f = false;
#pragma unroll
			for (int k = 0; k < WIN_MAXC; k++)
				f |= q[k] < 0;
			if (f) {
				mine |= 1u << u;
				cnt++;
			}
		}
	}
	int before, total;
	Scan(tmp).ExclusiveSum(cnt, before, total);
	__syncthreads();
	for (int u = 0; u < per; u++)
		if (mine & (1u << u)) {
			if (before < a.Wn) {
				a.slotlidx[before] = pos + tid * per + u;
				if (before == a.Wn - 1)
					s_newpos = pos + tid * per + u + 1;
			}
			before++;
		}
	if (tid == 0)
		s_total = min(total, a.Wn);
	__syncthreads();
	const int nslots = s_total;
	if (tid == 0) {
		*a.nslots = nslots;
		*a.pos = s_newpos;
		*a.nrep = 0;
	}
	/* candidate entries of the window, in row order; the first entry holding a column owns its vectors */
	for (int b = tid; b < nslots; b += 1024) {
		const int lidx = a.slotlidx[b];
		int ent[WIN_MAXC];
		load_row16(a.lrow, lidx, ent);
		int q[WIN_MAXC];
#pragma unroll
		for (int k = 0; k < WIN_MAXC; k++)
			q[k] = ent[k] >= 0 ? a.qinv[ent[k]] : 0;
		int nc = 0;
#pragma unroll
		for (int k = 0; k < WIN_MAXC; k++)
			if (q[k] < 0) {
				a.slotcol[b * WIN_MAXC + nc] = ent[k];
				atomicMin(&a.colmark[ent[k]], b * WIN_MAXC + nc);
				nc++;
			}
		a.slotnc[b] = nc;
		a.slotrow[b] = a.list[lidx];
		a.tent[b] = 0;
	}
	__threadfence();
	__syncthreads();
	for (int b = tid; b < nslots; b += 1024) {
		const int nc = a.slotnc[b];
		for (int k = 0; k < nc; k++) {
			const int e = b * WIN_MAXC + k;
			const int v = a.colmark[a.slotcol[e]];
			a.slotvec[e] = v;
			atomicOr(&a.Cvec[(size_t) v * a.W + (b >> 5)], 1u << (b & 31));
			if (v == e)
				a.replist[atomicAdd(a.nrep, 1)] = e;
		}
	}
}

/* ------------------------------------------------------------------ phase A: one search per window row */

template <int K>
__global__ void __launch_bounds__(WIN_BFS_THREADS) k_win_bfs(WinArgs a)
{
	extern __shared__ unsigned vis[];
	__shared__ int s_tail, s_nc, s_head;
	__shared__ int s_alive[2];
	__shared__ int s_cand[WIN_MAXC];
	__shared__ int ring[WIN_RING];
	const int slot = blockIdx.x, tid = threadIdx.x;
	if (slot >= *a.nslots)
		return;
	int *queue = a.queues + (size_t) slot * a.queue_cap;
	for (int w = tid; w < a.words; w += WIN_BFS_THREADS)
		vis[w] = 0;
	__syncthreads();
	if (tid == 0) {
		int ent[WIN_MAXC];
		load_row16(a.lrow, a.slotlidx[slot], ent);
		int q[WIN_MAXC];
#pragma unroll
		for (int k = 0; k < WIN_MAXC; k++)
			q[k] = ent[k] >= 0 ? a.qinv[ent[k]] : -1;
		int tail = 0, nc = 0;
#pragma unroll
		for (int k = 0; k < WIN_MAXC; k++) {
			if (ent[k] < 0)
				continue;
			if (q[k] >= 0) {                    /* pivotal at the snapshot: a source of the search */
				vis[ent[k] >> 5] |= 1u << (ent[k] & 31);
				ring[tail] = ent[k];
				queue[tail++] = ent[k];
			} else {
				s_cand[nc++] = ent[k];
			}
		}
		s_tail = tail;
		s_nc = nc;
		s_alive[0] = nc;
		s_alive[1] = nc;
	}
	__syncthreads();
	unsigned long long my_edges = 0;
	int head = 0;
	int tail = s_tail, alive = s_alive[0];
	for (int iter = 0;; iter++) {
		/* one slice of the frontier per iteration.  head is replicated; the tail and the survivor count are read
		 * between two barriers (nobody pushes then), the count lags by one slice and is double-buffered, so every
		 * thread takes the same decisions */
		if (head >= tail || alive <= 0)
			break;
		if (tail - head <= 32) {
			/* NARROW FRONTIER (deep, thin parts of the pivot DAG: hundreds of levels late in the search): warp 0 walks
			 * the levels on its own, one L2 round trip per level and no block barrier, for as long as the frontier
			 * fits the warp; the other warps wait at the barrier below */
			if (tid < 32) {
				int h = head, t = tail, al = alive;
				for (int rounds = 0; h < t && t - h <= 32 && al > 0 && rounds < 256; rounds++) {
					const int j = (h + tid < t) ? ring[(h + tid) & (WIN_RING - 1)] : -1;
					if (j >= 0) {
						int adj[K];
#pragma unroll
						for (int k = 0; k < K; k += 4) {
							const int4 v = *reinterpret_cast<const int4 *>(a.padj + (size_t) j * K + k);
							adj[k] = v.x; adj[k + 1] = v.y; adj[k + 2] = v.z; adj[k + 3] = v.w;
						}
						if (adj[0] != -2) {
							my_edges += 1;
#define WIN_PUSH(c) { const int at_ = atomicAdd(&s_tail, 1); ring[at_ & (WIN_RING - 1)] = (c); queue[at_] = (c); }
							WIN_EXPAND(adj, WIN_PUSH)
						}
					}
					__syncwarp();
					h = t;
					t = *((volatile int *) &s_tail);
					const bool live = tid < s_nc && !(*((volatile unsigned *) &vis[s_cand[tid] >> 5]) & (1u << (s_cand[tid] & 31)));
					al = __popc(__ballot_sync(0xffffffffu, live));
				}
				if (tid == 0) {
					s_head = h;
					s_alive[(iter + 1) & 1] = al;
				}
			}
			__syncthreads();
			head = s_head;
			tail = s_tail;
			alive = s_alive[(iter + 1) & 1];
			__syncthreads();
			continue;
		}
		const int cnt = min(WIN_BFS_PER * WIN_BFS_THREADS, tail - head);
		const bool in_ring = tail - head <= WIN_RING;
		/* up to WIN_BFS_PER frontier entries per thread: their adjacency records are loaded together (a slice costs one
		 * L2 round trip whatever its width).  The entries are taken out of the ring before anybody pushes: the pushes
		 * of a wide slice may wrap around it. */
		int js[WIN_BFS_PER];
#pragma unroll
		for (int u = 0; u < WIN_BFS_PER; u++) {
			const int q = head + tid + u * WIN_BFS_THREADS;
			js[u] = (q < head + cnt) ? (in_ring ? ring[q & (WIN_RING - 1)] : queue[q]) : -1;
		}
		__syncthreads();
		if (js[0] >= 0) {
			/* CHASE: of the columns an entry discovers, the thread keeps the first one for itself and expands it in the
			 * next round (up to WIN_CHASE rounds) instead of pushing it: a thin chain of the pivot DAG is then walked by
			 * one thread with one L2 round trip per hop, not one slice (three block barriers) per hop.  The order of a
			 * search does not matter, only the set it reaches. */
			for (int round = 0; round <= WIN_CHASE; round++) {
				int adj[WIN_BFS_PER][K];
				bool any = false;
#pragma unroll
				for (int u = 0; u < WIN_BFS_PER; u++) {
					if (js[u] >= 0) {
						any = true;
#pragma unroll
						for (int k = 0; k < K; k += 4) {
							const int4 v = *reinterpret_cast<const int4 *>(a.padj + (size_t) js[u] * K + k);
							adj[u][k] = v.x; adj[u][k + 1] = v.y; adj[u][k + 2] = v.z; adj[u][k + 3] = v.w;
						}
					} else {
						adj[u][0] = -2;
					}
				}
				if (!any)
					break;
				const bool last = round == WIN_CHASE;
#pragma unroll
				for (int u = 0; u < WIN_BFS_PER; u++) {
					js[u] = -1;
					if (adj[u][0] == -2)
						continue;
					my_edges += 1;
#define WIN_KEEP_OR_PUSH(c) { if (js[u] < 0 && !last) js[u] = (c); else WIN_PUSH(c) }
					WIN_EXPAND(adj[u], WIN_KEEP_OR_PUSH)
				}
			}
		}
		if (tid >= WIN_BFS_THREADS - 32) {
			/* the last warp recounts the survivors (marks of the earlier slices and part of this one) */
			__syncwarp();
			const int l = tid - (WIN_BFS_THREADS - 32);
			const bool live = l < s_nc && !(vis[s_cand[l] >> 5] & (1u << (s_cand[l] & 31)));
			const unsigned mask = __ballot_sync(0xffffffffu, live);
			if (l == 0)
				s_alive[(iter + 1) & 1] = __popc(mask);
		}
		head += cnt;
		__syncthreads();
		tail = s_tail;
		alive = s_alive[(iter + 1) & 1];
		__syncthreads();
	}
	__syncthreads();
	if (my_edges)
		atomicAdd(a.edges, my_edges);
	/* final recount with every mark in place */
	if (tid < 32) {
		const bool live = tid < s_nc && !(vis[s_cand[tid] >> 5] & (1u << (s_cand[tid] & 31)));
		const unsigned mask = __ballot_sync(0xffffffffu, live);
		if (tid == 0)
			s_alive[0] = __popc(mask);
	}
	__syncthreads();
	if (s_alive[0] <= 0)
		return;                                 /* every candidate is reached: failed for good */
	/* survivors.  The loop above stops early only when no candidate is left, so the search ran to exhaustion: it is
	 * closed under the snapshot.  Publish which candidate columns of the window it reaches. */
	if (tid == 0)
		a.tent[slot] = 1;
	const int nrep = *a.nrep;
	const unsigned mybit = 1u << (slot & 31);
	for (int t = tid; t < nrep; t += WIN_BFS_THREADS) {
		const int e = a.replist[t];
		const int c = a.slotcol[e];
		if (vis[c >> 5] & (1u << (c & 31)))
			atomicOr(&a.Rvec[(size_t) e * a.W + (slot >> 5)], mybit);
	}
}

/* ------------------------------------------------------------------ phase B: ordered resolution, one CTA, blockDim == Wn = 32 W */

template <int K, int W>
__global__ void __launch_bounds__(32 * W) k_win_resolve(WinArgs a)
{
	constexpr int Wn = 32 * W;
	constexpr int NWARPS = W;
	extern __shared__ unsigned dyn[];
	unsigned *chR = dyn;                                           /* WIN_ECAP * W */
	unsigned *chC = chR + (size_t) WIN_ECAP * W;
	unsigned *taken = chC + (size_t) WIN_ECAP * W;                 /* Wn * WIN_MAXC bits */
	__shared__ int tl[1024];                                       /* surviving rows, increasing */
	__shared__ unsigned char tnc[1024];                            /* their number of candidate entries */
	__shared__ int ch_vec[WIN_ECAP];
	__shared__ int ch_off[WIN_ECAP + 1];
	__shared__ int s_warpcnt[32];
	__shared__ unsigned s_committed[32];
	__shared__ unsigned s_reff[2][WIN_MAXC][W], s_c0[2][WIN_MAXC][W];   /* per candidate of the current row (double-buffered) */
	__shared__ unsigned s_hit[3];                                  /* bit k: candidate k is reached (or no candidate any more); rotated */
	__shared__ unsigned s_Eb[2][W];                                /* E[b] of the row being resolved (double-buffered) */
	__shared__ int s_T, s_rows;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int nslots = *a.nslots;
	if (nslots == 0)
		return;
	const bool tentative = tid < nslots && a.tent[tid];
	const unsigned bal = __ballot_sync(0xffffffffu, tentative);
	if (lane == 0)
		s_warpcnt[warp] = __popc(bal);
	/* E[t], the window rows whose searches row t inherits (itself included): ONE ROW PER THREAD, IN REGISTERS */
	unsigned Er[W];
#pragma unroll
	for (int w = 0; w < W; w++)
		Er[w] = (w == (tid >> 5)) ? (1u << (tid & 31)) : 0u;
	for (int idx = tid; idx < (Wn * WIN_MAXC) / 32; idx += Wn)
		taken[idx] = 0;
	if (tid < 32)
		s_committed[tid] = 0;
	if (tid < 3)
		s_hit[tid] = 0;
	__syncthreads();
	{
		int before = 0;
		for (int w = 0; w < warp; w++)
			before += s_warpcnt[w];
		if (tentative) {
			const int t = before + __popc(bal & ((1u << lane) - 1));
			tl[t] = tid;
			tnc[t] = (unsigned char) a.slotnc[tid];
		}
		if (tid == Wn - 1)
			s_T = before + __popc(bal);
	}
	__syncthreads();
	const int T = s_T;
	int mypick = -1;
	int iter = 0;                       /* the per-row scratch is double-buffered: a slow warp may still read the previous row's */
	if (T > 0 && tid == tl[0]) {
#pragma unroll
		for (int w = 0; w < W; w++)
			s_Eb[0][w] = Er[w];
	}
	for (int base = 0; base < T;) {
		/* ---- stage the vectors of the next rows (as many rows as fit WIN_ECAP entries) */
		if (tid == 0) {
			int rows = 0, ent = 0;
			while (base + rows < T && ent + tnc[base + rows] <= WIN_ECAP) {
				ch_off[rows] = ent;
				ent += tnc[base + rows];
				rows++;
			}
			ch_off[rows] = ent;
			s_rows = rows;
		}
		__syncthreads();
		const int rows_here = s_rows;
		const int nent = ch_off[rows_here];
		for (int r = tid; r < rows_here; r += Wn) {
			const int b = tl[base + r];
			for (int k = 0; k < tnc[base + r]; k++)
				ch_vec[ch_off[r] + k] = a.slotvec[b * WIN_MAXC + k];
		}
		__syncthreads();
		for (int idx = tid; idx < nent * W; idx += Wn) {
			const int v = ch_vec[idx / W];
			const int w = idx % W;
			chR[idx] = a.Rvec[(size_t) v * W + w];
			chC[idx] = a.Cvec[(size_t) v * W + w];
		}
		__syncthreads();
		/* ---- the rows of the chunk, in order; everything below touches shared memory only */
		for (int r = 0; r < rows_here; r++, iter++) {
			const int b = tl[base + r];
			const int nc = tnc[base + r], off = ch_off[r];
			const int pb = iter & 1, hb = iter % 3;
			/* the hit mask of the NEXT row is cleared here: its previous use (two rows ago) was read before the
			 * barrier of the previous row, and its next writers come after the barrier below */
			if (tid == 0)
				s_hit[(iter + 1) % 3] = 0;
			/* the row resolved next publishes its E row now (rewritten below if this row commits and changes it) */
			const int bnext = (base + r + 1 < T) ? tl[base + r + 1] : -1;
			if (tid == bnext) {
#pragma unroll
				for (int w = 0; w < W; w++)
					s_Eb[pb ^ 1][w] = Er[w];
			}
			/* one warp per candidate entry: is it reached by row b or by a search row b inherits? */
			for (int k = warp; k < nc; k += NWARPS) {
				const int v = ch_vec[off + k];
				const bool gone = (taken[v >> 5] >> (v & 31)) & 1u;              /* became pivotal inside this window */
				unsigned reff = 0, cv = 0, Ev = 0;
				if (lane < W) {
					cv = chC[(off + k) * W + lane];
					reff = chR[(off + k) * W + lane] | (cv & s_committed[lane]);
					Ev = s_Eb[pb][lane];
					s_reff[pb][k][lane] = reff;
					s_c0[pb][k][lane] = cv;
				}
				const bool hit = __any_sync(0xffffffffu, (reff & Ev) != 0) || gone;
				if (lane == 0 && hit)
					atomicOr(&s_hit[hb], 1u << k);
			}
			__syncthreads();
			const unsigned hitmask = s_hit[hb];
			const unsigned freemask = ~hitmask & ((nc >= 32) ? 0xffffffffu : ((1u << nc) - 1));
			if (freemask == 0)
				continue;                                       /* uniform: no survivor, row b fails */
			const int pick = __ffs(freemask) - 1;               /* first survivor in row order (pivots.c:233-237) */
			/* commit (b, pick): the later rows that reach the new pivot inherit the search of row b */
			if (tentative && tid > b) {
				bool hit = (s_c0[pb][pick][tid >> 5] >> (tid & 31)) & 1u;
#pragma unroll
				for (int w = 0; w < W; w++)
					hit |= (Er[w] & s_reff[pb][pick][w]) != 0;
				if (hit) {
#pragma unroll
					for (int w = 0; w < W; w++)
						Er[w] |= s_Eb[pb][w];
					if (tid == bnext) {
#pragma unroll
						for (int w = 0; w < W; w++)
							s_Eb[pb ^ 1][w] = Er[w];
					}
				}
			}
			if (tid == b)
				mypick = pick;
			if (tid == 0) {
				s_committed[b >> 5] |= 1u << (b & 31);
				const int v0 = ch_vec[off + pick];
				taken[v0 >> 5] |= 1u << (v0 & 31);
			}
			__syncthreads();
		}
		base += rows_here;
		__syncthreads();
	}
	/* publish the new pivots and their adjacency records */
	if (mypick >= 0) {
		const int row = a.slotrow[tid];
		const int c0 = a.slotcol[tid * WIN_MAXC + mypick];
		a.qinv[c0] = row;
		a.pinv[row] = c0;
		int ent[WIN_MAXC];
		load_row16(a.lrow, a.slotlidx[tid], ent);
		win_build_record<K>(ent, WIN_MAXC, c0, a.padj + (size_t) c0 * K, a.ovf, a.novf);
		atomicAdd(a.found, 1);
	}
}

/* reset what the window used (column marks, vectors) */
__global__ void k_win_cleanup(WinArgs a)
{
	const int nslots = *a.nslots;
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	const int b = e / WIN_MAXC, k = e - b * WIN_MAXC;
	if (b >= nslots || k >= a.slotnc[b])
		return;
	const int c = a.slotcol[e];
	a.colmark[c] = 0x7f7f7f7f;
	if (a.slotvec[e] == e)
		for (int w = 0; w < a.W; w++) {
			a.Rvec[(size_t) e * a.W + w] = 0;
			a.Cvec[(size_t) e * a.W + w] = 0;
		}
}

/* ------------------------------------------------------------------ driver */

template <int K, int W>
static void run_windows(WinArgs &a, size_t bfs_smem, size_t res_smem, int max_windows)
{
	cudaStream_t s = ctx().stream;
	CUDA_CHECK(cudaFuncSetAttribute(k_win_bfs<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bfs_smem));
	CUDA_CHECK(cudaFuncSetAttribute(k_win_resolve<K, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) res_smem));
	k_win_padj_init<K><<<cdiv(a.m, 256), 256, 0, s>>>(a.m, a.Ap, a.Aj, a.qinv, a.padj, a.ovf, a.novf);
	LAUNCHED(1);
	int done = 0;
	while (done < max_windows) {
		const int batch = std::min(16, max_windows - done);
		for (int w = 0; w < batch; w++) {
			k_win_form<<<1, 1024, 0, s>>>(a);
			k_win_bfs<K><<<a.Wn, WIN_BFS_THREADS, bfs_smem, s>>>(a);
			k_win_resolve<K, W><<<1, a.Wn, res_smem, s>>>(a);
			k_win_cleanup<<<cdiv((size_t) a.Wn * WIN_MAXC, 256), 256, 0, s>>>(a);
		}
		LAUNCHED(4 * batch);
		KERNEL_CHECK();
		done += batch;
		if (fetch(a.pos) >= a.nlist)
			break;
	}
}

/* returns false when the windowed search does not apply (long rows, bitmap too large): the caller takes the journal
 * kernels.  On success *d_found (device counter) is incremented by the number of new pivots. */
bool greedy_windowed(const DevCsr &A, int *d_pinv, int *d_qinv, i64 longest_row, int *d_found, unsigned long long *d_edges)
{
	cudaStream_t s = ctx().stream;
	const int n = A.n, m = A.m;
	static const bool off = getenv("SPASM_B200_GREEDY_JOURNAL") != NULL;
	if (off || longest_row > WIN_MAXC)
		return false;
	WinArgs a;
	a.n = n;
	a.m = m;
	a.words = (m + 31) / 32;
	const size_t bfs_smem = (size_t) a.words * sizeof(unsigned);
	if (bfs_smem > 160 * 1024)
		return false;
	int Wn = getenv("SPASM_B200_GREEDY_WINDOW") ? atoi(getenv("SPASM_B200_GREEDY_WINDOW")) : 512;
	Wn = Wn >= 1024 ? 1024 : (Wn >= 512 ? 512 : 256);
	a.Wn = Wn;
	a.W = Wn / 32;
	a.K = longest_row <= 5 ? 4 : 8;      /* longer rows (up to WIN_MAXC entries) spill into an overflow record */
	a.Ap = A.p;
	a.Aj = A.j;
	a.qinv = d_qinv;
	a.pinv = d_pinv;
	a.found = d_found;
	a.edges = d_edges;
	/* candidate rows */
	DevBuf<int> flag((size_t) n + 1), off_((size_t) n + 1), list((size_t) std::max(n, 1));
	CUDA_CHECK(cudaMemsetAsync(flag.ptr + n, 0, sizeof(int), s));
	k_win_flag_rows<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, A.j, d_pinv, d_qinv, flag.ptr);
	static DevBuf<char> tmp;
	size_t bytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, bytes, flag.ptr, off_.ptr, n + 1, s);
	tmp.ensure(bytes + 16);
	cub::DeviceScan::ExclusiveSum(tmp.ptr, bytes, flag.ptr, off_.ptr, n + 1, s);
	LAUNCHED(3);
	a.nlist = fetch(off_.ptr + n);
	if (a.nlist == 0)
		return true;
	/* Which formulation?  Measured on the BASELINE shapes: with short rows (3-5 entries: configs 1, 2, 4, 5) most
	 * searching rows commit and the windows win or tie (config 1: 6.0 vs 7.8 ms, config 2: 57 vs 54 ms) without any
	 * spin-wait; with 11 entries per row (config 3) 96 % of the searches fail after visiting two thirds of the graph, the
	 * work is pure traversal throughput and the journal kernel keeps more rows in flight (109 vs 189 ms). */
	{
		static const bool force = getenv("SPASM_B200_GREEDY_WINDOW") != NULL;
		const double avg_len = (double) A.nnz / std::max(1, n);
		if (!force && avg_len > 6.0)
			return false;
	}
	DevBuf<int> lrow((size_t) a.nlist * WIN_MAXC);
	k_win_list_rows<<<cdiv(n, 256), 256, 0, s>>>(n, flag.ptr, off_.ptr, A.p, A.j, list.ptr, lrow.ptr);
	LAUNCHED(1);
	a.list = list.ptr;
	a.lrow = lrow.ptr;
	a.queue_cap = m + WIN_MAXC;
	/* overflow records: one per row that does not fit an inline record (a row becomes a pivot row at most once) */
	DevBuf<int> nlong(2);
	nlong.zero(s);
	k_win_count_long<<<cdiv(n, 256), 256, 0, s>>>(n, A.p, a.K + 1, nlong.ptr);
	LAUNCHED(1);
	const int novf_cap = fetch(nlong.ptr);
	DevBuf<int> ovf((size_t) std::max(novf_cap, 1) * WIN_OVF);
	a.ovf = ovf.ptr;
	a.novf = nlong.ptr + 1;
	DevBuf<int> padj((size_t) m * a.K), counters(4), slotrow((size_t) Wn), slotlidx((size_t) Wn), slotnc((size_t) Wn),
	    slotcol((size_t) Wn * WIN_MAXC), slotvec((size_t) Wn * WIN_MAXC), tent((size_t) Wn), colmark((size_t) m),
	    replist((size_t) Wn * WIN_MAXC), queues((size_t) Wn * a.queue_cap);
	DevBuf<unsigned> Rvec((size_t) Wn * WIN_MAXC * a.W), Cvec((size_t) Wn * WIN_MAXC * a.W);
	counters.zero(s);
	colmark.fill_byte(0x7f, s);
	Rvec.zero(s);
	Cvec.zero(s);
	a.padj = padj.ptr;
	a.pos = counters.ptr + 0;
	a.nslots = counters.ptr + 1;
	a.nrep = counters.ptr + 2;
	a.slotrow = slotrow.ptr;
	a.slotlidx = slotlidx.ptr;
	a.slotnc = slotnc.ptr;
	a.slotcol = slotcol.ptr;
	a.slotvec = slotvec.ptr;
	a.tent = tent.ptr;
	a.colmark = colmark.ptr;
	a.replist = replist.ptr;
	a.Rvec = Rvec.ptr;
	a.Cvec = Cvec.ptr;
	a.queues = queues.ptr;
	const size_t res_smem = (2 * (size_t) WIN_ECAP * a.W + (size_t) Wn * WIN_MAXC / 32) * sizeof(unsigned);
	const int max_windows = (a.nlist + Wn - 1) / Wn;
#define WIN_DISPATCH(KK)                                                   \
	do {                                                                   \
		if (a.W == 8)                                                      \
			run_windows<KK, 8>(a, bfs_smem, res_smem, max_windows);        \
		else if (a.W == 16)                                                \
			run_windows<KK, 16>(a, bfs_smem, res_smem, max_windows);       \
		else                                                               \
			run_windows<KK, 32>(a, bfs_smem, res_smem, max_windows);       \
	} while (0)
	if (a.K == 4)
		WIN_DISPATCH(4);
	else
		WIN_DISPATCH(8);
	return true;
}

}  // namespace sb
