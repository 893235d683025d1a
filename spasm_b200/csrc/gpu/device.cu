/*
 * Device context, statistics and the small extra C ABI of include/spasm_b200.h.
 */
#include "common.cuh"
#include "stats.cuh"
#include <sys/mman.h>
#include <thread>
#include <unistd.h>

namespace sb {

static bool g_exiting = false;        /* set by an atexit handler: device_free becomes a no-op */
static Context g_ctx;
static Stats g_stats;
static int g_requested_device = -1;
static bool g_verbose = true;
static cudaMemPool_t g_pool = nullptr;

Stats &stats() { return g_stats; }

static std::thread *g_warm = nullptr;      /* context creation started ahead of the first GPU call (spasm_b200_warmup_async) */

static int pick_device(int count)
{
	int dev = g_requested_device;
	if (dev < 0) {
		const char *s = getenv("SPASM_B200_DEVICE");
		if (!s)
			s = getenv("LOCAL_RANK");
		dev = s ? atoi(s) % count : 0;
	}
	return dev;
}

Context &ctx()
{
	if (g_ctx.device >= 0)
		return g_ctx;
	if (g_warm) {                      /* the helper thread is creating (or has created) the primary context */
		g_warm->join();
		delete g_warm;
		g_warm = nullptr;
	}
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0)
		errx(1, "[spasm-b200] no usable CUDA device (%s). This library has no CPU fallback: "
		        "spasm_echelonize / spasm_rref / spasm_kernel / spasm_schur* run on a B200 (sm_100a) only.",
		     e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
	int dev = pick_device(count);
	CUDA_CHECK(cudaSetDevice(dev));
	cudaDeviceProp prop;
	CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
	if (prop.major != 10)
		errx(1, "[spasm-b200] device %d (%s) is compute capability %d.%d; this build contains sm_100a code only",
		     dev, prop.name, prop.major, prop.minor);
	g_ctx.device = dev;
	g_ctx.sm_count = prop.multiProcessorCount;
	g_ctx.smem_optin = prop.sharedMemPerBlockOptin;
	g_ctx.verbose = g_verbose;
	/* the library's stream has the highest priority: the dense echelon runs its wide trailing updates on a second,
	 * lowest-priority stream, and the panel factorisation queued behind them on this one must get the SMs they free */
	int prio_least = 0, prio_greatest = 0;
	CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
	CUDA_CHECK(cudaStreamCreateWithPriority(&g_ctx.stream, cudaStreamNonBlocking, prio_greatest));
	/* a memory pool OWNED by the library (the device's default pool is shared with every other cudaMallocAsync user of
	 * the process and is left alone): freed blocks stay in it instead of going back to the driver at every
	 * synchronisation; spasm_b200_trim() gives everything back */
	cudaMemPoolProps props;
	memset(&props, 0, sizeof(props));
	props.allocType = cudaMemAllocationTypePinned;
	props.handleTypes = cudaMemHandleTypeNone;
	props.location.type = cudaMemLocationTypeDevice;
	props.location.id = dev;
	CUDA_CHECK(cudaMemPoolCreate(&g_pool, &props));
	unsigned long long keep = ~0ull;
	CUDA_CHECK(cudaMemPoolSetAttribute(g_pool, cudaMemPoolAttrReleaseThreshold, &keep));
	atexit([] { g_exiting = true; });      /* registered after the CUDA runtime's own handlers: runs before them */
	return g_ctx;
}

/* memcpy by several threads: the destination is usually untouched memory, so this is where the page faults happen */
static void parallel_memcpy(char *dst, const char *src, size_t bytes, int nthreads)
{
	if (nthreads <= 1 || bytes < ((size_t) 1 << 20)) {
		memcpy(dst, src, bytes);
		return;
	}
	std::vector<std::thread> pool;
	size_t slice = ((bytes / nthreads) + 4095) & ~(size_t) 4095;
	for (int t = 1; t < nthreads; t++) {
		size_t lo = std::min(bytes, (size_t) t * slice), hi = std::min(bytes, lo + slice);
		if (hi > lo)
			pool.emplace_back([=] { memcpy(dst + lo, src + lo, hi - lo); });
	}
	memcpy(dst, src, std::min(bytes, slice));
	for (std::thread &th : pool)
		th.join();
}

/* ---- large-block cache (common.cuh) */
static const size_t BIG_BLOCK = (size_t) 32 << 20;
struct CachedBlock { void *ptr; size_t bytes; };
/* never destroyed: buffers with static storage in other translation units are released at exit, in unspecified order */
static std::vector<CachedBlock> &g_cache = *new std::vector<CachedBlock>();       /* free blocks */
static std::vector<CachedBlock> &g_live_big = *new std::vector<CachedBlock>();    /* blocks handed out from the cache (possibly larger than asked) */
static size_t g_cache_bytes = 0;
static size_t cache_limit()
{
	static size_t limit = 0;
	if (!limit) {
		const char *e = getenv("SPASM_B200_CACHE_GB");
		limit = (size_t) ((e ? atof(e) : 48.0) * 1e9);
	}
	return limit;
}

void *device_alloc(size_t bytes)
{
	void *p = nullptr;
	if (bytes >= BIG_BLOCK) {
		/* best fit among the cached blocks that are not wastefully large */
		int best = -1;
		for (int k = 0; k < (int) g_cache.size(); k++)
			if (g_cache[k].bytes >= bytes && g_cache[k].bytes <= 2 * bytes && (best < 0 || g_cache[k].bytes < g_cache[best].bytes))
				best = k;
		if (best >= 0) {
			p = g_cache[best].ptr;
			g_cache_bytes -= g_cache[best].bytes;
			g_live_big.push_back({p, g_cache[best].bytes});
			g_cache.erase(g_cache.begin() + best);
			return p;
		}
	}
	cudaStream_t st = ctx().stream;
	cudaError_t e = cudaMallocFromPoolAsync(&p, bytes, g_pool, st);
	if (e != cudaSuccess && !g_cache.empty()) {
		/* out of memory with blocks parked in the cache: give them back and retry */
		(void) cudaGetLastError();
		for (CachedBlock &b : g_cache)
			cudaFreeAsync(b.ptr, ctx().stream);
		g_cache.clear();
		g_cache_bytes = 0;
		CUDA_CHECK(cudaStreamSynchronize(ctx().stream));
		e = cudaMallocFromPoolAsync(&p, bytes, g_pool, ctx().stream);
	}
	CUDA_CHECK(e);
	return p;
}

void device_free(void *ptr, size_t bytes)
{
	if (g_exiting)
		return;                        /* the process is going away: the driver reclaims everything */
	if (bytes >= BIG_BLOCK) {
		size_t true_bytes = bytes;
		for (size_t k = 0; k < g_live_big.size(); k++)
			if (g_live_big[k].ptr == ptr) {
				true_bytes = g_live_big[k].bytes;
				g_live_big.erase(g_live_big.begin() + k);
				break;
			}
		if (g_cache_bytes + true_bytes <= cache_limit()) {
			g_cache.push_back({ptr, true_bytes});
			g_cache_bytes += true_bytes;
			return;
		}
	}
	cudaFreeAsync(ptr, ctx().stream);
}

size_t bulk_download_threshold()
{
	const char *e = getenv("SPASM_B200_BULK_MB");
	return e ? (size_t) (atof(e) * 1048576.0) : ((size_t) 2 << 20);
}

void download_bulk(void *host, const void *dev, size_t bytes)
{
	if (bytes == 0)
		return;
	cudaStream_t s = ctx().stream;
	constexpr int NB = 3;
	const size_t CH = (size_t) 32 << 20;
	static char *pin[NB] = {nullptr, nullptr, nullptr};
	static cudaEvent_t ev[NB];
	static int nthreads = 0;
	if (!pin[0]) {
		for (int k = 0; k < NB; k++) {
			CUDA_CHECK(cudaHostAlloc((void **) &pin[k], CH, cudaHostAllocDefault));
			CUDA_CHECK(cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming));
		}
		const char *e = getenv("SPASM_B200_COPY_THREADS");
		long online = sysconf(_SC_NPROCESSORS_ONLN);
		nthreads = e ? atoi(e) : (int) std::max(1L, std::min(8L, online / 2));
	}
#ifdef MADV_HUGEPAGE
	{
		/* 2 MB pages where the kernel grants them: 512 times fewer faults (no effect, and no harm, elsewhere) */
		const uintptr_t two_mb = (uintptr_t) 2 << 20;
		uintptr_t lo = ((uintptr_t) host + two_mb - 1) & ~(two_mb - 1), hi = ((uintptr_t) host + bytes) & ~(two_mb - 1);
		if (hi > lo)
			(void) madvise((void *) lo, hi - lo, MADV_HUGEPAGE);
	}
#endif
	const size_t nchunks = (bytes + CH - 1) / CH;
	for (size_t i = 0; i < nchunks + NB - 1; i++) {
		if (i < nchunks) {
			size_t off = i * CH, len = std::min(CH, bytes - off);
			CUDA_CHECK(cudaMemcpyAsync(pin[i % NB], (const char *) dev + off, len, cudaMemcpyDeviceToHost, s));
			CUDA_CHECK(cudaEventRecord(ev[i % NB], s));
		}
		if (i + 1 >= (size_t) NB) {
			size_t j = i + 1 - NB;          /* its slot is the one the next iteration refills */
			if (j < nchunks) {
				size_t off = j * CH, len = std::min(CH, bytes - off);
				CUDA_CHECK(cudaEventSynchronize(ev[j % NB]));
				parallel_memcpy((char *) host + off, pin[j % NB], len, nthreads);
			}
		}
	}
}

void DevCsr::upload(const struct spasm_csr *A)
{
	cudaStream_t s = ctx().stream;
	n = A->n;
	m = A->m;
	nnz = A->p[A->n];
	prime = A->field->p;
	p.upload(A->p, (size_t) n + 1, s);
	j.upload(A->j, (size_t) nnz, s);
	if (A->x)
		x.upload(A->x, (size_t) nnz, s);
	stats().pub.h2d_bytes += (i64) (n + 1) * 8 + nnz * 8;
}

}  // namespace sb

using namespace sb;

/* Creation of the primary CUDA context on a helper thread, so that it overlaps the host work that precedes the first
 * GPU call of a one-shot tool (parsing the matrix).  Errors are ignored here: ctx() reports them.  SPASM_B200_NO_WARMUP
 * turns it off. */
extern "C" void spasm_b200_warmup_async(void)
{
	static bool started = false;
	if (started || g_ctx.device >= 0 || getenv("SPASM_B200_NO_WARMUP"))
		return;
	started = true;
	/* a program that exits without ever touching the GPU waits for the helper instead of tearing the runtime down under it */
	atexit([] {
		if (g_warm && g_warm->joinable())
			g_warm->join();
	});
	g_warm = new std::thread([] {
		int count = 0;
		if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
			(void) cudaGetLastError();
			return;
		}
		if (cudaSetDevice(pick_device(count)) == cudaSuccess)
			(void) cudaFree(0);
		(void) cudaGetLastError();
	});
}


extern "C" {

int spasm_b200_device_count(void)
{
	int count = 0;
	if (cudaGetDeviceCount(&count) != cudaSuccess)
		return 0;
	return count;
}

void spasm_b200_set_device(int device) { g_requested_device = device; }

/* give the idle device memory back: the cached large blocks, then everything the library's pool holds but does not use */
void spasm_b200_trim(void)
{
	if (g_ctx.device < 0)
		return;
	for (CachedBlock &b : g_cache)
		cudaFreeAsync(b.ptr, g_ctx.stream);
	g_cache.clear();
	g_cache_bytes = 0;
	CUDA_CHECK(cudaStreamSynchronize(g_ctx.stream));
	if (g_pool)
		CUDA_CHECK(cudaMemPoolTrimTo(g_pool, 0));
}

void spasm_b200_set_verbose(int verbose)
{
	g_verbose = verbose != 0;
	if (g_ctx.device >= 0)
		g_ctx.verbose = g_verbose;
}

void spasm_b200_reset_stats(void)
{
	std::vector<int> rows, cols;
	g_stats = Stats();
}

void spasm_b200_get_stats(struct spasm_b200_stats *out) { *out = g_stats.pub; }

const char *spasm_b200_version(void) { return "spasm-b200 r1 (sm_100a)"; }

int spasm_b200_last_pivot_pairs(int *rows, int *cols, int *round_start)
{
	int n = (int) g_stats.pair_row.size();
	if (rows)
		memcpy(rows, g_stats.pair_row.data(), n * sizeof(int));
	if (cols)
		memcpy(cols, g_stats.pair_col.data(), n * sizeof(int));
	if (round_start)
		for (size_t k = 0; k < g_stats.pair_start.size(); k++)
			round_start[k] = g_stats.pair_start[k];
	return n;
}
}
