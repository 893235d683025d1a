/*
 * Shared plumbing of the CUDA side: error convention, device buffers, device CSR, timers.
 * Everything here is internal to libspasm_b200.so; the public boundary is include/spasm.h.
 */
#pragma once
#include <cuda_runtime.h>
#include <err.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

extern "C" {
#include "spasm.h"
}

/* The reference reports fatal conditions with err()/errx() and exit status 1
 * (reference: src/spasm_util.c:65-87); CUDA failures follow the same convention. */
#define CUDA_CHECK(call)                                                                              \
	do {                                                                                              \
		cudaError_t e_ = (call);                                                                      \
		if (e_ != cudaSuccess)                                                                        \
			errx(1, "[spasm-b200] CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, \
			     cudaGetErrorString(e_));                                                             \
	} while (0)

#define KERNEL_CHECK() CUDA_CHECK(cudaGetLastError())

namespace sb {

/* -------------------------------------------------------------------- device context */
struct Context {
	int device = -1;
	int sm_count = 0;
	size_t smem_optin = 0;
	cudaStream_t stream = nullptr;
	bool verbose = true;
};
Context &ctx();            /* lazily initialised; aborts when no usable sm_100 device */

/* -------------------------------------------------------------------- statistics (include/spasm_b200.h) */
struct Stats;
Stats &stats();

/* -------------------------------------------------------------------- large results to pageable host memory
 * A cudaMemcpy into freshly malloc'ed memory runs at the speed of the page faults it triggers (about 2 GB/s), which
 * made the download of a 2.4 GB echelon form cost more than its computation.  download_bulk asks for huge pages on
 * the destination, copies chunk by chunk into pinned staging buffers and lets several host threads move each chunk
 * to its place (they take the page faults in parallel) while the next chunk is in flight. */
size_t bulk_download_threshold();      /* 2 MB; SPASM_B200_BULK_MB overrides (read at every call: the tests toggle it) */
void download_bulk(void *host, const void *dev, size_t bytes);

/* -------------------------------------------------------------------- device memory
 * cudaMallocAsync from the device pool, plus a small cache of the LARGE blocks (>= 32 MB: panels, search queues, stacked
 * dense blocks) owned by the library: returned to the pool at every call they were split by the smaller allocations
 * that followed, and every few echelonizations the pool had to map a fresh 1.4 GB block for the next panel -- a 600 ms
 * stall in the middle of a 110 ms step.  Everything runs on one stream, so a cached block can be handed out again as
 * soon as it is released. */
void *device_alloc(size_t bytes);
void device_free(void *ptr, size_t bytes);

/* -------------------------------------------------------------------- device buffers */
template <typename T> struct DevBuf {
	T *ptr = nullptr;
	size_t count = 0;
	DevBuf() {}
	explicit DevBuf(size_t n) { alloc(n); }
	DevBuf(const DevBuf &) = delete;
	DevBuf &operator=(const DevBuf &) = delete;
	DevBuf(DevBuf &&o) noexcept : ptr(o.ptr), count(o.count) { o.ptr = nullptr; o.count = 0; }
	DevBuf &operator=(DevBuf &&o) noexcept {
		if (this != &o) {
			release();
			ptr = o.ptr; count = o.count;
			o.ptr = nullptr; o.count = 0;
		}
		return *this;
	}
	~DevBuf() { release(); }
	/* stream-ordered allocation from the device's memory pool: no device-wide synchronisation on free,
	 * freed blocks are reused by the next call (the drivers allocate many short-lived work arrays) */
	void alloc(size_t n) {
		release();
		count = n;
		if (n > 0)
			ptr = (T *) device_alloc(n * sizeof(T));
	}
	/* grow (contents are NOT preserved) */
	void ensure(size_t n) { if (n > count) alloc(n); }
	void release() {
		if (ptr)
			device_free(ptr, count * sizeof(T));
		ptr = nullptr;
		count = 0;
	}
	void zero(cudaStream_t s) { if (count) CUDA_CHECK(cudaMemsetAsync(ptr, 0, count * sizeof(T), s)); }
	void fill_byte(int b, cudaStream_t s) { if (count) CUDA_CHECK(cudaMemsetAsync(ptr, b, count * sizeof(T), s)); }
	void upload(const T *host, size_t n, cudaStream_t s) {
		ensure(n);
		if (n) CUDA_CHECK(cudaMemcpyAsync(ptr, host, n * sizeof(T), cudaMemcpyHostToDevice, s));
	}
	/* small transfers are asynchronous on s; large ones take the staged path (download_bulk) and are complete on return */
	void download(T *host, size_t n, cudaStream_t s) const {
		if (n * sizeof(T) >= bulk_download_threshold())
			download_bulk(host, ptr, n * sizeof(T));
		else if (n)
			CUDA_CHECK(cudaMemcpyAsync(host, ptr, n * sizeof(T), cudaMemcpyDeviceToHost, s));
	}
	operator T *() const { return ptr; }
};

inline void sync() { CUDA_CHECK(cudaStreamSynchronize(ctx().stream)); }

template <typename T> inline T fetch(const T *dptr) {
	T v;
	CUDA_CHECK(cudaMemcpyAsync(&v, dptr, sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
	sync();
	return v;
}

/* -------------------------------------------------------------------- device CSR */
struct DevCsr {
	int n = 0, m = 0;
	i64 nnz = 0;
	DevBuf<i64> p;
	DevBuf<int> j;
	DevBuf<i32> x;
	i64 prime = 0;
	void upload(const struct spasm_csr *A);
};

inline unsigned cdiv(size_t a, size_t b) { return (unsigned) ((a + b - 1) / b); }

/* CUDA-event timer on the library's stream */
struct GpuTimer {
	cudaEvent_t a, b;
	GpuTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
	~GpuTimer() { cudaEventDestroy(a); cudaEventDestroy(b); }
	void start() { cudaEventRecord(a, ctx().stream); }
	double stop_ms() {
		cudaEventRecord(b, ctx().stream);
		cudaEventSynchronize(b);
		float ms = 0;
		cudaEventElapsedTime(&ms, a, b);
		return ms;
	}
};

}  // namespace sb
