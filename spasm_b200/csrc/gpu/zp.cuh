/*
 * Arithmetic in Z/pZ on the device, balanced representatives in int32
 * (same value set as the reference's spasm_ZZp, src/spasm_ZZp.c).
 *
 * The reference reduces a*x+y with a double-precision quotient estimate and
 * one correction (ZZp.c:77-84).  On the GPU we accumulate exact 64-bit
 * integers and reduce with a 64-bit Barrett step; the balanced representative
 * of a residue is unique, so the stored values are bit-identical.
 *
 * Delayed reduction: a sum of `delay` products of balanced values fits an
 * int64, so inner loops reduce once every `delay` terms (delay is huge for
 * p = 42013 and 1 for 32-bit primes).
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sb {

struct Zp {
	uint32_t p;
	int64_t half;       /* largest representative  */
	int64_t mhalf;      /* smallest representative */
	uint64_t inv64;     /* floor(2^64 / p) */
	int delay;          /* products that may be accumulated on top of a reduced value */
};

inline Zp make_zp(int64_t prime)
{
	Zp F;
	F.p = (uint32_t) prime;
	F.half = prime / 2;
	F.mhalf = prime / 2 - prime + 1;
	F.inv64 = (uint64_t) ((((unsigned __int128) 1) << 64) / (unsigned __int128) prime);
	/* |v| <= A := max(half, -mhalf); need (delay + 1) * A^2 < 2^63 */
	int64_t A = F.half > -F.mhalf ? F.half : -F.mhalf;
	if (A < 1)
		A = 1;
	unsigned __int128 cap = ((unsigned __int128) 1 << 63) / ((unsigned __int128) A * (unsigned __int128) A);
	int64_t d = (cap > (unsigned __int128) (1 << 30)) ? (1 << 30) : (int64_t) cap;
	F.delay = (int) (d > 2 ? d - 1 : 1);
	return F;
}

/* any |v| < 2^63  ->  balanced representative */
__host__ __device__ __forceinline__ int32_t zp_reduce(int64_t v, const Zp &F)
{
	uint64_t a = v < 0 ? (uint64_t) (-v) : (uint64_t) v;
#ifdef __CUDA_ARCH__
	uint64_t q = __umul64hi(a, F.inv64);
#else
	uint64_t q = (uint64_t) (((unsigned __int128) a * F.inv64) >> 64);
#endif
	uint64_t r = a - q * (uint64_t) F.p;
	if (r >= F.p)
		r -= F.p;
	int64_t s = v < 0 ? -(int64_t) r : (int64_t) r;
	if (s > F.half)
		s -= F.p;
	else if (s < F.mhalf)
		s += F.p;
	return (int32_t) s;
}

__host__ __device__ __forceinline__ int32_t zp_mul(int32_t a, int32_t b, const Zp &F)
{
	return zp_reduce((int64_t) a * (int64_t) b, F);
}

__host__ __device__ __forceinline__ int32_t zp_add(int32_t a, int32_t b, const Zp &F)
{
	int64_t s = (int64_t) a + (int64_t) b;
	if (s > F.half)
		s -= F.p;
	else if (s < F.mhalf)
		s += F.p;
	return (int32_t) s;
}

/* extended Euclid, same element as the reference's spasm_ZZp_inverse (ZZp.c:49-74) */
__host__ __device__ inline int32_t zp_inverse(int32_t a, const Zp &F)
{
#ifdef __CUDA_ARCH__
	/* same remainder sequence with 32-bit unsigned divisions (p < 2^32): a 64-bit division is a ~100-instruction
	 * routine on the GPU and this runs on the critical path of the dense panel kernel */
	uint32_t u0 = (uint32_t) F.p, u1 = (uint32_t) (a < 0 ? (int64_t) a + F.p : (int64_t) a);
	int64_t s0 = 0, s1 = 1;
	while (u1 != 0) {
		uint32_t q = u0 / u1, u2 = u0 - q * u1;
		int64_t s2 = s0 - (int64_t) q * s1;
		u0 = u1; u1 = u2; s0 = s1; s1 = s2;
	}
	return zp_reduce(s0, F);
#endif
	int64_t r0 = F.p, r1 = a < 0 ? (int64_t) a + F.p : a, t0 = 0, t1 = 1;
	while (r1 != 0) {
		int64_t q = r0 / r1;
		int64_t r2 = r0 - q * r1, t2 = t0 - q * t1;
		r0 = r1; r1 = r2; t0 = t1; t1 = t2;
	}
	return zp_reduce(t0, F);
}

#ifdef __CUDACC__
/* x[idx] <- x[idx] + v (mod p), atomically; used where several sources may hit the same slot */
__device__ __forceinline__ void zp_atomic_add(int32_t *addr, int32_t v, const Zp &F)
{
	int old = *addr, assumed;
	do {
		assumed = old;
		old = atomicCAS(addr, assumed, zp_add(assumed, v, F));
	} while (old != assumed);
}
#endif

}  // namespace sb
