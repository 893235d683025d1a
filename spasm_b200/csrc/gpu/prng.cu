/*
 * Coefficients of the randomized Schur complement, generated on the device.
 *
 * The reference draws them on the CPU from a SHA-256 counter-mode PRNG re-seeded for every
 * output row k with (prime, seed = k, sequence = 0) (reference: src/spasm_schur.c:368-385,
 * src/spasm_prng.c).  The streams of different rows are independent, so one thread per row
 * reproduces them bit for bit: message block = seed[32] | prime | counter | sequence (big
 * endian), eight 32-bit outputs per hash, masked rejection sampling into [0, p), balanced
 * representative.  Known answers: tests/Expected/prng of the reference (checked through
 * spasm_b200_prng_stream in tests/test_gpu_components.py).
 */
#include "engine.cuh"
#include "stats.cuh"

namespace sb {

__constant__ uint32_t SHA_K[64] = {
	0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
	0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
	0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
	0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
	0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
	0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
	0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t rotr32(uint32_t x, int s) { return __funnelshift_r(x, x, s); }

/* SHA-256 of the 44-byte block (seed_lo, seed_hi, 0 x 6, prime, counter, seq), big-endian words */
__device__ void sha256_prng_block(uint32_t seed_lo, uint32_t seed_hi, uint32_t prime, uint32_t counter, uint32_t seq, uint32_t H[8])
{
	uint32_t w[64];
	w[0] = seed_lo; w[1] = seed_hi;
#pragma unroll
	for (int t = 2; t < 8; t++)
		w[t] = 0;
	w[8] = prime; w[9] = counter; w[10] = seq;
	w[11] = 0x80000000u; w[12] = 0; w[13] = 0; w[14] = 0; w[15] = 44 * 8;
#pragma unroll
	for (int t = 16; t < 64; t++)
		w[t] = w[t - 16] + w[t - 7] + (rotr32(w[t - 15], 7) ^ rotr32(w[t - 15], 18) ^ (w[t - 15] >> 3))
		     + (rotr32(w[t - 2], 17) ^ rotr32(w[t - 2], 19) ^ (w[t - 2] >> 10));
	uint32_t a = 0x6a09e667, b = 0xbb67ae85, c = 0x3c6ef372, d = 0xa54ff53a, e = 0x510e527f, f = 0x9b05688c, g = 0x1f83d9ab, h = 0x5be0cd19;
#pragma unroll
	for (int t = 0; t < 64; t++) {
		uint32_t t1 = h + (rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25)) + ((e & f) ^ (~e & g)) + SHA_K[t] + w[t];
		uint32_t t2 = (rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
		h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
	}
	H[0] = 0x6a09e667 + a; H[1] = 0xbb67ae85 + b; H[2] = 0x3c6ef372 + c; H[3] = 0xa54ff53a + d;
	H[4] = 0x510e527f + e; H[5] = 0x9b05688c + f; H[6] = 0x1f83d9ab + g; H[7] = 0x5be0cd19 + h;
}

/* warp k: out[k*stride + t] = t-th output of the stream (prime, seed0 + k, seq), t < count.  The generator is SHA-256 in
 * counter mode with rejection sampling (reference: src/spasm_prng.c): the 32 lanes hash 32 consecutive counters at once,
 * the accepted words are ranked with a warp prefix sum (lane order = counter order, word order inside a digest), and
 * the first `count` of them are the stream.  One thread per stream hashed ~50 blocks in sequence: 1 ms per call on the
 * critical path of every randomized block. */
__global__ void k_prng_streams(int N, int count, uint64_t seed0, uint32_t seq, uint32_t prime, uint32_t mask, Zp F, i32 *out, int stride)
{
	const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
	if (k >= N)
		return;
	const uint64_t seed = seed0 + (uint64_t) k;
	int produced = 0;
	uint32_t base = 0;
	while (produced < count) {
		uint32_t H[8];
		sha256_prng_block((uint32_t) seed, (uint32_t) (seed >> 32), prime, base + lane, seq, H);
		int mine = 0;
#pragma unroll
		for (int u = 0; u < 8; u++)
			mine += ((H[u] & mask) < prime);
		int pre = mine;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			int o = __shfl_up_sync(0xffffffffu, pre, d);
			if (lane >= d)
				pre += o;
		}
		const int total = __shfl_sync(0xffffffffu, pre, 31);
		int at = produced + pre - mine;
#pragma unroll
		for (int u = 0; u < 8; u++) {
			const uint32_t v = H[u] & mask;
			if (v < prime) {
				if (at < count)
					out[(size_t) k * stride + at] = zp_reduce((i64) v, F);
				at++;
			}
		}
		produced += total;
		base += 32;
	}
}

__global__ void k_fill_first(int N, int stride, i32 *out)
{
	int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < N)
		out[(size_t) k * stride] = 1;
}

/* coefficients of N random combinations of w rows: 1 for the first row, PRNG (prime, k, 0) for the others
 * (reference: src/spasm_schur.c:381-385) */
void prng_combo_coefficients(i64 prime, int N, int w, i32 *d_coef)
{
	if (N <= 0 || w <= 0)
		return;
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	i64 pow2 = 1;
	while (pow2 < prime)
		pow2 <<= 1;
	k_fill_first<<<cdiv(N, 128), 128, 0, s>>>(N, w, d_coef);
	/* the stream of row k starts at its SECOND coefficient: slot t >= 1 receives output number t - 1 */
	if (w > 1)
		k_prng_streams<<<cdiv((size_t) N * 32, 128), 128, 0, s>>>(N, w - 1, 0, 0, (uint32_t) prime, (uint32_t) (pow2 - 1), F, d_coef + 1, w);
	LAUNCHED(2);
	KERNEL_CHECK();
}

}  // namespace sb

extern "C" void spasm_b200_prng_stream(i64 prime, u64 seed, u32 seq, int count, i32 *out_host)
{
	using namespace sb;
	ctx();
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	i64 pow2 = 1;
	while (pow2 < prime)
		pow2 <<= 1;
	DevBuf<i32> d((size_t) std::max(count, 1));
	k_prng_streams<<<1, 32, 0, s>>>(1, count, seed, seq, (uint32_t) prime, (uint32_t) (pow2 - 1), F, d.ptr, count);
	LAUNCHED(1);
	d.download(out_host, (size_t) count, s);
	sb::sync();
}
