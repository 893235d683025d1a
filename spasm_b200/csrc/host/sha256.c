/*
 * SHA-256 (FIPS 180-4), written from the standard.  Keeps the context layout
 * and the three entry points of the reference (src/sha256.c, src/spasm.h:120-125,
 * :153-155): h = chaining value, (Nh:Nl) = message length in bits,
 * data = pending bytes, num = number of pending bytes.
 * Known answers: tests/Expected/hash of the reference (tests/sha.c:20-23).
 */
#include <string.h>
#include "spasm.h"
#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define SPASM_HAVE_SHANI 1
#endif

static const u32 K256[64] = {
	0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
	0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
	0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
	0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
	0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
	0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
	0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
	0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2
};

static inline u32 rotr(u32 x, int s) { return (x >> s) | (x << (32 - s)); }

/* one compression-function call on a 64-byte block */
static void compress(u32 state[8], const u8 *block)
{
	u32 w[64];
	for (int t = 0; t < 16; t++)
		w[t] = ((u32) block[4 * t] << 24) | ((u32) block[4 * t + 1] << 16) | ((u32) block[4 * t + 2] << 8) | block[4 * t + 3];
	for (int t = 16; t < 64; t++) {
		u32 s0 = rotr(w[t - 15], 7) ^ rotr(w[t - 15], 18) ^ (w[t - 15] >> 3);
		u32 s1 = rotr(w[t - 2], 17) ^ rotr(w[t - 2], 19) ^ (w[t - 2] >> 10);
		w[t] = w[t - 16] + s0 + w[t - 7] + s1;
	}
	u32 a = state[0], b = state[1], c = state[2], d = state[3];
	u32 e = state[4], f = state[5], g = state[6], h = state[7];
	for (int t = 0; t < 64; t++) {
		u32 S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25);
		u32 ch = (e & f) ^ (~e & g);
		u32 t1 = h + S1 + ch + K256[t] + w[t];
		u32 S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22);
		u32 maj = (a & b) ^ (a & c) ^ (b & c);
		u32 t2 = S0 + maj;
		h = g; g = f; f = e; e = d + t1;
		d = c; c = b; b = a; a = t1 + t2;
	}
	state[0] += a; state[1] += b; state[2] += c; state[3] += d;
	state[4] += e; state[5] += f; state[6] += g; state[7] += h;
}

#ifdef SPASM_HAVE_SHANI
/* The same compression function with the x86 SHA extensions (sha256rnds2 does two rounds, sha256msg1/msg2 the message
 * schedule), chosen at run time when the CPU has them: hashing the input file is 40 % of spasm_triplet_load with the
 * portable code (308 MB/s against ~1.8 GB/s).  State is kept in the (ABEF, CDGH) lane order the instructions expect. */
__attribute__((target("sha,sse4.1,ssse3")))
static void compress_shani(u32 state[8], const u8 *data, size_t nblocks)
{
	const __m128i bswap = _mm_set_epi64x(0x0c0d0e0f08090a0bULL, 0x0405060700010203ULL);
	__m128i tmp = _mm_loadu_si128((const __m128i *) &state[0]);          /* DCBA */
	__m128i st1 = _mm_loadu_si128((const __m128i *) &state[4]);          /* HGFE */
	tmp = _mm_shuffle_epi32(tmp, 0xB1);                                   /* CDAB */
	st1 = _mm_shuffle_epi32(st1, 0x1B);                                   /* EFGH */
	__m128i st0 = _mm_alignr_epi8(tmp, st1, 8);                           /* ABEF */
	st1 = _mm_blend_epi16(st1, tmp, 0xF0);                                /* CDGH */
	while (nblocks--) {
		const __m128i save0 = st0, save1 = st1;
		__m128i w[4];
		for (int i = 0; i < 16; i++) {
			__m128i cur;
			if (i < 4) {
				cur = _mm_shuffle_epi8(_mm_loadu_si128((const __m128i *) (data + 16 * i)), bswap);
			} else {
				/* W[t] = sigma1(W[t-2]) + W[t-7] + sigma0(W[t-15]) + W[t-16], four at a time */
				__m128i t = _mm_sha256msg1_epu32(w[(i - 4) & 3], w[(i - 3) & 3]);
				t = _mm_add_epi32(t, _mm_alignr_epi8(w[(i - 1) & 3], w[(i - 2) & 3], 4));
				cur = _mm_sha256msg2_epu32(t, w[(i - 1) & 3]);
			}
			w[i & 3] = cur;
			__m128i wk = _mm_add_epi32(cur, _mm_loadu_si128((const __m128i *) &K256[4 * i]));
			st1 = _mm_sha256rnds2_epu32(st1, st0, wk);
			wk = _mm_shuffle_epi32(wk, 0x0E);
			st0 = _mm_sha256rnds2_epu32(st0, st1, wk);
		}
		st0 = _mm_add_epi32(st0, save0);
		st1 = _mm_add_epi32(st1, save1);
		data += 64;
	}
	tmp = _mm_shuffle_epi32(st0, 0x1B);                                   /* FEBA */
	st1 = _mm_shuffle_epi32(st1, 0xB1);                                   /* DCHG */
	st0 = _mm_blend_epi16(tmp, st1, 0xF0);                                /* DCBA */
	st1 = _mm_alignr_epi8(st1, tmp, 8);                                   /* HGFE */
	_mm_storeu_si128((__m128i *) &state[0], st0);
	_mm_storeu_si128((__m128i *) &state[4], st1);
}

static int have_shani(void)
{
	static int known = -1;
	if (known < 0) {
		__builtin_cpu_init();
		known = __builtin_cpu_supports("sha") && __builtin_cpu_supports("sse4.1") && __builtin_cpu_supports("ssse3");
	}
	return known;
}
#endif

/* nblocks consecutive 64-byte blocks */
static void compress_blocks(u32 state[8], const u8 *data, size_t nblocks)
{
#ifdef SPASM_HAVE_SHANI
	if (have_shani()) {
		compress_shani(state, data, nblocks);
		return;
	}
#endif
	for (size_t b = 0; b < nblocks; b++)
		compress(state, data + 64 * b);
}

void spasm_SHA256_init(spasm_sha256_ctx *c)
{
	static const u32 iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
	memset(c, 0, sizeof(*c));
	memcpy(c->h, iv, sizeof(iv));
	c->md_len = 32;
}

void spasm_SHA256_update(spasm_sha256_ctx *c, const void *data, size_t len)
{
	const u8 *in = data;
	u8 *pending = (u8 *) c->data;

	/* bit counter, 64 bits split over (Nh, Nl) */
	u64 bits = (((u64) c->Nh) << 32) | c->Nl;
	bits += ((u64) len) << 3;
	c->Nl = (u32) bits;
	c->Nh = (u32) (bits >> 32);

	if (c->num > 0) {
		size_t room = 64 - c->num;
		size_t take = (len < room) ? len : room;
		memcpy(pending + c->num, in, take);
		c->num += take;
		in += take;
		len -= take;
		if (c->num < 64)
			return;
		compress_blocks(c->h, pending, 1);
		c->num = 0;
	}
	if (len >= 64) {
		size_t nblocks = len / 64;
		compress_blocks(c->h, in, nblocks);
		in += 64 * nblocks;
		len -= 64 * nblocks;
	}
	if (len > 0) {
		memcpy(pending, in, len);
		c->num = len;
	}
}

void spasm_SHA256_final(u8 *md, spasm_sha256_ctx *c)
{
	u8 *pending = (u8 *) c->data;
	size_t k = c->num;
	pending[k++] = 0x80;
	if (k > 56) {
		memset(pending + k, 0, 64 - k);
		compress_blocks(c->h, pending, 1);
		k = 0;
	}
	memset(pending + k, 0, 56 - k);
	for (int b = 0; b < 4; b++) {
		pending[56 + b] = (u8) (c->Nh >> (24 - 8 * b));
		pending[60 + b] = (u8) (c->Nl >> (24 - 8 * b));
	}
	compress_blocks(c->h, pending, 1);
	c->num = 0;
	for (int t = 0; t < 8; t++) {
		md[4 * t] = (u8) (c->h[t] >> 24);
		md[4 * t + 1] = (u8) (c->h[t] >> 16);
		md[4 * t + 2] = (u8) (c->h[t] >> 8);
		md[4 * t + 3] = (u8) c->h[t];
	}
}
