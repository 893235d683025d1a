/*
 * Deterministic PRNG = SHA-256 in counter mode over a 44-byte block
 * (32-byte seed | prime | counter | sequence number, all big endian), with
 * rejection sampling to get uniform elements of Z/pZ.
 * reference: src/spasm_prng.c.  Known answers: tests/Expected/prng.
 *
 * The randomized Schur complement (reference: src/spasm_schur.c:368-385) draws
 * its coefficients from this stream, re-seeded per output row, so the stream
 * is part of the result and must match the reference bit for bit.
 */
#include "spasm.h"

static inline u32 be32(u32 x)       /* host order <-> big endian */
{
	const union { u32 word; u8 byte[4]; } probe = {1};
	if (probe.byte[0] == 0)
		return x;
	return (x >> 24) | ((x >> 8) & 0xff00) | ((x << 8) & 0xff0000) | (x << 24);
}

/* refill ctx->hash with SHA256(block), then bump the counter (reference: src/spasm_prng.c:7-16) */
static void refill(spasm_prng_ctx *ctx)
{
	spasm_sha256_ctx h;
	spasm_SHA256_init(&h);
	spasm_SHA256_update(&h, ctx->block, 44);
	spasm_SHA256_final((u8 *) ctx->hash, &h);
	ctx->counter += 1;
	ctx->block[9] = be32((u32) ctx->counter);
	ctx->i = 0;
}

/* reference: src/spasm_prng.c:21-28 */
u32 spasm_prng_u32(spasm_prng_ctx *ctx)
{
	if (ctx->i == 8)
		refill(ctx);
	return be32(ctx->hash[ctx->i++]);
}

/* reference: src/spasm_prng.c:33-40 */
spasm_ZZp spasm_prng_ZZp(spasm_prng_ctx *ctx)
{
	u32 x;
	do {
		x = spasm_prng_u32(ctx) & ctx->mask;
	} while (x >= ctx->prime);
	return spasm_ZZp_init(ctx->field, x);
}

/* reference: src/spasm_prng.c:45-61 */
void spasm_prng_seed(const u8 *seed, i64 prime, u32 seq, spasm_prng_ctx *ctx)
{
	u8 *bytes = (u8 *) ctx->block;
	for (int k = 0; k < 32; k++)
		bytes[k] = seed[k];
	ctx->prime = (u32) prime;
	i64 pow2 = 1;
	while (pow2 < prime)
		pow2 <<= 1;
	ctx->mask = (u32) (pow2 - 1);
	ctx->block[8] = be32((u32) prime);
	ctx->block[9] = 0;
	ctx->block[10] = be32(seq);
	ctx->counter = 0;
	spasm_field_init(prime, ctx->field);
	refill(ctx);
}

/* reference: src/spasm_prng.c:66-73 */
void spasm_prng_seed_simple(i64 prime, u64 seed, u32 seq, spasm_prng_ctx *ctx)
{
	u32 words[8] = {be32((u32) (seed & 0xffffffff)), be32((u32) (seed >> 32)), 0, 0, 0, 0, 0, 0};
	spasm_prng_seed((const u8 *) words, prime, seq, ctx);
}
