/*
 * Host-side arithmetic in Z/pZ with balanced representatives.
 * Mirrors the semantics of the reference's src/spasm_ZZp.c (same results for
 * every input the reference accepts), written with exact integer arithmetic:
 * where the reference estimates the quotient in double precision
 * (spasm_ZZp.c:44, :79) we take the mathematical remainder, which is the
 * same element since the balanced representative is unique.
 *
 * These are scalar helpers for I/O, U-row normalisation bookkeeping and the
 * tests; the bulk arithmetic of the hot path runs on the GPU (gpu/zp.cuh).
 */
#include <assert.h>
#include "spasm.h"

/* reference: src/spasm_ZZp.c:5-15 */
void spasm_field_init(i64 p, spasm_field F)
{
	F->p = p;
	if (p < 0)
		return;            /* "pattern only" matrices carry p == -1 */
	assert(p >= 2);
	assert(p <= 0xfffffffbLL);
	F->halfp = p / 2;
	F->mhalfp = p / 2 - p + 1;
	F->dinvp = 1.0 / (double) p;
}

/* bring any x with |x| < p into the balanced range */
static inline spasm_ZZp balance(const spasm_field F, i64 x)
{
	if (x > F->halfp)
		x -= F->p;
	else if (x < F->mhalfp)
		x += F->p;
	return (spasm_ZZp) x;
}

/* reference: src/spasm_ZZp.c:26-30 */
spasm_ZZp spasm_ZZp_init(const spasm_field F, i64 x)
{
	return balance(F, x % F->p);
}

/* reference: src/spasm_ZZp.c:32-40 */
spasm_ZZp spasm_ZZp_add(const spasm_field F, spasm_ZZp a, spasm_ZZp b)
{
	return balance(F, (i64) a + (i64) b);
}

spasm_ZZp spasm_ZZp_sub(const spasm_field F, spasm_ZZp a, spasm_ZZp b)
{
	return balance(F, (i64) a - (i64) b);
}

/* reference: src/spasm_ZZp.c:42-46.  |a*b| < 2^62 so the product is exact in i64. */
spasm_ZZp spasm_ZZp_mul(const spasm_field F, spasm_ZZp a, spasm_ZZp b)
{
	return balance(F, ((i64) a * (i64) b) % F->p);
}

/* reference: src/spasm_ZZp.c:49-74 -- extended Euclid on (a mod p, p) */
spasm_ZZp spasm_ZZp_inverse(const spasm_field F, spasm_ZZp a)
{
	i64 r0 = F->p, r1 = a;
	if (r1 < 0)
		r1 += F->p;
	i64 t0 = 0, t1 = 1;
	while (r1 != 0) {
		i64 q = r0 / r1;
		i64 r2 = r0 - q * r1;
		i64 t2 = t0 - q * t1;
		r0 = r1; r1 = r2;
		t0 = t1; t1 = t2;
	}
	/* r0 == gcd == 1 for a != 0; |t0| < p */
	return balance(F, t0 % F->p);
}

/* reference: src/spasm_ZZp.c:77-84 : a*x + y */
spasm_ZZp spasm_ZZp_axpy(const spasm_field F, spasm_ZZp a, spasm_ZZp x, spasm_ZZp y)
{
	return balance(F, ((i64) a * (i64) x + (i64) y) % F->p);
}
