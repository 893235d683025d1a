/*
 * Entry points of the reference that lie outside the B200 hot path
 * (the L / PLUQ path, linear solves, rank certificates).
 * They are exported so that the reference's tools link unchanged
 * (tools/rank.c references the certificate functions behind --certificate);
 * calling one aborts in the reference's own error style (src/spasm_util.c:65-87).
 */
#include <err.h>
#include "spasm.h"

#define NOT_IN_SCOPE(what) errx(1, "[spasm-b200] %s is not part of the B200 echelonization path (see DESIGN.md, out of scope)", what)

bool spasm_solve(const struct spasm_lu *fact, const spasm_ZZp *b, spasm_ZZp *x)
{
	(void) fact; (void) b; (void) x;
	NOT_IN_SCOPE("spasm_solve");
}

struct spasm_csr *spasm_gesv(const struct spasm_lu *fact, const struct spasm_csr *B, bool *ok)
{
	(void) fact; (void) B; (void) ok;
	NOT_IN_SCOPE("spasm_gesv");
}

struct spasm_rank_certificate *spasm_certificate_rank_create(const struct spasm_csr *A, const u8 *hash, const struct spasm_lu *fact)
{
	(void) A; (void) hash; (void) fact;
	NOT_IN_SCOPE("spasm_certificate_rank_create");
}

bool spasm_certificate_rank_verify(const struct spasm_csr *A, const u8 *hash, const struct spasm_rank_certificate *proof)
{
	(void) A; (void) hash; (void) proof;
	NOT_IN_SCOPE("spasm_certificate_rank_verify");
}

void spasm_rank_certificate_save(const struct spasm_rank_certificate *proof, FILE *f)
{
	(void) proof; (void) f;
	NOT_IN_SCOPE("spasm_rank_certificate_save");
}

bool spasm_rank_certificate_load(FILE *f, struct spasm_rank_certificate *proof)
{
	(void) f; (void) proof;
	NOT_IN_SCOPE("spasm_rank_certificate_load");
}

bool spasm_factorization_verify(const struct spasm_csr *A, const struct spasm_lu *fact, u64 seed)
{
	(void) A; (void) fact; (void) seed;
	NOT_IN_SCOPE("spasm_factorization_verify");
}

int spasm_ffpack_LU(i64 prime, int n, int m, void *A, int ldA, spasm_datatype datatype, size_t *p, size_t *qinv)
{
	(void) prime; (void) n; (void) m; (void) A; (void) ldA; (void) datatype; (void) p; (void) qinv;
	NOT_IN_SCOPE("spasm_ffpack_LU (dense PLUQ, opts->L path)");
}
