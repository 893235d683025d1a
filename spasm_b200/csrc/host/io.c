/*
 * Text I/O: SMS ("n m M" header, 1-based "i j x" lines, "0 0 0" terminator)
 * and MatrixMarket "matrix coordinate integer general".
 * reference: src/spasm_io.c:59-196.  The optional SHA-256 covers the raw bytes
 * of every line read, exactly like the reference (io.c:23-24), because rank
 * certificates are bound to that digest.
 */
#include <stdlib.h>
#include <stdio.h>
#include <string.h>
#include <ctype.h>
#include <assert.h>
#include <err.h>
#include "spasm.h"

#define LINE_MAX_LEN 1024

struct reader {
	FILE *f;
	spasm_sha256_ctx sha;
	bool hashing;
	i64 lineno;
	char buf[LINE_MAX_LEN];
};

/* fetch the next line; returns false at end of file (reference: src/spasm_io.c:11-26) */
static bool next_line(struct reader *r)
{
	if (fgets(r->buf, LINE_MAX_LEN, r->f) == NULL) {
		if (feof(r->f))
			return false;
		err(1, "[spasm_triplet_load] impossible to read line %" PRId64, r->lineno);
	}
	size_t len = strlen(r->buf);
	if (len == 0)
		errx(1, "[spasm_triplet_load] empty line %" PRId64, r->lineno);
	if (r->buf[len - 1] != '\n' && !feof(r->f))
		errx(1, "[spasm_triplet_load] line %" PRId64 " too long (> %d)", r->lineno, LINE_MAX_LEN);
	if (r->hashing)
		spasm_SHA256_update(&r->sha, r->buf, len);
	return true;
}

static void lowercase(char *s)
{
	for (; *s; s++)
		*s = (char) tolower((unsigned char) *s);
}

/* reference: src/spasm_io.c:28-52 */
static void check_matrixmarket_banner(const char *line)
{
	char object[LINE_MAX_LEN], format[LINE_MAX_LEN], field[LINE_MAX_LEN], symmetry[LINE_MAX_LEN];
	if (sscanf(line, "%%%%MatrixMarket %s %s %s %s", object, format, field, symmetry) != 4)
		errx(1, "incomplete MatrixMarket header");
	lowercase(object);
	lowercase(format);
	lowercase(field);
	lowercase(symmetry);
	if (strcmp(object, "matrix") != 0)
		errx(1, "unsupported MatrixMarket object type %s (I only know about ``matrix'')", object);
	if (strcmp(format, "coordinate") != 0)
		errx(1, "unsupported MatrixMarket format %s (I only know about ``coordinate'')", format);
	if (strcmp(field, "integer") != 0)
		errx(1, "unsupported MatrixMarket data type %s (I only know about ``integer'')", field);
	if (strcmp(symmetry, "general") != 0)
		errx(1, "unsupported MatrixMarket storage scheme %s (I only know about ``general'')", symmetry);
}

/* parse "i j x"; returns false on malformed input */
static bool parse_entry(const char *s, int *i, int *j, i64 *x)
{
	char *end;
	long a = strtol(s, &end, 10);
	if (end == s)
		return false;
	s = end;
	long b = strtol(s, &end, 10);
	if (end == s)
		return false;
	s = end;
	long long c = strtoll(s, &end, 10);
	if (end == s)
		return false;
	*i = (int) a;
	*j = (int) b;
	*x = (i64) c;
	return true;
}

/*
 * reference: src/spasm_io.c:59-159.  prime == -1 loads the pattern only.
 */
/* csrc/gpu/device.cu: starts the creation of the CUDA context on a helper thread (once; no effect without a GPU) */
void spasm_b200_warmup_async(void);

struct spasm_triplet *spasm_triplet_load(FILE *f, i64 prime, u8 *hash)
{
	assert(f != NULL);
	/* a program that loads a matrix is about to echelonize it (tools/rank.c, echelonize.c, kernel.c): the CUDA context
	 * (about a second for a fresh process) is created while the text is parsed */
	spasm_b200_warmup_async();
	double start = spasm_wtime();
	struct reader r;
	r.f = f;
	r.hashing = (hash != NULL);
	r.lineno = 0;
	spasm_SHA256_init(&r.sha);

	if (!next_line(&r))
		errx(1, "[spasm_triplet_load] empty file\n");

	int n, m;
	i64 declared_nnz = 1;
	bool mm = (strncmp(r.buf, "%%MatrixMarket", 14) == 0);
	char hnnz[16];
	if (mm) {
		check_matrixmarket_banner(r.buf);
		do {    /* skip the comment block */
			r.lineno += 1;
			if (!next_line(&r))
				errx(1, "premature EOF on line %" PRId64 " (expected matrix dimensions)", r.lineno);
		} while (r.buf[0] == '%');
		if (sscanf(r.buf, "%d %d %" SCNd64 "\n", &n, &m, &declared_nnz) != 3)
			errx(1, "[spasm_triplet_load] bad MatrixMarking dimensions (line %" PRId64 ")\n", r.lineno);
		spasm_human_format(declared_nnz, hnnz);
		fprintf(stderr, "[IO] loading %d x %d MatrixMarket matrix modulo %" PRId64 " with %s non-zero... ", n, m, prime, hnnz);
	} else {
		char type;
		if (sscanf(r.buf, "%d %d %c\n", &n, &m, &type) != 3)
			errx(1, "[spasm_triplet_load] bad SMS file (header)\n");
		if (prime != -1 && type != 'M')
			errx(1, "[spasm_triplet_load] only ``Modular'' type supported\n");
		fprintf(stderr, "[IO] loading %d x %d SMS matrix modulo %" PRId64 "... ", n, m, prime);
	}
	fflush(stderr);

	struct spasm_triplet *T = spasm_triplet_alloc(n, m, declared_nnz, prime, prime != -1);
	bool finished = false;
	i64 entries = 0;
	for (;;) {
		r.lineno += 1;
		bool got = next_line(&r);
		if (!got) {
			if (finished)
				break;
			errx(1, "[spasm_triplet_load] premature end of file (line %" PRId64 ", read %" PRId64 " nz)", r.lineno, entries);
		}
		if (finished) {
			warn("[spasm_load] garbage detected near end of file");
			continue;
		}
		int i, j;
		i64 x;
		if (!parse_entry(r.buf, &i, &j, &x))
			errx(1, "parse error line %" PRId64, r.lineno);
		if (i == 0 && j == 0 && x == 0) {
			if (mm)
				errx(1, "SMS end marker in MatrixMarket file");
			finished = true;
		} else {
			spasm_add_entry(T, i - 1, j - 1, x);
			entries += 1;
		}
		if (mm && entries == declared_nnz)
			finished = true;
	}

	if (mm) {
		fprintf(stderr, "[%.1fs]\n", spasm_wtime() - start);
	} else {
		spasm_triplet_realloc(T, -1);
		spasm_human_format(T->nz, hnnz);
		fprintf(stderr, "%s non-zero [%.1fs]\n", hnnz, spasm_wtime() - start);
	}

	if (hash != NULL) {
		i64 size = (((i64) r.sha.Nh) << 29) + r.sha.Nl / 8;
		spasm_SHA256_final(hash, &r.sha);
		fprintf(stderr, "[spasm_triplet_load] sha256(matrix) = ");
		for (int k = 0; k < 32; k++)
			fprintf(stderr, "%02x", hash[k]);
		fprintf(stderr, " / size = %" PRId64 " bytes\n", size);
	}
	return T;
}

/* reference: src/spasm_io.c:164-180 */
void spasm_csr_save(const struct spasm_csr *A, FILE *f)
{
	assert(f != NULL);
	fprintf(f, "%d %d M\n", A->n, A->m);
	for (int i = 0; i < A->n; i++)
		for (i64 k = A->p[i]; k < A->p[i + 1]; k++) {
			i64 v = (A->x != NULL) ? A->x[k] : 1;
			fprintf(f, "%d %d %" PRId64 "\n", i + 1, A->j[k] + 1, v);
		}
	fprintf(f, "0 0 0\n");
}

/* reference: src/spasm_io.c:185-196 */
void spasm_triplet_save(const struct spasm_triplet *A, FILE *f)
{
	assert(f != NULL);
	fprintf(f, "%d %d M\n", A->n, A->m);
	for (i64 k = 0; k < A->nz; k++)
		fprintf(f, "%d %d %d\n", A->i[k] + 1, A->j[k] + 1, (A->x != NULL) ? A->x[k] : 1);
	fprintf(f, "0 0 0\n");
}
