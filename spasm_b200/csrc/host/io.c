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

/* The stream is read to its end first (the reference reads to EOF too, warning about what follows the terminator):
 * the SHA-256 is then one pass over the buffer and the entry lines can go to the GPU in one piece (csrc/gpu/ingest.cu). */
struct reader {
	const char *text;
	size_t bytes, pos;
	i64 lineno;
	char buf[LINE_MAX_LEN];
};

static char *slurp(FILE *f, size_t *bytes)
{
	size_t cap = (size_t) 1 << 20, len = 0;
	char *text = spasm_malloc(cap);
	for (;;) {
		size_t got = fread(text + len, 1, cap - len, f);
		len += got;
		if (got == 0) {
			if (ferror(f))
				err(1, "[spasm_triplet_load] impossible to read the input");
			break;
		}
		if (len == cap) {
			cap *= 2;
			text = spasm_realloc(text, cap);
		}
	}
	*bytes = len;
	return text;
}

/* fetch the next line into r->buf; returns false at end of input (reference: src/spasm_io.c:11-26) */
static bool next_line(struct reader *r)
{
	if (r->pos >= r->bytes)
		return false;
	const char *line = r->text + r->pos;
	const char *nl = memchr(line, '\n', r->bytes - r->pos);
	size_t len = nl ? (size_t) (nl - line) + 1 : r->bytes - r->pos;
	if (len > LINE_MAX_LEN - 1)       /* what fgets(buf, LINE_MAX_LEN) cannot hold in one piece */
		errx(1, "[spasm_triplet_load] line %" PRId64 " too long (> %d)", r->lineno, LINE_MAX_LEN);
	memcpy(r->buf, line, len);
	r->buf[len] = 0;
	r->pos += len;
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
int spasm_b200_device_count(void);
/* csrc/gpu/ingest.cu: the entry lines parsed on the GPU */
i64 spasm_b200_ingest_entries(const char *text, size_t bytes, struct spasm_triplet *T, bool matrixmarket, i64 declared, i64 first_lineno);

/* Entry lines go to the GPU when there is at least this much text (SPASM_B200_INGEST_MIN_BYTES, default 1 MB: below
 * that the host loop is faster than the round trip) and a CUDA device is present; SPASM_B200_INGEST=host / gpu forces
 * one side.  I/O is not part of the echelonization path: programs that only read, transpose or print matrices keep
 * working without a GPU. */
static bool ingest_on_gpu(size_t bytes)
{
	if (bytes >= ((size_t) 1 << 32))
		return false;          /* the device pass indexes the text with 32-bit offsets: larger inputs keep the host loop */
	const char *mode = getenv("SPASM_B200_INGEST");
	if (mode != NULL && strcmp(mode, "host") == 0)
		return false;
	if (mode != NULL && strcmp(mode, "gpu") == 0)
		return true;
	const char *min = getenv("SPASM_B200_INGEST_MIN_BYTES");
	size_t threshold = min ? (size_t) atoll(min) : ((size_t) 1 << 20);
	return bytes >= threshold && spasm_b200_device_count() > 0;
}

struct spasm_triplet *spasm_triplet_load(FILE *f, i64 prime, u8 *hash)
{
	assert(f != NULL);
	/* a program that loads a matrix is about to echelonize it (tools/rank.c, echelonize.c, kernel.c): the CUDA context
	 * (about a second for a fresh process) is created while the text is parsed */
	spasm_b200_warmup_async();
	double start = spasm_wtime();
	struct reader r;
	size_t bytes = 0;
	char *text = slurp(f, &bytes);
	r.text = text;
	r.bytes = bytes;
	r.pos = 0;
	r.lineno = 0;

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
	if (ingest_on_gpu(r.bytes - r.pos)) {
		i64 garbage = spasm_b200_ingest_entries(r.text + r.pos, r.bytes - r.pos, T, mm, declared_nnz, r.lineno + 1);
		if (garbage > 0)
			warnx("[spasm_load] garbage detected near end of file");
		r.pos = r.bytes;
		finished = true;
	}
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
		/* the digest covers every byte read, like the reference's line-by-line updates (io.c:23-24) */
		spasm_sha256_ctx sha;
		spasm_SHA256_init(&sha);
		spasm_SHA256_update(&sha, text, bytes);
		i64 size = (i64) bytes;
		spasm_SHA256_final(hash, &sha);
		fprintf(stderr, "[spasm_triplet_load] sha256(matrix) = ");
		for (int k = 0; k < 32; k++)
			fprintf(stderr, "%02x", hash[k]);
		fprintf(stderr, " / size = %" PRId64 " bytes\n", size);
	}
	free(text);
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
