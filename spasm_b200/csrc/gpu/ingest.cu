/*
 * GPU-side ingest: the entry lines of an SMS / MatrixMarket text ("i j x", one per line) parsed on
 * the device.  reference: src/spasm_io.c:104-141 (the entry loop of spasm_triplet_load), src/spasm_triplet.c:7-24
 * (spasm_add_entry: reduction into the balanced range, multiples of p dropped, dimensions grown to fit).
 *
 * The host (csrc/host/io.c) reads the stream, hashes it and parses the header; this file receives the bytes that
 * follow the header.  Three passes, all streaming: (1) a flag per byte "a line starts here" and a compaction of the
 * flagged positions; (2) one thread per line: three integers with strtol's syntax (blanks, optional sign, digits),
 * status = entry / "0 0 0" terminator / malformed / too long; the position of the first line that is not an entry;
 * (3) the entries before that line whose value is not a multiple of p, compacted IN FILE ORDER (spasm_compress keeps
 * the file order inside a row and the pivot search reads it, src/spasm_pivots.c:104-116) and copied to the triplet.
 * Algorithmic bytes: the text once (pass 1) + once (pass 2) + 12 B per entry out.
 */
#include <cub/cub.cuh>
#include "common.cuh"
#include "stats.cuh"

namespace sb {

#define INGEST_LINE_MAX 1024          /* the reference reads lines with fgets into a buffer of this size (io.c:9,13) */

enum { LINE_ENTRY = 0, LINE_TERMINATOR = 1, LINE_MALFORMED = 2, LINE_TOO_LONG = 3 };

struct ByteToCount {
	__host__ __device__ i64 operator()(const unsigned char &b) const { return (i64) b; }
};

__global__ void k_ingest_flag_starts(const char *__restrict__ text, size_t bytes, unsigned char *flag)
{
	size_t b = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
	if (b < bytes)
		flag[b] = (b == 0 || text[b - 1] == '\n') ? 1 : 0;
}

/* strtol / strtoll on [s, end): blanks, optional sign, at least one digit.  Returns false when no digit is found.
 * Saturates like strtoll (the value is reduced modulo p afterwards, a saturated value is garbage in the reference too). */
__device__ __forceinline__ bool ingest_number(const char *&s, const char *end, long long &out)
{
	while (s < end && (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\v' || *s == '\f'))
		s++;
	bool neg = false;
	if (s < end && (*s == '-' || *s == '+')) {
		neg = (*s == '-');
		s++;
	}
	if (s >= end || *s < '0' || *s > '9')
		return false;
	unsigned long long acc = 0;
	bool overflow = false;
	while (s < end && *s >= '0' && *s <= '9') {
		unsigned d = (unsigned) (*s - '0');
		if (acc > (0x7fffffffffffffffULL - d) / 10)
			overflow = true;
		else
			acc = acc * 10 + d;
		s++;
	}
	if (overflow)
		out = neg ? (long long) 0x8000000000000000ULL : 0x7fffffffffffffffLL;
	else
		out = neg ? -(long long) acc : (long long) acc;
	return true;
}

__global__ void k_ingest_parse(const char *__restrict__ text, size_t bytes, const unsigned *__restrict__ start, int nlines, i64 prime,
                               int *ti, int *tj, i32 *tx, unsigned char *status, int *keep, int *first_stop)
{
	int l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= nlines)
		return;
	const size_t b0 = start[l], b1 = (l + 1 < nlines) ? start[l + 1] : bytes;      /* [b0, b1): the line with its '\n' (if any) */
	const char *s = text + b0, *end = text + b1;
	int st = LINE_ENTRY;
	long long a = 0, b = 0, c = 0;
	if (b1 - b0 > (size_t) INGEST_LINE_MAX - 1)
		st = LINE_TOO_LONG;            /* fgets(buf, 1024) would have cut the line (io.c:19-20) */
	else if (!ingest_number(s, end, a) || !ingest_number(s, end, b) || !ingest_number(s, end, c))
		st = LINE_MALFORMED;
	else if ((int) a == 0 && (int) b == 0 && c == 0)
		st = LINE_TERMINATOR;
	i32 v = 1;
	if (st == LINE_ENTRY && prime > 0) {
		long long r = c % prime;       /* spasm_ZZp_init (ZZp.c:26-30): C remainder, then one correction */
		const long long half = prime / 2, mhalf = prime / 2 - prime + 1;
		if (r > half)
			r -= prime;
		else if (r < mhalf)
			r += prime;
		v = (i32) r;
	}
	ti[l] = (int) a - 1;
	tj[l] = (int) b - 1;
	tx[l] = v;
	status[l] = (unsigned char) st;
	keep[l] = (st == LINE_ENTRY && v != 0) ? 1 : 0;
	if (st != LINE_ENTRY)
		atomicMin(first_stop, l);
}

/* entries [0, count) with keep set, in order: out[off[l]] = in[l]; also the largest row / column index and whether an
 * index is negative */
__global__ void k_ingest_compact(int count, const int *__restrict__ keep, const int *__restrict__ off, const int *__restrict__ ti,
                                 const int *__restrict__ tj, const i32 *__restrict__ tx, int *oi, int *oj, i32 *ox, int *maxima)
{
	int l = blockIdx.x * blockDim.x + threadIdx.x;
	if (l >= count || !keep[l])
		return;
	const int o = off[l];
	oi[o] = ti[l];
	oj[o] = tj[l];
	if (ox)
		ox[o] = tx[l];
	atomicMax(&maxima[0], ti[l]);
	atomicMax(&maxima[1], tj[l]);
	if (ti[l] < 0 || tj[l] < 0)
		maxima[2] = 1;
}

}  // namespace sb

using namespace sb;

/*
 * Entry lines `text[0:bytes)` of a matrix file -> entries appended to T (which must be empty and is resized to fit).
 * matrixmarket: the entries end after `declared` lines; SMS: they end with the line "0 0 0".  first_lineno = number of
 * the first line of `text` in the file (error messages).  Errors follow the reference: errx(1, ...).
 * Returns the number of lines that follow the end of the entries ("garbage").
 */
extern "C" i64 spasm_b200_ingest_entries(const char *text, size_t bytes, struct spasm_triplet *T, bool matrixmarket, i64 declared, i64 first_lineno)
{
	cudaStream_t s = ctx().stream;
	if (bytes >= ((size_t) 1 << 32))
		errx(1, "[spasm-b200] GPU ingest: %zu bytes of text exceed the 4 GB limit of one batch", bytes);
	const i64 prime = T->field->p;
	const bool valued = T->x != NULL;
	DevBuf<char> d_text(bytes + 1);
	CUDA_CHECK(cudaMemcpyAsync(d_text.ptr, text, bytes, cudaMemcpyHostToDevice, s));
	stats().pub.h2d_bytes += (i64) bytes;
	/* 1. line starts */
	DevBuf<unsigned char> flag(bytes + 1);
	k_ingest_flag_starts<<<cdiv(bytes, 256), 256, 0, s>>>(d_text.ptr, bytes, flag.ptr);
	DevBuf<unsigned> start;
	DevBuf<int> nsel(1);
	size_t tmp_bytes = 0;
	cub::CountingInputIterator<unsigned> positions(0);
	static DevBuf<char> tmp;
	/* count first: the number of lines sizes everything that follows */
	DevBuf<i64> nl(1);
	cub::TransformInputIterator<i64, ByteToCount, const unsigned char *> counts(flag.ptr, ByteToCount());
	{
		size_t tb = 0;
		cub::DeviceReduce::Sum(nullptr, tb, counts, nl.ptr, (i64) bytes, s);
		tmp.ensure(tb + 16);
		cub::DeviceReduce::Sum(tmp.ptr, tb, counts, nl.ptr, (i64) bytes, s);
	}
	const i64 nlines64 = fetch(nl.ptr);
	if (nlines64 >= ((i64) 1 << 31))
		errx(1, "[spasm-b200] GPU ingest: too many lines");
	const int nlines = (int) nlines64;
	start.ensure((size_t) nlines + 1);
	cub::DeviceSelect::Flagged(nullptr, tmp_bytes, positions, flag.ptr, start.ptr, nsel.ptr, (i64) bytes, s);
	tmp.ensure(tmp_bytes + 16);
	cub::DeviceSelect::Flagged(tmp.ptr, tmp_bytes, positions, flag.ptr, start.ptr, nsel.ptr, (i64) bytes, s);
	LAUNCHED(3);
	/* 2. parse */
	const size_t room = (size_t) std::max(nlines, 1);
	DevBuf<int> ti(room), tj(room), first_stop(1), off(room + 1), maxima(3);
	DevBuf<i32> tx(room);
	DevBuf<unsigned char> status(room);
	DevBuf<int> keep(room + 1);
	int init = nlines;
	CUDA_CHECK(cudaMemcpyAsync(first_stop.ptr, &init, sizeof(int), cudaMemcpyHostToDevice, s));
	if (nlines > 0)
		k_ingest_parse<<<cdiv((size_t) nlines, 256), 256, 0, s>>>(d_text.ptr, bytes, start.ptr, nlines, prime, ti.ptr, tj.ptr, tx.ptr, status.ptr, keep.ptr, first_stop.ptr);
	LAUNCHED(1);
	const int stop = fetch(first_stop.ptr);
	int stop_status = LINE_ENTRY;
	if (stop < nlines) {
		unsigned char st;
		CUDA_CHECK(cudaMemcpyAsync(&st, status.ptr + stop, 1, cudaMemcpyDeviceToHost, s));
		sb::sync();
		stop_status = st;
	}
	/* where the entries end, with the reference's diagnostics (io.c:108-141) */
	i64 nentry_lines;
	if (matrixmarket) {
		nentry_lines = std::min<i64>(declared, nlines);
		if (stop < nentry_lines) {
			if (stop_status == LINE_TERMINATOR)
				errx(1, "SMS end marker in MatrixMarket file");
			errx(1, "parse error line %" PRId64, first_lineno + stop);
		}
		if (nentry_lines < declared)
			errx(1, "[spasm_triplet_load] premature end of file (line %" PRId64 ", read %" PRId64 " nz)", first_lineno + nlines, nentry_lines);
	} else {
		if (stop >= nlines)
			errx(1, "[spasm_triplet_load] premature end of file (line %" PRId64 ", read %d nz)", first_lineno + nlines, nlines);
		if (stop_status == LINE_TOO_LONG)
			errx(1, "[spasm_triplet_load] line %" PRId64 " too long (> %d)", first_lineno + stop, INGEST_LINE_MAX);
		if (stop_status != LINE_TERMINATOR)
			errx(1, "parse error line %" PRId64, first_lineno + stop);
		nentry_lines = stop;
	}
	const i64 garbage = nlines - nentry_lines - (matrixmarket ? 0 : 1);
	/* 3. keep the non-zero entries, in file order */
	const int count = (int) nentry_lines;
	i64 kept = 0;
	if (count > 0) {
		size_t tb = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, tb, keep.ptr, off.ptr, count + 1, s);
		tmp.ensure(tb + 16);
		cub::DeviceScan::ExclusiveSum(tmp.ptr, tb, keep.ptr, off.ptr, count + 1, s);
		kept = fetch(off.ptr + count);
	}
	const i64 nzmax = matrixmarket ? std::max<i64>(declared, 1) : std::max<i64>(kept, 1);
	spasm_triplet_realloc(T, nzmax);
	if (kept > 0) {
		DevBuf<int> oi((size_t) kept), oj((size_t) kept);
		DevBuf<i32> ox((size_t) kept);
		int h_max[3] = {-1, -1, 0};
		CUDA_CHECK(cudaMemcpyAsync(maxima.ptr, h_max, sizeof(h_max), cudaMemcpyHostToDevice, s));
		k_ingest_compact<<<cdiv((size_t) count, 256), 256, 0, s>>>(count, keep.ptr, off.ptr, ti.ptr, tj.ptr, tx.ptr, oi.ptr, oj.ptr, valued ? ox.ptr : nullptr, maxima.ptr);
		LAUNCHED(2);
		oi.download(T->i, (size_t) kept, s);
		oj.download(T->j, (size_t) kept, s);
		if (valued)
			ox.download(T->x, (size_t) kept, s);
		CUDA_CHECK(cudaMemcpyAsync(h_max, maxima.ptr, sizeof(h_max), cudaMemcpyDeviceToHost, s));
		sb::sync();
		stats().pub.d2h_bytes += kept * (valued ? 12 : 8);
		if (h_max[2])
			errx(1, "[spasm_triplet_load] entry with a zero row or column index (indices are 1-based)");
		if (h_max[0] >= T->n)
			T->n = h_max[0] + 1;
		if (h_max[1] >= T->m)
			T->m = h_max[1] + 1;
	}
	T->nz = kept;
	return garbage;
}
