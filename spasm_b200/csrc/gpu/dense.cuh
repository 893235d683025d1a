/*
 * Dense linear algebra mod p on the device: the replacement of the
 * FFLAS-FFPACK boundary (reference: src/spasm_ffpack.cpp:22-44, :78-86).
 *
 * dense_rref: block Gauss-Jordan with column-rank-profile pivoting on a
 * row-major int32 matrix.  Panels of 32 columns are factorised by one CTA
 * (CUDA cores, latency bound); the trailing update  S[:, c0:] -= W * P  is a
 * dense product mod p and goes through dense_gemm_sub.
 * The pivot columns are the column rank profile (what a reduced row echelon
 * form defines uniquely), in increasing order; the RREF rows are unique.
 */
#pragma once
#include "common.cuh"
#include "zp.cuh"

namespace sb {

/* C (M x N, ldc) -= A (M x K, lda) * B (K x N, ldb)  over Z/pZ, all row-major balanced int32.
 * If d_K != NULL the inner dimension is read from device memory (<= K). */
void dense_gemm_sub(i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F, const int *d_K = nullptr);

/* the same product on the int8 tensor cores (umma_gemm.cu): tcgen05.mma kind::i8 on signed byte limbs */
bool umma_gemm_available(const Zp &F);
void umma_gemm_sub(i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F);
/* operands pre-split into int8 limb planes, tile-packed in the UMMA shared-memory layout (umma_gemm.cu, version 2) */
size_t umma_packed_bytes(int rows, int K, int L);
int umma_limbs(const Zp &F);
void umma_pack(const i32 *src, int ld, int rows, int K, bool rowmajor_k, int8_t *dst, const Zp &F);
void umma_gemm_sub_packed(i32 *C, int ldc, const int8_t *Ap, const int8_t *Bp, int M, int N, int K, const Zp &F);

struct RrefResult {
	int rank = 0;
	std::vector<int> pivcol;    /* pivot column of RREF row t, increasing */
	std::vector<int> pivrow;    /* physical row of S holding RREF row t */
};

/* In-place reduced row echelon form of S (n x m, leading dimension ld). */
RrefResult dense_rref(i32 *S, int n, int m, int ld, const Zp &F);

/* dst[r*ldd + t] = src[r*lds + cols[t]] */
void dense_gather_columns(const i32 *src, int lds, int rows, const int *d_cols, int ncols, i32 *dst, int ldd);
/* dst[t*ldd + :] = src[rows[t]*lds + :]  (width entries) */
void dense_gather_rows(const i32 *src, int lds, const int *d_rows, int nrows, int width, i32 *dst, int ldd);

/* dst[t*ldd + cols[c]] = src[rows[t]*lds + c]  for c < ncols (rows of a column-compacted matrix expanded back) */
void dense_scatter_rows(const i32 *src, int lds, const int *d_rows, int nrows, const int *d_cols, int ncols, i32 *dst, int ldd);

/* sparse rows (reference: src/spasm_echelonize.c:192-223, update_U_after_rref): for each of the nrows rows of D
 * (ld), the entries on columns c with skip[c] == 0 that are non-zero, as (colmap[c], value), by increasing c,
 * preceded by (colmap[pivcol[t]], 1).  d_skip == NULL skips the row's own pivot column only (rows that are not reduced
 * against each other: the L mode).  Output CSR arrays on the device. */
void dense_rows_to_csr(const i32 *D, int ld, int nrows, int width, const int *d_pivcol, const unsigned char *d_skip,
                       const int *d_colmap, DevBuf<i64> &Rp, DevBuf<int> &Rj, DevBuf<i32> &Rx, i64 &nnz);

}  // namespace sb
