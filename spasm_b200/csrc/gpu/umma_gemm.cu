/*
 * C -= A * B  over Z/pZ on the 5th-generation tensor cores (tcgen05, kind::i8), sm_100a.
 *
 * This is the dense contraction of the echelonization path: the trailing update of the dense
 * echelon form and the reduction of a new block by the rows of the earlier dense blocks
 * (reference: inside FFLAS-FFPACK for src/spasm_ffpack.cpp, and -- entry by entry -- in
 * spasm_scatter for src/spasm_schur.c:291-304, :389-394).
 *
 * Exact modular arithmetic on int8 tensor cores by limb splitting with delayed reduction:
 *   a balanced residue v (|v| <= (p-1)/2) is written  v = sum_i l_i * 256^i,  l_i in [-128, 127]
 *   (L = 2 limbs for p = 42013, 4 limbs for 31/32-bit primes), so
 *       A*B = sum_{i,j} 256^(i+j) * (A_i * B_j)
 *   and every A_i * B_j is an int8 x int8 -> int32 tensor-core product.  Products of equal weight i+j
 *   share one accumulator in tensor memory (2L-1 accumulators of BM x BN int32); |l_i * l_j| <= 2^14, so
 *   a K-extent of up to 2^31 / (L * 2^14) is accumulated without overflow (K is chunked at 16384).
 *   The epilogue recombines the weight classes in 64-bit integers and applies the Barrett reduction.
 *
 * Kernel anatomy (one CTA = one 128 x BN tile of C, 256 threads):
 *   - all threads stage the operands: 16 consecutive K-entries of one row (A) / one column (B) are
 *     loaded as int32 from HBM, split into limbs in registers, and each limb plane is written as one
 *     16-byte store into shared memory in the canonical K-major no-swizzle UMMA layout
 *     (8 x 16-byte core matrices; LBO = 128 B between the K halves, SBO = (BK/16) * 128 B between
 *     8-row groups).  TMA cannot be used for this copy: the limb split happens on the way.
 *   - two shared-memory stages; an mbarrier per stage is armed by tcgen05.commit, so the copy of
 *     K-block k+1 overlaps the MMAs of block k;
 *   - one elected thread issues the L*L*(BK/32) tcgen05.mma instructions of a K-block;
 *   - epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> recombine ->
 *     C = (C - sum) mod p, written back to HBM.
 */
#include "dense.cuh"
#include "stats.cuh"

namespace sb {

#define UM_BM 128
#define UM_BK 64
#define UM_THREADS 256

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "WAIT_LOOP:\n\t"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
	    "@p bra DONE;\n\t"
	    "bra WAIT_LOOP;\n\t"
	    "DONE:\n\t"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

/* shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor layout):
 * bits [0,14) start address >> 4, [16,30) leading byte offset >> 4, [32,46) stride byte offset >> 4,
 * [46,48) version = 1 (Blackwell), [61,64) layout type = 0 (SWIZZLE_NONE) */
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
	uint64_t d = 0;
	d |= (uint64_t) ((saddr & 0x3ffff) >> 4);
	d |= (uint64_t) ((lbo_bytes >> 4) & 0x3fff) << 16;
	d |= (uint64_t) ((sbo_bytes >> 4) & 0x3fff) << 32;
	d |= (uint64_t) 1 << 46;
	return d;
}

/* instruction descriptor for kind::i8 (cute::UMMA::InstrDescriptor): c = S32 (2) at [4,6), a,b signed int8 (1) at
 * [7,10) and [10,13), both operands K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29) */
__device__ __forceinline__ uint32_t umma_idesc_i8(int M, int N)
{
	return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate)
{
	asm volatile(
	    "{\n\t"
	    ".reg .pred p;\n\t"
	    "setp.ne.b32 p, %4, 0;\n\t"
	    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
	    "}\n" ::"r"(tmem_c),
	    "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
	    : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
	asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

/* split 16 balanced residues into L signed byte planes, one 16-byte vector per plane */
template <int L> __device__ __forceinline__ void split16(const i32 (&v)[16], uint4 (&plane)[L])
{
	uint32_t w[L][4];
#pragma unroll
	for (int l = 0; l < L; l++)
#pragma unroll
		for (int q = 0; q < 4; q++)
			w[l][q] = 0;
#pragma unroll
	for (int e = 0; e < 16; e++) {
		i32 x = v[e];
#pragma unroll
		for (int l = 0; l < L; l++) {
			i32 limb = (i32) (int8_t) (x & 0xff);      /* in [-128, 127] */
			x = (x - limb) >> 8;                        /* exact: x - limb is a multiple of 256 */
			w[l][e >> 2] |= ((uint32_t) (limb & 0xff)) << (8 * (e & 3));
		}
	}
#pragma unroll
	for (int l = 0; l < L; l++)
		plane[l] = make_uint4(w[l][0], w[l][1], w[l][2], w[l][3]);
}

/*
 * L limbs, BN columns per CTA.  Shared memory per stage: L planes of A (128 x 64 B) and L planes of B (BN x 64 B).
 * Plane layout (bytes): core matrix (g = row / 8, kb = k / 16) at (g * (UM_BK / 16) + kb) * 128, row r % 8 at + 16 * (r % 8).
 */
template <int L, int BN>
__global__ void __launch_bounds__(UM_THREADS, 1)
k_umma_gemm_sub(i32 *__restrict__ C, int ldc, const i32 *__restrict__ A, int lda, const i32 *__restrict__ B, int ldb,
                int M, int N, int K, Zp F, i32 w256_0, i32 w256_1, i32 w256_2, i32 w256_3, i32 w256_4, i32 w256_5, i32 w256_6)
{
	constexpr int NCLASS = 2 * L - 1;
	constexpr int A_PLANE = UM_BM * UM_BK;           /* bytes */
	constexpr int B_PLANE = BN * UM_BK;
	constexpr int STAGE = L * (A_PLANE + B_PLANE);
	constexpr int TMEM_COLS = (NCLASS * BN <= 32) ? 32 : (NCLASS * BN <= 64) ? 64 : (NCLASS * BN <= 128) ? 128 : (NCLASS * BN <= 256) ? 256 : 512;
	static_assert(NCLASS * BN <= 512, "accumulators do not fit tensor memory");
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t bar[2];
	__shared__ uint32_t tmem_base_slot;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;

	if (warp == 0) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t) TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 0) {
		mbar_init(&bar[0], 1);
		mbar_init(&bar[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = tmem_base_slot;
	const uint32_t idesc = umma_idesc_i8(UM_BM, BN);

	const int nkb = (K + UM_BK - 1) / UM_BK;
	uint32_t phase[2] = {0, 0};
	for (int kb = 0; kb < nkb; kb++) {
		const int st = kb & 1;
		unsigned char *stage = smem + (size_t) st * STAGE;
		/* the MMAs that read this stage two iterations ago must have completed */
		if (kb >= 2) {
			mbar_wait(&bar[st], phase[st]);
			phase[st] ^= 1;
		}
		const int k0 = kb * UM_BK;
		/* ---- stage A: 128 rows x 4 groups of 16 k */
		for (int item = tid; item < UM_BM * (UM_BK / 16); item += UM_THREADS) {
			const int row = item / (UM_BK / 16), kq = item % (UM_BK / 16);
			const int gm = m0 + row, gk = k0 + kq * 16;
			i32 v[16];
			if (gm < M && gk + 15 < K && ((lda & 3) == 0)) {
				const int4 *src = reinterpret_cast<const int4 *>(A + (size_t) gm * lda + gk);
#pragma unroll
				for (int q = 0; q < 4; q++) {
					int4 t = src[q];
					v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
				}
			} else {
#pragma unroll
				for (int e = 0; e < 16; e++)
					v[e] = (gm < M && gk + e < K) ? A[(size_t) gm * lda + gk + e] : 0;
			}
			uint4 plane[L];
			split16<L>(v, plane);
			const int off = ((row >> 3) * (UM_BK / 16) + kq) * 128 + (row & 7) * 16;
#pragma unroll
			for (int l = 0; l < L; l++)
				*reinterpret_cast<uint4 *>(stage + l * A_PLANE + off) = plane[l];
		}
		/* ---- stage B (K x N row-major in HBM, written K-major): column n, 16 consecutive k */
		for (int item = tid; item < BN * (UM_BK / 16); item += UM_THREADS) {
			const int col = item % BN, kq = item / BN;         /* consecutive threads -> consecutive columns: coalesced */
			const int gn = n0 + col, gk = k0 + kq * 16;
			i32 v[16];
#pragma unroll
			for (int e = 0; e < 16; e++)
				v[e] = (gn < N && gk + e < K) ? B[(size_t) (gk + e) * ldb + gn] : 0;
			uint4 plane[L];
			split16<L>(v, plane);
			const int off = ((col >> 3) * (UM_BK / 16) + kq) * 128 + (col & 7) * 16;
#pragma unroll
			for (int l = 0; l < L; l++)
				*reinterpret_cast<uint4 *>(stage + L * A_PLANE + l * B_PLANE + off) = plane[l];
		}
		/* generic-proxy writes -> visible to the tensor core (async proxy) */
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		__syncthreads();
		if (tid == 0) {
			asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
			const uint32_t sa = smem_u32(stage), sb_ = sa + L * A_PLANE;
#pragma unroll
			for (int ks = 0; ks < UM_BK / 32; ks++) {
#pragma unroll
				for (int i = 0; i < L; i++) {
					const uint64_t da = umma_desc(sa + i * A_PLANE + ks * 256, 128, (UM_BK / 16) * 128);
#pragma unroll
					for (int j = 0; j < L; j++) {
						const uint64_t db = umma_desc(sb_ + j * B_PLANE + ks * 256, 128, (UM_BK / 16) * 128);
						/* first product ever written to a weight class overwrites, the others accumulate */
						const bool first = (kb == 0 && ks == 0 && (i == 0 || j == L - 1));
						umma_i8(tmem_base + (uint32_t) ((i + j) * BN), da, db, idesc, first ? 0u : 1u);
					}
				}
			}
			umma_commit(&bar[st]);
		}
	}
	/* wait for the last commits of both stages */
	for (int st = 0; st < 2; st++) {
		int uses = (nkb + 1 - st) / 2;              /* K-blocks that used this stage */
		if (uses > 0) {
			mbar_wait(&bar[st], phase[st]);
			phase[st] ^= 1;
		}
	}
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

	/* ---- epilogue: warp w reads TMEM lanes 32*(w%4) .. +31 (its quarter), columns split between w/4 = 0, 1 */
	const i32 w256[7] = {w256_0, w256_1, w256_2, w256_3, w256_4, w256_5, w256_6};
	const int lane_base = 32 * (warp & 3);
	const int row = m0 + lane_base + lane;
	constexpr int CHUNK = 16;
	for (int c0 = (warp >> 2) * CHUNK; c0 < BN; c0 += 2 * CHUNK) {
		uint32_t acc[NCLASS][CHUNK];
#pragma unroll
		for (int cl = 0; cl < NCLASS; cl++) {
			const uint32_t taddr = tmem_base + ((uint32_t) lane_base << 16) + (uint32_t) (cl * BN + c0);
			asm volatile(
			    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
			    : "=r"(acc[cl][0]), "=r"(acc[cl][1]), "=r"(acc[cl][2]), "=r"(acc[cl][3]), "=r"(acc[cl][4]), "=r"(acc[cl][5]),
			      "=r"(acc[cl][6]), "=r"(acc[cl][7]), "=r"(acc[cl][8]), "=r"(acc[cl][9]), "=r"(acc[cl][10]), "=r"(acc[cl][11]),
			      "=r"(acc[cl][12]), "=r"(acc[cl][13]), "=r"(acc[cl][14]), "=r"(acc[cl][15])
			    : "r"(taddr));
		}
		asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
		if (row < M) {
#pragma unroll
			for (int e = 0; e < CHUNK; e++) {
				const int gn = n0 + c0 + e;
				if (gn >= N)
					continue;
				i64 sum;
				if (L <= 2) {
					sum = 0;
#pragma unroll
					for (int cl = 0; cl < NCLASS; cl++)
						sum += ((i64) (i32) acc[cl][e]) << (8 * cl);      /* < 2^31 * 2^16 * 3 */
				} else {
					sum = 0;
#pragma unroll
					for (int cl = 0; cl < NCLASS; cl++)
						sum += (i64) zp_reduce((i64) zp_reduce((i64) (i32) acc[cl][e], F) * (i64) w256[cl], F);
				}
				const size_t at = (size_t) row * ldc + gn;
				C[at] = zp_reduce((i64) C[at] - (i64) zp_reduce(sum, F), F);
			}
		}
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 0)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) TMEM_COLS));
}

/* ======================================================================================================
 * Version 2: operands pre-split into limb planes in HBM, stored tile by tile in the very byte layout the tensor
 * core reads from shared memory, so that a pipeline stage is filled by plain 1-D bulk copies (cp.async.bulk, the TMA
 * engine without a tensor map) and no thread touches the operands:
 *
 *   plane l of an operand with `rows` rows (M or N side) and K columns (K-major)
 *     = tiles of 64 rows x 64 k bytes (4 KB), tile (kb, rb) at byte offset ((kb * nRB + rb) * 4096),
 *       inside a tile: core matrix (g = r/8, kq = k/16) at (g*4 + kq)*128, row r%8 at +16*(r%8), byte k%16.
 *   Two consecutive row tiles are contiguous, so a 128-row stage of one plane is ONE 8 KB bulk copy and the UMMA
 *   descriptor (LBO 128 B, SBO 512 B) spans both.
 *
 * Warp roles (320 threads): warp 0 = producer (one lane issues the bulk copies, mbarrier expect_tx), warp 1 = MMA
 * issuer (one lane; tcgen05.commit frees the stage), warps 2-9 = epilogue (TMEM -> registers -> shared memory -> C),
 * two per TMEM lane quarter.  4 to 6 stages.
 * The packing kernels read the int32 matrices once; the dense rows of a block are packed when the block is created
 * and reused by every later block.
 */
#define P_TILE 4096
#define P_THREADS 320

size_t umma_packed_bytes(int rows, int K, int L)
{
	size_t nRB = 2 * (size_t) ((rows + 127) / 128), nKB = (size_t) ((K + 63) / 64);
	return (size_t) L * nRB * nKB * P_TILE;
}

/* ROWMAJOR_K = true : src is rows x K, K contiguous (the A operand)
 * ROWMAJOR_K = false: src is K x rows, rows contiguous (the B operand, K x N row-major) */
template <int L, bool ROWMAJOR_K>
__global__ void k_umma_pack(const i32 *__restrict__ src, int ld, int rows, int K, int8_t *__restrict__ dst, int nRB, int nKB)
{
	const size_t plane = (size_t) nRB * nKB * P_TILE;
	const int nkq = nKB * 4;
	const long total = (long) nRB * 64 * nkq;
	for (long item = (long) blockIdx.x * blockDim.x + threadIdx.x; item < total; item += (long) gridDim.x * blockDim.x) {
		int r, kq;
		if (ROWMAJOR_K) {
			r = (int) (item / nkq);
			kq = (int) (item % nkq);
		} else {
			r = (int) (item % (nRB * 64));      /* consecutive threads -> consecutive rows: coalesced reads of K x rows */
			kq = (int) (item / (nRB * 64));
		}
		const int gk = kq * 16;
		i32 v[16];
#pragma unroll
		for (int e = 0; e < 16; e++) {
			const int k = gk + e;
			i32 x = 0;
			if (r < rows && k < K)
				x = ROWMAJOR_K ? src[(size_t) r * ld + k] : src[(size_t) k * ld + r];
			v[e] = x;
		}
		uint4 pl[L];
		split16<L>(v, pl);
		const int rb = r >> 6, rr = r & 63, kb = kq >> 2, kqq = kq & 3;
		const size_t off = ((size_t) kb * nRB + rb) * P_TILE + ((rr >> 3) * 4 + kqq) * 128 + (rr & 7) * 16;
#pragma unroll
		for (int l = 0; l < L; l++)
			*reinterpret_cast<uint4 *>(dst + l * plane + off) = pl[l];
	}
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

__host__ __device__ constexpr int TMEM_COLS_FOR(int nclass, int bn)
{
	return (nclass * bn <= 32) ? 32 : (nclass * bn <= 64) ? 64 : (nclass * bn <= 128) ? 128 : (nclass * bn <= 256) ? 256 : 512;
}

/* development: cycles spent per phase, summed over CTAs (SPASM_B200_UMMA_PROFILE builds only) */
#ifdef SPASM_B200_UMMA_PROFILE
__device__ unsigned long long g_umma_prof[8];
#define PROF_MARK(k) do { if ((threadIdx.x & 31) == 0) { long long now_ = clock64(); atomicAdd(&g_umma_prof[k], (unsigned long long) (now_ - t_prof)); t_prof = now_; } } while (0)
#else
#define PROF_MARK(k) do { } while (0)
#endif

/* OCC = CTAs meant to be resident per SM (they share the 227 KB of shared memory and the 512 TMEM columns) */
__host__ __device__ constexpr int packed_stages(int stage_bytes, int occ)
{
	return (200 * 1024 / occ) / stage_bytes > 6 ? 6 : (200 * 1024 / occ) / stage_bytes;
}

template <int L, int BN, int OCC>
__global__ void __launch_bounds__(P_THREADS, OCC)
k_umma_gemm_packed(i32 *__restrict__ C, int ldc, const int8_t *__restrict__ Ap, int nRB_A, const int8_t *__restrict__ Bp, int nRB_B,
                   int nKB, int kb_begin, int kb_end, int M, int N, Zp F)
{
	constexpr int NCLASS = 2 * L - 1;
	constexpr int A_STAGE = 2 * P_TILE;                   /* 128 rows of one plane */
	constexpr int B_STAGE = (BN / 64) * P_TILE;
	constexpr int STAGE = L * (A_STAGE + B_STAGE);
	constexpr int P_STAGES = packed_stages(STAGE, OCC);
	static_assert(P_STAGES >= 2, "pipeline too shallow");
	static_assert(OCC * TMEM_COLS_FOR(2 * L - 1, BN) <= 512, "co-resident CTAs must fit the tensor memory");
	constexpr int TMEM_COLS = TMEM_COLS_FOR(NCLASS, BN);
	extern __shared__ __align__(1024) unsigned char smem[];
	__shared__ uint64_t full_bar[P_STAGES], empty_bar[P_STAGES], tmem_full_bar;
	__shared__ uint32_t tmem_base_slot;

	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#ifdef SPASM_B200_UMMA_PROFILE
	long long t_prof = clock64();
#endif
	const int m0 = blockIdx.y * UM_BM, n0 = blockIdx.x * BN;
	const size_t planeA = (size_t) nRB_A * nKB * P_TILE, planeB = (size_t) nRB_B * nKB * P_TILE;

	if (warp == 1) {
		asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t) TMEM_COLS));
		asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
	}
	if (tid == 0) {
		for (int s = 0; s < P_STAGES; s++) {
			mbar_init(&full_bar[s], 1);
			mbar_init(&empty_bar[s], 1);
		}
		mbar_init(&tmem_full_bar, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
	const uint32_t tmem_base = tmem_base_slot;
	const int nkb = kb_end - kb_begin;
	if (warp == 2)
		PROF_MARK(0);      /* setup: barriers, TMEM allocation */

	if (warp == 0) {
		/* ===== producer ===== */
		if (lane == 0) {
			const int rbA = (m0 >> 6), rbB = (n0 >> 6);
			for (int it = 0; it < nkb; it++) {
				const int s = it % P_STAGES;
				const uint32_t ph = (uint32_t) ((it / P_STAGES) & 1);
				mbar_wait(&empty_bar[s], ph ^ 1);                 /* passes at once on the first round */
				unsigned char *stage = smem + (size_t) s * STAGE;
				mbar_expect_tx(&full_bar[s], (uint32_t) STAGE);
				const int kb = kb_begin + it;
#pragma unroll
				for (int l = 0; l < L; l++) {
					bulk_copy_g2s(stage + l * A_STAGE, Ap + l * planeA + ((size_t) kb * nRB_A + rbA) * P_TILE, A_STAGE, &full_bar[s]);
					bulk_copy_g2s(stage + L * A_STAGE + l * B_STAGE, Bp + l * planeB + ((size_t) kb * nRB_B + rbB) * P_TILE, B_STAGE, &full_bar[s]);
				}
			}
		}
	} else if (warp == 1) {
		/* ===== MMA issuer ===== */
		if (lane == 0) {
			const uint32_t idesc = umma_idesc_i8(UM_BM, BN);
			for (int it = 0; it < nkb; it++) {
				const int s = it % P_STAGES;
				const uint32_t ph = (uint32_t) ((it / P_STAGES) & 1);
				mbar_wait(&full_bar[s], ph);
				asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
				const uint32_t sa = smem_u32(smem + (size_t) s * STAGE), sb_ = sa + L * A_STAGE;
#pragma unroll
				for (int ks = 0; ks < 2; ks++) {
#pragma unroll
					for (int i = 0; i < L; i++) {
						const uint64_t da = umma_desc(sa + i * A_STAGE + ks * 256, 128, 512);
#pragma unroll
						for (int j = 0; j < L; j++) {
							const uint64_t db = umma_desc(sb_ + j * B_STAGE + ks * 256, 128, 512);
							const bool first = (it == 0 && ks == 0 && (i == 0 || j == L - 1));
							umma_i8(tmem_base + (uint32_t) ((i + j) * BN), da, db, idesc, first ? 0u : 1u);
						}
					}
				}
				umma_commit(&empty_bar[s]);                       /* the stage may be refilled once these MMAs have read it */
			}
			umma_commit(&tmem_full_bar);
		}
	} else {
		/* ===== epilogue: warps 2..9, TMEM lane quarter = warp % 4 ===== */
		mbar_wait(&tmem_full_bar, 0);
		asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
		if (warp == 2)
			PROF_MARK(1);  /* main loop */
		const int quarter = warp & 3, half = (warp - 2) >> 2;       /* two warps per TMEM lane quarter */
		const int lane_base = 32 * quarter;
		/* TMEM gives a thread one ROW of the tile; C wants a warp on one row at a time.  The recombined products go
		 * through shared memory (the pipeline stages are free now: every MMA has completed) as 64-bit integers, so
		 * that the only modular reduction is the one of C - product; row stride BN + 1 elements keeps both the
		 * column-wise writes and the row-wise reads conflict-free.  Each warp converts half of the columns of its
		 * quarter, then updates half of its rows. */
		i64 *tile = reinterpret_cast<i64 *>(smem) + (size_t) quarter * 32 * (BN + 1);
		constexpr int CHUNK = 16;
		for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += CHUNK) {
			uint32_t acc[NCLASS][CHUNK];
#pragma unroll
			for (int cl = 0; cl < NCLASS; cl++) {
				const uint32_t taddr = tmem_base + ((uint32_t) lane_base << 16) + (uint32_t) (cl * BN + c0);
				asm volatile(
				    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
				    : "=r"(acc[cl][0]), "=r"(acc[cl][1]), "=r"(acc[cl][2]), "=r"(acc[cl][3]), "=r"(acc[cl][4]), "=r"(acc[cl][5]),
				      "=r"(acc[cl][6]), "=r"(acc[cl][7]), "=r"(acc[cl][8]), "=r"(acc[cl][9]), "=r"(acc[cl][10]), "=r"(acc[cl][11]),
				      "=r"(acc[cl][12]), "=r"(acc[cl][13]), "=r"(acc[cl][14]), "=r"(acc[cl][15])
				    : "r"(taddr));
			}
			asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
			for (int e = 0; e < CHUNK; e++) {
				i64 sum = 0;
				if (L <= 2) {
#pragma unroll
					for (int cl = 0; cl < NCLASS; cl++)
						sum += ((i64) (i32) acc[cl][e]) << (8 * cl);
				} else {
					/* Horner in base 256, one reduction per class: |sum| * 256 + |acc| < 2^40 */
#pragma unroll
					for (int cl = NCLASS - 1; cl >= 0; cl--)
						sum = (i64) zp_reduce(sum * 256 + (i64) (i32) acc[cl][e], F);
				}
				tile[lane * (BN + 1) + c0 + e] = sum;
			}
		}
		asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");      /* the two warps of this quarter */
		if (warp == 2)
			PROF_MARK(2);  /* TMEM -> shared memory */
		const int rows_here = min(32, M - (m0 + lane_base));
		/* read-modify-write of C, 8 rows at a time: every load of a group is issued before the first store (a store
		 * orders the next load behind it), and the reductions of a group are computed branch-free so that their
		 * dependency chains interleave */
		constexpr int RG = 8, NG = BN / 32;
		for (int r0 = half * 16; r0 < min(rows_here, (half + 1) * 16); r0 += RG) {
			i32 cv[RG][NG];
#pragma unroll
			for (int rr = 0; rr < RG; rr++) {
				const i32 *crow = C + (size_t) (m0 + lane_base + r0 + rr) * ldc + n0;
#pragma unroll
				for (int g = 0; g < NG; g++) {
					const int col = g * 32 + lane;
					cv[rr][g] = (r0 + rr < rows_here && n0 + col < N) ? crow[col] : 0;
				}
			}
			/* scheduling barrier: without it ptxas sinks each load next to its first use and the group pays one
			 * memory latency per load instead of one per group */
			__syncwarp();
#pragma unroll
			for (int rr = 0; rr < RG; rr++)
#pragma unroll
				for (int g = 0; g < NG; g++)
					cv[rr][g] = zp_reduce((i64) cv[rr][g] - tile[(r0 + rr) * (BN + 1) + g * 32 + lane], F);
#pragma unroll
			for (int rr = 0; rr < RG; rr++) {
				i32 *crow = C + (size_t) (m0 + lane_base + r0 + rr) * ldc + n0;
#pragma unroll
				for (int g = 0; g < NG; g++) {
					const int col = g * 32 + lane;
					if (r0 + rr < rows_here && n0 + col < N)
						crow[col] = cv[rr][g];
				}
			}
		}
	}
	if (warp == 2)
		PROF_MARK(3);      /* update of C */
	asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
	__syncthreads();
	if (warp == 1)
		asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) TMEM_COLS));
	if (warp == 2)
		PROF_MARK(4);      /* drain */
}

/* number of signed byte limbs needed for balanced residues mod p, or 0 if more than 4 */
static int limbs_for(const Zp &F)
{
	i64 A = F.half > -F.mhalf ? F.half : -F.mhalf;
	i64 cap = 127;
	for (int L = 1; L <= 4; L++) {
		if (A <= cap)
			return L;
		cap = cap * 256 + 127;
	}
	return 0;
}

bool umma_gemm_available(const Zp &F)
{
	static int disabled = -1;
	if (disabled < 0)
		disabled = getenv("SPASM_B200_NO_TENSOR") ? 1 : 0;
	return !disabled && limbs_for(F) > 0;
}

template <int L, int BN>
static void launch_umma(i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F)
{
	constexpr int STAGE = L * (UM_BM * UM_BK + BN * UM_BK);
	size_t smem = 2 * (size_t) STAGE + 1024;
	CUDA_CHECK(cudaFuncSetAttribute(k_umma_gemm_sub<L, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	i32 w[7];
	i64 pw = 1;
	for (int c = 0; c < 7; c++) {
		w[c] = zp_reduce(pw, F);
		pw = (i64) zp_reduce(pw * 256, F);
	}
	dim3 grid(cdiv(N, BN), cdiv(M, UM_BM));
	/* K is chunked so that an int32 accumulator cannot overflow: L * K * 2^14 < 2^31 */
	const int kmax = 16384;
	for (int k0 = 0; k0 < K; k0 += kmax) {
		int kk = std::min(kmax, K - k0);
		k_umma_gemm_sub<L, BN><<<grid, UM_THREADS, smem, ctx().stream>>>(C, ldc, A + k0, lda, B + (size_t) k0 * ldb, ldb, M, N, kk, F,
		                                                                w[0], w[1], w[2], w[3], w[4], w[5], w[6]);
		LAUNCHED(1);
	}
	KERNEL_CHECK();
	stats().pub.gemm_int8_ops += 2.0 * M * (double) N * K * L * L;
}

template <int L>
static void pack_operand(const i32 *src, int ld, int rows, int K, bool rowmajor_k, int8_t *dst)
{
	int nRB = 2 * ((rows + 127) / 128), nKB = (K + 63) / 64;
	long total = (long) nRB * 64 * nKB * 4;
	unsigned blocks = std::min<unsigned>(cdiv((size_t) total, 256), 148u * 16);
	if (rowmajor_k)
		k_umma_pack<L, true><<<blocks, 256, 0, ctx().stream>>>(src, ld, rows, K, dst, nRB, nKB);
	else
		k_umma_pack<L, false><<<blocks, 256, 0, ctx().stream>>>(src, ld, rows, K, dst, nRB, nKB);
	LAUNCHED(1);
}

void umma_pack(const i32 *src, int ld, int rows, int K, bool rowmajor_k, int8_t *dst, const Zp &F)
{
	switch (limbs_for(F)) {
	case 1: pack_operand<1>(src, ld, rows, K, rowmajor_k, dst); break;
	case 2: pack_operand<2>(src, ld, rows, K, rowmajor_k, dst); break;
	case 3: pack_operand<3>(src, ld, rows, K, rowmajor_k, dst); break;
	case 4: pack_operand<4>(src, ld, rows, K, rowmajor_k, dst); break;
	default: errx(1, "[spasm-b200] internal: umma_pack for a prime that needs more than 4 limbs");
	}
	KERNEL_CHECK();
}

int umma_limbs(const Zp &F) { return limbs_for(F); }

template <int L, int BN, int OCC>
static void launch_packed(i32 *C, int ldc, const int8_t *Ap, const int8_t *Bp, int M, int N, int K, const Zp &F)
{
	constexpr int STAGE = L * (2 * P_TILE + (BN / 64) * P_TILE);
	size_t smem = std::max<size_t>((size_t) packed_stages(STAGE, OCC) * STAGE, (size_t) 128 * (BN + 1) * 8) + 1024;   /* stages, reused by the epilogue tile */
	CUDA_CHECK(cudaFuncSetAttribute(k_umma_gemm_packed<L, BN, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	int nRB_A = 2 * ((M + 127) / 128), nRB_B = 2 * ((N + 127) / 128), nKB = (K + 63) / 64;
	dim3 grid(cdiv(N, BN), cdiv(M, UM_BM));
	const int kb_max = 16384 / 64;              /* L * K * 2^14 < 2^31 */
	for (int kb0 = 0; kb0 < nKB; kb0 += kb_max) {
		int kb1 = std::min(nKB, kb0 + kb_max);
		k_umma_gemm_packed<L, BN, OCC><<<grid, P_THREADS, smem, ctx().stream>>>(C, ldc, Ap, nRB_A, Bp, nRB_B, nKB, kb0, kb1, M, N, F);
		LAUNCHED(1);
	}
	KERNEL_CHECK();
	stats().pub.gemm_int8_ops += 2.0 * M * (double) N * K * L * L;
}

/* C -= A * B with both operands already packed (A: M x K, B: N x K) */
void umma_gemm_sub_packed(i32 *C, int ldc, const int8_t *Ap, const int8_t *Bp, int M, int N, int K, const Zp &F)
{
	/* 128 x 64 tiles, two CTAs per SM (one's epilogue overlaps the other's main loop) for the short products of the
	 * block elimination; 128 x 128 tiles, which move fewer bytes from L2 per MMA, when K is large enough for the
	 * main loop to dominate.  SPASM_B200_UMMA_TILE=64|128 forces one (development knob). */
	static const char *force = getenv("SPASM_B200_UMMA_TILE");
	bool narrow = force ? atoi(force) == 64 : K < 4096;
	switch (limbs_for(F)) {
	case 1: narrow ? launch_packed<1, 64, 2>(C, ldc, Ap, Bp, M, N, K, F) : launch_packed<1, 128, 1>(C, ldc, Ap, Bp, M, N, K, F); break;
	case 2: narrow ? launch_packed<2, 64, 2>(C, ldc, Ap, Bp, M, N, K, F) : launch_packed<2, 128, 1>(C, ldc, Ap, Bp, M, N, K, F); break;
	case 3: launch_packed<3, 64, 1>(C, ldc, Ap, Bp, M, N, K, F); break;
	case 4: launch_packed<4, 64, 1>(C, ldc, Ap, Bp, M, N, K, F); break;
	default: errx(1, "[spasm-b200] internal: umma_gemm_sub_packed for a prime that needs more than 4 limbs");
	}
}

/* same contract as dense_gemm_sub (dense.cuh): packs both operands, then runs the packed kernel.
 * SPASM_B200_UMMA_V1=1 selects the first version (operands staged by the threads). */
void umma_gemm_sub(i32 *C, int ldc, const i32 *A, int lda, const i32 *B, int ldb, int M, int N, int K, const Zp &F)
{
	static const bool v1 = getenv("SPASM_B200_UMMA_V1") != NULL;
	int L = limbs_for(F);
	if (L < 1 || L > 4)
		errx(1, "[spasm-b200] internal: umma_gemm_sub called for a prime that needs more than 4 limbs");
	if (v1) {
		if (L <= 1)
			launch_umma<1, 128>(C, ldc, A, lda, B, ldb, M, N, K, F);
		else if (L == 2)
			launch_umma<2, 128>(C, ldc, A, lda, B, ldb, M, N, K, F);
		else if (L == 3)
			launch_umma<3, 64>(C, ldc, A, lda, B, ldb, M, N, K, F);
		else
			launch_umma<4, 64>(C, ldc, A, lda, B, ldb, M, N, K, F);
		return;
	}
	DevBuf<int8_t> Ap(umma_packed_bytes(M, K, L)), Bp(umma_packed_bytes(N, K, L));
	umma_pack(A, lda, M, K, true, Ap.ptr, F);
	umma_pack(B, ldb, N, K, false, Bp.ptr, F);
	umma_gemm_sub_packed(C, ldc, Ap.ptr, Bp.ptr, M, N, K, F);
}

}  // namespace sb

/* test hook (include/spasm_b200.h): C -= A*B on host matrices through either implementation */
extern "C" void spasm_b200_gemm_sub(int64_t prime, int M, int N, int K, int32_t *C, const int32_t *A, const int32_t *B, int use_tensor)
{
	using namespace sb;
	ctx();
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	DevBuf<i32> dA, dB, dC;
	dA.upload(A, (size_t) M * K, s);
	dB.upload(B, (size_t) K * N, s);
	dC.upload(C, (size_t) M * N, s);
	if (use_tensor) {
		if (!umma_gemm_available(F))
			errx(1, "[spasm-b200] tensor-core product not available for this prime");
		umma_gemm_sub(dC.ptr, N, dA.ptr, K, dB.ptr, N, M, N, K, F);
	} else {
		dense_gemm_sub(dC.ptr, N, dA.ptr, K, dB.ptr, N, M, N, K, F);
	}
	dC.download(C, (size_t) M * N, s);
	sb::sync();
}

/* timing hook (include/spasm_b200.h): average milliseconds of C -= A*B on device-resident pseudo-random operands.
 * mode 0 = CUDA cores, 1 = tensor cores with both operands packed inside the timed region,
 * 2 = tensor cores, B packed beforehand (the dense rows of a block), 3 = packed kernel alone */
__global__ void k_fill_residues(i32 *x, size_t n, unsigned seed, sb::Zp F)
{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
		unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
		z ^= z >> 29;
		z *= 0xBF58476D1CE4E5B9ull;
		z ^= z >> 32;
		x[i] = sb::zp_reduce((i64) (z >> 2), F);
	}
}

extern "C" double spasm_b200_gemm_time(int64_t prime, int M, int N, int K, int mode, int reps)
{
	using namespace sb;
	ctx();
	cudaStream_t s = ctx().stream;
	Zp F = make_zp(prime);
	DevBuf<i32> dA((size_t) M * K), dB((size_t) K * N), dC((size_t) M * N);
	k_fill_residues<<<1184, 256, 0, s>>>(dA.ptr, dA.count, 1u, F);
	k_fill_residues<<<1184, 256, 0, s>>>(dB.ptr, dB.count, 2u, F);
	k_fill_residues<<<1184, 256, 0, s>>>(dC.ptr, dC.count, 3u, F);
	int L = mode ? limbs_for(F) : 1;
	if (mode && (L < 1 || L > 4))
		errx(1, "[spasm-b200] tensor-core product not available for this prime");
	DevBuf<int8_t> Ap, Bp;
	if (mode >= 2) {
		Ap.alloc(umma_packed_bytes(M, K, L));
		Bp.alloc(umma_packed_bytes(N, K, L));
		umma_pack(dA.ptr, K, M, K, true, Ap.ptr, F);
		umma_pack(dB.ptr, N, N, K, false, Bp.ptr, F);
	}
	GpuTimer t;
	double total = 0;
	for (int it = -1; it < reps; it++) {
		t.start();
		switch (mode) {
		case 0: dense_gemm_sub(dC.ptr, N, dA.ptr, K, dB.ptr, N, M, N, K, F); break;
		case 1: umma_gemm_sub(dC.ptr, N, dA.ptr, K, dB.ptr, N, M, N, K, F); break;
		case 2:
			umma_pack(dA.ptr, K, M, K, true, Ap.ptr, F);
			umma_gemm_sub_packed(dC.ptr, N, Ap.ptr, Bp.ptr, M, N, K, F);
			break;
		default: umma_gemm_sub_packed(dC.ptr, N, Ap.ptr, Bp.ptr, M, N, K, F); break;
		}
		double ms = t.stop_ms();
		if (it >= 0)
			total += ms;
	}
#ifdef SPASM_B200_UMMA_PROFILE
	unsigned long long prof[8];
	cudaMemcpyFromSymbol(prof, g_umma_prof, sizeof(prof));
	double ctas = (double) (reps + 1) * ((N + 63) / 64) * ((M + 127) / 128);
	fprintf(stderr, "[umma profile, cycles per CTA assuming 64-wide tiles] setup %.0f main %.0f tmem->smem %.0f C update %.0f drain %.0f\n",
	        prof[0] / ctas, prof[1] / ctas, prof[2] / ctas, prof[3] / ctas, prof[4] / ctas);
	unsigned long long zero[8] = {0};
	cudaMemcpyToSymbol(g_umma_prof, zero, sizeof(zero));
#endif
	return total / reps;
}
