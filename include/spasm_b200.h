/*
 * spasm_b200.h -- extra C-ABI entry points of libspasm_b200.so.
 *
 * These do not exist in the reference; they expose what the benchmark and the
 * tests need to *measure* the CUDA path (device selection, per-phase CUDA-event
 * timers, kernel launch counts, algorithmic byte counters of SURVEY.md 8d).
 * The drop-in surface proper is include/spasm.h.
 */
#ifndef _SPASM_B200_H
#define _SPASM_B200_H
#include <stdint.h>
#include "spasm.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-process counters, reset with spasm_b200_reset_stats().  All times are CUDA-event
 * milliseconds measured on the library's stream. */
struct spasm_b200_stats {
	int64_t kernel_launches;      /* kernels of this library launched since the last reset */
	double ms_pivots;             /* FL + FL-columns + greedy + reorder + U extraction */
	double ms_pivots_greedy;      /*   of which the greedy cycle-free search */
	double ms_solve;              /* batched triangular solves (all callers) */
	double ms_dense;              /* dense echelon + dense updates */
	double ms_dense_gemm;         /*   of which the trailing-update / block-update GEMMs */
	double ms_total_echelonize;   /* host wall clock of the last echelonize call */
	double ms_device_echelonize;  /* the same call bracketed by CUDA events on the library's stream */
	double ms_k_greedy;           /* the greedy search kernel alone */
	double ms_k_panel_solve;      /* the panel-solve kernel launches alone */
	/* algorithmic work */
	double solve_bytes;           /* sum over solved rows of 8*nnz(B[k]) + 8*sum nnz(U rows reached) + output bytes */
	int64_t solve_rows;
	int64_t solve_batches;
	double solve_traffic_model;   /* bytes the batched formulation itself has to move (panel reads + writes) */
	double gemm_fieldops;         /* 2*M*N*K of every dense update executed */
	double gemm_int8_ops;         /* limb products actually issued to the int8 tensor pipe (0 when on CUDA cores) */
	int64_t greedy_edges;         /* pivot-row entries traversed by the greedy search */
	int64_t h2d_bytes, d2h_bytes; /* bytes copied across PCIe by the library */
	int64_t nccl_bytes;           /* bytes received through NCCL collectives by this rank */
	/* trace of the last spasm_echelonize call, for parity with the oracle */
	int nrounds;
	int found_FL[64], found_FLcol[64], found_greedy[64];
	double density[64];
	int finish;                   /* 0 none, 1 low-rank, 2 dense, 3 GPLU (executed through the dense path) */
	int nblocks;
	int block_Sn[4096], block_Sm[4096], block_rr[4096], block_w[4096];
	int dag_depth;                /* number of levels of the last structural U */
};

int  spasm_b200_device_count(void);
void spasm_b200_set_device(int device);      /* default: SPASM_B200_DEVICE or LOCAL_RANK or 0 */
void spasm_b200_reset_stats(void);
void spasm_b200_get_stats(struct spasm_b200_stats *out);
const char *spasm_b200_version(void);
void spasm_b200_set_verbose(int verbose);    /* 0 silences the stderr progress lines */
/* The library allocates from a memory pool of its own (never from the device's default pool) and keeps freed blocks
 * for the next call (blocks of 32 MB and more in a cache of up to SPASM_B200_CACHE_GB = 48 GB).  spasm_b200_trim()
 * returns all idle device memory to the driver; call it between echelonizations when another library of the process
 * needs the HBM.  All entry points of the library are single-threaded: call them from one thread at a time. */
void spasm_b200_trim(void);

/* Device-resident operation, for measuring the kernels without PCIe traffic: upload once, echelonize many times.
 * The echelon form of the resident variant stays on the device and is discarded; the rank is returned and
 * *ms_device receives the CUDA-event time of the call. */
void *spasm_b200_upload_csr(const struct spasm_csr *A);
void  spasm_b200_free_csr(void *handle);
int   spasm_b200_echelonize_resident(void *handle, struct echelonize_opts *opts, double *ms_device);
void  spasm_b200_flush_l2(void);             /* writes 512 MB: cold L2 for the next timed step */

/* first `count` outputs of the device implementation of the reference's PRNG stream (prime, seed, seq);
 * for the known-answer test against tests/Expected/prng */
void spasm_b200_prng_stream(int64_t prime, uint64_t seed, uint32_t seq, int count, int32_t *out_host);

/* test hook: C (M x N) -= A (M x K) * B (K x N) mod prime on host row-major matrices, through the CUDA-core product
 * (use_tensor = 0) or the tcgen05 int8 limb-split product (use_tensor = 1) */
void spasm_b200_gemm_sub(int64_t prime, int M, int N, int K, int32_t *C, const int32_t *A, const int32_t *B, int use_tensor);
/* average ms of one C -= A*B on device-resident pseudo-random operands; mode 0 = CUDA cores, 1 = tensor cores (operands
 * split into int8 limb planes inside the timed region), 2 = B planes prepared beforehand, 3 = tensor kernel alone */
double spasm_b200_gemm_time(int64_t prime, int M, int N, int K, int mode, int reps);

/* Multi-GPU (one process per GPU).  Rank 0 creates a 128-byte NCCL unique id, the caller distributes it (e.g.
 * torch.distributed.broadcast), every rank calls spasm_b200_comm_init.  From then on spasm_echelonize shards the
 * rows of its solve batches across the ranks and exchanges the dense blocks with ncclAllGather; every rank returns
 * the same echelon form. */
void spasm_b200_comm_unique_id(void *out128);
void spasm_b200_comm_init(int rank, int world, const void *unique_id128);
void spasm_b200_comm_destroy(void);
int  spasm_b200_comm_world(void);
/* root >= 0: spasm_rref and spasm_kernel GATHER their sharded result on that rank (ncclSend / ncclRecv); the other ranks
 * return a matrix of the right shape with no entries and skip the download of a result they do not use.  root < 0
 * (default): every rank returns the full result (one group of ncclBroadcast). */
void spasm_b200_comm_result_root(int root);

/* structural pivot pairs (row of the ORIGINAL matrix, column) of the last spasm_echelonize call,
 * round after round; returns their number.  Pass NULL to query the count. */
int spasm_b200_last_pivot_pairs(int *rows, int *cols, int *round_start /* size nrounds+1 */);

#ifdef __cplusplus
}
#endif
#endif
