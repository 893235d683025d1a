/*
 * Multi-GPU support: one process per GPU (torchrun), NCCL over NVLink/NVSwitch.
 *
 * What shards: given U, the right-hand sides of a solve batch are independent, so the rows of a
 * dense block (or the output rows of the randomized Schur complement, whose PRNG is seeded per output row) are
 * split across the ranks; every rank solves its slice against its own replica of U and the slices of the dense
 * block are exchanged with ONE ncclAllGather (the only data-path collective; 4 * rows * Sm0 bytes, e.g. 9.5 MB for
 * a 1000 x 2371 block).  Pivot search has ordered commits and is run redundantly by every rank (deterministic, so
 * the replicas of U agree bit for bit without a broadcast); the dense echelon of a block is small and is also
 * replicated, which keeps every rank's state identical.
 *
 * NCCL is loaded with dlopen at communicator creation: a single-GPU process never needs it, and in a process that
 * already holds torch's NCCL the same library instance is reused.
 */
#include <dlfcn.h>
#include "engine.cuh"
#include "stats.cuh"

namespace sb {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2 };      /* ncclDataType_t: ncclInt8 0, ncclUint8 1, ncclInt32 2 (nccl.h) */

static struct {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	ncclComm_t comm = nullptr;
	int rank = 0, world = 1;
	int result_root = -1;      /* >= 0: spasm_rref / spasm_kernel materialise their result on this rank only */
} g_nccl;

static void nccl_load()
{
	if (g_nccl.handle)
		return;
	const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
	for (int k = 0; names[k] && !g_nccl.handle; k++)
		g_nccl.handle = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
	if (!g_nccl.handle)
		errx(1, "[spasm-b200] cannot load NCCL (%s): multi-GPU operation needs libnccl.so.2", dlerror());
	*(void **) &g_nccl.GetUniqueId = dlsym(g_nccl.handle, "ncclGetUniqueId");
	*(void **) &g_nccl.CommInitRank = dlsym(g_nccl.handle, "ncclCommInitRank");
	*(void **) &g_nccl.CommDestroy = dlsym(g_nccl.handle, "ncclCommDestroy");
	*(void **) &g_nccl.AllGather = dlsym(g_nccl.handle, "ncclAllGather");
	*(void **) &g_nccl.GetErrorString = dlsym(g_nccl.handle, "ncclGetErrorString");
	*(void **) &g_nccl.Broadcast = dlsym(g_nccl.handle, "ncclBroadcast");
	*(void **) &g_nccl.Send = dlsym(g_nccl.handle, "ncclSend");
	*(void **) &g_nccl.Recv = dlsym(g_nccl.handle, "ncclRecv");
	*(void **) &g_nccl.GroupStart = dlsym(g_nccl.handle, "ncclGroupStart");
	*(void **) &g_nccl.GroupEnd = dlsym(g_nccl.handle, "ncclGroupEnd");
	if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllGather || !g_nccl.Broadcast || !g_nccl.GroupStart || !g_nccl.GroupEnd)
		errx(1, "[spasm-b200] NCCL library lacks the expected entry points");
}

#define NCCL_CHECK(call)                                                                                    \
	do {                                                                                                    \
		ncclResult_t r_ = (call);                                                                           \
		if (r_ != 0)                                                                                        \
			errx(1, "[spasm-b200] NCCL error %d at %s:%d: %s", r_, __FILE__, __LINE__,                       \
			     g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?");                                  \
	} while (0)

int comm_world() { return g_nccl.world; }
int comm_rank() { return g_nccl.rank; }

/* rows [begin, end) of a block of `total` rows owned by this rank; every rank owns `chunk` rows (the last ones padded) */
void comm_slice(int total, int *chunk, int *begin, int *end)
{
	int c = (total + g_nccl.world - 1) / g_nccl.world;
	*chunk = c;
	*begin = std::min(total, g_nccl.rank * c);
	*end = std::min(total, *begin + c);
}

/* in-place all-gather of a row-major block: rank r owns rows [r*chunk, (r+1)*chunk); buffer holds world*chunk rows */
void comm_allgather_rows(i32 *B, int chunk, int ld)
{
	if (g_nccl.world == 1)
		return;
	size_t count = (size_t) chunk * ld;
	NCCL_CHECK(g_nccl.AllGather(B + (size_t) g_nccl.rank * count, B, count, ncclInt32, g_nccl.comm, ctx().stream));
	stats().pub.nccl_bytes += (i64) count * 4 * (g_nccl.world - 1);
}

/* in-place all-gather of `bytes` bytes per rank (rank r's part at offset r * bytes) */
void comm_allgather_bytes(void *buf, size_t bytes)
{
	if (g_nccl.world == 1)
		return;
	NCCL_CHECK(g_nccl.AllGather((char *) buf + (size_t) g_nccl.rank * bytes, buf, bytes, 0 /* ncclInt8 */, g_nccl.comm, ctx().stream));
	stats().pub.nccl_bytes += (i64) bytes * (g_nccl.world - 1);
}

/* broadcasts of variable-length pieces, fused into one NCCL group by the caller (comm_group_begin / _end) */
void comm_group_begin() { if (g_nccl.world > 1) NCCL_CHECK(g_nccl.GroupStart()); }
void comm_group_end() { if (g_nccl.world > 1) NCCL_CHECK(g_nccl.GroupEnd()); }
void comm_bcast_bytes(void *buf, size_t bytes, int root)
{
	if (g_nccl.world == 1 || bytes == 0)
		return;
	NCCL_CHECK(g_nccl.Broadcast(buf, buf, bytes, 0 /* ncclInt8 */, root, g_nccl.comm, ctx().stream));
	if (root != g_nccl.rank)
		stats().pub.nccl_bytes += (i64) bytes;
}

/* point-to-point halves of a gather to one rank (inside a group, like the broadcasts) */
int comm_result_root() { return g_nccl.world > 1 ? g_nccl.result_root : -1; }
void comm_send_bytes(const void *buf, size_t bytes, int peer)
{
	if (bytes == 0)
		return;
	if (!g_nccl.Send)
		errx(1, "[spasm-b200] NCCL library lacks ncclSend");
	NCCL_CHECK(g_nccl.Send(buf, bytes, 0 /* ncclInt8 */, peer, g_nccl.comm, ctx().stream));
}
void comm_recv_bytes(void *buf, size_t bytes, int peer)
{
	if (bytes == 0)
		return;
	if (!g_nccl.Recv)
		errx(1, "[spasm-b200] NCCL library lacks ncclRecv");
	NCCL_CHECK(g_nccl.Recv(buf, bytes, 0 /* ncclInt8 */, peer, g_nccl.comm, ctx().stream));
	stats().pub.nccl_bytes += (i64) bytes;
}

}  // namespace sb

using namespace sb;

extern "C" {

void spasm_b200_comm_unique_id(void *out128)
{
	nccl_load();
	ncclUniqueId id;
	NCCL_CHECK(g_nccl.GetUniqueId(&id));
	memcpy(out128, &id, sizeof(id));
}

void spasm_b200_comm_init(int rank, int world, const void *unique_id128)
{
	nccl_load();
	ctx();
	ncclUniqueId id;
	memcpy(&id, unique_id128, sizeof(id));
	NCCL_CHECK(g_nccl.CommInitRank(&g_nccl.comm, world, id, rank));
	g_nccl.rank = rank;
	g_nccl.world = world;
}

void spasm_b200_comm_destroy(void)
{
	if (g_nccl.comm) {
		sb::sync();
		g_nccl.CommDestroy(g_nccl.comm);
		g_nccl.comm = nullptr;
	}
	g_nccl.rank = 0;
	g_nccl.world = 1;
}

int spasm_b200_comm_world(void) { return g_nccl.world; }

void spasm_b200_comm_result_root(int root) { g_nccl.result_root = root; }
}
