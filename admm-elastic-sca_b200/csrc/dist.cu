// dist.cu -- single mesh partitioned over several GPUs (one process per GPU), SURVEY.md 8(e) rows 2 and 3.
//
// Setup is replicated: every rank is given the whole scene, computes the same nested-dissection order, and OWNS the
// contiguous chunk [own0, own1) of it.  Per ADMM iteration a rank
//   * runs the local step for the forces that touch an owned node (boundary forces are evaluated on both sides,
//     bit-identically, so the owned rows of the right-hand side are complete without any exchange),
//   * direct solver (default): all-gathers the right-hand side and runs the solve sharded by subtrees of the elimination
//     tree (direct_solve.cu, shard_owners) -- one all-reduce of the top separator rows, one all-reduce of x -- or
//     replicated (deterministic mode / ADMMB_DIST_SOLVE=replicated),
//   * PCG: runs Jacobi-PCG on its rows of A_n; per CG iteration the data-path exchanges are a neighbour-only halo of the
//     preconditioned residual u (grouped ncclSend / ncclRecv of exactly the nodes a peer's rows reference: the separator
//     surfaces between the chunks, dist_halo_exchange) and ONE all-reduce of 9 doubles for the fused dot products; the
//     solution is all-gathered once per solve so that every rank holds curr_x for the next local step.
// x and v stay replicated (frame begin / end are evaluated redundantly), so the C ABI is unchanged: every rank makes
// the same calls with the same data.
#include <dlfcn.h>
#include <nccl.h>

#include "common.h"

using namespace admmb;

// NCCL is resolved at run time (dlopen) instead of being a link-time dependency of libadmm_b200.so: a process that
// also imports PyTorch must end up with ONE libnccl.so.2, and PyTorch's bundled build is newer than the system one
// (loading the system library first breaks `import torch`).  dlopen by soname returns whichever copy the process
// already holds, or the system library otherwise.  Single-GPU use never touches NCCL.
namespace {
struct NcclApi {
	void *handle = nullptr;
	ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
	ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
	ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
	ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
	ncclResult_t (*GroupStart)() = nullptr;
	ncclResult_t (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(ncclResult_t) = nullptr;
	bool ok = false;
} g_nccl;

bool nccl_load() {
	if (g_nccl.ok) return true;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (!h) return false;
	g_nccl.handle = h;
	g_nccl.GetUniqueId = (ncclResult_t(*)(ncclUniqueId *))dlsym(h, "ncclGetUniqueId");
	g_nccl.CommInitRank = (ncclResult_t(*)(ncclComm_t *, int, ncclUniqueId, int))dlsym(h, "ncclCommInitRank");
	g_nccl.CommDestroy = (ncclResult_t(*)(ncclComm_t))dlsym(h, "ncclCommDestroy");
	g_nccl.AllGather = (ncclResult_t(*)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllGather");
	g_nccl.AllReduce = (ncclResult_t(*)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t))dlsym(h, "ncclAllReduce");
	g_nccl.Send = (ncclResult_t(*)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclSend");
	g_nccl.Recv = (ncclResult_t(*)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t))dlsym(h, "ncclRecv");
	g_nccl.GroupStart = (ncclResult_t(*)())dlsym(h, "ncclGroupStart");
	g_nccl.GroupEnd = (ncclResult_t(*)())dlsym(h, "ncclGroupEnd");
	g_nccl.GetErrorString = (const char *(*)(ncclResult_t))dlsym(h, "ncclGetErrorString");
	g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllGather && g_nccl.AllReduce && g_nccl.Send && g_nccl.Recv &&
	            g_nccl.GroupStart && g_nccl.GroupEnd && g_nccl.GetErrorString;
	return g_nccl.ok;
}
} // namespace

#define ADMMB_NCCL(ctx, call)                                                                                        \
	do {                                                                                                             \
		ncclResult_t _r = (call);                                                                                    \
		if (_r != ncclSuccess) ADMMB_FAIL(ctx, ADMMB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); \
	} while (0)

extern "C" int admmb_dist_unique_id(char *out128) {
	if (!out128) return ADMMB_E_ARG;
	if (!nccl_load()) return ADMMB_E_CUDA;
	ncclUniqueId id;
	if (g_nccl.GetUniqueId(&id) != ncclSuccess) return ADMMB_E_CUDA;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
	memcpy(out128, &id, sizeof(id));
	return ADMMB_OK;
}

extern "C" int admmb_dist_init(admmb_ctx *ctx, int rank, int world, const char *id128) {
	if (!ctx) return ADMMB_E_ARG;
	cudaSetDevice(ctx->device);
	if (ctx->finalized) ADMMB_FAIL(ctx, ADMMB_E_STATE, "admmb_dist_init must precede admmb_finalize");
	if (world < 1 || rank < 0 || rank >= world || !id128) ADMMB_FAIL(ctx, ADMMB_E_ARG, "bad rank / world / id");
	ctx->dist_rank = rank;
	ctx->dist_world = world;
	if (world == 1) return ADMMB_OK;
	if (!nccl_load()) ADMMB_FAIL(ctx, ADMMB_E_CUDA, "libnccl.so.2 could not be loaded: %s", dlerror());
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t comm;
	ADMMB_NCCL(ctx, g_nccl.CommInitRank(&comm, world, id, rank));
	ctx->nccl_comm = (void *)comm;
	return ADMMB_OK;
}

namespace admmb {

int dist_allgather_nodes(admmb_ctx *ctx, double *vec) {
	if (ctx->dist_world == 1) return ADMMB_OK;
	const size_t cnt = 3 * (size_t)ctx->chunk;
	ADMMB_NCCL(ctx, g_nccl.AllGather(vec + cnt * ctx->dist_rank, vec, cnt, ncclDouble, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	return ADMMB_OK;
}

int dist_allreduce_sum(admmb_ctx *ctx, double *dev, int count) {
	if (ctx->dist_world == 1) return ADMMB_OK;
	ADMMB_NCCL(ctx, g_nccl.AllReduce(dev, dev, count, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
	return ADMMB_OK;
}

// Neighbour-only halo exchange: one grouped send / receive per peer this rank shares matrix entries with.  send_buf holds the
// packed values for all peers (3 doubles per node), peer q's slice at node offset send_off[q]; the values from q land at
// recv_base + 3 * recv_off[q].
int dist_halo_exchange(admmb_ctx *ctx, const double *send_buf, const int *send_off, const int *send_cnt, double *recv_base, const int *recv_off, const int *recv_cnt) {
	if (ctx->dist_world == 1) return ADMMB_OK;
	ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
	ADMMB_NCCL(ctx, g_nccl.GroupStart());
	for (int q = 0; q < ctx->dist_world; ++q) {
		if (q == ctx->dist_rank) continue;
		if (send_cnt[q] > 0) ADMMB_NCCL(ctx, g_nccl.Send(send_buf + 3 * (size_t)send_off[q], 3 * (size_t)send_cnt[q], ncclDouble, q, comm, ctx->stream));
		if (recv_cnt[q] > 0) ADMMB_NCCL(ctx, g_nccl.Recv(recv_base + 3 * (size_t)recv_off[q], 3 * (size_t)recv_cnt[q], ncclDouble, q, comm, ctx->stream));
	}
	ADMMB_NCCL(ctx, g_nccl.GroupEnd());
	return ADMMB_OK;
}

// every rank contributes `bytes` host bytes; `all` receives world x bytes in rank order (setup-time exchange of IPC handles)
int dist_allgather_host(admmb_ctx *ctx, const void *mine, void *all, size_t bytes) {
	if (ctx->dist_world == 1) { memcpy(all, mine, bytes); return ADMMB_OK; }
	DevBuf<char> d;
	ADMMB_CUDA(ctx, d.alloc(bytes * ctx->dist_world));
	cudaError_t e = cudaMemcpyAsync(d.p + bytes * ctx->dist_rank, mine, bytes, cudaMemcpyHostToDevice, ctx->stream);
	ncclResult_t r = ncclSuccess;
	if (e == cudaSuccess) r = g_nccl.AllGather(d.p + bytes * ctx->dist_rank, d.p, bytes, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream);
	if (r == ncclSuccess && e == cudaSuccess) e = cudaMemcpyAsync(all, d.p, bytes * ctx->dist_world, cudaMemcpyDeviceToHost, ctx->stream);
	if (r == ncclSuccess && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	d.free();
	if (r != ncclSuccess) ADMMB_FAIL(ctx, ADMMB_E_CUDA, "ncclAllGather -> %s", g_nccl.GetErrorString(r));
	ADMMB_CUDA(ctx, e);
	return ADMMB_OK;
}

// sum of one host int over the ranks (setup-time agreement / barrier); synchronises the stream
int dist_allreduce_host_int(admmb_ctx *ctx, int *value) {
	if (ctx->dist_world == 1) return ADMMB_OK;
	DevBuf<double> d;
	ADMMB_CUDA(ctx, d.alloc(1));
	const double v = (double)*value;
	double out = 0.0;
	cudaError_t e = cudaMemcpyAsync(d.p, &v, sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
	ncclResult_t r = ncclSuccess;
	if (e == cudaSuccess) r = g_nccl.AllReduce(d.p, d.p, 1, ncclDouble, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream);
	if (r == ncclSuccess && e == cudaSuccess) e = cudaMemcpyAsync(&out, d.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
	if (r == ncclSuccess && e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	d.free();
	if (r != ncclSuccess) ADMMB_FAIL(ctx, ADMMB_E_CUDA, "ncclAllReduce -> %s", g_nccl.GetErrorString(r));
	ADMMB_CUDA(ctx, e);
	*value = (int)out;
	return ADMMB_OK;
}

void dist_destroy(admmb_ctx *ctx) {
	if (ctx->nccl_comm && g_nccl.ok) g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
	ctx->nccl_comm = nullptr;
}

} // namespace admmb
