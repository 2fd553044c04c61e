// kernels_local.cu -- the ADMM local step on the device: one fused kernel per force class that replaces,
// for every force of a batch,
//     Dx = m_D * curr_x                    (System.cpp:54, rows of this force only)
//     forces[i]->project(dt, Dx, u, z)     (System.cpp:57-58; prox + dual update u += D_i x - z)
// and the force's share of the right-hand side
//     dt^2 D_i^T W_i^2 (z_i - u_i)         (System.cpp:61)
// which is written per (force, node) "slot" into the contribution buffer P and summed per node by
// k_rhs_gather (kernels_global.cu) -- a gather, so no atomics and a deterministic sum.
//
// Layout: one thread per force; every per-force array is structure-of-arrays [component][count] so that
// a warp's loads and stores are coalesced; node positions are gathered from the interleaved x[n][3]
// (nodes are in nested-dissection order and forces in Morton order, so a block's gathers hit a few
// compact ranges).  All arithmetic is FP64 and follows elastic_math.h (the literal restatement of the
// reference's Eigen/cppoptlib arithmetic); this translation unit is compiled with -fmad=false so that no
// multiply-add is contracted where the reference (x86-64, no FMA) rounds twice.
#include "common.h"
#include "local_bodies.h"

namespace admmb {

#ifndef LOCAL_THREADS
#define LOCAL_THREADS 128
#endif
// Hyperelastic kernel shape, measured on the 1 M-tet cube (steady state, ms per launch): 256 x 3 blocks (80 registers,
// 24 warps / SM) 0.920; 320 x 2 (96 registers, 20 warps) 0.898; 256 x 2 (128 registers) 0.919; 288 x 2 0.979;
// 192 x 3 0.977; 256 x 4 (64 registers) 1.047; 128 x 6 1.007.  Fewer spills beat more warps up to 96 registers.
#ifndef HYPER_THREADS
#define HYPER_THREADS 320
#endif
#define PARK_SLOTS 18 // per thread: U and V of the SVD while the optimiser runs (local_bodies.h)
#ifndef HYPER_MIN_BLOCKS
#define HYPER_MIN_BLOCKS 2
#endif

template <int KIND, int MH>
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_tets(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_tet<KIND, MH>(a, e);
}
// hyperelastic tets: U, V of the SVD are parked in shared memory while the optimiser runs (local_bodies.h)
template <class Model, int MH>
__global__ void __launch_bounds__(HYPER_THREADS, HYPER_MIN_BLOCKS) k_local_tets_hyper(const LocalArgs a) {
	extern __shared__ double park[]; // PARK_SLOTS * HYPER_THREADS doubles (dynamic: more than the 48 KB a static array may have)
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_tet_hyper<Model, MH>(a, e, park + threadIdx.x, HYPER_THREADS);
}
template <int KIND>
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_tris(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_tri<KIND>(a, e);
}
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_springs(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_spring(a, e);
}
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_bends(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_bend(a, e);
}
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_anchors(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_anchor(a, e);
}
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_collision(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_collision(a, e);
}

// ---- small scenes: every force batch of the ADMM iteration in ONE launch ---------------------------------------------
// The reference's shipped scenes have 2.5-6 k elements in 1-3 force classes: each per-class kernel is a handful of CTAs
// and an iteration is bound by launch latency, not by work.  One kernel walks all batches (block range -> batch -> the
// same per-force body as the per-class kernels), launched with programmatic dependent launch so that its CTAs are already
// resident when the previous iteration's solve drains.  Used when the system has at most FUSED_MAX_FORCES forces.
#define FUSED_MAX 8
#define FUSED_MAX_FORCES 32768
struct FusedArgs {
	int nb;
	int block_first[FUSED_MAX + 1];
	int type[FUSED_MAX], kind[FUSED_MAX];
	LocalArgs a[FUSED_MAX];
};
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_fused(const __grid_constant__ FusedArgs F) {
	asm volatile("griddepcontrol.wait;" ::: "memory"); // curr_x of the previous solve is complete (no-op without PDL)
	__shared__ double park[PARK_SLOTS * LOCAL_THREADS];
	int b = 0;
	while (b + 1 < F.nb && (int)blockIdx.x >= F.block_first[b + 1]) ++b;
	const LocalArgs &a = F.a[b];
	const int e = ((int)blockIdx.x - F.block_first[b]) * LOCAL_THREADS + threadIdx.x;
	if (e >= a.count) return;
	const int kind = F.kind[b];
	switch (F.type[b]) {
	case BT_TETS:
		if (kind == ADMMB_TET_LINEAR_STRAIN) local_tet<ADMMB_TET_LINEAR_STRAIN, 1>(a, e);
		else if (kind == ADMMB_TET_VOLUME) local_tet<ADMMB_TET_VOLUME, 1>(a, e);
		else if (kind == ADMMB_TET_NEOHOOKEAN) {
			if (a.max_iterations <= 5) local_tet_hyper<NHModel, 5>(a, e, park + threadIdx.x, LOCAL_THREADS);
			else local_tet_hyper<NHModel, 10>(a, e, park + threadIdx.x, LOCAL_THREADS);
		} else {
			if (a.max_iterations <= 5) local_tet_hyper<StVKModel, 5>(a, e, park + threadIdx.x, LOCAL_THREADS);
			else local_tet_hyper<StVKModel, 10>(a, e, park + threadIdx.x, LOCAL_THREADS);
		}
		break;
	case BT_TRIS:
		if (kind == ADMMB_TRI_LIMITED_STRAIN) local_tri<ADMMB_TRI_LIMITED_STRAIN>(a, e);
		else if (kind == ADMMB_TRI_AREA) local_tri<ADMMB_TRI_AREA>(a, e);
		else local_tri<ADMMB_TRI_FUNG>(a, e);
		break;
	case BT_SPRINGS: local_spring(a, e); break;
	case BT_BENDS: local_bend(a, e); break;
	case BT_STATIC_ANCHORS:
	case BT_MOVING_ANCHORS: local_anchor(a, e); break;
	case BT_COLLISION: local_collision(a, e); break;
	}
}

#define HYPER_SMEM (PARK_SLOTS * HYPER_THREADS * sizeof(double))
// the hyperelastic kernels park more than the default 48 KB of dynamic shared memory per block: opt in, once per device
static int hyper_smem_opt_in(admmb_ctx *ctx) {
	static bool done[64] = { false };
	if (ctx->device >= 0 && ctx->device < 64 && done[ctx->device]) return ADMMB_OK;
	ADMMB_CUDA(ctx, cudaFuncSetAttribute(k_local_tets_hyper<NHModel, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HYPER_SMEM));
	ADMMB_CUDA(ctx, cudaFuncSetAttribute(k_local_tets_hyper<NHModel, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HYPER_SMEM));
	ADMMB_CUDA(ctx, cudaFuncSetAttribute(k_local_tets_hyper<StVKModel, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HYPER_SMEM));
	ADMMB_CUDA(ctx, cudaFuncSetAttribute(k_local_tets_hyper<StVKModel, 10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HYPER_SMEM));
	if (ctx->device >= 0 && ctx->device < 64) done[ctx->device] = true;
	return ADMMB_OK;
}

static void fill_local_args(admmb_ctx *ctx, Batch &b, const double *d_x, LocalArgs &a) {
	memset(&a, 0, sizeof(a));
	a.count = b.nlocal;
	a.idx = b.d_idx.p; a.S = b.d_S.p; a.w = b.d_w.p; a.wdt2 = b.d_wdt2.p; a.kk = b.d_kk.p; a.aux = b.d_aux.p;
	a.u = b.d_u.p; a.z = b.d_z.p; a.state = b.d_state.p; a.its = b.d_its.p; a.trips = b.d_trips.p; a.active = b.d_active.p;
	a.x = d_x;
	a.P = ctx->d_P.p + 3 * (size_t)b.slot_base;
	a.p0 = b.p0; a.p1 = b.p1; a.p2 = b.p2; a.max_iterations = b.max_iterations; a.flag = b.flag;
	a.kprox = std::min(b.p0, b.p1);
	a.shape_kind = b.d_shape_kind.p; a.shape_params = b.d_shape_params.p; a.nshapes = (int)b.shape_kind.size();
}

// The local step of every batch (System.cpp:57-58).  Small systems: one fused launch; otherwise one launch per batch.
int launch_local_all(admmb_ctx *ctx, const double *d_x) {
	long total = 0;
	int nb = 0;
	for (Batch &b : ctx->batches) { total += b.nlocal; nb += b.nlocal > 0; }
	if (nb == 0) return ADMMB_OK;
	if (ctx->fused_local && nb <= FUSED_MAX && total <= FUSED_MAX_FORCES) {
		FusedArgs F;
		memset(&F, 0, sizeof(F));
		int blocks = 0;
		for (Batch &b : ctx->batches) {
			if (b.nlocal == 0) continue;
			F.block_first[F.nb] = blocks;
			F.type[F.nb] = b.type; F.kind[F.nb] = b.kind;
			fill_local_args(ctx, b, d_x, F.a[F.nb]);
			blocks += (b.nlocal + LOCAL_THREADS - 1) / LOCAL_THREADS;
			F.nb++;
		}
		F.block_first[F.nb] = blocks;
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(LOCAL_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
		cudaLaunchAttribute attr[1];
		attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr[0].val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = attr; cfg.numAttrs = ctx->use_pdl ? 1 : 0;
		ADMMB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_local_fused, F));
		ctx->launches++;
		return ADMMB_OK;
	}
	for (Batch &b : ctx->batches) {
		int rc = launch_local_step(ctx, b, d_x, 0.0);
		if (rc) return rc;
	}
	return ADMMB_OK;
}

int launch_local_step(admmb_ctx *ctx, Batch &b, const double *d_x, double dt2) {
	(void)dt2;
	if (b.nlocal == 0) return ADMMB_OK;
	if (b.type == BT_TETS && (b.kind == ADMMB_TET_NEOHOOKEAN || b.kind == ADMMB_TET_STVK)) {
		int rc = hyper_smem_opt_in(ctx);
		if (rc) return rc;
	}
	LocalArgs a;
	fill_local_args(ctx, b, d_x, a);
	const int grid = (b.nlocal + LOCAL_THREADS - 1) / LOCAL_THREADS;
	cudaStream_t s = ctx->stream;
	switch (b.type) {
	case BT_TETS:
		switch (b.kind) {
		case ADMMB_TET_LINEAR_STRAIN: k_local_tets<ADMMB_TET_LINEAR_STRAIN, 1><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TET_VOLUME: k_local_tets<ADMMB_TET_VOLUME, 1><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TET_NEOHOOKEAN: {
			const int g = (b.nlocal + HYPER_THREADS - 1) / HYPER_THREADS;
			if (b.max_iterations <= 5) k_local_tets_hyper<NHModel, 5><<<g, HYPER_THREADS, HYPER_SMEM, s>>>(a);
			else k_local_tets_hyper<NHModel, 10><<<g, HYPER_THREADS, HYPER_SMEM, s>>>(a);
			break;
		}
		case ADMMB_TET_STVK: {
			const int g = (b.nlocal + HYPER_THREADS - 1) / HYPER_THREADS;
			if (b.max_iterations <= 5) k_local_tets_hyper<StVKModel, 5><<<g, HYPER_THREADS, HYPER_SMEM, s>>>(a);
			else k_local_tets_hyper<StVKModel, 10><<<g, HYPER_THREADS, HYPER_SMEM, s>>>(a);
			break;
		}
		}
		break;
	case BT_TRIS:
		switch (b.kind) {
		case ADMMB_TRI_LIMITED_STRAIN: a.p1 = b.p1; k_local_tris<ADMMB_TRI_LIMITED_STRAIN><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TRI_AREA: k_local_tris<ADMMB_TRI_AREA><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TRI_FUNG: k_local_tris<ADMMB_TRI_FUNG><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		}
		break;
	case BT_SPRINGS: k_local_springs<<<grid, LOCAL_THREADS, 0, s>>>(a); break;
	case BT_BENDS: k_local_bends<<<grid, LOCAL_THREADS, 0, s>>>(a); break;
	case BT_STATIC_ANCHORS:
	case BT_MOVING_ANCHORS: k_local_anchors<<<grid, LOCAL_THREADS, 0, s>>>(a); break;
	case BT_COLLISION: k_local_collision<<<grid, LOCAL_THREADS, 0, s>>>(a); break;
	}
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

// ExplicitForce::project (ExplicitForce.cpp:29-39): v += dt * direction, over all nodes or over `indices`.
__global__ void __launch_bounds__(LOCAL_THREADS) k_explicit_all(int n3, double g0, double g1, double g2, double dt, double *__restrict__ v) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n3) return;
	const int j = t % 3;
	v[t] += (dt * (j == 0 ? g0 : (j == 1 ? g1 : g2)));
}
// idx holds each listed node once, followed (second half) by how often the caller listed it: a node listed k times
// receives the increment k times, one rounding each, as in the reference's loop.
__global__ void __launch_bounds__(LOCAL_THREADS) k_explicit_subset(int count, const int *__restrict__ idx, double g0, double g1, double g2,
                                                                   double dt, double *__restrict__ v) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= count) return;
	double *vv = v + 3 * (size_t)idx[t];
	for (int k = idx[count + t]; k > 0; --k) { vv[0] += (dt * g0); vv[1] += (dt * g1); vv[2] += (dt * g2); }
}

// ---- explicit forces that need their own launch (frame begin) -------------------------------------------
// WindForce::project (ExplicitForce.cpp:42-98) with the semantics of the reference run on one thread: triangle t reads
// the velocities already updated by every earlier triangle that shares a node with it.  That dependency chain is
// resolved at upload time into wavefronts (level(t) = 1 + max level of the earlier triangles on its three nodes), so
// one CTA walks the levels with a barrier in between and the triangles of a level run in parallel -- same values,
// same order of additions per node, bit for bit.  (The reference's OpenMP version races on v; this is its serial meaning.)
#define WIND_THREADS 1024
__global__ void __launch_bounds__(WIND_THREADS) k_wind(int n_levels, const int *__restrict__ level_ptr, const int *__restrict__ tri, int count,
                                                       double d0, double d1, double d2, double dt, const double *__restrict__ x, double *v) {
	const double dir[3] = { d0, d1, d2 };
	for (int l = 0; l < n_levels; ++l) {
		const int t1 = level_ptr[l + 1];
		for (int t = level_ptr[l] + threadIdx.x; t < t1; t += WIND_THREADS) {
			const size_t a = 3 * (size_t)tri[t], b = 3 * (size_t)tri[count + t], c = 3 * (size_t)tri[2 * count + t];
			double f[3];
			wind_triangle(x + a, x + b, x + c, v + a, v + b, v + c, dir, dt, f);
			// v[idx[j]] += force for j = 0, 1, 2 in corner order: a corner listed twice receives the force twice
			v[a] += f[0]; v[a + 1] += f[1]; v[a + 2] += f[2];
			v[b] += f[0]; v[b + 1] += f[1]; v[b + 2] += f[2];
			v[c] += f[0]; v[c + 1] += f[1]; v[c + 2] += f[2];
		}
		__syncthreads();
	}
}

int upload_explicit(admmb_ctx *ctx, ExplicitEntry &e) {
	cudaStream_t s = ctx->stream;
	if (e.kind == 0 || e.count == 0) return ADMMB_OK;
	if (e.kind == 1) {
		std::vector<int> mult(ctx->n, 0), nodes;
		for (size_t i = 0; i < e.idx.size(); ++i)
			if (mult[e.idx[i]]++ == 0) nodes.push_back(e.idx[i]);
		e.n_levels = (int)nodes.size(); // unique nodes
		std::vector<int> ii(2 * nodes.size());
		for (size_t i = 0; i < nodes.size(); ++i) { ii[i] = ctx->node_iperm[nodes[i]]; ii[nodes.size() + i] = mult[nodes[i]]; }
		ADMMB_CUDA(ctx, e.d_idx.upload(ii, s));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
		return ADMMB_OK;
	}
	// wind: wavefront levels of the serial dependency chain, triangles stably regrouped by level
	const int T = e.count;
	std::vector<int> ptr, order;
	wind_wavefronts(ctx->n, T, e.idx.data(), ptr, order);
	const int depth = (int)ptr.size() - 1;
	std::vector<int> tri(3 * (size_t)T);
	for (int p = 0; p < T; ++p)
		for (int c = 0; c < 3; ++c) tri[(size_t)c * T + p] = ctx->node_iperm[e.idx[3 * (size_t)order[p] + c]];
	e.n_levels = depth;
	ADMMB_CUDA(ctx, e.d_idx.upload(tri, s));
	ADMMB_CUDA(ctx, e.d_level_ptr.upload(ptr, s));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	return ADMMB_OK;
}

int launch_explicit(admmb_ctx *ctx, ExplicitEntry &e) {
	cudaStream_t s = ctx->stream;
	if (e.kind == 0) {
		const int n = ctx->n;
		k_explicit_all<<<(3 * n + LOCAL_THREADS - 1) / LOCAL_THREADS, LOCAL_THREADS, 0, s>>>(3 * n, e.dir[0], e.dir[1], e.dir[2], ctx->dt, ctx->d_v.p);
	} else if (e.kind == 1) {
		if (e.count == 0) return ADMMB_OK;
		k_explicit_subset<<<(e.n_levels + LOCAL_THREADS - 1) / LOCAL_THREADS, LOCAL_THREADS, 0, s>>>(e.n_levels, e.d_idx.p, e.dir[0], e.dir[1], e.dir[2], ctx->dt, ctx->d_v.p);
	} else {
		if (e.count == 0) return ADMMB_OK;
		k_wind<<<1, WIND_THREADS, 0, s>>>(e.n_levels, e.d_level_ptr.p, e.d_idx.p, e.count, e.dir[0], e.dir[1], e.dir[2], ctx->dt, ctx->d_x.p, ctx->d_v.p);
	}
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

} // namespace admmb
