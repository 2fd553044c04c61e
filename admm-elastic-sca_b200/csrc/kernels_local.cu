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
#ifndef HYPER_THREADS
#define HYPER_THREADS 256
#endif
#ifndef HYPER_MIN_BLOCKS
#define HYPER_MIN_BLOCKS 3
#endif

template <int KIND, int MH>
__global__ void __launch_bounds__(LOCAL_THREADS) k_local_tets(const LocalArgs a) {
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	if (e < a.count) local_tet<KIND, MH>(a, e);
}
// hyperelastic tets: U, V of the SVD are parked in shared memory while the optimiser runs (local_bodies.h)
template <class Model, int MH>
__global__ void __launch_bounds__(HYPER_THREADS, HYPER_MIN_BLOCKS) k_local_tets_hyper(const LocalArgs a) {
	__shared__ double park[18 * HYPER_THREADS];
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

int launch_local_step(admmb_ctx *ctx, Batch &b, const double *d_x, double dt2) {
	(void)dt2;
	if (b.nlocal == 0) return ADMMB_OK;
	LocalArgs a;
	memset(&a, 0, sizeof(a));
	a.count = b.nlocal;
	a.idx = b.d_idx.p; a.S = b.d_S.p; a.w = b.d_w.p; a.wdt2 = b.d_wdt2.p; a.kk = b.d_kk.p; a.aux = b.d_aux.p;
	a.u = b.d_u.p; a.z = b.d_z.p; a.state = b.d_state.p; a.its = b.d_its.p; a.active = b.d_active.p;
	a.x = d_x;
	a.P = ctx->d_P.p + 3 * (size_t)b.slot_base;
	a.p0 = b.p0; a.p1 = b.p1; a.p2 = b.p2; a.max_iterations = b.max_iterations; a.flag = b.flag;
	a.shape_kind = b.d_shape_kind.p; a.shape_params = b.d_shape_params.p; a.nshapes = (int)b.shape_kind.size();
	const int grid = (b.nlocal + LOCAL_THREADS - 1) / LOCAL_THREADS;
	cudaStream_t s = ctx->stream;
	switch (b.type) {
	case BT_TETS:
		switch (b.kind) {
		case ADMMB_TET_LINEAR_STRAIN: k_local_tets<ADMMB_TET_LINEAR_STRAIN, 1><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TET_VOLUME: k_local_tets<ADMMB_TET_VOLUME, 1><<<grid, LOCAL_THREADS, 0, s>>>(a); break;
		case ADMMB_TET_NEOHOOKEAN: {
			const int g = (b.nlocal + HYPER_THREADS - 1) / HYPER_THREADS;
			if (b.max_iterations <= 5) k_local_tets_hyper<NHModel, 5><<<g, HYPER_THREADS, 0, s>>>(a);
			else k_local_tets_hyper<NHModel, 10><<<g, HYPER_THREADS, 0, s>>>(a);
			break;
		}
		case ADMMB_TET_STVK: {
			const int g = (b.nlocal + HYPER_THREADS - 1) / HYPER_THREADS;
			if (b.max_iterations <= 5) k_local_tets_hyper<StVKModel, 5><<<g, HYPER_THREADS, 0, s>>>(a);
			else k_local_tets_hyper<StVKModel, 10><<<g, HYPER_THREADS, 0, s>>>(a);
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

} // namespace admmb
