// kernels_global.cu -- per-frame vector kernels, the right-hand-side gather of the global step and the
// warm-started Jacobi-PCG alternative to the direct solve.
//
//   frame begin   explicit forces, x_bar = x + dt v, M x_bar, curr_x = x_bar      (System.cpp:37-48)
//   rhs gather    b = M x_bar + dt^2 D^T W^2 (z - u)                              (System.cpp:61)
//   pcg           curr_x = A^{-1} b, A = M + dt^2 D^T W^2 D (scalar n x n, 3 RHS) (System.cpp:62)
//   frame end     v = (curr_x - x) / dt, x = curr_x                               (System.cpp:70-71)
//
// All node vectors are [n][3] interleaved, internal (nested-dissection) node order.
#include <algorithm>

#include "common.h"
#include "dist_plan.h"

namespace admmb {

#define VEC_THREADS 256

__global__ void __launch_bounds__(VEC_THREADS) k_permute_rows3(int n, const int *__restrict__ perm,
                                                               const double *__restrict__ src, double *__restrict__ dst,
                                                               int scatter) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= 3 * n) return;
	const int i = t / 3, j = t - 3 * i;
	if (scatter) dst[3 * (size_t)perm[i] + j] = src[t]; // internal -> user
	else dst[t] = src[3 * (size_t)perm[i] + j];         // user -> internal
}

__global__ void __launch_bounds__(VEC_THREADS) k_permute_rows3_f32(int n, const int *__restrict__ perm, const double *__restrict__ src,
                                                                   float *__restrict__ dst) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= 3 * n) return;
	const int i = t / 3, j = t - 3 * i;
	dst[3 * (size_t)perm[i] + j] = __double2float_rn(src[t]); // internal -> user, (float) as the host cast rounds
}

int launch_permute_out_f32(admmb_ctx *ctx, const double *d_src_internal, float *d_dst_user) {
	const int n = ctx->n;
	k_permute_rows3_f32<<<(3 * n + VEC_THREADS - 1) / VEC_THREADS, VEC_THREADS, 0, ctx->stream>>>(n, ctx->d_node_perm.p, d_src_internal, d_dst_user);
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

int launch_permute_in(admmb_ctx *ctx, const double *d_src_user, double *d_dst_internal) {
	const int n = ctx->n;
	k_permute_rows3<<<(3 * n + VEC_THREADS - 1) / VEC_THREADS, VEC_THREADS, 0, ctx->stream>>>(n, ctx->d_node_perm.p, d_src_user, d_dst_internal, 0);
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}
int launch_permute_out(admmb_ctx *ctx, const double *d_src_internal, double *d_dst_user) {
	const int n = ctx->n;
	k_permute_rows3<<<(3 * n + VEC_THREADS - 1) / VEC_THREADS, VEC_THREADS, 0, ctx->stream>>>(n, ctx->d_node_perm.p, d_src_internal, d_dst_user, 1);
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

struct GravityList { int count; double g[8][3]; };

// ExplicitForce::project (ExplicitForce.cpp:29-39) for every registered direction, then System.cpp:46-48.
__global__ void __launch_bounds__(VEC_THREADS) k_frame_begin(int n3, double dt, GravityList G, const double *__restrict__ x,
                                                             double *__restrict__ v, const double *__restrict__ m,
                                                             double *__restrict__ xbar, double *__restrict__ Mxbar,
                                                             double *__restrict__ currx) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n3) return;
	const int i = t / 3, j = t - 3 * i;
	double vv = v[t];
	// __dmul_rn / __dadd_rn: never contracted into an FMA -- the reference rounds the product first (x86-64, no FMA)
	for (int g = 0; g < G.count; ++g) vv = __dadd_rn(vv, __dmul_rn(dt, G.g[g][j]));
	v[t] = vv;
	const double xb = __dadd_rn(x[t], __dmul_rn(dt, vv));
	xbar[t] = xb;
	Mxbar[t] = m[i] * xb;
	currx[t] = xb;
}

__global__ void __launch_bounds__(VEC_THREADS) k_frame_end(int n3, double inv_dt, double *__restrict__ x, double *__restrict__ v,
                                                           const double *__restrict__ currx, int *__restrict__ bad) {
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n3) return;
	const double cx = currx[t];
	v[t] = (cx - x[t]) * inv_dt;
	x[t] = cx;
	if (!isfinite(cx)) *bad = 1; // failure detection (admmb_set_check_finite): every writer stores the same value
}

int launch_frame_begin(admmb_ctx *ctx) {
	const int n3 = 3 * ctx->n;
	GravityList G;
	G.count = 0;
	// Explicit forces run in registration order (System.cpp:37-39).  The common case -- a few plain ExplicitForces --
	// is folded into the frame kernel; anything else (node subsets, wind) gets its own launch, in order.
	bool plain = ctx->explicit_forces.size() <= 8;
	for (const ExplicitEntry &e : ctx->explicit_forces) plain = plain && (e.kind == 0 || !e.enabled);
	if (plain) {
		for (const ExplicitEntry &e : ctx->explicit_forces) {
			if (!e.enabled) continue;
			for (int j = 0; j < 3; ++j) G.g[G.count][j] = e.dir[j];
			G.count++;
		}
	} else {
		for (ExplicitEntry &e : ctx->explicit_forces) {
			if (!e.enabled) continue;
			int rc = launch_explicit(ctx, e);
			if (rc) return rc;
		}
	}
	k_frame_begin<<<(n3 + VEC_THREADS - 1) / VEC_THREADS, VEC_THREADS, 0, ctx->stream>>>(n3, ctx->dt, G, ctx->d_x.p, ctx->d_v.p, ctx->d_m.p,
	                                                                                   ctx->d_xbar.p, ctx->d_Mxbar.p, ctx->d_currx.p);
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

int launch_frame_end(admmb_ctx *ctx) {
	const int n3 = 3 * ctx->n;
	k_frame_end<<<(n3 + VEC_THREADS - 1) / VEC_THREADS, VEC_THREADS, 0, ctx->stream>>>(n3, 1.0 / ctx->dt, ctx->d_x.p, ctx->d_v.p, ctx->d_currx.p, ctx->d_bad.p);
	ctx->launches++;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

// b_v = M x_bar_v + sum over the (force, corner) slots incident to node v of P[slot]   -- gather, deterministic.
__global__ void __launch_bounds__(VEC_THREADS) k_rhs_gather(int n, const int *__restrict__ vptr, const int *__restrict__ vslots,
                                                            const double *__restrict__ P, const double *__restrict__ Mxbar,
                                                            double *__restrict__ b, double *__restrict__ zero_y, double *__restrict__ zero_x) {
	asm volatile("griddepcontrol.wait;" ::: "memory"); // every local kernel has completed (no-op without PDL)
	const int v = blockIdx.x * blockDim.x + threadIdx.x;
	if (v >= n) return;
	// the direct solve accumulates into y and curr_x: cleared here instead of by two memset nodes per iteration
	if (zero_y) { zero_y[3 * (size_t)v + 0] = 0.0; zero_y[3 * (size_t)v + 1] = 0.0; zero_y[3 * (size_t)v + 2] = 0.0; }
	if (zero_x) { zero_x[3 * (size_t)v + 0] = 0.0; zero_x[3 * (size_t)v + 1] = 0.0; zero_x[3 * (size_t)v + 2] = 0.0; }
	double s0 = 0.0, s1 = 0.0, s2 = 0.0;
	const int p1 = vptr[v + 1];
	for (int p = vptr[v]; p < p1; ++p) {
		const double *c = P + 3 * (size_t)vslots[p];
		s0 += c[0]; s1 += c[1]; s2 += c[2];
	}
	b[3 * (size_t)v + 0] = Mxbar[3 * (size_t)v + 0] + s0;
	b[3 * (size_t)v + 1] = Mxbar[3 * (size_t)v + 1] + s1;
	b[3 * (size_t)v + 2] = Mxbar[3 * (size_t)v + 2] + s2;
}

int launch_rhs(admmb_ctx *ctx) {
	const int n = ctx->n;
	double *zy = nullptr, *zx = nullptr;
	if (ctx->solver == ADMMB_SOLVER_DIRECT && direct_vectors(ctx, &zy)) zx = ctx->d_currx.p;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((n + VEC_THREADS - 1) / VEC_THREADS); cfg.blockDim = dim3(VEC_THREADS); cfg.dynamicSmemBytes = 0; cfg.stream = ctx->stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr; cfg.numAttrs = ctx->use_pdl ? 1 : 0;
	ADMMB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_rhs_gather, n, (const int *)ctx->d_vert_ptr.p, (const int *)ctx->d_vert_slots.p, (const double *)ctx->d_P.p,
	                                   (const double *)ctx->d_Mxbar.p, ctx->d_b.p, zy, zx));
	ctx->solve_vectors_zeroed = (zx != nullptr);
	ctx->launches++;
	return ADMMB_OK;
}

// =====================================================================================================
// Jacobi-preconditioned conjugate gradients on A_n with three right-hand sides (x, y, z columns) sharing every matrix
// read, warm-started from the previous ADMM iterate already in curr_x.  Single-reduction (Chronopoulos-Gear) form:
//     u = Dinv r,  w = A u,  gamma = r.u,  delta = w.u                      (ONE fused reduction per iteration)
//     beta = gamma / gamma_old,  alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p,  s = w + beta s,  x += alpha p,  r -= alpha s
// i.e. two kernels per iteration (vector update; SpMV + the nine dot products) instead of three, and on a mesh partitioned
// over ranks one all-gather of u (the halo exchange in its simplest form) + ONE all-reduce of nine doubles instead of
// one all-gather + two all-reduces.  The iterations run as CUDA-graph chunks of PCG_CHUNK iterations with a device-side
// "done" flag (kernels of a finished solve return at once); the host looks at the flag once per batch of chunks, sized
// from the previous solve's iteration count, so a solve normally costs one host synchronisation.
// =====================================================================================================
#define PCG_MAXW 16
// Peer-to-peer exchange state handed to the CG kernels by value (world == 1: every loop below vanishes).
// flag ints (S.flag): [0] done, [1] iterations, [2] stagnation count, [3] p2p time-out, [4] halo pushes made, [5] dot pushes
// made, [6] blocks finished (last-block detection).  All ranks run the same kernel sequence, so "the n-th push" is the same
// event everywhere: a rank waits until a peer's flag reaches its OWN push count.
struct PeerArgs {
	int world = 1, rank = 0;
	const int *send_idx = nullptr;        // owned nodes the peers' rows reference, grouped by peer
	int send_off[PCG_MAXW + 1] = {};      // peer q's slice of send_idx
	int recv_cnt[PCG_MAXW] = {};          // nodes received from q (0: nothing to wait for)
	double *peer_halo[PCG_MAXW] = {};     // where this rank's slice lands in q's u (q's slots for this rank)
	double *peer_dots[PCG_MAXW] = {};     // q's mail + rank * 24: this rank's partial dots, [parity][12]
	int *peer_flags[PCG_MAXW] = {};       // q's flag words: [rank] halo, [world + 2 rank + parity] dots
	const double *dots_in = nullptr;      // own mail
	int *flags_in = nullptr;              // own flag words
};

struct PcgSolver {
	// partitioned rows: neighbour-only halo of u.  Columns owned by a peer are remapped (idxu) to slots behind the padded
	// vector, u[3 * (npad_nodes + slot)], which the peer's values are received into directly; send_idx lists the owned nodes
	// the peers need, grouped by peer.  ADMMB_PCG_HALO=0 falls back to an all-gather of the whole vector (round 1).
	bool halo = false;
	// p2p: the halo values and the partial dot products are written straight into the peers' memory over NVLink by the CG
	// kernels themselves (CUDA IPC mappings, release / acquire flags) -- no NCCL call inside the CG loop.  Falls back to the
	// grouped ncclSend / ncclRecv halo + ncclAllReduce when the mappings cannot be made (ADMMB_PCG_P2P=0 forces that).
	bool p2p = false;
	DevBuf<double> mail;                 // [world][2 parities][12] partial dots written by the peers, then the flags (ints)
	std::vector<void *> peer_u, peer_mail; // IPC mappings of the peers' u and mail
	std::vector<int> peer_recv_off;      // where this rank's values go in peer q's halo slots
	PeerArgs peer;                       // what the kernels get (world = 1 unless p2p)
	DevBuf<int> idxu, send_idx;
	DevBuf<double> send_buf;
	std::vector<int> send_off, send_cnt, recv_off, recv_cnt;
	int send_total = 0, recv_total = 0;
	size_t npad_nodes = 0;
	DevBuf<int> ptr, idx;
	DevBuf<double> val, dinv;
	DevBuf<double> r, u, w, p, s;
	DevBuf<double> scal; // see S_* below
	DevBuf<int> flag;    // [0] done, [1] iterations used, [2] iterations without progress
	int *h_flag = nullptr;
	int grid = 0;
	cudaGraphExec_t chunk_exec = nullptr;
	bool graph_failed = false;
	int last_iters = 0;
};

#define PCG_THREADS 256
#define PCG_CHUNK 16
// scal: [0..8] gamma, delta, rr of parity 0; [9..11] b.b; [12..20] the same dots of parity 1; [21..26] alpha_old, gamma_old
// read by parity 0 (written by parity 1); [27..32] the same read by parity 1; [33] best residual seen (stagnation guard);
// [34..36] b.b over all ranks
#define S_DOTS(q) (12 * (q))
#define S_BB 9
#define S_STATE(q) (21 + 6 * (q))
#define S_BEST 33
#define S_BBG 34 /* b.b summed over the ranks (peer-to-peer exchange) */
#define S_SIZE 40

__device__ __forceinline__ void st_release_sys(int *p, int v) { asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
	int v;
	asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
// bounded spin (2^28 polls of >= 64 ns: 20-30 s, far beyond any host-side stall between two ranks' launches): a peer that
// never arrives sets flag[3], which makes every later wait return at once, so the kernels drain and the host reports an
// error instead of the GPU hanging
__device__ __noinline__ void peer_wait(const int *word, int want, int *flag) {
	if (flag[3]) return;
	for (int spin = 0; spin < (1 << 28); ++spin) {
		if (ld_acquire_sys(word) >= want) return;
		__nanosleep(64);
	}
	flag[3] = 1;
}
// true in exactly one block: the last one to get here (call from all threads after the block's global writes)
__device__ __forceinline__ bool last_block_done(int *counter) {
	__shared__ int last;
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0) {
		const int t = atomicAdd(counter, 1);
		last = (t == (int)gridDim.x - 1);
		if (last) *counter = 0;
	}
	__syncthreads();
	if (last) __threadfence();
	return last != 0;
}
// last block of the kernel that produced u: write the nodes the peers need into their halo slots, then raise their flags
__device__ __forceinline__ void push_halo(const PeerArgs &P, const double *u, int *flag) {
	for (int q = 0; q < P.world; ++q) {
		const int c0 = 3 * P.send_off[q], c1 = 3 * P.send_off[q + 1];
		for (int t = c0 + (int)threadIdx.x; t < c1; t += blockDim.x) P.peer_halo[q][t - c0] = __ldcg(u + 3 * (size_t)P.send_idx[t / 3] + t % 3);
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const int v = flag[4] + 1;
		flag[4] = v;
		for (int q = 0; q < P.world; ++q)
			if (P.send_off[q + 1] > P.send_off[q]) st_release_sys(P.peer_flags[q] + P.rank, v);
	}
}
// every block, before it reads halo columns of u
__device__ __forceinline__ void wait_halo(const PeerArgs &P, int *flag) {
	if (P.world == 1) return;
	if ((int)threadIdx.x < P.world && (int)threadIdx.x != P.rank && P.recv_cnt[threadIdx.x] > 0) peer_wait(P.flags_in + threadIdx.x, flag[4], flag);
	__syncthreads();
}
// last block of the SpMV: this rank's 12 partial sums of parity q (dots, and b.b behind parity 0) to every peer
__device__ __forceinline__ void push_dots(const PeerArgs &P, const double *scal, int q, int *flag) {
	if (threadIdx.x < 12) {
		const double v = __ldcg(scal + 12 * q + threadIdx.x);
		for (int o = 0; o < P.world; ++o)
			if (o != P.rank) P.peer_dots[o][12 * q + threadIdx.x] = v;
	}
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const int v = flag[5] + 1;
		flag[5] = v;
		for (int o = 0; o < P.world; ++o)
			if (o != P.rank) st_release_sys(P.peer_flags[o] + P.world + 2 * P.rank + q, v);
	}
}

__device__ __forceinline__ void block_reduce3_atomic(double a0, double a1, double a2, double *dst) {
	__shared__ double sh[3][PCG_THREADS / 32];
	for (int o = 16; o > 0; o >>= 1) {
		a0 += __shfl_down_sync(0xffffffffu, a0, o);
		a1 += __shfl_down_sync(0xffffffffu, a1, o);
		a2 += __shfl_down_sync(0xffffffffu, a2, o);
	}
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (lane == 0) { sh[0][wid] = a0; sh[1][wid] = a1; sh[2][wid] = a2; }
	__syncthreads();
	if (wid == 0) {
		a0 = (lane < PCG_THREADS / 32) ? sh[0][lane] : 0.0;
		a1 = (lane < PCG_THREADS / 32) ? sh[1][lane] : 0.0;
		a2 = (lane < PCG_THREADS / 32) ? sh[2][lane] : 0.0;
		for (int o = 4; o > 0; o >>= 1) {
			a0 += __shfl_down_sync(0xffffffffu, a0, o);
			a1 += __shfl_down_sync(0xffffffffu, a1, o);
			a2 += __shfl_down_sync(0xffffffffu, a2, o);
		}
		if (lane == 0) { atomicAdd(dst + 0, a0); atomicAdd(dst + 1, a1); atomicAdd(dst + 2, a2); }
	}
	__syncthreads();
}

// r = b - A x; u = Dinv r; p = s = 0; bb = b.b; resets the scalars and the flags (rows [r0, n) of this rank; ptr / dinv are
// indexed by the local row i - r0, everything else by the global node)
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_init(int r0, int n, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                          const double *__restrict__ val, const double *__restrict__ dinv,
                                                          const double *__restrict__ b, const double *__restrict__ x,
                                                          double *__restrict__ r, double *__restrict__ u, double *__restrict__ p,
                                                          double *__restrict__ s, double *scal, int *flag, const PeerArgs P) {
	double bb[3] = { 0, 0, 0 };
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double a[3] = { 0, 0, 0 };
		for (int q = ptr[i - r0]; q < ptr[i - r0 + 1]; ++q) {
			const double av = val[q];
			const double *xc = x + 3 * (size_t)idx[q];
			a[0] += av * xc[0]; a[1] += av * xc[1]; a[2] += av * xc[2];
		}
		const double di = dinv[i - r0];
		for (int j = 0; j < 3; ++j) {
			const size_t t = 3 * (size_t)i + j;
			const double bj = b[t];
			const double rj = bj - a[j];
			r[t] = rj;
			u[t] = di * rj;
			p[t] = 0.0;
			s[t] = 0.0;
			bb[j] += bj * bj;
		}
	}
	block_reduce3_atomic(bb[0], bb[1], bb[2], scal + S_BB);
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (int j = 0; j < 3; ++j) { scal[S_STATE(0) + j] = 1.0; scal[S_STATE(0) + 3 + j] = __longlong_as_double(0x7ff0000000000000LL); } // alpha_old = 1, gamma_old = inf: beta = 0, alpha = gamma / delta
		flag[0] = 0; flag[1] = 0; flag[2] = 0;
	}
	if (P.world > 1 && last_block_done(flag + 6)) push_halo(P, u, flag);
}

// w = A u; dots[q] += (r.u, w.u, r.r)
// (u is deliberately NOT `const __restrict__`: on a partitioned mesh its halo slots are written by the peers while this kernel
// is already resident -- before wait_halo returns -- so its loads must not take the non-coherent LDG.CONSTANT path, whose
// contract is "unchanged for the lifetime of the kernel" and which the compiler may hoist above the wait)
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_spmv(int r0, int n, int q, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                          const double *__restrict__ val, const double *u,
                                                          const double *__restrict__ r, double *__restrict__ w, double *scal, int *flag, const PeerArgs P) {
	if (flag[0]) return;
	wait_halo(P, flag);
	double g[3] = { 0, 0, 0 }, d[3] = { 0, 0, 0 }, rr[3] = { 0, 0, 0 };
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double a[3] = { 0, 0, 0 };
		for (int e = ptr[i - r0]; e < ptr[i - r0 + 1]; ++e) {
			const double av = val[e];
			const double *uc = u + 3 * (size_t)idx[e];
			a[0] += av * uc[0]; a[1] += av * uc[1]; a[2] += av * uc[2];
		}
		for (int j = 0; j < 3; ++j) {
			const size_t t = 3 * (size_t)i + j;
			const double rj = r[t], uj = u[t];
			w[t] = a[j];
			g[j] += rj * uj; d[j] += a[j] * uj; rr[j] += rj * rj;
		}
	}
	block_reduce3_atomic(g[0], g[1], g[2], scal + S_DOTS(q) + 0);
	block_reduce3_atomic(d[0], d[1], d[2], scal + S_DOTS(q) + 3);
	block_reduce3_atomic(rr[0], rr[1], rr[2], scal + S_DOTS(q) + 6);
	if (P.world > 1 && last_block_done(flag + 6)) push_dots(P, scal, q, flag);
}

// convergence test on dots[q]; beta, alpha; p = u + beta p; s = w + beta s; x += alpha p; r -= alpha s; u = Dinv r
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_update(int r0, int n, int q, const double *__restrict__ dinv, double *__restrict__ u,
                                                            const double *__restrict__ w, double *__restrict__ p, double *__restrict__ s,
                                                            double *__restrict__ x, double *__restrict__ r, double *scal, int *flag, double tol2,
                                                            const PeerArgs P) {
	if (flag[0]) return;
	const int nq = q ^ 1;
	// the 9 dots of parity q and b.b: own partial sums, plus -- on a partitioned mesh with peer-to-peer exchange -- the peers'
	// from the mailbox, added in rank order so that every rank gets the same bits (and takes the same decisions)
	__shared__ double tot[12];
	if (threadIdx.x < 12) {
		const int t = threadIdx.x;
		const bool bbslot = t >= 9;
		double acc;
		if (P.world == 1) acc = scal[bbslot ? S_BB + t - 9 : S_DOTS(q) + t];
		else if (bbslot && q == 1) acc = scal[S_BBG + t - 9]; // b.b is exchanged behind parity 0 only; kept by the parity-0 update
		else {
			acc = 0.0;
			for (int o = 0; o < P.world; ++o) {
				if (o == P.rank) { acc += scal[12 * q + t]; continue; }
				peer_wait(P.flags_in + P.world + 2 * o + q, flag[5], flag);
				acc += __ldcg(P.dots_in + 24 * o + 12 * q + t);
			}
		}
		tot[t] = acc;
	}
	__syncthreads();
	double alpha[3], beta[3], gamma[3], rrs = 0.0;
	bool done = true;
	for (int j = 0; j < 3; ++j) {
		gamma[j] = tot[j];
		const double delta = tot[3 + j], rr = tot[6 + j];
		const double a_old = scal[S_STATE(q) + j], g_old = scal[S_STATE(q) + 3 + j];
		done = done && (rr <= tol2 * tot[9 + j]);
		rrs += rr;
		beta[j] = (g_old > 0.0) ? gamma[j] / g_old : 0.0;
		const double den = delta - beta[j] * gamma[j] / a_old;
		alpha[j] = (den > 0.0) ? gamma[j] / den : 0.0;   // breakdown (rounding-level residual): no step in this column
	}
	if (done) {
		if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) flag[0] = 1;
		return;
	}
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const double di = dinv[i - r0];
		for (int j = 0; j < 3; ++j) {
			const size_t t = 3 * (size_t)i + j;
			const double pj = u[t] + beta[j] * p[t];
			const double sj = w[t] + beta[j] * s[t];
			p[t] = pj; s[t] = sj;
			x[t] += alpha[j] * pj;
			const double rj = r[t] - alpha[j] * sj;
			r[t] = rj;
			u[t] = di * rj;
		}
	}
	if (blockIdx.x == 0 && threadIdx.x < 9) scal[S_DOTS(nq) + threadIdx.x] = 0.0; // the SpMV of this iteration accumulates here
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (int j = 0; j < 3; ++j) { scal[S_STATE(nq) + j] = alpha[j]; scal[S_STATE(nq) + 3 + j] = gamma[j]; }
		if (P.world > 1 && q == 0) for (int j = 0; j < 3; ++j) scal[S_BBG + j] = tot[9 + j];
		const int k = flag[1];
		flag[1] = k + 1;
		// stagnation guard: once the residual sits at rounding level CG must not be iterated further (the recurrences break
		// down and the iterate blows up).  The residual 2-norm of CG is NOT monotone -- on the 1 M-tet cube it rises for
		// dozens of iterations after a large step -- so the window is long: 300 iterations without a 1 % improvement of the
		// best residual seen.  A NaN stops at once.
		if (k == 0 || rrs < 0.99 * scal[S_BEST]) { scal[S_BEST] = rrs; flag[2] = 0; }
		else if (++flag[2] >= 300 || !(rrs == rrs)) flag[0] = 1;
	}
	if (P.world > 1 && last_block_done(flag + 6)) push_halo(P, u, flag);
}

// send_buf[t] = u[send_idx[t]] (3 doubles per node)
__global__ void __launch_bounds__(PCG_THREADS) k_halo_pack(int count, const int *__restrict__ send_idx, const double *__restrict__ u, double *__restrict__ out, const int *flag) {
	if (flag[0]) return;
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= 3 * count) return;
	out[t] = u[3 * (size_t)send_idx[t / 3] + t % 3];
}

static void pcg_p2p_teardown(admmb_ctx *ctx, PcgSolver &S) {
	bool any = false;
	for (void *q : S.peer_u) if (q) { cudaIpcCloseMemHandle(q); any = true; }
	for (void *q : S.peer_mail) if (q) { cudaIpcCloseMemHandle(q); any = true; }
	S.peer_u.clear(); S.peer_mail.clear();
	// nobody frees memory a peer still has mapped: all ranks tear down together (setup and destroy are collective calls)
	if ((any || S.p2p) && ctx->nccl_comm) { int one = 1; dist_allreduce_host_int(ctx, &one); }
	S.p2p = false;
	S.peer = PeerArgs();
}

// Maps every peer's u and mailbox (CUDA IPC), checks the mappings with a round of test writes and fills S.peer.  Any failure,
// on any rank, leaves p2p off on ALL ranks (the NCCL halo path is used instead); only a failing collective is an error.
static int pcg_p2p_setup(admmb_ctx *ctx, PcgSolver &S) {
	const int W = ctx->dist_world, me = ctx->dist_rank;
	S.p2p = false;
	S.peer = PeerArgs();
	if (W == 1 || !S.halo || W > PCG_MAXW) return ADMMB_OK;
	if (const char *e = getenv("ADMMB_PCG_P2P")) if (e[0] == '0') return ADMMB_OK;
	struct Table { cudaIpcMemHandle_t hu, hm; int recv_off[PCG_MAXW]; int ok; };
	Table mine;
	memset(&mine, 0, sizeof(mine));
	const size_t mail_doubles = (size_t)W * 24 + (3 * (size_t)W + 1) / 2 + 1;
	int ok = 1, rc;
	if (S.mail.alloc(mail_doubles) != cudaSuccess || cudaMemset(S.mail.p, 0, mail_doubles * sizeof(double)) != cudaSuccess) ok = 0;
	if (ok && cudaIpcGetMemHandle(&mine.hu, S.u.p) != cudaSuccess) ok = 0;
	if (ok && cudaIpcGetMemHandle(&mine.hm, S.mail.p) != cudaSuccess) ok = 0;
	for (int q = 0; q < W; ++q) mine.recv_off[q] = S.recv_off[q];
	mine.ok = ok;
	cudaGetLastError();
	std::vector<Table> all(W);
	if ((rc = dist_allgather_host(ctx, &mine, all.data(), sizeof(Table)))) return rc;
	for (int q = 0; q < W; ++q) ok = ok && all[q].ok;
	S.peer_u.assign(W, nullptr); S.peer_mail.assign(W, nullptr); S.peer_recv_off.assign(W, 0);
	for (int q = 0; q < W && ok; ++q) {
		if (q == me) continue;
		if (cudaIpcOpenMemHandle(&S.peer_u[q], all[q].hu, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { S.peer_u[q] = nullptr; ok = 0; break; }
		if (cudaIpcOpenMemHandle(&S.peer_mail[q], all[q].hm, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { S.peer_mail[q] = nullptr; ok = 0; break; }
		S.peer_recv_off[q] = all[q].recv_off[me];
	}
	cudaGetLastError();
	// test writes: rank r leaves 1000 + r in slot 23 of its block of every peer's mailbox (and in the last halo double it owns there)
	if (ok) {
		for (int q = 0; q < W; ++q) {
			if (q == me) continue;
			const double v = 1000.0 + me;
			if (cudaMemcpy((double *)S.peer_mail[q] + 24 * me + 23, &v, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) ok = 0;
			if (S.send_cnt[q] > 0 &&
			    cudaMemcpy((double *)S.peer_u[q] + 3 * (S.npad_nodes + S.peer_recv_off[q] + S.send_cnt[q]) - 1, &v, sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) ok = 0;
		}
		cudaDeviceSynchronize();
		cudaGetLastError();
	}
	int bad = ok ? 0 : 1;
	if ((rc = dist_allreduce_host_int(ctx, &bad))) return rc; // also the barrier between the test writes and the checks
	if (!bad) {
		for (int q = 0; q < W; ++q) {
			if (q == me) continue;
			double got[2] = { 0, 1000.0 + q };
			cudaMemcpy(&got[0], S.mail.p + 24 * q + 23, sizeof(double), cudaMemcpyDeviceToHost);
			if (S.recv_cnt[q] > 0) cudaMemcpy(&got[1], S.u.p + 3 * (S.npad_nodes + S.recv_off[q] + S.recv_cnt[q]) - 1, sizeof(double), cudaMemcpyDeviceToHost);
			if (got[0] != 1000.0 + q || got[1] != 1000.0 + q) ok = 0;
		}
		bad = ok ? 0 : 1;
		if ((rc = dist_allreduce_host_int(ctx, &bad))) return rc;
	}
	if (bad) {
		if (getenv("ADMMB_VERBOSE")) fprintf(stderr, "[setup] rank %d: peer-to-peer mappings not available, PCG exchanges through NCCL\n", me);
		pcg_p2p_teardown(ctx, S);
		return ADMMB_OK;
	}
	cudaMemset(S.mail.p, 0, mail_doubles * sizeof(double));
	PeerArgs &P = S.peer;
	P.world = W; P.rank = me;
	P.send_idx = S.send_idx.p;
	for (int q = 0; q < W; ++q) {
		P.send_off[q] = S.send_off[q];
		P.recv_cnt[q] = S.recv_cnt[q];
		if (q == me) continue;
		P.peer_halo[q] = (double *)S.peer_u[q] + 3 * (S.npad_nodes + S.peer_recv_off[q]);
		P.peer_dots[q] = (double *)S.peer_mail[q] + 24 * me;
		P.peer_flags[q] = (int *)((double *)S.peer_mail[q] + 24 * W);
	}
	P.send_off[W] = S.send_total;
	P.dots_in = S.mail.p;
	P.flags_in = (int *)(S.mail.p + 24 * W);
	S.p2p = true;
	int one = 1;
	if ((rc = dist_allreduce_host_int(ctx, &one))) return rc; // every mailbox is cleared before anyone starts pushing
	if (getenv("ADMMB_VERBOSE")) fprintf(stderr, "[setup] rank %d: PCG exchanges peer to peer (CUDA IPC over NVLink), no NCCL call inside the CG loop\n", me);
	return ADMMB_OK;
}

// u of the nodes the peers' rows reference: pack, grouped send / receive into the slots behind the vector
static int pcg_exchange_u(admmb_ctx *ctx, PcgSolver &S) {
	if (ctx->dist_world == 1 || S.p2p) return ADMMB_OK; // p2p: the kernel that produced u has pushed it
	if (!S.halo) return dist_allgather_nodes(ctx, S.u.p);
	if (S.send_total > 0) {
		k_halo_pack<<<(3 * S.send_total + PCG_THREADS - 1) / PCG_THREADS, PCG_THREADS, 0, ctx->stream>>>(S.send_total, S.send_idx.p, S.u.p, S.send_buf.p, S.flag.p);
		ctx->launches += 1;
	}
	return dist_halo_exchange(ctx, S.send_buf.p, S.send_off.data(), S.send_cnt.data(), S.u.p + 3 * S.npad_nodes, S.recv_off.data(), S.recv_cnt.data());
}

int pcg_setup(admmb_ctx *ctx) {
	if (!ctx->pcg) ctx->pcg = new PcgSolver();
	PcgSolver &S = *ctx->pcg;
	const int n = ctx->n, r0 = ctx->own0, r1 = ctx->own1;
	pcg_p2p_teardown(ctx, S); // a repeated setup (new weights) reallocates what the peers have mapped
	// this rank's rows of A_n (all rows when the mesh is not partitioned); column indices stay global
	std::vector<int> ptr(1, 0), idx;
	std::vector<double> val, dinv(std::max(r1 - r0, 1), 1.0);
	for (int i = r0; i < r1; ++i) {
		for (int q = ctx->A_ptr[i]; q < ctx->A_ptr[i + 1]; ++q) {
			idx.push_back(ctx->A_idx[q]);
			val.push_back(ctx->A_val[q]);
			if (ctx->A_idx[q] == i) {
				if (!(ctx->A_val[q] > 0.0)) ADMMB_FAIL(ctx, ADMMB_E_NUMERIC, "system matrix has a non-positive diagonal at node %d (zero mass?)", ctx->node_perm[i]);
				dinv[i - r0] = 1.0 / ctx->A_val[q];
			}
		}
		ptr.push_back((int)idx.size());
	}
	if (idx.empty()) { idx.push_back(0); val.push_back(0.0); }
	const size_t npad = 3 * (size_t)ctx->chunk * ctx->dist_world;
	// halo of the partitioned rows.  A_n is structurally symmetric, so "the owned nodes peer q's rows reference" is "the owned
	// rows that reference a node of q": both sides derive the same sorted list from their own rows.
	S.halo = ctx->dist_world > 1;
	if (const char *e = getenv("ADMMB_PCG_HALO")) if (e[0] == '0') S.halo = false;
	S.npad_nodes = (size_t)ctx->chunk * ctx->dist_world;
	S.send_total = S.recv_total = 0;
	if (S.halo) {
		HaloPlan H;
		plan_halo(ctx->A_ptr.data(), ctx->A_idx.data(), ctx->chunk, ctx->dist_world, r0, r1, H);
		S.send_off = H.send_off; S.send_cnt = H.send_cnt; S.recv_off = H.recv_off; S.recv_cnt = H.recv_cnt;
		S.send_total = H.send_total; S.recv_total = H.recv_total;
		std::vector<int> send_idx = H.send_idx;
		std::vector<int> idxu(idx.size());
		for (size_t e = 0; e < idx.size(); ++e) {
			const int j = idx[e];
			if ((j >= r0 && j < r1) || r1 <= r0) { idxu[e] = j; continue; }
			idxu[e] = (int)S.npad_nodes + H.slot_of(j, ctx->chunk);
		}
		if (send_idx.empty()) send_idx.push_back(0);
		ADMMB_CUDA(ctx, S.idxu.upload(idxu, ctx->stream));
		ADMMB_CUDA(ctx, S.send_idx.upload(send_idx, ctx->stream));
		ADMMB_CUDA(ctx, S.send_buf.alloc(3 * (size_t)std::max(S.send_total, 1)));
		if (getenv("ADMMB_VERBOSE")) fprintf(stderr, "[setup] rank %d: PCG halo: %d nodes sent, %d received per CG iteration (all-gather: %zu)\n", ctx->dist_rank, S.send_total, S.recv_total, S.npad_nodes);
	}
	ADMMB_CUDA(ctx, S.ptr.upload(ptr, ctx->stream));
	ADMMB_CUDA(ctx, S.idx.upload(idx, ctx->stream));
	ADMMB_CUDA(ctx, S.val.upload(val, ctx->stream));
	ADMMB_CUDA(ctx, S.dinv.upload(dinv, ctx->stream));
	ADMMB_CUDA(ctx, S.r.alloc(npad));
	ADMMB_CUDA(ctx, S.u.alloc(npad + 3 * (size_t)S.recv_total));
	ADMMB_CUDA(ctx, S.w.alloc(npad));
	ADMMB_CUDA(ctx, S.p.alloc(npad));
	ADMMB_CUDA(ctx, S.s.alloc(npad));
	ADMMB_CUDA(ctx, S.u.zero(ctx->stream));
	ADMMB_CUDA(ctx, S.scal.alloc(S_SIZE));
	ADMMB_CUDA(ctx, S.flag.alloc(8));
	ADMMB_CUDA(ctx, S.flag.zero(ctx->stream));
	if (!S.h_flag) ADMMB_CUDA(ctx, cudaMallocHost((void **)&S.h_flag, 4 * sizeof(int)));
	if (S.chunk_exec) { cudaGraphExecDestroy(S.chunk_exec); S.chunk_exec = nullptr; }
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	const int want = (r1 - r0 + PCG_THREADS - 1) / PCG_THREADS;
	S.grid = want < sms * 4 ? (want > 0 ? want : 1) : sms * 4;
	(void)n;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return pcg_p2p_setup(ctx, S);
}

// one CG iteration entering with the dots of parity q complete
static int pcg_enqueue_iteration(admmb_ctx *ctx, PcgSolver &S, int q, double tol2) {
	const int r0 = ctx->own0, r1 = ctx->own1;
	cudaStream_t s = ctx->stream;
	int rc;
	k_pcg_update<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, q, S.dinv.p, S.u.p, S.w.p, S.p.p, S.s.p, ctx->d_currx.p, S.r.p, S.scal.p, S.flag.p, tol2, S.peer);
	if ((rc = pcg_exchange_u(ctx, S))) return rc;
	k_pcg_spmv<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, q ^ 1, S.ptr.p, S.halo ? S.idxu.p : S.idx.p, S.val.p, S.u.p, S.r.p, S.w.p, S.scal.p, S.flag.p, S.peer);
	if (!S.p2p && (rc = dist_allreduce_sum(ctx, S.scal.p + S_DOTS(q ^ 1), 9))) return rc;
	ctx->launches += 2;
	return ADMMB_OK;
}

int pcg_solve(admmb_ctx *ctx) {
	PcgSolver &S = *ctx->pcg;
	const int r0 = ctx->own0, r1 = ctx->own1;
	cudaStream_t s = ctx->stream;
	const double tol = ctx->cg_tol < 1e-15 ? 1e-15 : ctx->cg_tol; // below ~4 eps the relative residual is not attainable
	const double tol2 = tol * tol;
	int rc;
	ADMMB_CUDA(ctx, S.scal.zero(s));
	k_pcg_init<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, S.ptr.p, S.idx.p, S.val.p, S.dinv.p, ctx->d_b.p, ctx->d_currx.p, S.r.p, S.u.p, S.p.p, S.s.p, S.scal.p, S.flag.p, S.peer);
	if ((rc = pcg_exchange_u(ctx, S))) return rc;
	k_pcg_spmv<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, 0, S.ptr.p, S.halo ? S.idxu.p : S.idx.p, S.val.p, S.u.p, S.r.p, S.w.p, S.scal.p, S.flag.p, S.peer);
	if (!S.p2p && (rc = dist_allreduce_sum(ctx, S.scal.p + S_DOTS(0), 12))) return rc; // the dots of parity 0 and b.b
	ctx->launches += 2;
	// chunk of PCG_CHUNK iterations (even: the parities repeat), captured once
	if (!S.chunk_exec && !S.graph_failed && ctx->use_graph) {
		cudaGraph_t g = nullptr;
		const long before = ctx->launches;
		cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
		rc = ADMMB_OK;
		if (e == cudaSuccess) {
			for (int c = 0; c < PCG_CHUNK && !rc; ++c) rc = pcg_enqueue_iteration(ctx, S, c & 1, tol2);
			e = cudaStreamEndCapture(s, &g);
		}
		if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&S.chunk_exec, g, 0);
		if (g) cudaGraphDestroy(g);
		ctx->launches = before;
		if (e != cudaSuccess || rc) { S.chunk_exec = nullptr; S.graph_failed = true; cudaGetLastError(); }
	}
	int k = 0;
	// chunks enqueued before the host looks at the flag: enough for the previous solve's iteration count
	int batch = std::max(1, (S.last_iters + PCG_CHUNK - 1) / PCG_CHUNK);
	S.h_flag[0] = 0; S.h_flag[1] = 0;
	while (k < ctx->cg_max_iters) {
		for (int c = 0; c < batch && k < ctx->cg_max_iters; ++c, k += PCG_CHUNK) {
			if (S.chunk_exec) { ADMMB_CUDA(ctx, cudaGraphLaunch(S.chunk_exec, s)); ctx->launches += (2 + ((S.halo && !S.p2p && S.send_total > 0) ? 1 : 0)) * PCG_CHUNK; }
			else for (int i = 0; i < PCG_CHUNK; ++i) if ((rc = pcg_enqueue_iteration(ctx, S, i & 1, tol2))) return rc;
		}
		ADMMB_CUDA(ctx, cudaMemcpyAsync(S.h_flag, S.flag.p, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
		if (S.h_flag[3]) ADMMB_FAIL(ctx, ADMMB_E_CUDA, "PCG peer-to-peer exchange timed out waiting for another rank (rank %d)", ctx->dist_rank);
		if (S.h_flag[0]) break;
		batch = 1;
	}
	S.last_iters = S.h_flag[1];
	ctx->cg_iters_total += S.h_flag[1];
	// every rank needs the whole iterate for the next local step
	if ((rc = dist_allgather_nodes(ctx, ctx->d_currx.p))) return rc;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

void pcg_destroy(admmb_ctx *ctx) {
	if (!ctx->pcg) return;
	PcgSolver &S = *ctx->pcg;
	pcg_p2p_teardown(ctx, S);
	S.mail.free();
	S.idxu.free(); S.send_idx.free(); S.send_buf.free();
	S.ptr.free(); S.idx.free(); S.val.free(); S.dinv.free(); S.r.free(); S.u.free(); S.w.free(); S.p.free(); S.s.free(); S.scal.free(); S.flag.free();
	if (S.chunk_exec) cudaGraphExecDestroy(S.chunk_exec);
	if (S.h_flag) cudaFreeHost(S.h_flag);
	delete ctx->pcg;
	ctx->pcg = nullptr;
}

} // namespace admmb
