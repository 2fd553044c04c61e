// kernels_global.cu -- per-frame vector kernels, the right-hand-side gather of the global step and the
// warm-started Jacobi-PCG alternative to the direct solve.
//
//   frame begin   explicit forces, x_bar = x + dt v, M x_bar, curr_x = x_bar      (System.cpp:37-48)
//   rhs gather    b = M x_bar + dt^2 D^T W^2 (z - u)                              (System.cpp:61)
//   pcg           curr_x = A^{-1} b, A = M + dt^2 D^T W^2 D (scalar n x n, 3 RHS) (System.cpp:62)
//   frame end     v = (curr_x - x) / dt, x = curr_x                               (System.cpp:70-71)
//
// All node vectors are [n][3] interleaved, internal (nested-dissection) node order.
#include "common.h"

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
// Jacobi-preconditioned conjugate gradients on A_n with three right-hand sides (x, y, z columns) sharing
// every matrix read.  Warm-started from the previous ADMM iterate already in curr_x.
// =====================================================================================================
struct PcgSolver {
	DevBuf<int> ptr, idx;
	DevBuf<double> val, dinv;
	DevBuf<double> r, p, Ap;
	DevBuf<double> scal; // [0..5] pAp[2][3]; parity q: rz[3] at 6+6q, rr[3] at 9+6q (adjacent: one all-reduce); [18..20] bb[3]; [21] best rr
	DevBuf<int> flag;    // [0] done, [1] iterations used, [2] iterations without progress
	int *h_flag = nullptr;
	int grid = 0;
};

#define PCG_THREADS 256
#define S_PAP(q) (3 * (q))
#define S_RZ(q) (6 + 6 * (q))
#define S_RR(q) (9 + 6 * (q))
#define S_BB 18
#define S_BEST 21

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

// r = b - A x; p = Dinv r; rz[0] = r.p; rr[0] = r.r; bb = b.b
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_init(int r0, int n, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                          const double *__restrict__ val, const double *__restrict__ dinv,
                                                          const double *__restrict__ b, const double *__restrict__ x,
                                                          double *__restrict__ r, double *__restrict__ p, double *scal) {
	// rows [r0, n) of this rank; ptr / dinv are indexed by the local row (i - r0), everything else by the global node
	double rz[3] = { 0, 0, 0 }, rr[3] = { 0, 0, 0 }, bb[3] = { 0, 0, 0 };
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double a[3] = { 0, 0, 0 };
		for (int q = ptr[i - r0]; q < ptr[i - r0 + 1]; ++q) {
			const double av = val[q];
			const double *xc = x + 3 * (size_t)idx[q];
			a[0] += av * xc[0]; a[1] += av * xc[1]; a[2] += av * xc[2];
		}
		const double di = dinv[i - r0];
		for (int j = 0; j < 3; ++j) {
			const double bj = b[3 * (size_t)i + j];
			const double rj = bj - a[j];
			r[3 * (size_t)i + j] = rj;
			const double zj = di * rj;
			p[3 * (size_t)i + j] = zj;
			rz[j] += rj * zj; rr[j] += rj * rj; bb[j] += bj * bj;
		}
	}
	block_reduce3_atomic(rz[0], rz[1], rz[2], scal + S_RZ(0));
	block_reduce3_atomic(rr[0], rr[1], rr[2], scal + S_RR(0));
	block_reduce3_atomic(bb[0], bb[1], bb[2], scal + S_BB);
}

__global__ void k_pcg_check0(double *scal, int *flag, double tol2) {
	// converged before the first iteration?
	bool done = true;
	for (int j = 0; j < 3; ++j) done = done && (scal[S_RR(0) + j] <= tol2 * scal[S_BB + j]);
	flag[0] = done ? 1 : 0;
	flag[1] = 0;
	flag[2] = 0;
}

// Ap = A p; pAp[k&1] += p.Ap; zero the accumulators the rest of this iteration adds into
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_spmv(int r0, int n, int k, const int *__restrict__ ptr, const int *__restrict__ idx,
                                                          const double *__restrict__ val, const double *__restrict__ p,
                                                          double *__restrict__ Ap, double *scal, const int *flag) {
	if (flag[0]) return;
	const int cur = k & 1, nxt = cur ^ 1;
	double s[3] = { 0, 0, 0 };
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		double a[3] = { 0, 0, 0 };
		for (int q = ptr[i - r0]; q < ptr[i - r0 + 1]; ++q) {
			const double av = val[q];
			const double *pc = p + 3 * (size_t)idx[q];
			a[0] += av * pc[0]; a[1] += av * pc[1]; a[2] += av * pc[2];
		}
		for (int j = 0; j < 3; ++j) { Ap[3 * (size_t)i + j] = a[j]; s[j] += a[j] * p[3 * (size_t)i + j]; }
	}
	if (blockIdx.x == 0 && threadIdx.x < 3) { scal[S_RZ(nxt) + threadIdx.x] = 0.0; scal[S_RR(nxt) + threadIdx.x] = 0.0; }
	block_reduce3_atomic(s[0], s[1], s[2], scal + S_PAP(cur));
}

// alpha = rz/pAp; x += alpha p; r -= alpha Ap; rz[nxt] += r.Dinv r; rr[nxt] += r.r
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_update(int r0, int n, int k, const double *__restrict__ dinv, const double *__restrict__ p,
                                                            const double *__restrict__ Ap, double *__restrict__ x,
                                                            double *__restrict__ r, double *scal, const int *flag) {
	if (flag[0]) return;
	const int cur = k & 1, nxt = cur ^ 1;
	double alpha[3];
	for (int j = 0; j < 3; ++j) {
		const double pAp = scal[S_PAP(cur) + j];
		alpha[j] = (pAp > 0.0) ? scal[S_RZ(cur) + j] / pAp : 0.0;
	}
	double rz[3] = { 0, 0, 0 }, rr[3] = { 0, 0, 0 };
	for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const double di = dinv[i - r0];
		for (int j = 0; j < 3; ++j) {
			const size_t t = 3 * (size_t)i + j;
			x[t] += alpha[j] * p[t];
			const double rj = r[t] - alpha[j] * Ap[t];
			r[t] = rj;
			rz[j] += rj * (di * rj); rr[j] += rj * rj;
		}
	}
	block_reduce3_atomic(rz[0], rz[1], rz[2], scal + S_RZ(nxt));
	block_reduce3_atomic(rr[0], rr[1], rr[2], scal + S_RR(nxt));
}

// beta = rz_new/rz_old; p = Dinv r + beta p; convergence test; zero pAp for the next iteration
__global__ void __launch_bounds__(PCG_THREADS) k_pcg_direction(int r0, int n, int k, const double *__restrict__ dinv, const double *__restrict__ r,
                                                               double *__restrict__ p, double *scal, int *flag, double tol2) {
	if (flag[0]) return;
	const int cur = k & 1, nxt = cur ^ 1;
	double beta[3];
	bool done = true;
	for (int j = 0; j < 3; ++j) {
		const double o = scal[S_RZ(cur) + j];
		beta[j] = (o > 0.0) ? scal[S_RZ(nxt) + j] / o : 0.0;
		done = done && (scal[S_RR(nxt) + j] <= tol2 * scal[S_BB + j]);
	}
	if (!done) {
		for (int i = r0 + blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
			const double di = dinv[i - r0];
			for (int j = 0; j < 3; ++j) {
				const size_t t = 3 * (size_t)i + j;
				p[t] = di * r[t] + beta[j] * p[t];
			}
		}
	}
	// All blocks have read scal[*cur*] above only through registers; the writes below touch slots that no
	// block of THIS kernel reads (pAp[nxt]) or that are only read by later kernels (flag).
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		scal[S_PAP(nxt) + 0] = 0.0; scal[S_PAP(nxt) + 1] = 0.0; scal[S_PAP(nxt) + 2] = 0.0;
		flag[1] = k + 1;
		// stagnation guard: once the residual sits at rounding level CG must not be iterated further (the recurrences
		// break down: p.Ap underflows and the iterate blows up).  The residual 2-norm of CG is NOT monotone -- on the
		// 1 M-tet cube it rises for dozens of iterations after a large step -- so the window is long: 300 iterations
		// without a 1 % improvement of the best residual seen.  A NaN stops at once.
		const double rr = scal[S_RR(nxt)] + scal[S_RR(nxt) + 1] + scal[S_RR(nxt) + 2];
		if (k == 0 || rr < 0.99 * scal[S_BEST]) { scal[S_BEST] = rr; flag[2] = 0; }
		else if (++flag[2] >= 300 || !(rr == rr)) flag[0] = 1;
	}
	if (done && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) flag[0] = 1;
}

int pcg_setup(admmb_ctx *ctx) {
	if (!ctx->pcg) ctx->pcg = new PcgSolver();
	PcgSolver &S = *ctx->pcg;
	const int n = ctx->n, r0 = ctx->own0, r1 = ctx->own1;
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
	ADMMB_CUDA(ctx, S.ptr.upload(ptr, ctx->stream));
	ADMMB_CUDA(ctx, S.idx.upload(idx, ctx->stream));
	ADMMB_CUDA(ctx, S.val.upload(val, ctx->stream));
	ADMMB_CUDA(ctx, S.dinv.upload(dinv, ctx->stream));
	ADMMB_CUDA(ctx, S.r.alloc(npad));
	ADMMB_CUDA(ctx, S.p.alloc(npad));
	ADMMB_CUDA(ctx, S.Ap.alloc(npad));
	ADMMB_CUDA(ctx, S.p.zero(ctx->stream));
	ADMMB_CUDA(ctx, S.scal.alloc(24));
	ADMMB_CUDA(ctx, S.flag.alloc(4));
	if (!S.h_flag) ADMMB_CUDA(ctx, cudaMallocHost((void **)&S.h_flag, 2 * sizeof(int)));
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
	const int want = (r1 - r0 + PCG_THREADS - 1) / PCG_THREADS;
	S.grid = want < sms * 4 ? (want > 0 ? want : 1) : sms * 4;
	(void)n;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
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
	k_pcg_init<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, S.ptr.p, S.idx.p, S.val.p, S.dinv.p, ctx->d_b.p, ctx->d_currx.p, S.r.p, S.p.p, S.scal.p);
	if ((rc = dist_allreduce_sum(ctx, S.scal.p + S_RZ(0), 15))) return rc; // rz[0], rr[0], (rz[1], rr[1] still zero), bb
	if ((rc = dist_allgather_nodes(ctx, S.p.p))) return rc;
	k_pcg_check0<<<1, 1, 0, s>>>(S.scal.p, S.flag.p, tol2);
	ctx->launches += 2;
	const int chunk = 32;
	int k = 0;
	while (k < ctx->cg_max_iters) {
		for (int c = 0; c < chunk && k < ctx->cg_max_iters; ++c, ++k) {
			const int cur = k & 1, nxt = cur ^ 1;
			k_pcg_spmv<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, k, S.ptr.p, S.idx.p, S.val.p, S.p.p, S.Ap.p, S.scal.p, S.flag.p);
			if ((rc = dist_allreduce_sum(ctx, S.scal.p + S_PAP(cur), 3))) return rc;
			k_pcg_update<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, k, S.dinv.p, S.p.p, S.Ap.p, ctx->d_currx.p, S.r.p, S.scal.p, S.flag.p);
			if ((rc = dist_allreduce_sum(ctx, S.scal.p + S_RZ(nxt), 6))) return rc;
			k_pcg_direction<<<S.grid, PCG_THREADS, 0, s>>>(r0, r1, k, S.dinv.p, S.r.p, S.p.p, S.scal.p, S.flag.p, tol2);
			if ((rc = dist_allgather_nodes(ctx, S.p.p))) return rc;
			ctx->launches += 3;
		}
		ADMMB_CUDA(ctx, cudaMemcpyAsync(S.h_flag, S.flag.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
		if (S.h_flag[0]) break;
	}
	ctx->cg_iters_total += S.h_flag[1];
	// every rank needs the whole iterate for the next local step
	if ((rc = dist_allgather_nodes(ctx, ctx->d_currx.p))) return rc;
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

void pcg_destroy(admmb_ctx *ctx) {
	if (!ctx->pcg) return;
	PcgSolver &S = *ctx->pcg;
	S.ptr.free(); S.idx.free(); S.val.free(); S.dinv.free(); S.r.free(); S.p.free(); S.Ap.free(); S.scal.free(); S.flag.free();
	if (S.h_flag) cudaFreeHost(S.h_flag);
	delete ctx->pcg;
	ctx->pcg = nullptr;
}

} // namespace admmb
