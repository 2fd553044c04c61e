// abi.cu -- the extern "C" surface of libadmm_b200.so (include/admm_b200.h): context lifetime, system
// construction, finalize (host setup + upload), the per-frame step driver and state access.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <numeric>
#include <omp.h>

#include "common.h"

using namespace admmb;

static std::string g_create_error;

extern "C" const char *admmb_version(void) { return "admm-elastic-sca_b200 0.1 (sm_100a)"; }

extern "C" const char *admmb_last_error(const admmb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

extern "C" int admmb_create(int device, admmb_ctx **out) {
	if (!out) return ADMMB_E_ARG;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0) {
		g_create_error = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU path";
		return ADMMB_E_CUDA;
	}
	if (device < 0 || device >= count) { g_create_error = "device index out of range"; return ADMMB_E_ARG; }
	e = cudaSetDevice(device);
	if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return ADMMB_E_CUDA; }
	admmb_ctx *ctx = new admmb_ctx();
	ctx->device = device;
	if (getenv("ADMMB_NO_GRAPH")) ctx->use_graph = false;
	if (const char *e = getenv("ADMMB_DETERMINISTIC")) ctx->deterministic = e[0] && e[0] != '0';
	if (const char *e = getenv("ADMMB_NO_FUSED_LOCAL")) ctx->fused_local = !(e[0] && e[0] != '0');
	if (const char *e = getenv("ADMMB_NO_PDL")) ctx->use_pdl = !(e[0] && e[0] != '0');
	e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
	if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete ctx; return ADMMB_E_CUDA; }
	*out = ctx;
	return ADMMB_OK;
}

static void drop_iteration_graph(admmb_ctx *ctx);

static void free_batch(Batch &b) {
	b.d_idx.free(); b.d_S.free(); b.d_w.free(); b.d_wdt2.free(); b.d_kk.free(); b.d_aux.free(); b.d_u.free(); b.d_z.free();
	b.d_state.free(); b.d_active.free(); b.d_its.free(); b.d_trips.free(); b.d_shape_kind.free(); b.d_shape_params.free();
}

extern "C" int admmb_destroy(admmb_ctx *ctx) {
	if (!ctx) return ADMMB_OK;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	for (Batch &b : ctx->batches) free_batch(b);
	for (ExplicitEntry &e : ctx->explicit_forces) { e.d_idx.free(); e.d_level_ptr.free(); }
	ctx->d_x.free(); ctx->d_v.free(); ctx->d_xbar.free(); ctx->d_Mxbar.free(); ctx->d_currx.free(); ctx->d_b.free(); ctx->d_m.free();
	ctx->d_bad.free(); ctx->d_io.free(); ctx->d_node_perm.free(); ctx->d_P.free(); ctx->d_vert_ptr.free(); ctx->d_vert_slots.free();
	drop_iteration_graph(ctx);
	pcg_destroy(ctx);
	direct_destroy(ctx);
	dist_destroy(ctx);
	for (cudaEvent_t ev : ctx->timing.ev) cudaEventDestroy(ev);
	for (int k = 0; k < 2; ++k) if (ctx->ev_region[k]) cudaEventDestroy(ctx->ev_region[k]);
	for (const auto &r : ctx->host_regs) cudaHostUnregister(r.first);
	if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
	if (ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
	return ADMMB_OK;
}

#define CHECK_CTX(ctx) do { if (!(ctx)) return ADMMB_E_ARG; cudaSetDevice((ctx)->device); } while (0)
#define CHECK_BUILDING(ctx) do { CHECK_CTX(ctx); if ((ctx)->finalized) ADMMB_FAIL(ctx, ADMMB_E_STATE, "system already finalized"); } while (0)
#define CHECK_READY(ctx) do { CHECK_CTX(ctx); if (!(ctx)->finalized) ADMMB_FAIL(ctx, ADMMB_E_STATE, "admmb_finalize has not been called"); \
	if ((ctx)->broken) ADMMB_FAIL(ctx, ADMMB_E_STATE, "the context is unusable: admmb_recompute_weights failed, the factor no longer matches the weights (call admmb_recompute_weights again with valid weights)"); } while (0)
// Between admmb_step_async and admmb_sync the context's staging areas (pinned host block, device I/O block) and the
// pending-download record hold the step's results: every entry point that would overwrite them refuses to run.
#define CHECK_NO_PENDING(ctx, what) do { if ((ctx)->pending_out[0] || (ctx)->pending_out[1]) \
	ADMMB_FAIL(ctx, ADMMB_E_STATE, what ": an asynchronous step is pending, collect it with admmb_sync first"); } while (0)

extern "C" int admmb_set_nodes(admmb_ctx *ctx, int n, const double *x3n, const double *m3n) {
	CHECK_BUILDING(ctx);
	if (n < 1 || !x3n || !m3n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "set_nodes: need n >= 1 and non-null x, m");
	if (ctx->n) ADMMB_FAIL(ctx, ADMMB_E_STATE, "set_nodes may be called once");
	ctx->h_m.resize(n);
	for (int i = 0; i < n; ++i) {
		// A = A_n (x) I_3 needs equal masses per coordinate (always true for ForceBuilder, ForceBuilder.hpp:139-145)
		if (m3n[3 * i] != m3n[3 * i + 1] || m3n[3 * i] != m3n[3 * i + 2])
			ADMMB_FAIL(ctx, ADMMB_E_NUMERIC, "node %d has different masses per coordinate; not supported", i);
		ctx->h_m[i] = m3n[3 * i];
	}
	ctx->n = n;
	ctx->h_x0.assign(x3n, x3n + 3 * (size_t)n);
	return ADMMB_OK;
}

static int new_batch(admmb_ctx *ctx, int type, int kind, int count, int nv, int rows, int nsel, int naux, int nstate, const int *idx) {
	if (ctx->n == 0) ADMMB_FAIL(ctx, ADMMB_E_STATE, "call admmb_set_nodes first");
	if (count < 0 || (count > 0 && !idx && type != BT_COLLISION)) ADMMB_FAIL(ctx, ADMMB_E_ARG, "bad count / idx");
	ctx->batches.emplace_back();
	Batch &b = ctx->batches.back();
	b.type = type; b.kind = kind; b.count = count; b.nv = nv; b.rows = rows; b.nsel = nsel; b.naux = naux; b.nstate = nstate;
	if (idx) b.idx.assign(idx, idx + (size_t)count * nv);
	for (size_t i = 0; i < b.idx.size(); ++i)
		if (b.idx[i] < 0 || b.idx[i] >= ctx->n) {
			const int bad = b.idx[i];
			ctx->batches.pop_back();
			ADMMB_FAIL(ctx, ADMMB_E_ARG, "node index %d out of range [0,%d)", bad, ctx->n);
		}
	return (int)ctx->batches.size() - 1;
}

extern "C" int admmb_add_tets(admmb_ctx *ctx, int kind, int count, const int *idx4, double p0, double p1, double p2, int max_iterations) {
	CHECK_BUILDING(ctx);
	if (kind < 0 || kind > 3) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown tet kind %d", kind);
	const bool hyper = (kind == ADMMB_TET_NEOHOOKEAN || kind == ADMMB_TET_STVK);
	if (hyper && max_iterations < 1) ADMMB_FAIL(ctx, ADMMB_E_ARG, "max_iterations must be >= 1");
	const int id = new_batch(ctx, BT_TETS, kind, count, 4, 9, 12, 0, hyper ? 4 : 0, idx4);
	if (id < 0) return id;
	Batch &b = ctx->batches[id];
	b.p0 = p0; b.p1 = p1; b.p2 = p2; b.max_iterations = max_iterations;
	return id;
}

extern "C" int admmb_add_tris(admmb_ctx *ctx, int kind, int count, const int *idx3, double stiffness, double limit_min, double limit_max, int flag) {
	CHECK_BUILDING(ctx);
	if (kind < 0 || kind > 2) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown triangle kind %d", kind);
	const int id = new_batch(ctx, BT_TRIS, kind, count, 3, 6, 6, 0, kind == ADMMB_TRI_FUNG ? 1 : 0, idx3);
	if (id < 0) return id;
	Batch &b = ctx->batches[id];
	b.p0 = stiffness; b.p1 = limit_min; b.p2 = limit_max; b.flag = flag;
	return id;
}

extern "C" int admmb_add_springs(admmb_ctx *ctx, int count, const int *idx2, const double *stiffness) {
	CHECK_BUILDING(ctx);
	if (count > 0 && !stiffness) ADMMB_FAIL(ctx, ADMMB_E_ARG, "springs need a stiffness array");
	const int id = new_batch(ctx, BT_SPRINGS, 0, count, 2, 3, 0, 1, 0, idx2);
	if (id < 0) return id;
	ctx->batches[id].stiffness.assign(stiffness, stiffness + count);
	return id;
}

extern "C" int admmb_add_bends(admmb_ctx *ctx, int count, const int *idx4, double stiffness) {
	CHECK_BUILDING(ctx);
	const int id = new_batch(ctx, BT_BENDS, 0, count, 4, 9, 0, 4, 0, idx4);
	if (id < 0) return id;
	ctx->batches[id].p0 = stiffness;
	return id;
}

extern "C" int admmb_add_static_anchors(admmb_ctx *ctx, int count, const int *idx, double weight) {
	CHECK_BUILDING(ctx);
	const int id = new_batch(ctx, BT_STATIC_ANCHORS, 0, count, 1, 3, 0, 3, 0, idx);
	if (id < 0) return id;
	ctx->batches[id].anchor_weight = weight;
	return id;
}

extern "C" int admmb_add_moving_anchors(admmb_ctx *ctx, int count, const int *idx, const double *pos3, double weight) {
	CHECK_BUILDING(ctx);
	if (count > 0 && !pos3) ADMMB_FAIL(ctx, ADMMB_E_ARG, "moving anchors need control point positions");
	const int id = new_batch(ctx, BT_MOVING_ANCHORS, 0, count, 1, 3, 0, 3, 0, idx);
	if (id < 0) return id;
	Batch &b = ctx->batches[id];
	b.anchor_weight = weight;
	b.stiffness.assign(pos3, pos3 + 3 * (size_t)count); // staged; becomes aux in compute_rest_state
	b.active.assign(count, 1);
	return id;
}

extern "C" int admmb_add_collision(admmb_ctx *ctx, int nshapes, const int *shape_kind, const double *params4, double weight) {
	CHECK_BUILDING(ctx);
	if (nshapes < 0 || (nshapes > 0 && (!shape_kind || !params4))) ADMMB_FAIL(ctx, ADMMB_E_ARG, "bad shape list");
	for (int i = 0; i < nshapes; ++i)
		if (shape_kind[i] < 0 || shape_kind[i] > 2) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown shape kind %d", shape_kind[i]);
	std::vector<int> ids(ctx->n);
	std::iota(ids.begin(), ids.end(), 0);
	const int id = new_batch(ctx, BT_COLLISION, 0, ctx->n, 1, 3, 0, 0, 0, ids.data());
	if (id < 0) return id;
	Batch &b = ctx->batches[id];
	b.anchor_weight = weight;
	b.shape_kind.assign(shape_kind, shape_kind + nshapes);
	b.shape_params.assign(params4, params4 + 4 * (size_t)nshapes);
	for (int i = 0; i < nshapes; ++i)
		if (shape_kind[i] == ADMMB_SHAPE_CYLINDER) b.shape_params[4 * i + 2] = 0.0; // CollisionCylinder.hpp:45
	return id;
}

// ---- explicit forces (System::explicit_forces, applied in registration order at the start of every frame) ----
static int add_explicit(admmb_ctx *ctx, int kind, int count, const int *idx, int per, const double *dir3, const char *what) {
	CHECK_BUILDING(ctx);
	if (!dir3) ADMMB_FAIL(ctx, ADMMB_E_ARG, "%s: null direction", what);
	if (count < 0 || (count > 0 && !idx)) ADMMB_FAIL(ctx, ADMMB_E_ARG, "%s: bad count / null indices", what);
	if (ctx->n == 0) ADMMB_FAIL(ctx, ADMMB_E_STATE, "%s: call admmb_set_nodes first", what);
	for (long i = 0; i < (long)count * per; ++i)
		if (idx[i] < 0 || idx[i] >= ctx->n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "%s: node index %d out of range", what, idx[i]);
	ExplicitEntry e;
	e.kind = kind;
	e.count = count;
	for (int j = 0; j < 3; ++j) e.dir[j] = dir3[j];
	if (count) e.idx.assign(idx, idx + (size_t)count * per);
	ctx->explicit_forces.push_back(e);
	return (int)ctx->explicit_forces.size() - 1;
}

extern "C" int admmb_set_gravity(admmb_ctx *ctx, int id, const double *dir3) {
	CHECK_CTX(ctx);
	if (!dir3) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null direction");
	if (id < 0) {
		// a plain ExplicitForce over all nodes may also be added after finalize (it needs no device data)
		ExplicitEntry e;
		for (int j = 0; j < 3; ++j) e.dir[j] = dir3[j];
		ctx->explicit_forces.push_back(e);
		return (int)ctx->explicit_forces.size() - 1;
	}
	if ((size_t)id >= ctx->explicit_forces.size()) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown explicit force id %d", id);
	for (int j = 0; j < 3; ++j) ctx->explicit_forces[id].dir[j] = dir3[j];
	return id;
}

extern "C" int admmb_enable_explicit(admmb_ctx *ctx, int id, int on) {
	CHECK_CTX(ctx);
	if (id < 0 || (size_t)id >= ctx->explicit_forces.size()) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown explicit force id %d", id);
	ctx->explicit_forces[id].enabled = on != 0;
	return ADMMB_OK;
}

extern "C" int admmb_add_explicit_subset(admmb_ctx *ctx, int count, const int *idx, const double *dir3) {
	if (!ctx) return ADMMB_E_ARG;
	if (count < 1) { CHECK_CTX(ctx); ADMMB_FAIL(ctx, ADMMB_E_ARG, "explicit subset: need count >= 1 (an empty index list means ALL nodes in the reference: use admmb_set_gravity)"); }
	return add_explicit(ctx, 1, count, idx, 1, dir3, "explicit subset");
}

extern "C" int admmb_add_wind(admmb_ctx *ctx, int ntris, const int *tris3, const double *dir3) {
	if (!ctx) return ADMMB_E_ARG;
	return add_explicit(ctx, 2, ntris, tris3, 3, dir3, "wind");
}

extern "C" int admmb_set_check_finite(admmb_ctx *ctx, int on) {
	CHECK_CTX(ctx);
	ctx->check_finite = on != 0;
	return ADMMB_OK;
}

// after a synchronisation point: has any frame since the last check produced a non-finite position?
static int check_finite(admmb_ctx *ctx) {
	if (!ctx->check_finite) return ADMMB_OK;
	int bad = 0;
	ADMMB_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (!bad) return ADMMB_OK;
	ADMMB_CUDA(ctx, ctx->d_bad.zero(ctx->stream));
	ADMMB_FAIL(ctx, ADMMB_E_NUMERIC, "positions are not finite after the step (diverged, or NaN / Inf in the inputs)");
}

extern "C" int admmb_set_deterministic(admmb_ctx *ctx, int on) {
	CHECK_BUILDING(ctx);
	ctx->deterministic = on != 0;
	return ADMMB_OK;
}

extern "C" int admmb_set_host_threads(int n) {
	omp_set_num_threads(n > 0 ? n : omp_get_num_procs());
	return ADMMB_OK;
}

extern "C" int admmb_set_solver(admmb_ctx *ctx, int solver, double tol, int max_cg_iters) {
	CHECK_BUILDING(ctx);
	if (solver != ADMMB_SOLVER_DIRECT && solver != ADMMB_SOLVER_PCG) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown solver %d", solver);
	ctx->solver = solver;
	if (tol > 0.0) ctx->cg_tol = tol;
	if (max_cg_iters > 0) ctx->cg_max_iters = max_cg_iters;
	return ADMMB_OK;
}

// ---- finalize -----------------------------------------------------------------------------------------
template <class T>
static void to_soa(const std::vector<T> &src, int ncomp, const std::vector<int> &perm, std::vector<T> &dst) {
	const size_t cnt = perm.size();
	dst.resize(cnt * ncomp);
	for (int k = 0; k < ncomp; ++k)
		for (size_t p = 0; p < cnt; ++p) dst[(size_t)k * cnt + p] = src[(size_t)perm[p] * ncomp + k];
}

static int upload_batch_weights(admmb_ctx *ctx, Batch &b) {
	std::vector<double> w, wdt2(b.nlocal);
	to_soa(b.w, 1, b.perm, w);
	const double dt2 = ctx->dt * ctx->dt;
	for (int p = 0; p < b.nlocal; ++p) wdt2[p] = dt2 * w[p] * w[p];
	ADMMB_CUDA(ctx, b.d_w.upload(w, ctx->stream));
	ADMMB_CUDA(ctx, b.d_wdt2.upload(wdt2, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ADMMB_OK;
}

static int upload_batch(admmb_ctx *ctx, Batch &b) {
	cudaStream_t s = ctx->stream;
	std::vector<int> idx_int(b.idx.size()), idx_soa;
	for (size_t i = 0; i < b.idx.size(); ++i) idx_int[i] = ctx->node_iperm[b.idx[i]];
	to_soa(idx_int, b.nv, b.perm, idx_soa);
	ADMMB_CUDA(ctx, b.d_idx.upload(idx_soa, s));
	std::vector<double> tmp;
	if (b.nsel) { to_soa(b.S, b.nsel, b.perm, tmp); ADMMB_CUDA(ctx, b.d_S.upload(tmp, s)); }
	to_soa(b.kk, 1, b.perm, tmp);
	ADMMB_CUDA(ctx, b.d_kk.upload(tmp, s));
	if (b.naux) { to_soa(b.aux, b.naux, b.perm, tmp); ADMMB_CUDA(ctx, b.d_aux.upload(tmp, s)); }
	if (b.type == BT_MOVING_ANCHORS) {
		std::vector<int> act;
		to_soa(b.active, 1, b.perm, act);
		ADMMB_CUDA(ctx, b.d_active.upload(act, s));
	}
	ADMMB_CUDA(ctx, b.d_u.alloc((size_t)b.rows * b.nlocal));
	ADMMB_CUDA(ctx, b.d_z.alloc((size_t)b.rows * b.nlocal));
	ADMMB_CUDA(ctx, b.d_u.zero(s)); // curr_u.setZero()  System.cpp:146
	ADMMB_CUDA(ctx, b.d_z.zero(s));
	if (b.nstate) {
		// HyperElasticTet: last_prox_result = (1,1,1) (TetForce.hpp:127-128), init_hess = 1 (cppoptlib/meta.h:33)
		std::vector<double> st((size_t)b.nstate * b.nlocal, 1.0);
		ADMMB_CUDA(ctx, b.d_state.upload(st, s));
		ADMMB_CUDA(ctx, b.d_its.alloc(b.nlocal));
		ADMMB_CUDA(ctx, b.d_its.zero(s));
		ADMMB_CUDA(ctx, b.d_trips.alloc(b.nlocal));
		ADMMB_CUDA(ctx, b.d_trips.zero(s));
	}
	if (b.type == BT_COLLISION) {
		ADMMB_CUDA(ctx, b.d_shape_kind.upload(b.shape_kind, s));
		ADMMB_CUDA(ctx, b.d_shape_params.upload(b.shape_params, s));
	}
	ADMMB_CUDA(ctx, cudaStreamSynchronize(s)); // host vectors above go out of scope
	return upload_batch_weights(ctx, b);
}

static int setup_solver(admmb_ctx *ctx) {
	auto t0 = std::chrono::steady_clock::now();
	assemble_system(ctx);
	auto t1 = std::chrono::steady_clock::now();
	int rc = (ctx->solver == ADMMB_SOLVER_PCG) ? pcg_setup(ctx) : direct_setup(ctx);
	ctx->factor_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (getenv("ADMMB_FACTOR_VERBOSE"))
		fprintf(stderr, "[setup] assemble A_n %.3f s, solver setup %.3f s\n", std::chrono::duration<double>(t1 - t0).count(),
		        ctx->factor_seconds - std::chrono::duration<double>(t1 - t0).count());
	return rc;
}

extern "C" int admmb_finalize(admmb_ctx *ctx, double timestep_s) {
	CHECK_BUILDING(ctx);
	if (ctx->n == 0) ADMMB_FAIL(ctx, ADMMB_E_STATE, "no nodes");
	if (timestep_s <= 0.0) timestep_s = 0.04; // System.cpp:103-107
	ctx->dt = timestep_s;
	const int n = ctx->n;
	cudaStream_t s = ctx->stream;

	for (Batch &b : ctx->batches) {
		int rc = compute_rest_state(ctx, b);
		if (rc) return rc;
	}
	// node order (nested dissection) from the node graph
	{
		std::vector<int> gp, gi;
		build_node_graph(ctx, gp, gi);
		std::vector<int> blocks;
		int leaf = 64; // dissection stops at sub-domains of this many nodes (= the leaf supernodes); ADMMB_ND_LEAF overrides (tuning knob)
		// small systems (the reference's shipped scenes: ~1 000 nodes): no dissection at all -- ONE dense supernode, so the
		// solve is two launches (forward, backward) over a factor that lives in the L2 instead of 2 x 5 latency-bound levels
		if (n <= 2048) leaf = n;
		if (const char *e = getenv("ADMMB_ND_LEAF")) { const int v = atoi(e); if (v >= 8 && v <= 8192) leaf = v; }
		compute_node_order(n, ctx->h_x0.data(), gp, gi, leaf, ctx->node_perm, blocks);
		ctx->node_iperm.assign(n, -1);
		for (int i = 0; i < n; ++i) ctx->node_iperm[ctx->node_perm[i]] = i;
		direct_set_blocks(ctx, blocks); // dissection blocks = supernode partition of the direct solver
	}
	// node ownership when the mesh is partitioned over ranks: contiguous chunks of the internal order
	ctx->chunk = (n + ctx->dist_world - 1) / ctx->dist_world;
	ctx->own0 = std::min(n, ctx->dist_rank * ctx->chunk);
	ctx->own1 = std::min(n, ctx->own0 + ctx->chunk);
	// (both solvers run on a partitioned mesh: PCG partitions its rows; the direct solve does not shard -- sparse triangular
	// solves are a dependency chain -- so the right-hand side is all-gathered and every rank solves redundantly)
	long row = 0, slot = 0;
	for (Batch &b : ctx->batches) {
		if (b.type == BT_COLLISION) { b.perm = ctx->node_perm; }
		else {
			morton_order_elements(ctx, b);
			if (ctx->dist_world > 1) {
				// keep the forces that touch an owned node: their contributions complete the owned rows of the right-hand
				// side without any exchange (boundary forces are evaluated on both sides, deterministically)
				std::vector<int> keep;
				for (int e : b.perm) {
					bool mine = false;
					for (int c = 0; c < b.nv && !mine; ++c) {
						const int v = ctx->node_iperm[b.idx[(size_t)e * b.nv + c]];
						mine = (v >= ctx->own0 && v < ctx->own1);
					}
					if (mine) keep.push_back(e);
				}
				b.perm.swap(keep);
			}
		}
		b.nlocal = (int)b.perm.size();
		b.row_base = row; b.slot_base = slot;
		row += (long)b.rows * b.count;
		slot += (long)b.nv * b.nlocal;
		int rc = upload_batch(ctx, b);
		if (rc) return rc;
	}
	ctx->n_rows = row;
	ctx->n_slots = slot;
	ADMMB_CUDA(ctx, ctx->d_P.alloc(3 * (size_t)std::max<long>(slot, 1)));
	ADMMB_CUDA(ctx, ctx->d_P.zero(s));
	// node -> slots adjacency (gather lists of the right-hand side)
	{
		std::vector<int> vptr(n + 1, 0), vslots((size_t)slot);
		for (const Batch &b : ctx->batches)
			for (int p = 0; p < b.nlocal; ++p)
				for (int c = 0; c < b.nv; ++c) vptr[ctx->node_iperm[b.idx[(size_t)b.perm[p] * b.nv + c]] + 1]++;
		for (int i = 0; i < n; ++i) vptr[i + 1] += vptr[i];
		std::vector<int> fill(vptr.begin(), vptr.end() - 1);
		for (const Batch &b : ctx->batches)
			for (int p = 0; p < b.nlocal; ++p)
				for (int c = 0; c < b.nv; ++c) {
					const int v = ctx->node_iperm[b.idx[(size_t)b.perm[p] * b.nv + c]];
					vslots[fill[v]++] = (int)(b.slot_base + (long)p * b.nv + c);
				}
		ADMMB_CUDA(ctx, ctx->d_vert_ptr.upload(vptr, s));
		ADMMB_CUDA(ctx, ctx->d_vert_slots.alloc(std::max<size_t>(vslots.size(), 1)));
		ADMMB_CUDA(ctx, ctx->d_vert_slots.upload(vslots.data(), vslots.size(), s));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	}
	// node vectors
	{
		const size_t npad = (size_t)ctx->chunk * ctx->dist_world; // >= n; padding keeps the all-gather chunks equal
		std::vector<double> xi(3 * npad, 0.0), mi(n);
		for (int i = 0; i < n; ++i) {
			const int u = ctx->node_perm[i];
			for (int j = 0; j < 3; ++j) xi[3 * (size_t)i + j] = ctx->h_x0[3 * (size_t)u + j];
			mi[i] = ctx->h_m[u];
		}
		ADMMB_CUDA(ctx, ctx->d_x.upload(xi, s));
		ADMMB_CUDA(ctx, ctx->d_m.upload(mi, s));
		ADMMB_CUDA(ctx, ctx->d_node_perm.upload(ctx->node_perm, s));
		ADMMB_CUDA(ctx, ctx->d_v.alloc(3 * npad));
		ADMMB_CUDA(ctx, ctx->d_v.zero(s)); // m_v.setZero()  System.cpp:113
		ADMMB_CUDA(ctx, ctx->d_xbar.alloc(3 * npad));
		ADMMB_CUDA(ctx, ctx->d_Mxbar.alloc(3 * npad));
		ADMMB_CUDA(ctx, ctx->d_currx.upload(xi, s));
		ADMMB_CUDA(ctx, ctx->d_b.alloc(3 * npad));
		ADMMB_CUDA(ctx, ctx->d_bad.alloc(1));
		ADMMB_CUDA(ctx, ctx->d_bad.zero(s));
		ADMMB_CUDA(ctx, ctx->d_io.alloc(6 * (size_t)n));
		if (ctx->h_pin) { cudaFreeHost(ctx->h_pin); ctx->h_pin = nullptr; } // a retry after a failed finalize
		ADMMB_CUDA(ctx, cudaMallocHost((void **)&ctx->h_pin, 6 * (size_t)n * sizeof(double)));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	}
	for (ExplicitEntry &e : ctx->explicit_forces) {
		int rc = upload_explicit(ctx, e);
		if (rc) return rc;
	}
	int rc = setup_solver(ctx);
	if (rc) return rc;
	ctx->finalized = true;
	ctx->launches = 0;
	return ADMMB_OK;
}

// ---- step -----------------------------------------------------------------------------------------------
static cudaEvent_t next_event(admmb_ctx *ctx) {
	Timing &T = ctx->timing;
	if (T.used == T.ev.size()) {
		cudaEvent_t e;
		cudaEventCreate(&e);
		T.ev.push_back(e);
	}
	cudaEvent_t e = T.ev[T.used++];
	cudaEventRecord(e, ctx->stream);
	return e;
}

struct DumpTarget { double *x_it, *z_it, *u_it; };

static void drop_iteration_graph(admmb_ctx *ctx) {
	if (ctx->iter_graph_exec) cudaGraphExecDestroy(ctx->iter_graph_exec);
	if (ctx->iter_graph) cudaGraphDestroy(ctx->iter_graph);
	ctx->iter_graph_exec = nullptr;
	ctx->iter_graph = nullptr;
	for (int ph = 0; ph < 3; ++ph) {
		if (ctx->phase_graph_exec[ph]) cudaGraphExecDestroy(ctx->phase_graph_exec[ph]);
		ctx->phase_graph_exec[ph] = nullptr;
	}
}

// Right-hand side (System.cpp:61).  On a mesh partitioned over ranks every rank has evaluated the forces that touch its own
// nodes, so its OWN rows of b are complete (and bit-identical to the single-GPU sum: same slots, same order); with the
// direct solver, which every rank runs redundantly on the whole vector, the owned chunks are all-gathered in place
// (NCCL over NVLink, 3n doubles per ADMM iteration: 4.2 MB at 1 M tets).  PCG only needs the owned rows.
static int rhs_phase(admmb_ctx *ctx) {
	int rc = launch_rhs(ctx);
	if (rc) return rc;
	if (ctx->dist_world > 1 && ctx->solver == ADMMB_SOLVER_DIRECT) rc = dist_allgather_nodes(ctx, ctx->d_b.p);
	return rc;
}
// Solve (System.cpp:62).  Partitioned mesh, direct solver: by default the solve is sharded by subtrees of the elimination
// tree (direct_solve.cu: every rank streams only its subtrees' tiles and the replicated top of the tree, one small
// all-reduce of the top rows in between, one all-reduce of x at the end -- all ranks hold the same curr_x bit for bit).
// With ADMMB_DIST_SOLVE=replicated every rank solves the whole system; the default (atomic) solve then agrees between ranks
// only to rounding, so every rank keeps ITS chunk and the chunks are all-gathered.  With the deterministic solve (always
// replicated) the ranks' solutions are already identical and the exchange is skipped.
static int solve_phase(admmb_ctx *ctx) {
	if (ctx->solver == ADMMB_SOLVER_PCG) return pcg_solve(ctx);
	int rc = direct_solve(ctx);
	if (rc) return rc;
	if (ctx->dist_world > 1 && !ctx->deterministic && !direct_is_sharded(ctx)) rc = dist_allgather_nodes(ctx, ctx->d_currx.p);
	return rc;
}

// One ADMM iteration (System.cpp:51-66): local step of every batch, right-hand side, solve.
static int enqueue_iteration(admmb_ctx *ctx) {
	const double dt2 = ctx->dt * ctx->dt;
	int rc = launch_local_all(ctx, ctx->d_currx.p);
	if (rc) return rc;
	if ((rc = rhs_phase(ctx))) return rc;
	return solve_phase(ctx);
}

static int run_iterations(admmb_ctx *ctx, int admm_iters, const DumpTarget *dump = nullptr) {
	const bool timed = ctx->timing.on && !dump;
	const double dt2 = ctx->dt * ctx->dt;
	// Fast path: the iteration is a fixed sequence of ~35 small launches with fixed arguments, so it is captured
	// once into a CUDA graph and replayed (the PCG solve polls a convergence flag from the host and cannot be).
	if (!timed && !dump && ctx->use_graph && ctx->solver == ADMMB_SOLVER_DIRECT && admm_iters > 0) {
		if (!ctx->iter_graph_exec) {
			const long before = ctx->launches;
			ADMMB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
			int rc = enqueue_iteration(ctx);
			cudaError_t e = cudaStreamEndCapture(ctx->stream, &ctx->iter_graph);
			if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&ctx->iter_graph_exec, ctx->iter_graph, 0);
			ctx->iter_graph_launches = ctx->launches - before;
			ctx->launches = before;
			if (rc || e != cudaSuccess) {
				drop_iteration_graph(ctx);
				cudaGetLastError();
				if (ctx->dist_world > 1) {
					// the collectives of a partitioned mesh could not be captured (NCCL build without graph support): plain launches
					ctx->use_graph = false;
					return run_iterations(ctx, admm_iters, dump);
				}
				if (rc) return rc;
				ADMMB_CUDA(ctx, e);
			}
		}
		for (int it = 0; it < admm_iters; ++it) ADMMB_CUDA(ctx, cudaGraphLaunch(ctx->iter_graph_exec, ctx->stream));
		ctx->launches += ctx->iter_graph_launches * admm_iters;
		return ADMMB_OK;
	}
	// Timed mode with the direct solver: the frame's iterations are captured ONCE into a graph whose kernel nodes are the
	// production ones and whose phase boundaries are event-record nodes (cudaEventRecordExternal), so that the per-phase
	// timings see the launch behaviour of the production path -- no extra graph launches or host work between phases.
	if (timed && !dump && ctx->use_graph && ctx->solver == ADMMB_SOLVER_DIRECT && admm_iters > 0) {
		Timing &T = ctx->timing;
		const size_t first = T.used; // events [first, first + 4 * admm_iters) belong to this frame's iterations
		if (!ctx->phase_graph_exec[0] || ctx->timed_graph_iters != admm_iters || ctx->timed_graph_first != first) {
			if (ctx->phase_graph_exec[0]) { cudaGraphExecDestroy(ctx->phase_graph_exec[0]); ctx->phase_graph_exec[0] = nullptr; }
			while (T.ev.size() < first + 4 * (size_t)admm_iters) { cudaEvent_t e; cudaEventCreate(&e); T.ev.push_back(e); }
			const long before = ctx->launches;
			cudaGraph_t g = nullptr;
			ADMMB_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
			int rc = ADMMB_OK;
			cudaError_t e = cudaSuccess;
			size_t k = first;
			for (int it = 0; it < admm_iters && !rc && e == cudaSuccess; ++it) {
				e = cudaEventRecordWithFlags(T.ev[k++], ctx->stream, cudaEventRecordExternal);
				if (!rc) rc = launch_local_all(ctx, ctx->d_currx.p);
				if (e == cudaSuccess) e = cudaEventRecordWithFlags(T.ev[k++], ctx->stream, cudaEventRecordExternal);
				if (!rc) rc = rhs_phase(ctx);
				if (e == cudaSuccess) e = cudaEventRecordWithFlags(T.ev[k++], ctx->stream, cudaEventRecordExternal);
				if (!rc) rc = solve_phase(ctx);
				if (e == cudaSuccess) e = cudaEventRecordWithFlags(T.ev[k++], ctx->stream, cudaEventRecordExternal);
			}
			cudaError_t e2 = cudaStreamEndCapture(ctx->stream, &g);
			if (e == cudaSuccess) e = e2;
			if (e == cudaSuccess && !rc) e = cudaGraphInstantiate(&ctx->phase_graph_exec[0], g, 0);
			if (g) cudaGraphDestroy(g);
			ctx->timed_graph_launches = ctx->launches - before;
			ctx->launches = before;
			if (rc || e != cudaSuccess) {
				drop_iteration_graph(ctx); // nothing half-built may survive
				cudaGetLastError();
				if (ctx->dist_world > 1) { ctx->use_graph = false; return run_iterations(ctx, admm_iters, dump); }
				if (rc) return rc;
				ADMMB_CUDA(ctx, e);
			}
			ctx->timed_graph_iters = admm_iters;
			ctx->timed_graph_first = first;
		}
		ADMMB_CUDA(ctx, cudaGraphLaunch(ctx->phase_graph_exec[0], ctx->stream));
		T.used = first + 4 * (size_t)admm_iters;
		ctx->launches += ctx->timed_graph_launches;
		return ADMMB_OK;
	}
	for (int it = 0; it < admm_iters; ++it) {
		if (timed) next_event(ctx);
		if (dump && dump->x_it) {
			int rc = admmb_get_state(ctx, ADMMB_STATE_X, dump->x_it + (size_t)it * 3 * ctx->n);
			if (rc) return rc;
		}
		{
			int rc = launch_local_all(ctx, ctx->d_currx.p);
			if (rc) return rc;
		}
		if (dump && dump->z_it) {
			int rc = admmb_get_state(ctx, ADMMB_STATE_Z, dump->z_it + (size_t)it * ctx->n_rows);
			if (rc) return rc;
		}
		if (dump && dump->u_it) {
			int rc = admmb_get_state(ctx, ADMMB_STATE_U, dump->u_it + (size_t)it * ctx->n_rows);
			if (rc) return rc;
		}
		if (timed) next_event(ctx);
		int rc = rhs_phase(ctx);
		if (rc) return rc;
		if (timed) next_event(ctx);
		rc = solve_phase(ctx);
		if (rc) return rc;
		if (timed) next_event(ctx);
	}
	return ADMMB_OK;
}

static int collect_timing(admmb_ctx *ctx, int admm_iters, cudaEvent_t e_begin, cudaEvent_t e_end) {
	Timing &T = ctx->timing;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	float ms = 0.f;
	// events: [e_begin] then 4 per iteration, then [e_end]; e_begin is index 0
	size_t k = 1;
	for (int it = 0; it < admm_iters; ++it, k += 4) {
		cudaEventElapsedTime(&ms, T.ev[k], T.ev[k + 1]); T.ms[0] += ms;
		cudaEventElapsedTime(&ms, T.ev[k + 1], T.ev[k + 2]); T.ms[1] += ms;
		cudaEventElapsedTime(&ms, T.ev[k + 2], T.ev[k + 3]); T.ms[2] += ms;
	}
	cudaEventElapsedTime(&ms, e_begin, e_end);
	T.ms[3] += ms;
	T.iters += admm_iters;
	T.used = 0;
	return ADMMB_OK;
}

// Host copies between the caller's (pageable) buffers and the pinned staging area are the largest host-side cost of
// admmb_step at 1 M tets (4 x 4.2 MB per frame at single-thread memcpy speed): split them over a few OpenMP threads.
static void staged_copy(void *dst, const void *src, size_t bytes) {
	const size_t chunk = (size_t)1 << 18;
	if (bytes < 4 * chunk) { memcpy(dst, src, bytes); return; }
	const long nchunks = (long)((bytes + chunk - 1) / chunk);
#pragma omp parallel for schedule(static) num_threads(8)
	for (long c = 0; c < nchunks; ++c) {
		const size_t o = (size_t)c * chunk;
		memcpy((char *)dst + o, (const char *)src + o, std::min(chunk, bytes - o));
	}
}

static bool is_registered(const admmb_ctx *ctx, const void *p, size_t bytes) {
	for (const auto &r : ctx->host_regs)
		if ((const char *)p >= r.first && (const char *)p + bytes <= r.first + r.second) return true;
	return false;
}

extern "C" int admmb_register_host_buffer(admmb_ctx *ctx, void *ptr, long bytes) {
	CHECK_CTX(ctx);
	if (!ptr || bytes <= 0) ADMMB_FAIL(ctx, ADMMB_E_ARG, "register_host_buffer: null pointer / empty range");
	if (is_registered(ctx, ptr, (size_t)bytes)) return ADMMB_OK;
	ADMMB_CUDA(ctx, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
	ctx->host_regs.push_back(std::make_pair((char *)ptr, (size_t)bytes));
	return ADMMB_OK;
}

extern "C" int admmb_unregister_host_buffer(admmb_ctx *ctx, void *ptr) {
	CHECK_CTX(ctx);
	for (size_t i = 0; i < ctx->host_regs.size(); ++i)
		if (ctx->host_regs[i].first == (char *)ptr) {
			if (ctx->stream) cudaStreamSynchronize(ctx->stream);
			cudaError_t e = cudaHostUnregister(ptr);
			ctx->host_regs.erase(ctx->host_regs.begin() + i);
			ADMMB_CUDA(ctx, e);
			return ADMMB_OK;
		}
	ADMMB_FAIL(ctx, ADMMB_E_ARG, "unregister_host_buffer: %p was not registered", ptr);
}

extern "C" int admmb_upload_xv(admmb_ctx *ctx, const double *x3n, const double *v3n) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "upload_xv");
	const size_t n3 = 3 * (size_t)ctx->n;
	cudaStream_t s = ctx->stream;
	const double *src[2] = { x3n, v3n };
	double *dst[2] = { ctx->d_x.p, ctx->d_v.p };
	for (int k = 0; k < 2; ++k) {
		if (!src[k]) continue;
		const double *from = src[k];
		if (!is_registered(ctx, from, n3 * sizeof(double))) {
			staged_copy(ctx->h_pin + k * n3, from, n3 * sizeof(double));
			from = ctx->h_pin + k * n3;
		}
		ADMMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_io.p + k * n3, from, n3 * sizeof(double), cudaMemcpyHostToDevice, s));
		int rc = launch_permute_in(ctx, ctx->d_io.p + k * n3, dst[k]);
		if (rc) return rc;
	}
	return ADMMB_OK;
}

// download of x / v in two halves: enqueue (permute + device-to-host copy on the stream) and, after the stream has been
// synchronised, finish (copy out of the staging area for destinations that are not page-locked)
static int enqueue_download(admmb_ctx *ctx, double *x3n, double *v3n) {
	const size_t n3 = 3 * (size_t)ctx->n;
	cudaStream_t s = ctx->stream;
	double *out[2] = { x3n, v3n };
	const double *src[2] = { ctx->d_x.p, ctx->d_v.p };
	for (int k = 0; k < 2; ++k) {
		ctx->pending_out[k] = out[k];
		if (!out[k]) continue;
		int rc = launch_permute_out(ctx, src[k], ctx->d_io.p + k * n3);
		if (rc) return rc;
		ctx->pending_staged[k] = !is_registered(ctx, out[k], n3 * sizeof(double));
		ADMMB_CUDA(ctx, cudaMemcpyAsync(ctx->pending_staged[k] ? ctx->h_pin + k * n3 : out[k], ctx->d_io.p + k * n3, n3 * sizeof(double),
		                                cudaMemcpyDeviceToHost, s));
	}
	return ADMMB_OK;
}
static void finish_download(admmb_ctx *ctx) {
	const size_t n3 = 3 * (size_t)ctx->n;
	for (int k = 0; k < 2; ++k) {
		if (ctx->pending_out[k] && ctx->pending_staged[k]) staged_copy(ctx->pending_out[k], ctx->h_pin + k * n3, n3 * sizeof(double));
		ctx->pending_out[k] = nullptr;
		ctx->pending_staged[k] = false;
	}
}

extern "C" int admmb_download_xv(admmb_ctx *ctx, double *x3n, double *v3n) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "download_xv");
	int rc = enqueue_download(ctx, x3n, v3n);
	if (rc) return rc;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	finish_download(ctx);
	return ADMMB_OK;
}

extern "C" int admmb_download_x_f32(admmb_ctx *ctx, float *x3n) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "download_x_f32");
	if (!x3n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "download_x_f32: null output");
	const size_t n3 = 3 * (size_t)ctx->n;
	float *d_tmp = reinterpret_cast<float *>(ctx->d_io.p);
	int rc = launch_permute_out_f32(ctx, ctx->d_x.p, d_tmp);
	if (rc) return rc;
	const bool direct = is_registered(ctx, x3n, n3 * sizeof(float));
	float *h_tmp = reinterpret_cast<float *>(ctx->h_pin);
	ADMMB_CUDA(ctx, cudaMemcpyAsync(direct ? x3n : h_tmp, d_tmp, n3 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (!direct) staged_copy(x3n, h_tmp, n3 * sizeof(float));
	return ADMMB_OK;
}

extern "C" int admmb_step(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "step");
	if (admm_iters < 0 || !x3n_inout || !v3n_inout) ADMMB_FAIL(ctx, ADMMB_E_ARG, "step: bad arguments");
	const bool timed = ctx->timing.on;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	if (timed) { ctx->timing.used = 0; e0 = next_event(ctx); }
	int rc = admmb_upload_xv(ctx, x3n_inout, v3n_inout);
	if (rc) return rc;
	if ((rc = launch_frame_begin(ctx))) return rc;
	if ((rc = run_iterations(ctx, admm_iters))) return rc;
	if ((rc = launch_frame_end(ctx))) return rc;
	ctx->elapsed_s += ctx->dt;
	rc = admmb_download_xv(ctx, x3n_inout, v3n_inout);
	if (rc) return rc;
	if (timed) { e1 = next_event(ctx); if ((rc = collect_timing(ctx, admm_iters, e0, e1))) return rc; }
	return check_finite(ctx);
}

extern "C" int admmb_step_async(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout) {
	CHECK_READY(ctx);
	if (admm_iters < 0 || !x3n_inout || !v3n_inout) ADMMB_FAIL(ctx, ADMMB_E_ARG, "step_async: bad arguments");
	if (ctx->pending_out[0] || ctx->pending_out[1]) ADMMB_FAIL(ctx, ADMMB_E_STATE, "step_async: the previous asynchronous step has not been collected with admmb_sync");
	int rc = admmb_upload_xv(ctx, x3n_inout, v3n_inout);
	if (rc) return rc;
	if ((rc = launch_frame_begin(ctx))) return rc;
	// phase timing is collected by the synchronous calls only: an asynchronous step records no events (the pool would
	// otherwise grow by 4 events per iteration and call, never reset)
	const bool timing_was_on = ctx->timing.on;
	ctx->timing.on = false;
	rc = run_iterations(ctx, admm_iters);
	ctx->timing.on = timing_was_on;
	if (rc) return rc;
	if ((rc = launch_frame_end(ctx))) return rc;
	ctx->elapsed_s += ctx->dt;
	return enqueue_download(ctx, x3n_inout, v3n_inout);
}

extern "C" int admmb_step_dump(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout, double *x_it, double *z_it,
                               double *u_it) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "step_dump");
	if (admm_iters < 0 || !x3n_inout || !v3n_inout) ADMMB_FAIL(ctx, ADMMB_E_ARG, "step_dump: bad arguments");
	DumpTarget d = { x_it, z_it, u_it };
	int rc = admmb_upload_xv(ctx, x3n_inout, v3n_inout);
	if (rc) return rc;
	if ((rc = launch_frame_begin(ctx))) return rc;
	if ((rc = run_iterations(ctx, admm_iters, &d))) return rc;
	if ((rc = launch_frame_end(ctx))) return rc;
	ctx->elapsed_s += ctx->dt;
	return admmb_download_xv(ctx, x3n_inout, v3n_inout);
}

// Diagnostics for "teacher-forced" parity tests: run exactly one half of an ADMM iteration on given inputs.
extern "C" int admmb_debug_local_step(admmb_ctx *ctx, const double *x3n) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "debug_local_step");
	if (!x3n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null x");
	const size_t n3 = 3 * (size_t)ctx->n;
	memcpy(ctx->h_pin, x3n, n3 * sizeof(double));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_io.p, ctx->h_pin, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	int rc = launch_permute_in(ctx, ctx->d_io.p, ctx->d_currx.p);
	if (rc) return rc;
	if ((rc = launch_local_all(ctx, ctx->d_currx.p))) return rc;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ADMMB_OK;
}

extern "C" int admmb_debug_global_step(admmb_ctx *ctx, const double *xbar3n) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "debug_global_step");
	if (!xbar3n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null x_bar");
	const size_t n3 = 3 * (size_t)ctx->n;
	// M x_bar on the host exactly as System.cpp:47 (m_masses.asDiagonal() * x_bar), then permuted in
	for (int i = 0; i < ctx->n; ++i)
		for (int j = 0; j < 3; ++j) ctx->h_pin[3 * (size_t)i + j] = ctx->h_m[i] * xbar3n[3 * (size_t)i + j];
	ADMMB_CUDA(ctx, cudaMemcpyAsync(ctx->d_io.p, ctx->h_pin, n3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	int rc = launch_permute_in(ctx, ctx->d_io.p, ctx->d_Mxbar.p);
	if (rc) return rc;
	if ((rc = rhs_phase(ctx))) return rc;
	rc = solve_phase(ctx);
	if (rc) return rc;
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ADMMB_OK;
}

extern "C" int admmb_step_resident_async(admmb_ctx *ctx, int admm_iters, int frames) {
	CHECK_READY(ctx);
	if (admm_iters < 0 || frames < 1) ADMMB_FAIL(ctx, ADMMB_E_ARG, "step_resident: bad arguments");
	const bool timed = ctx->timing.on;
	if (!ctx->ev_region[0]) { cudaEventCreate(&ctx->ev_region[0]); cudaEventCreate(&ctx->ev_region[1]); }
	ADMMB_CUDA(ctx, cudaEventRecord(ctx->ev_region[0], ctx->stream));
	for (int f = 0; f < frames; ++f) {
		cudaEvent_t e0 = nullptr, e1 = nullptr;
		if (timed) { ctx->timing.used = 0; e0 = next_event(ctx); }
		int rc;
		if ((rc = launch_frame_begin(ctx))) return rc;
		if ((rc = run_iterations(ctx, admm_iters))) return rc;
		if ((rc = launch_frame_end(ctx))) return rc;
		ctx->elapsed_s += ctx->dt;
		if (timed) { e1 = next_event(ctx); if ((rc = collect_timing(ctx, admm_iters, e0, e1))) return rc; } // (phase timing synchronises)
	}
	ADMMB_CUDA(ctx, cudaEventRecord(ctx->ev_region[1], ctx->stream));
	return ADMMB_OK;
}

extern "C" int admmb_sync(admmb_ctx *ctx) {
	CHECK_READY(ctx);
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	finish_download(ctx);
	if (ctx->ev_region[0]) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, ctx->ev_region[0], ctx->ev_region[1]) == cudaSuccess) ctx->last_region_ms = ms;
		else cudaGetLastError();
	}
	return check_finite(ctx);
}

extern "C" int admmb_step_resident(admmb_ctx *ctx, int admm_iters, int frames) {
	int rc = admmb_step_resident_async(ctx, admm_iters, frames);
	if (rc) return rc;
	return admmb_sync(ctx);
}

extern "C" int admmb_last_region_ms(admmb_ctx *ctx, double *ms) {
	CHECK_CTX(ctx);
	if (!ms) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null output");
	*ms = ctx->last_region_ms;
	return ADMMB_OK;
}

// ---- runtime changes ------------------------------------------------------------------------------------
static int check_batch(admmb_ctx *ctx, int batch) {
	if (batch < 0 || batch >= (int)ctx->batches.size()) ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown batch %d", batch);
	return ADMMB_OK;
}

static int sync_anchor_targets_from_device(admmb_ctx *ctx, Batch &b) {
	std::vector<double> soa((size_t)3 * b.nlocal);
	ADMMB_CUDA(ctx, cudaMemcpyAsync(soa.data(), b.d_aux.p, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (int k = 0; k < 3; ++k)
		for (int p = 0; p < b.nlocal; ++p) b.aux[(size_t)b.perm[p] * 3 + k] = soa[(size_t)k * b.nlocal + p];
	return ADMMB_OK;
}

extern "C" int admmb_update_anchor_targets(admmb_ctx *ctx, int batch, int first, int count, const double *pos3, const int *active) {
	CHECK_READY(ctx);
	int rc = check_batch(ctx, batch);
	if (rc) return rc;
	Batch &b = ctx->batches[batch];
	if (b.type != BT_MOVING_ANCHORS) ADMMB_FAIL(ctx, ADMMB_E_ARG, "batch %d is not a moving-anchor batch", batch);
	if (first < 0 || count < 0 || first + count > b.count) ADMMB_FAIL(ctx, ADMMB_E_ARG, "anchor range out of bounds");
	if ((rc = sync_anchor_targets_from_device(ctx, b))) return rc; // keep positions the device wrote for inactive points
	for (int i = 0; i < count; ++i) {
		if (pos3) for (int j = 0; j < 3; ++j) b.aux[(size_t)(first + i) * 3 + j] = pos3[3 * i + j];
		if (active) b.active[first + i] = active[i] ? 1 : 0;
	}
	std::vector<double> tmp;
	std::vector<int> act;
	to_soa(b.aux, 3, b.perm, tmp);
	to_soa(b.active, 1, b.perm, act);
	ADMMB_CUDA(ctx, b.d_aux.upload(tmp.data(), tmp.size(), ctx->stream));
	ADMMB_CUDA(ctx, b.d_active.upload(act.data(), act.size(), ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ADMMB_OK;
}

extern "C" int admmb_get_anchor_targets(admmb_ctx *ctx, int batch, int first, int count, double *pos3, int *active) {
	CHECK_READY(ctx);
	int rc = check_batch(ctx, batch);
	if (rc) return rc;
	Batch &b = ctx->batches[batch];
	if (b.type != BT_MOVING_ANCHORS) ADMMB_FAIL(ctx, ADMMB_E_ARG, "batch %d is not a moving-anchor batch", batch);
	if (first < 0 || count < 0 || first + count > b.count) ADMMB_FAIL(ctx, ADMMB_E_ARG, "anchor range out of bounds");
	if ((rc = sync_anchor_targets_from_device(ctx, b))) return rc;
	for (int i = 0; i < count; ++i) {
		if (pos3) for (int j = 0; j < 3; ++j) pos3[3 * i + j] = b.aux[(size_t)(first + i) * 3 + j];
		if (active) active[i] = b.active[first + i];
	}
	return ADMMB_OK;
}

extern "C" int admmb_set_batch_weights(admmb_ctx *ctx, int batch, const double *weights) {
	CHECK_CTX(ctx);
	int rc = check_batch(ctx, batch);
	if (rc) return rc;
	if (!weights) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null weights");
	Batch &b = ctx->batches[batch];
	b.w.assign(weights, weights + b.count);
	return ADMMB_OK;
}

extern "C" int admmb_get_batch_weights(admmb_ctx *ctx, int batch, double *weights) {
	CHECK_READY(ctx);
	int rc = check_batch(ctx, batch);
	if (rc) return rc;
	const Batch &b = ctx->batches[batch];
	std::copy(b.w.begin(), b.w.end(), weights);
	return ADMMB_OK;
}

extern "C" int admmb_recompute_weights(admmb_ctx *ctx) {
	CHECK_CTX(ctx);
	if (!ctx->finalized) ADMMB_FAIL(ctx, ADMMB_E_STATE, "admmb_finalize has not been called");
	CHECK_NO_PENDING(ctx, "recompute_weights");
	drop_iteration_graph(ctx); // weight arrays and the factor are re-allocated
	// Until BOTH the weight arrays and the factor are those of the new weights the context must not step: a failure in
	// between (e.g. weights that make the matrix indefinite) would otherwise leave right-hand-side weights and factor tiles
	// of different systems behind.  A later successful call repairs it.
	ctx->broken = true;
	for (Batch &b : ctx->batches) {
		int rc = upload_batch_weights(ctx, b);
		if (rc) return rc;
	}
	int rc = setup_solver(ctx);
	if (rc) return rc;
	ctx->broken = false;
	return ADMMB_OK;
}

// ---- state access -----------------------------------------------------------------------------------------
static bool is_hyper(const Batch &b) { return b.type == BT_TETS && (b.kind == ADMMB_TET_NEOHOOKEAN || b.kind == ADMMB_TET_STVK); }

extern "C" long admmb_state_size(admmb_ctx *ctx, int which) {
	if (!ctx) return ADMMB_E_ARG;
	long c = 0;
	switch (which) {
	case ADMMB_STATE_X: return 3L * ctx->n;
	case ADMMB_STATE_Z:
	case ADMMB_STATE_U: for (const Batch &b : ctx->batches) c += (long)b.rows * b.count; return c;
	case ADMMB_STATE_PROX: for (const Batch &b : ctx->batches) if (is_hyper(b)) c += 4L * b.count; return c;
	case ADMMB_STATE_PROX_ITERS:
	case ADMMB_STATE_PROX_TRIALS: for (const Batch &b : ctx->batches) if (is_hyper(b)) c += b.count; return c;
	}
	return ADMMB_E_ARG;
}

static int soa_download(admmb_ctx *ctx, const Batch &b, const double *d, int ncomp, double *out_aos) {
	// a mesh partitioned over ranks: this rank holds only the forces that touch its nodes; the entries of the others are
	// NaN here (merge the ranks' exports by taking the non-NaN entries; duplicated boundary forces are bit-identical)
	if (b.nlocal < b.count) std::fill(out_aos, out_aos + (size_t)ncomp * b.count, std::nan(""));
	std::vector<double> soa((size_t)ncomp * b.nlocal);
	ADMMB_CUDA(ctx, cudaMemcpyAsync(soa.data(), d, soa.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (int k = 0; k < ncomp; ++k)
		for (int p = 0; p < b.nlocal; ++p) out_aos[(size_t)b.perm[p] * ncomp + k] = soa[(size_t)k * b.nlocal + p];
	return ADMMB_OK;
}

static int soa_upload(admmb_ctx *ctx, const Batch &b, double *d, int ncomp, const double *in_aos) {
	std::vector<double> soa((size_t)ncomp * b.nlocal);
	for (int k = 0; k < ncomp; ++k)
		for (int p = 0; p < b.nlocal; ++p) soa[(size_t)k * b.nlocal + p] = in_aos[(size_t)b.perm[p] * ncomp + k];
	ADMMB_CUDA(ctx, cudaMemcpyAsync(d, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return ADMMB_OK;
}

extern "C" int admmb_get_state(admmb_ctx *ctx, int which, double *out) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "get_state");
	if (!out) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null output");
	int rc;
	switch (which) {
	case ADMMB_STATE_X: {
		const size_t n3 = 3 * (size_t)ctx->n;
		if ((rc = launch_permute_out(ctx, ctx->d_currx.p, ctx->d_io.p))) return rc;
		ADMMB_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_io.p, n3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		return ADMMB_OK;
	}
	case ADMMB_STATE_Z:
	case ADMMB_STATE_U:
		for (const Batch &b : ctx->batches) {
			if (b.count == 0) continue;
			if ((rc = soa_download(ctx, b, which == ADMMB_STATE_Z ? b.d_z.p : b.d_u.p, b.rows, out + b.row_base))) return rc;
		}
		return ADMMB_OK;
	case ADMMB_STATE_PROX: {
		long o = 0;
		for (const Batch &b : ctx->batches) {
			if (!is_hyper(b) || b.count == 0) continue;
			if ((rc = soa_download(ctx, b, b.d_state.p, 4, out + o))) return rc;
			o += 4L * b.count;
		}
		return ADMMB_OK;
	}
	case ADMMB_STATE_PROX_ITERS:
	case ADMMB_STATE_PROX_TRIALS: {
		long o = 0;
		for (const Batch &b : ctx->batches) {
			if (!is_hyper(b) || b.count == 0) continue;
			std::vector<int> its(b.nlocal);
			if (b.nlocal < b.count) std::fill(out + o, out + o + b.count, std::nan("")); // partitioned mesh: see soa_download
			ADMMB_CUDA(ctx, cudaMemcpyAsync(its.data(), which == ADMMB_STATE_PROX_ITERS ? b.d_its.p : b.d_trips.p, its.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
			ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			for (int p = 0; p < b.nlocal; ++p) out[o + b.perm[p]] = (double)its[p];
			o += b.count;
		}
		return ADMMB_OK;
	}
	}
	ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown state selector %d", which);
}

extern "C" int admmb_set_state(admmb_ctx *ctx, int which, const double *in) {
	CHECK_READY(ctx);
	CHECK_NO_PENDING(ctx, "set_state");
	if (!in) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null input");
	int rc;
	switch (which) {
	case ADMMB_STATE_U:
		for (const Batch &b : ctx->batches) {
			if (b.count == 0) continue;
			if ((rc = soa_upload(ctx, b, b.d_u.p, b.rows, in + b.row_base))) return rc;
		}
		return ADMMB_OK;
	case ADMMB_STATE_PROX: {
		long o = 0;
		for (const Batch &b : ctx->batches) {
			if (!is_hyper(b) || b.count == 0) continue;
			if ((rc = soa_upload(ctx, b, b.d_state.p, 4, in + o))) return rc;
			o += 4L * b.count;
		}
		return ADMMB_OK;
	}
	}
	ADMMB_FAIL(ctx, ADMMB_E_ARG, "state selector %d cannot be set (use admmb_upload_xv for x and v)", which);
}

// ---- introspection -----------------------------------------------------------------------------------------
extern "C" int admmb_get_info(admmb_ctx *ctx, admmb_info *out) {
	CHECK_CTX(ctx);
	if (!out) ADMMB_FAIL(ctx, ADMMB_E_ARG, "null output");
	memset(out, 0, sizeof(*out));
	out->n_nodes = ctx->n;
	out->n_batches = (int)ctx->batches.size();
	out->solver = ctx->solver;
	out->n_rows = ctx->n_rows;
	out->nnz_A = (long)ctx->A_idx.size();
	out->factor_seconds = ctx->factor_seconds;
	out->cg_iters_total = ctx->cg_iters_total;
	out->launches_total = ctx->launches;
	out->elapsed_s = ctx->elapsed_s;
	direct_fill_info(ctx, out);
	return ADMMB_OK;
}

extern "C" int admmb_timing_enable(admmb_ctx *ctx, int on) {
	CHECK_CTX(ctx);
	ctx->timing.on = on != 0;
	return ADMMB_OK;
}

extern "C" int admmb_timing_read(admmb_ctx *ctx, double *ms4, long *iters, int reset) {
	CHECK_CTX(ctx);
	Timing &T = ctx->timing;
	if (ms4) for (int k = 0; k < 4; ++k) ms4[k] = T.ms[k];
	if (iters) *iters = T.iters;
	if (reset) { for (int k = 0; k < 4; ++k) T.ms[k] = 0.0; T.iters = 0; }
	return ADMMB_OK;
}
