// common.h -- context and batch structures shared by the translation units of libadmm_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/admm_b200.h"

namespace admmb {

enum BatchType { BT_TETS = 0, BT_TRIS, BT_SPRINGS, BT_BENDS, BT_STATIC_ANCHORS, BT_MOVING_ANCHORS, BT_COLLISION };

// Raw device allocation with size bookkeeping.
template <class T>
struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	cudaError_t alloc(size_t count) {
		free();
		n = count;
		if (count == 0) return cudaSuccess;
		return cudaMalloc((void **)&p, count * sizeof(T));
	}
	void free() {
		if (p) cudaFree(p);
		p = nullptr;
		n = 0;
	}
	cudaError_t upload(const T *h, size_t count, cudaStream_t s) {
		if (count == 0) return cudaSuccess;
		return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
	}
	cudaError_t upload(const std::vector<T> &h, cudaStream_t s) {
		cudaError_t e = alloc(h.size());
		if (e != cudaSuccess) return e;
		return upload(h.data(), h.size(), s);
	}
	cudaError_t zero(cudaStream_t s) {
		if (n == 0) return cudaSuccess;
		return cudaMemsetAsync(p, 0, n * sizeof(T), s);
	}
	size_t bytes() const { return n * sizeof(T); }
};

// One batch of forces of the same class and material (see admm_b200.h add_*).
struct Batch {
	int type = 0, kind = 0;
	int count = 0;  // number of forces
	int nlocal = 0; // forces resident on this device (= count unless the mesh is partitioned over ranks); = perm.size()
	int nv = 0;     // nodes per force
	int rows = 0;   // live rows of D per force
	int nsel = 0;   // selector coefficients per force stored in S (tets 12, tris 6, others 0)
	int naux = 0;   // auxiliary doubles per force (spring: rest length; bend: alpha[4]; anchors: pos[3])
	int nstate = 0; // persistent optimiser state per force (hyperelastic tets 4, fung 1)
	double p0 = 0, p1 = 0, p2 = 0;
	int max_iterations = 0, flag = 0;
	double anchor_weight = -1.0;
	long row_base = 0;   // offset of this batch in the compact z/u export
	long slot_base = 0;  // offset (in slots) of this batch in the contribution buffer

	// host, USER element order
	std::vector<int> idx;            // count*nv, user node ids
	std::vector<double> stiffness;   // per-force stiffness (springs)
	std::vector<double> S;           // count*nsel
	std::vector<double> w;           // Force::weight
	std::vector<double> kk;          // blend constant k (stiffness*volume etc.)
	std::vector<double> aux;         // count*naux
	std::vector<int> active;         // moving anchors
	std::vector<int> perm;           // internal position -> user element
	// collision shapes
	std::vector<int> shape_kind;
	std::vector<double> shape_params;

	// device, INTERNAL element order, structure-of-arrays [component][count]
	DevBuf<int> d_idx;      // [nv][count] internal node ids
	DevBuf<double> d_S, d_w, d_wdt2, d_kk, d_aux, d_u, d_z, d_state;
	DevBuf<int> d_active, d_its, d_trips;
	DevBuf<int> d_shape_kind;
	DevBuf<double> d_shape_params;
};

// Explicit forces applied on the device at the start of every frame, in registration order (System.cpp:37-39).
struct ExplicitEntry {
	int kind = 0;                 // 0 ExplicitForce over all nodes, 1 ExplicitForce over a node subset, 2 WindForce
	bool enabled = true;          // admmb_enable_explicit
	double dir[3] = { 0, 0, 0 };
	std::vector<int> idx;         // kind 1: user node ids; kind 2: 3 user node ids per triangle, reference order
	int count = 0;                // nodes / triangles
	// device: kind 1 internal node ids; kind 2 triangles regrouped into dependency wavefronts, [3][count] internal ids
	DevBuf<int> d_idx, d_level_ptr;
	int n_levels = 0;             // kind 2: wavefronts; kind 1: distinct nodes
};

struct DirectSolver; // direct_solve.cu
struct PcgSolver;    // kernels_global.cu

struct Timing {
	bool on = false;
	std::vector<cudaEvent_t> ev; // pool
	size_t used = 0;
	std::vector<int> tag;        // phase tag per interval start
	double ms[4] = { 0, 0, 0, 0 };
	long iters = 0;
};

} // namespace admmb

struct admmb_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;
	std::string err;
	bool finalized = false;
	bool fused_local = true;    // small systems: all force batches in one launch (ADMMB_NO_FUSED_LOCAL=1 disables)
	bool use_pdl = true;        // programmatic dependent launch between the kernels of an iteration (ADMMB_NO_PDL=1 disables)
	bool solve_vectors_zeroed = false; // the right-hand-side kernel has just cleared y and curr_x for the direct solve
	bool broken = false;        // a refactorisation failed after the new weights were uploaded: factor and weights disagree

	int n = 0;
	double dt = 0.0;
	double elapsed_s = 0.0;
	std::vector<double> h_x0;   // rest positions (user order), 3n
	std::vector<double> h_m;    // per-node mass (user order), n
	std::vector<int> node_perm;  // internal -> user
	std::vector<int> node_iperm; // user -> internal
	std::vector<admmb::Batch> batches;
	std::vector<admmb::ExplicitEntry> explicit_forces;

	int solver = ADMMB_SOLVER_DIRECT;
	double cg_tol = 1e-12;
	int cg_max_iters = 5000;

	// device node data, INTERNAL node order, interleaved xyz
	admmb::DevBuf<double> d_x, d_v, d_xbar, d_Mxbar, d_currx, d_b, d_m;
	admmb::DevBuf<double> d_io;      // staging for permuted host copies (3n)
	admmb::DevBuf<int> d_node_perm;  // internal -> user
	admmb::DevBuf<double> d_P;       // contribution buffer, [slots][3]
	admmb::DevBuf<int> d_vert_ptr, d_vert_slots;
	long n_slots = 0;
	long n_rows = 0;

	// pinned host staging
	double *h_pin = nullptr;
	double *pending_out[2] = { nullptr, nullptr }; // admmb_step_async: host destinations of the enqueued x / v download
	bool pending_staged[2] = { false, false };
	std::vector<std::pair<char *, size_t>> host_regs; // caller buffers page-locked by admmb_register_host_buffer

	// scalar system matrix (host CSR, internal order, full symmetric pattern)
	std::vector<int> A_ptr, A_idx;
	std::vector<double> A_val;

	admmb::DirectSolver *direct = nullptr;
	admmb::PcgSolver *pcg = nullptr;

	long launches = 0;
	long cg_iters_total = 0;
	double factor_seconds = 0.0;
	admmb::Timing timing;

	// single mesh partitioned over ranks (PCG only): replicated setup, owned node range [own0, own1) of the internal order
	int dist_rank = 0, dist_world = 1;
	int own0 = 0, own1 = 0, chunk = 0;  // chunk = ceil(n / world); node vectors are allocated world * chunk long
	void *nccl_comm = nullptr;
	bool check_finite = false;  // admmb_set_check_finite: steps fail with ADMMB_E_NUMERIC once x is not finite
	admmb::DevBuf<int> d_bad;   // set by k_frame_end when a position is not finite
	bool deterministic = false; // admmb_set_deterministic: atomic-free (bit-reproducible) direct solve
	bool use_graph = true;
	cudaGraph_t iter_graph = nullptr;          // one captured ADMM iteration (direct solver)
	cudaGraphExec_t iter_graph_exec = nullptr;
	cudaGraphExec_t phase_graph_exec[3] = { nullptr, nullptr, nullptr }; // [0]: timed mode, one frame's iterations with event-record nodes at the phase boundaries
	int timed_graph_iters = 0;       // iterations captured in phase_graph_exec[0]
	size_t timed_graph_first = 0;    // first event of the pool its record nodes use
	long timed_graph_launches = 0;
	long iter_graph_launches = 0;
	cudaEvent_t ev_region[2] = { nullptr, nullptr }; // around the last admmb_step_resident call
	double last_region_ms = 0.0;
};

#define ADMMB_FAIL(ctx, code, ...)                                  \
	do {                                                            \
		char _buf[512];                                             \
		snprintf(_buf, sizeof(_buf), __VA_ARGS__);                  \
		(ctx)->err = _buf;                                          \
		return (code);                                              \
	} while (0)

#define ADMMB_CUDA(ctx, call)                                                                          \
	do {                                                                                               \
		cudaError_t _e = (call);                                                                       \
		if (_e != cudaSuccess)                                                                         \
			ADMMB_FAIL(ctx, ADMMB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
	} while (0)

namespace admmb {
// rest_state.cpp
int compute_rest_state(admmb_ctx *ctx, Batch &b);
// Serial WindForce loop -> dependency wavefronts: order[level_ptr[l] .. level_ptr[l+1]) are the triangles of level l (no
// two of them share a node); triangle t is placed one level after the latest earlier triangle on any of its nodes.
void wind_wavefronts(int n, int ntris, const int *tris3, std::vector<int> &level_ptr, std::vector<int> &order);
// ordering.cpp
void compute_node_order(int n, const double *x3n, const std::vector<int> &adj_ptr, const std::vector<int> &adj_idx,
                        int leaf_size, std::vector<int> &perm /*internal->user*/, std::vector<int> &sep_tree);
void morton_order_elements(const admmb_ctx *ctx, Batch &b);
// assemble.cpp
void assemble_system(admmb_ctx *ctx);
void build_node_graph(const admmb_ctx *ctx, std::vector<int> &ptr, std::vector<int> &idx);
// kernels_local.cu
int launch_local_step(admmb_ctx *ctx, Batch &b, const double *d_x, double dt2);
bool direct_vectors(admmb_ctx *ctx, double **y);           // direct_solve.cu
bool direct_is_sharded(const admmb_ctx *ctx);              // partitioned mesh: the solve itself is sharded (it ends with its own all-reduce of x)
int launch_local_all(admmb_ctx *ctx, const double *d_x);   // every batch: one fused launch for small systems
int launch_explicit(admmb_ctx *ctx, ExplicitEntry &e); // subset ExplicitForce / WindForce on d_x, d_v
int upload_explicit(admmb_ctx *ctx, ExplicitEntry &e); // after the node order is known
// kernels_global.cu
int launch_frame_begin(admmb_ctx *ctx);
int launch_frame_end(admmb_ctx *ctx);
int launch_rhs(admmb_ctx *ctx);
int launch_permute_in(admmb_ctx *ctx, const double *d_src_user, double *d_dst_internal);
int launch_permute_out(admmb_ctx *ctx, const double *d_src_internal, double *d_dst_user);
int launch_permute_out_f32(admmb_ctx *ctx, const double *d_src_internal, float *d_dst_user);
int pcg_setup(admmb_ctx *ctx);
int pcg_solve(admmb_ctx *ctx);
void pcg_destroy(admmb_ctx *ctx);
// dist.cu
int dist_allgather_nodes(admmb_ctx *ctx, double *vec);            // in place: every rank contributes its owned rows
int dist_allreduce_sum(admmb_ctx *ctx, double *dev, int count);
int dist_halo_exchange(admmb_ctx *ctx, const double *send_buf, const int *send_off, const int *send_cnt, double *recv_base, const int *recv_off, const int *recv_cnt);
int dist_allgather_host(admmb_ctx *ctx, const void *mine, void *all, size_t bytes);  // setup-time, synchronises
int dist_allreduce_host_int(admmb_ctx *ctx, int *value);                                  // setup-time sum / barrier
void dist_destroy(admmb_ctx *ctx);
// front_gpu.cu
struct FrontBackend;
FrontBackend *make_device_front_backend(admmb_ctx *ctx);
// direct_solve.cu
void direct_set_blocks(admmb_ctx *ctx, const std::vector<int> &block_end);
void direct_fill_info(const admmb_ctx *ctx, admmb_info *out);
int direct_setup(admmb_ctx *ctx);
int direct_solve(admmb_ctx *ctx);
void direct_destroy(admmb_ctx *ctx);
} // namespace admmb
