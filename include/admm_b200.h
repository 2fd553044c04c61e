/* admm_b200.h -- C ABI of libadmm_b200.so, the B200 (sm_100a) ADMM-PD elastic solver hot path.
 *
 * This is the drop-in boundary for the solver path of mattoverby/admm-elastic-sca: everything that
 * admm::System::step() does per frame (reference A/src/system/System.cpp:26-75, A/ = deps/admm-elastic-sca)
 * runs on the device behind these calls; scene loading, Force construction and explicit user callbacks stay
 * on the host (see INTEGRATION.md for the C++ binding that replaces System.cpp).
 *
 * Conventions
 *  - every function returns 0 on success or a negative ADMMB_E_* code; admmb_last_error() gives the text.
 *    No exception crosses this boundary.
 *  - the caller owns every host buffer passed in; it is copied before the call returns.
 *  - a context owns all device memory, is bound to one CUDA device and one stream, and is not thread-safe;
 *    distinct contexts may be driven from distinct host threads (scene ensembles, one context per GPU).
 *  - node vectors are "scaled x3" exactly like System::m_x / m_v / m_masses (System.hpp:47-49): length 3n,
 *    xyz interleaved.
 *  - there is no CPU fallback: without a CUDA device admmb_create() fails with ADMMB_E_CUDA.
 */
#ifndef ADMM_B200_H
#define ADMM_B200_H 1

#ifdef __cplusplus
extern "C" {
#endif

typedef struct admmb_ctx admmb_ctx;

enum {
	ADMMB_OK = 0,
	ADMMB_E_ARG = -1,      /* bad argument / out-of-range index */
	ADMMB_E_STATE = -2,    /* call not valid in this state (e.g. add_* after finalize) */
	ADMMB_E_CUDA = -3,     /* CUDA runtime error, message has the details */
	ADMMB_E_NUMERIC = -4,  /* matrix not positive definite / masses differ per coordinate */
	ADMMB_E_NOMEM = -5
};

/* Force kinds (names follow the reference classes). */
enum {
	ADMMB_TET_LINEAR_STRAIN = 0, /* LinearTetStrain  TetForce.hpp:31-47   p0 = stiffness                       */
	ADMMB_TET_NEOHOOKEAN = 1,    /* HyperElasticTet "nh"   TetForce.hpp:114-147 p0 = mu, p1 = lambda, max_iterations */
	ADMMB_TET_STVK = 2,          /* HyperElasticTet "stvk"                  p0 = mu, p1 = lambda, max_iterations */
	ADMMB_TET_VOLUME = 3         /* TetVolume        TetForce.hpp:52-66   p0 = stiffness, p1 = limit_min, p2 = limit_max */
};
enum {
	ADMMB_TRI_LIMITED_STRAIN = 0, /* LimitedTriangleStrain TriangleForce.hpp:30-45  flag = strain_limiting */
	ADMMB_TRI_AREA = 1,           /* TriArea               TriangleForce.hpp:83-90  flag = iters           */
	ADMMB_TRI_FUNG = 2            /* FungTriangle          TriangleForce.hpp:63-81  stiffness = mu         */
};
enum { ADMMB_SHAPE_SPHERE = 0, ADMMB_SHAPE_CYLINDER = 1, ADMMB_SHAPE_FLOOR = 2 }; /* A/src/collision/ headers */

enum { ADMMB_SOLVER_DIRECT = 0, ADMMB_SOLVER_PCG = 1 };

/* which = state vector selector for admmb_get_state / admmb_set_state */
enum {
	ADMMB_STATE_X = 0,     /* curr_x of the last ADMM iteration, 3n                                   */
	ADMMB_STATE_Z = 1,     /* curr_z (System.hpp:97), live rows only, force order (see below)          */
	ADMMB_STATE_U = 2,     /* curr_u (System.hpp:96), same layout                                      */
	ADMMB_STATE_PROX = 3,  /* per HyperElasticTet: last_prox_result[3] + solver init_hess (TetForce.hpp:145,
	                          cppoptlib/meta.h:33), 4 doubles each, force order                        */
	ADMMB_STATE_PROX_ITERS = 4, /* per HyperElasticTet: L-BFGS n_iters of the last project(), as doubles */
	ADMMB_STATE_PROX_TRIALS = 5 /* per HyperElasticTet: line-search trial points of the last project() (MoreThuente nfev summed
	                               over its L-BFGS iterations), as doubles: diagnostics of the local step's lane divergence */
};

/* ---- lifetime ------------------------------------------------------------------------------- */
int admmb_create(int device, admmb_ctx **out);
int admmb_destroy(admmb_ctx *ctx);
const char *admmb_last_error(const admmb_ctx *ctx); /* ctx may be NULL: error of the last failed create */
const char *admmb_version(void);

/* ---- building the system: replaces System::add_nodes + forces.push_back (System.hpp:52-58) ---- */
/* n nodes; x3n, m3n are length 3n.  May be called once. */
int admmb_set_nodes(admmb_ctx *ctx, int n, const double *x3n, const double *m3n);

/* Each add_* call appends one BATCH of forces, equivalent to pushing that many Force objects on
 * System::forces in the given order.  The batch order defines the "force order" of z/u exports.
 * Returns the batch id (>= 0) or a negative error. */
int admmb_add_tets(admmb_ctx *ctx, int kind, int count, const int *idx4, double p0, double p1, double p2,
                   int max_iterations);
int admmb_add_tris(admmb_ctx *ctx, int kind, int count, const int *idx3, double stiffness, double limit_min,
                   double limit_max, int flag);
int admmb_add_springs(admmb_ctx *ctx, int count, const int *idx2, const double *stiffness /* per spring */);
int admmb_add_bends(admmb_ctx *ctx, int count, const int *idx4, double stiffness);
/* StaticAnchor (AnchorForce.hpp:55-71): pos = x at finalize; weight <= 0 selects the default 1000. */
int admmb_add_static_anchors(admmb_ctx *ctx, int count, const int *idx, double weight);
/* MovingAnchor + ControlPoint (AnchorForce.hpp:78-106): pos3 = initial control point positions, all active. */
int admmb_add_moving_anchors(admmb_ctx *ctx, int count, const int *idx, const double *pos3, double weight);
/* CollisionForce over all nodes (CollisionForce.hpp:28-41); shapes in list order, params4 = {cx,cy,cz,radius}. */
int admmb_add_collision(admmb_ctx *ctx, int nshapes, const int *shape_kind, const double *params4, double weight);

/* Explicit forces (System::explicit_forces): applied on the device to v at the start of every step, in the order
 * they were registered (System.cpp:37-39).  Each call returns the force's id.
 * admmb_set_gravity: ExplicitForce over all nodes (ExplicitForce.cpp:29-39); id < 0 adds a new one (allowed after
 *   finalize too), id >= 0 changes the direction of ANY registered explicit force (wind included) between steps.
 * admmb_add_explicit_subset: ExplicitForce restricted to `indices` (ExplicitForce.hpp:54-55).
 * admmb_add_wind: WindForce (ExplicitForce.hpp:61-68, ExplicitForce.cpp:42-98) over ntris triangles, tris3 = 3 node ids
 *   each.  Semantics = the reference on ONE OpenMP thread: triangle t sees the velocities already updated by the
 *   earlier triangles sharing a node with it (its multi-threaded loop races on v); reproduced bit for bit on the device
 *   by walking the dependency wavefronts.  Subsets and wind must be registered before admmb_finalize. */
int admmb_set_gravity(admmb_ctx *ctx, int id, const double *dir3);
/* Skip (on = 0) or re-apply (on = 1) a registered explicit force: what erasing / re-inserting an entry of
 * System::explicit_forces between steps does in the reference (System.cpp:37-39 walks the vector every frame). */
int admmb_enable_explicit(admmb_ctx *ctx, int id, int on);
int admmb_add_explicit_subset(admmb_ctx *ctx, int count, const int *idx, const double *dir3);
int admmb_add_wind(admmb_ctx *ctx, int ntris, const int *tris3, const double *dir3);

/* Host threads (OpenMP) used by the setup path of every context in this process (ordering, factorisation, tile
 * packing); n <= 0 restores the OpenMP default.  Several ranks on one node should share the cores: cores / ranks each. */
int admmb_set_host_threads(int n);

/* Failure detection (default off: System::step() of the reference always returns true, System.cpp:74): when on,
 * admmb_step / admmb_sync return ADMMB_E_NUMERIC once a frame has produced a position that is not finite. */
int admmb_set_check_finite(admmb_ctx *ctx, int on);

/* Bit-reproducible direct solve (call before admmb_finalize; default off, or on with ADMMB_DETERMINISTIC=1 in the
 * environment).  By default the tile products of the triangular solves are accumulated with floating-point atomics, whose
 * order -- and therefore the last bits of x -- vary from run to run; with this option every tile stores its partial
 * sums and the last tile to arrive at an output row adds them in a fixed order: two runs on the same input are
 * bit-identical, like the reference's serial solve, at about +35 % solve time (+10 % per ADMM iteration at 1 M tets).  The local step, the
 * right-hand-side assembly and the explicit forces are deterministic in either mode. */
int admmb_set_deterministic(admmb_ctx *ctx, int on);

/* Solver choice and tolerances; call before admmb_finalize.  tol and max_cg_iters apply to PCG only. */
int admmb_set_solver(admmb_ctx *ctx, int solver, double tol, int max_cg_iters);

/* One mesh partitioned over several GPUs, one process per GPU.  Rank 0 creates an id with admmb_dist_unique_id (an
 * ncclUniqueId, 128 bytes) and distributes it out of band; every rank then calls admmb_dist_init before admmb_finalize
 * and afterwards makes the SAME calls with the SAME data as in the single-GPU case: setup, x and v are replicated and
 * the local step is partitioned (every rank evaluates the forces that touch its chunk of the nodes).
 *   ADMMB_SOLVER_DIRECT: the owned rows of the right-hand side are all-gathered (NCCL, 3n doubles per ADMM iteration)
 *     and every rank runs the prefactored solve redundantly -- sparse triangular solves do not shard -- then
 *     the ranks' chunks of the solution are all-gathered so that all ranks continue from bit-identical positions.
 *   ADMMB_SOLVER_PCG: the rows of the solve are partitioned too; NCCL carries the all-gather of the CG search direction /
 *     solution and the dot-product all-reduces.
 * admmb_get_state(Z / U / PROX / PROX_ITERS) on a partitioned mesh returns NaN for the forces this rank does not hold
 * (merge the ranks' exports by taking the non-NaN entries); admmb_set_state ignores them. */
int admmb_dist_unique_id(char *out128);
int admmb_dist_init(admmb_ctx *ctx, int rank, int world, const char *id128);

/* Replaces System::initialize() (System.cpp:98-156): computes rest shapes / weights from the node positions
 * given to admmb_set_nodes, builds A = M + dt^2 D^T W^2 D in its scalar n x n form, orders and factors it
 * (direct) or builds the preconditioner (PCG), uploads everything and zeroes u and v. */
int admmb_finalize(admmb_ctx *ctx, double timestep_s);

/* ---- per frame: replaces System::step() (System.cpp:26-75) ------------------------------------ */
/* x3n_inout / v3n_inout: host System::m_x / m_v, read at entry (callers may have modified them, e.g.
 * singletet.cpp:40) and written at exit.  Explicit forces registered with admmb_set_gravity / admmb_add_explicit_subset /
 * admmb_add_wind are applied on the device; any other (user-defined) ExplicitForce must have been applied to v by the
 * caller beforehand. */
int admmb_step(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout);

/* Diagnostic variant of admmb_step for parity dumps: additionally copies, for every ADMM iteration it,
 * curr_x entering the iteration to x_it[it*3n..], and z / u after its local step to z_it / u_it[it*rows..]
 * (rows = admmb_state_size(ADMMB_STATE_Z)).  Any of the three may be NULL.  Synchronises every iteration. */
int admmb_step_dump(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout, double *x_it,
                    double *z_it, double *u_it);

/* Diagnostics for teacher-forced parity tests: one half of an ADMM iteration on caller-supplied inputs.
 * local:  curr_x := x3n, then every force's project() (z, u, prox state updated; read them with admmb_get_state).
 * global: b = M x_bar + dt^2 D^T W^2 (z - u) from the current z, u, then the solve; result in ADMMB_STATE_X. */
int admmb_debug_local_step(admmb_ctx *ctx, const double *x3n);
int admmb_debug_global_step(admmb_ctx *ctx, const double *xbar3n);

/* Same step with x and v kept resident on the device (no host copies); frames >= 1 consecutive steps. */
int admmb_step_resident(admmb_ctx *ctx, int admm_iters, int frames);
/* The same without the final synchronisation: the frames are only enqueued on the context's stream, so one host thread
 * can keep several contexts (scene ensembles: one context per scene) running concurrently on one GPU.  admmb_sync
 * waits for the context's stream and fills admmb_last_region_ms. */
int admmb_step_resident_async(admmb_ctx *ctx, int admm_iters, int frames);
int admmb_sync(admmb_ctx *ctx);
/* admmb_step without the final wait: host x / v are read at the call (staged at once unless page-locked), the step is
 * enqueued, and x / v are written when admmb_sync returns -- the buffers must stay valid and untouched until then.  With
 * page-locked buffers (admmb_register_host_buffer) several contexts overlap their transfers and their compute. */
int admmb_step_async(admmb_ctx *ctx, int admm_iters, double *x3n_inout, double *v3n_inout);
int admmb_upload_xv(admmb_ctx *ctx, const double *x3n, const double *v3n);   /* either may be NULL */
int admmb_download_xv(admmb_ctx *ctx, double *x3n, double *v3n);             /* either may be NULL */

/* Render hand-off (SimContext::update, src/SimContext.cpp:176-195: m_x -> float `trimesh::point`s): current positions
 * rounded to float on the device, user node order; half the bytes of admmb_download_xv(x, NULL) and no v. */
int admmb_download_x_f32(admmb_ctx *ctx, float *x3n);

/* Optional: page-lock a caller-owned host buffer (e.g. the storage of System::m_x / m_v) so that admmb_step /
 * admmb_upload_xv / admmb_download_xv transfer it directly instead of staging it through the context's own pinned
 * area.  The buffer must stay allocated until it is unregistered (admmb_destroy unregisters what is left).  Buffers
 * that are not registered keep working -- they are staged. */
int admmb_register_host_buffer(admmb_ctx *ctx, void *ptr, long bytes);
int admmb_unregister_host_buffer(admmb_ctx *ctx, void *ptr);

/* ---- runtime changes ------------------------------------------------------------------------ */
/* ControlPoint::pos / active of moving-anchor batch `batch` (poordillo.cpp:56,95,199): count entries from `first`.
 * pos3 may be NULL (keep), active may be NULL (keep). */
int admmb_update_anchor_targets(admmb_ctx *ctx, int batch, int first, int count, const double *pos3,
                                const int *active);
/* Reads back ControlPoint::pos (inactive anchors follow the mesh, AnchorForce.cpp:82). */
int admmb_get_anchor_targets(admmb_ctx *ctx, int batch, int first, int count, double *pos3, int *active);
/* Force::weight of every force of a batch (count = batch size), then System::recompute_weights()
 * (System.cpp:159-179): refactors A. */
int admmb_set_batch_weights(admmb_ctx *ctx, int batch, const double *weights);
int admmb_get_batch_weights(admmb_ctx *ctx, int batch, double *weights);
int admmb_recompute_weights(admmb_ctx *ctx);

/* ---- state access (parity dumps, checkpoint/resume) ------------------------------------------- */
/* Number of doubles admmb_get_state(which) writes. */
long admmb_state_size(admmb_ctx *ctx, int which);
/* z/u layout: batches in the order added, forces in the order given within a batch, live rows only
 * (tet 9, triangle 6, spring 3, bend 9, anchor 3, collision 3n) -- the reference pads every tet to 36 rows
 * (TetForce.cpp:61,73,122-124); the 27 dead rows are not exported. */
int admmb_get_state(admmb_ctx *ctx, int which, double *out);
int admmb_set_state(admmb_ctx *ctx, int which, const double *in);

/* ---- introspection / measurement -------------------------------------------------------------- */
typedef struct admmb_info {
	int n_nodes, n_batches, solver;
	long n_rows;            /* live rows of D */
	long nnz_A;             /* scalar n x n system matrix, full (both triangles) */
	long nnz_L;             /* scalar Cholesky factor incl. diagonal (direct) */
	int n_supernodes, n_levels;
	long factor_bytes;      /* device bytes of the factor as stored (dense supernode panels) */
	double factor_seconds;  /* host ordering + symbolic + numeric time of the last (re)factorisation */
	long cg_iters_total;    /* PCG iterations since finalize */
	long launches_total;    /* kernel launches issued by this context since finalize */
	double elapsed_s;       /* System::elapsed_s */
	int device_fronts;      /* fronts of the last factorisation whose dense work ran on the device (0 = all on the host) */
} admmb_info;
int admmb_get_info(admmb_ctx *ctx, admmb_info *out);

/* Phase timing with CUDA events on the context's stream.  When enabled, every admmb_step* records events
 * around the local and global phases of each ADMM iteration (this adds event records but no host syncs
 * inside the loop).  ms[0]=local step, ms[1]=rhs assembly, ms[2]=solve, ms[3]=whole step incl. copies;
 * accumulated since the last reset; iters = ADMM iterations accumulated. */
int admmb_timing_enable(admmb_ctx *ctx, int on);
int admmb_timing_read(admmb_ctx *ctx, double *ms4, long *iters, int reset);
/* Device time (CUDA events on the context's stream) of the whole last admmb_step_resident call, all frames. */
int admmb_last_region_ms(admmb_ctx *ctx, double *ms);

/* ---- probes (no context) ----------------------------------------------------------------------- */
/* FP64 denominators of the local step's roofline, measured on `device` (SURVEY 8d asks for a measured FP64 peak):
 * out6[0..2] = 1e12 thread-instructions/s of independent DFMA / DADD / DMUL chains (flop/s = 2x for DFMA);
 * out6[3] = SM cycles per DFMA of ONE dependent chain (the latency a warp without instruction-level parallelism pays);
 * out6[4] = number of SMs; out6[5] = maximum SM clock in MHz.  The local step is compiled without multiply-add
 * contraction (the reference's x86 arithmetic rounds twice), so its attainable rate is the DADD / DMUL one. */
int admmb_probe_fp64(int device, double *out6);
/* Self-test of the exact fast paths for division / reciprocal / square root used by the local step (csrc/elastic_math.h)
 * against the plain operators on `samples` pseudo-random operand sets of every class (any bit pattern, moderate,
 * near 1, extreme exponents, zeros and powers of two).  counts8[0..3] = results that differ in any bit from a / y, three
 * quotients sharing one reciprocal, 1 / x, sqrt(x) (must all be 0); counts8[4..7] = how often each took its fallback. */
int admmb_debug_fastmath_selftest(int device, unsigned long long seed, long samples, unsigned long long *counts8);

#ifdef __cplusplus
}
#endif
#endif
