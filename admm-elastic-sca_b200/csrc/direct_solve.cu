// direct_solve.cu -- the prefactored global solve  curr_x = (M + dt^2 D^T W^2 D)^{-1} b  on the device
// (reference: solver.solve(solver_termB), System.cpp:62, Eigen SimplicialLDLT::solve, SimplicialCholesky.h:153-177),
// on the scalar n x n matrix with the three coordinate columns solved at once.
//
// The factor comes from direct_factor.cpp in "inverse multifrontal" form: for every supernode J of the
// nested-dissection elimination tree,  T_J = [ inv(L_JJ) ; L_RJ inv(L_JJ) ].  With it both triangular solves are
// plain dense panel products with no dependency inside a supernode:
//     forward  (levels ascending):   y_J  = inv(L_JJ) b_J ,      b_R -= (L_RJ inv(L_JJ)) b_J
//     backward (levels descending):  x_J  = inv(L_JJ)^T y_J  -  (L_RJ inv(L_JJ))^T x_R
// Supernodes of one tree level are independent, so a level is ONE kernel launch over all of its tiles.
// The panels are cut into 64 x 64 tiles stored contiguously in exactly the order they are streamed (a second,
// transposed copy serves the backward pass, so both passes read memory linearly).  A CTA fetches its tile with ONE
// bulk-async copy (cp.async.bulk -> UBLKCP, completion on an mbarrier) issued before griddepcontrol.wait -- the factor
// does not depend on the vectors, so with programmatic dependent launch the next level's tiles stream in during the
// tail of the current level -- gathers its slice of the right-hand sides meanwhile, and multiplies from shared memory
// (k_solve_level_tma).  The kernel is bandwidth bound: 8 B of factor per 3 FMAs.  k_solve_level_pf is the
// register-streaming fallback, k_solve_level_det the bit-reproducible (atomic-free) variant.  Variants that lost their
// measurements were removed in round 2 (plain per-level kernel, register-prefetch PDL kernel, persistent ring kernel
// with grid barriers, ring kernel per level: DESIGN.md 3 keeps the numbers).
// A mesh partitioned over several ranks shards the solve by subtrees of the elimination tree (shard_owners).
// Node vectors are in elimination order already (the device node order IS the nested-dissection order), so
// there is no permutation step around the solve.
#include <algorithm>
#include <chrono>

#include <memory>

#include "common.h"
#include "direct_factor.h"
#include "dist_plan.h"

namespace admmb {

#define TILE_R 64
#define TILE_C 64

enum { TF_IN_LIST = 1, TF_OUT_LIST = 2, TF_NEG = 4, TF_IN_SHIFT = 4, TF_OUT_SHIFT = 6 }; // vector ids: 0 = b, 1 = y, 2 = x

struct SolveTile {
	unsigned long long off; // offset of the tile in the packed factor (doubles)
	int nrows, ncols;
	int in_idx, out_idx;    // contiguous base node, or offset into the index pool
	int flags, pad;
};

struct DirectSolver {
	std::vector<int> block_end;
	SupernodalFactor F;
	DevBuf<double> d_data;        // forward tiles then backward tiles
	DevBuf<SolveTile> d_tiles;
	DevBuf<int> d_pool;           // row structures (index lists)
	DevBuf<double> d_y;
	std::vector<int> fwd_first, fwd_count, bwd_first, bwd_count; // per level, into d_tiles
	size_t data_doubles = 0;
	size_t n_tiles = 0;
	int mode = 4;   // 4 (default) = launch per level, tile by bulk-async copy, programmatic dependent launch between levels; 3 = register-streaming fallback
	int unroll = 16; // loads in flight per thread of the per-level kernel (ADMMB_SOLVE_UNROLL = 8 | 16 | 32)
	int slots = 0;   // resident CTAs of the per-level kernel on the whole device
	int split = 0;   // ADMMB_SOLVE_SPLIT: 0 = choose per level, else log2 of the forced column split + 1
	bool no_pdl = false; // ADMMB_SOLVE_NO_PDL=1: mode 4 without programmatic dependent launch (A/B)
	// One mesh over several ranks, solve sharded by subtrees of the elimination tree (see shard_owners): this rank holds
	// the tiles of ITS subtrees and of the replicated top of the tree only.  Phases, each a list of per-level tile ranges:
	// own subtrees forward -> all-reduce of the top rows of b -> top forward -> top backward -> own subtrees backward ->
	// all-reduce of x.
	bool sharded = false;
	struct Range { int first, count; };
	std::vector<Range> sh_sub_fwd, sh_top_fwd, sh_top_bwd, sh_sub_bwd;
	DevBuf<int> d_top_rows;    // nodes (columns) of the replicated top supernodes
	DevBuf<double> d_top_buf;  // their 3 coordinates, packed for the all-reduce
	int n_top_rows = 0;
	double top_fraction = 0.0; // share of the factor that is replicated
	// bit-reproducible variant (ctx->deterministic): tiles store their partial sums, a second kernel per level adds them
	// to the vectors in a fixed order (no floating-point atomics)
	bool det = false;
	DevBuf<double> d_part;                 // [3][rows of the widest phase]
	DevBuf<int> d_red_key, d_red_ptr, d_red_slot; // per phase: targets (vector id * n + row), CSR into the slot list
	DevBuf<int> d_red_tgt, d_red_counter;  // partial row -> target (per phase, concatenated); arrivals per target
	std::vector<int> red_first, red_count, red_row_first; // per phase: range of targets, first entry of d_red_tgt
};

// fire-and-forget FP64 add at L2 (RED.E.ADD.F64): the generic atomicAdd would also emit a shared-memory CAS path
__device__ __forceinline__ void red_add(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

// Register-streaming variant (ADMMB_SOLVE_MODE=3; the fallback when the bulk-copy kernel below cannot get its shared
// memory): one thread per output row of a tile, UNROLL loads in flight per thread; the first UNROLL columns of the tile are
// requested BEFORE the right-hand-side gather.
// The factor does not depend on the vectors, so the two latency chains (descriptor -> index -> vector gather, and
// descriptor -> factor columns) overlap instead of following each other; the small tiles of the lower tree levels are
// one such chain long, so this is where the time goes there (profiles/r1c_solve.txt: 13 MB levels take 11 us).
template <int UNROLL>
__global__ void __launch_bounds__(TILE_R) k_solve_level_pf(const SolveTile *__restrict__ tiles, const double *__restrict__ data,
                                                           const int *__restrict__ pool, double *vb, double *vy, double *vx, int split_log2) {
	// split > 1: the columns of a tile are shared by `split` CTAs (levels with fewer tiles than the machine has CTA
	// slots are latency bound: a CTA walks its columns in rounds of UNROLL loads, so fewer columns = fewer rounds)
	const SolveTile t = tiles[blockIdx.x >> split_log2];
	const int part = blockIdx.x & ((1 << split_log2) - 1);
	const int per = (((t.ncols + (1 << split_log2) - 1) >> split_log2) + 7) & ~7;
	const int c0 = part * per;
	const int c1 = min(t.ncols, c0 + per);
	if (c0 >= c1) return;
	__shared__ double sv[TILE_C][3];
	const int r = threadIdx.x;
	const bool active = r < t.nrows;
	const int ld = t.nrows;
	const double *M = data + t.off + (active ? r : 0) + (size_t)c0 * ld;
	const int nc = c1 - c0;
	double m[UNROLL];
#pragma unroll
	for (int k = 0; k < UNROLL; ++k) m[k] = (active && k < nc) ? __ldcs(M + (size_t)k * ld) : 0.0;
	int go = 0;
	if (active) go = (t.flags & TF_OUT_LIST) ? pool[t.out_idx + r] : t.out_idx + r;
	const int si = (t.flags >> TF_IN_SHIFT) & 3, so = (t.flags >> TF_OUT_SHIFT) & 3;
	const double *vin = si == 0 ? vb : (si == 1 ? vy : vx);
	double *vout = so == 0 ? vb : (so == 1 ? vy : vx);
	for (int c = threadIdx.x; c < TILE_C; c += TILE_R) {
		if (c < nc) {
			const int gi = (t.flags & TF_IN_LIST) ? pool[t.in_idx + c0 + c] : t.in_idx + c0 + c;
			sv[c][0] = vin[3 * (size_t)gi + 0];
			sv[c][1] = vin[3 * (size_t)gi + 1];
			sv[c][2] = vin[3 * (size_t)gi + 2];
		} else {
			sv[c][0] = 0.0; sv[c][1] = 0.0; sv[c][2] = 0.0;
		}
	}
	__syncthreads();
	if (!active) return;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
	for (int k = 0; k < UNROLL; ++k) {
		a0 += m[k] * sv[k][0];
		a1 += m[k] * sv[k][1];
		a2 += m[k] * sv[k][2];
	}
	int c = UNROLL;
	for (; c + UNROLL <= nc; c += UNROLL) {
#pragma unroll
		for (int k = 0; k < UNROLL; ++k) m[k] = __ldcs(M + (size_t)(c + k) * ld);
#pragma unroll
		for (int k = 0; k < UNROLL; ++k) {
			a0 += m[k] * sv[c + k][0];
			a1 += m[k] * sv[c + k][1];
			a2 += m[k] * sv[c + k][2];
		}
	}
	for (; c < nc; ++c) {
		const double mm = __ldcs(M + (size_t)c * ld);
		a0 += mm * sv[c][0];
		a1 += mm * sv[c][1];
		a2 += mm * sv[c][2];
	}
	if (t.flags & TF_NEG) { a0 = -a0; a1 = -a1; a2 = -a2; }
	red_add(vout + 3 * (size_t)go + 0, a0);
	red_add(vout + 3 * (size_t)go + 1, a1);
	red_add(vout + 3 * (size_t)go + 2, a2);
}

// Bit-reproducible variant: no floating-point atomics.  Every tile stores its partial sums in its own rows of `part`
// (t.pad = first partial row of the tile within its phase); each output row of a phase has a counter of arrivals, and the
// thread that completes it -- whichever tile that happens to be -- adds the row's partial sums to the vector in ASCENDING
// partial-row order, i.e. always in the same order.  ("Last block reduces": partial stores, __threadfence, counter.)
template <int UNROLL>
__global__ void __launch_bounds__(TILE_R) k_solve_level_det(const SolveTile *__restrict__ tiles, const double *__restrict__ data,
                                                            const int *__restrict__ pool, double *vb, double *vy, double *vx,
                                                            double *part, const int *__restrict__ tgt_of_row, const int *__restrict__ tgt_key,
                                                            const int *__restrict__ tgt_ptr, const int *__restrict__ tgt_slot, int *counter, int n) {
	const SolveTile t = tiles[blockIdx.x];
	__shared__ double sv[TILE_C][3];
	const int r = threadIdx.x;
	const bool active = r < t.nrows;
	const int ld = t.nrows;
	const double *M = data + t.off + (active ? r : 0);
	double m[UNROLL];
#pragma unroll
	for (int k = 0; k < UNROLL; ++k) m[k] = (active && k < t.ncols) ? __ldcs(M + (size_t)k * ld) : 0.0;
	const int tj = active ? tgt_of_row[t.pad + r] : 0; // target (output row of this phase) this partial row belongs to
	const int si = (t.flags >> TF_IN_SHIFT) & 3;
	const double *vin = si == 0 ? vb : (si == 1 ? vy : vx);
	for (int c = threadIdx.x; c < TILE_C; c += TILE_R) {
		if (c < t.ncols) {
			const int gi = (t.flags & TF_IN_LIST) ? pool[t.in_idx + c] : t.in_idx + c;
			sv[c][0] = vin[3 * (size_t)gi + 0];
			sv[c][1] = vin[3 * (size_t)gi + 1];
			sv[c][2] = vin[3 * (size_t)gi + 2];
		} else {
			sv[c][0] = 0.0; sv[c][1] = 0.0; sv[c][2] = 0.0;
		}
	}
	__syncthreads();
	if (!active) return;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
	for (int k = 0; k < UNROLL; ++k) {
		a0 += m[k] * sv[k][0];
		a1 += m[k] * sv[k][1];
		a2 += m[k] * sv[k][2];
	}
	int c = UNROLL;
	for (; c + UNROLL <= t.ncols; c += UNROLL) {
#pragma unroll
		for (int k = 0; k < UNROLL; ++k) m[k] = __ldcs(M + (size_t)(c + k) * ld);
#pragma unroll
		for (int k = 0; k < UNROLL; ++k) {
			a0 += m[k] * sv[c + k][0];
			a1 += m[k] * sv[c + k][1];
			a2 += m[k] * sv[c + k][2];
		}
	}
	for (; c < t.ncols; ++c) {
		const double mm = __ldcs(M + (size_t)c * ld);
		a0 += mm * sv[c][0];
		a1 += mm * sv[c][1];
		a2 += mm * sv[c][2];
	}
	if (t.flags & TF_NEG) { a0 = -a0; a1 = -a1; a2 = -a2; }
	const int p0 = tgt_ptr[tj], p1 = tgt_ptr[tj + 1];
	const int kk = tgt_key[tj], vec = kk / n, row = kk - vec * n;
	double *v = (vec == 0 ? vb : (vec == 1 ? vy : vx)) + 3 * (size_t)row;
	if (p1 - p0 == 1) { // the only contribution to this row in this phase: nobody else touches it
		v[0] += a0; v[1] += a1; v[2] += a2;
		return;
	}
	double *p = part + 3 * ((size_t)t.pad + r);
	__stcg(p + 0, a0); __stcg(p + 1, a1); __stcg(p + 2, a2);
	__threadfence();
	if (atomicAdd(counter + tj, 1) != p1 - p0 - 1) return;
	__threadfence();
	counter[tj] = 0; // ready for the next solve
	double s0 = 0.0, s1 = 0.0, s2 = 0.0;
	for (int q = p0; q < p1; ++q) {
		const double *pq = part + 3 * (size_t)tgt_slot[q];
		s0 += __ldcg(pq + 0); s1 += __ldcg(pq + 1); s2 += __ldcg(pq + 2);
	}
	v[0] += s0; v[1] += s1; v[2] += s2;
}

// bulk-async copy (cp.async.bulk -> UBLKCP) and mbarrier helpers of the tile kernel below
#define PS_TILE_BYTES (TILE_R * TILE_C * 8)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
	             "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
	asm volatile(
	    "{\n"
	    ".reg .pred p;\n"
	    "MBAR_WAIT:\n"
	    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
	    "@p bra MBAR_DONE;\n"
	    "bra MBAR_WAIT;\n"
	    "MBAR_DONE:\n"
	    "}\n" ::"r"(smem_u32(bar)),
	    "r"(parity)
	    : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Launch per level, tile by bulk-async copy, levels chained by programmatic dependent launch (ADMMB_SOLVE_MODE=4).
//
// A level of the 1 M-tet cube is 1 200 ... 5 100 tiles of 32 KB: about one wave of CTAs, so a level lasts as long as ONE
// CTA's chain of dependent memory round trips (descriptor -> index -> vector gather -> four rounds of factor loads ->
// atomics, profiles/r2a_solve.txt: 11-27 us per level at 17-60 % of the DRAM bandwidth).  Here one thread of the CTA
// issues a single cp.async.bulk (UBLKCP) for the whole tile into shared memory -- one round trip, no registers held,
// completion on an mbarrier -- while the other threads gather the right-hand-side slice; and because the factor does
// not depend on the vectors, the copy is issued BEFORE griddepcontrol.wait: with programmatic dependent launch the CTAs
// of level p + 1 start as soon as every CTA of level p is resident, so their tiles stream in during the tail of level p
// instead of after it.  256 threads multiply the tile from shared memory (4 column groups x 64 rows).
// ---------------------------------------------------------------------------------------------------------
#ifndef LT_THREADS
#define LT_THREADS 256
#endif
#define LT_GROUPS (LT_THREADS / TILE_R)
#define LT_SMEM (PS_TILE_BYTES + TILE_C * 3 * 8 + 16) // 34 320 B: six CTAs per SM (the partial sums reuse the tile's storage)

__global__ void __launch_bounds__(LT_THREADS) k_solve_level_tma(const SolveTile *__restrict__ tiles, const double *__restrict__ data,
                                                                const int *__restrict__ pool, double *vb, double *vy, double *vx) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	double *tile = reinterpret_cast<double *>(smem_raw);                   // 64 x 64, column-major, ld = nrows
	double *sv = tile + TILE_R * TILE_C;                                    // TILE_C x 3
	double *part = tile;                                                    // LT_GROUPS x 64 x 3 partial sums, once the tile has been consumed
	unsigned long long *full = reinterpret_cast<unsigned long long *>(sv + TILE_C * 3);
	const int tid = threadIdx.x;
	const SolveTile t = tiles[blockIdx.x];
	if (tid == 0) {
		mbar_init(full, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		const unsigned bytes = ((unsigned)(t.nrows * t.ncols * 8) + 15u) & ~15u;
		mbar_expect_tx(full, bytes);
		bulk_g2s(tile, data + t.off, bytes, full);
	}
	// the next level may start (and fetch ITS tiles) once every CTA of this level has got this far
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
	double *vecs[3] = { vb, vy, vx };
	// index lists do not depend on the previous level
	static_assert(LT_THREADS >= 3 * TILE_C, "one thread per right-hand-side element of the tile");
	int gi = 0;
	const int c_in = tid / 3, k_in = tid - 3 * c_in;
	if (tid < 3 * t.ncols) gi = (t.flags & TF_IN_LIST) ? pool[t.in_idx + c_in] : t.in_idx + c_in;
	int go = 0;
	if (tid < 3 * t.nrows) { const int r = tid / 3; go = (t.flags & TF_OUT_LIST) ? pool[t.out_idx + r] : t.out_idx + r; }
	// the vectors do: wait until the previous level has completed and its sums are visible
	asm volatile("griddepcontrol.wait;" ::: "memory");
	if (tid < 3 * TILE_C) sv[tid] = (tid < 3 * t.ncols) ? __ldcg(vecs[(t.flags >> TF_IN_SHIFT) & 3] + 3 * (size_t)gi + k_in) : 0.0;
	__syncthreads(); // sv complete; the mbarrier initialisation is visible to every thread
	mbar_wait(full, 0u);
	{
		const int row = tid & (TILE_R - 1), grp = tid / TILE_R;
		const double *M = tile + row;
		const int ld = t.nrows;
		double a0 = 0.0, a1 = 0.0, a2 = 0.0;
		if (row < t.nrows) {
#pragma unroll 4
			for (int c = grp; c < t.ncols; c += LT_GROUPS) {
				const double m = M[c * ld];
				a0 += m * sv[3 * c + 0];
				a1 += m * sv[3 * c + 1];
				a2 += m * sv[3 * c + 2];
			}
		}
		__syncthreads(); // every thread has read its part of the tile: its storage now takes the partial sums
		part[(grp * TILE_R + row) * 3 + 0] = a0;
		part[(grp * TILE_R + row) * 3 + 1] = a1;
		part[(grp * TILE_R + row) * 3 + 2] = a2;
	}
	__syncthreads();
	if (tid < 3 * t.nrows) {
		const int r = tid / 3, k = tid - 3 * r;
		double a = part[(0 * TILE_R + r) * 3 + k];
#pragma unroll
		for (int gq = 1; gq < LT_GROUPS; ++gq) a += part[(gq * TILE_R + r) * 3 + k];
		if (t.flags & TF_NEG) a = -a;
		red_add(vecs[(t.flags >> TF_OUT_SHIFT) & 3] + 3 * (size_t)go + k, a);
	}
}

void direct_set_blocks(admmb_ctx *ctx, const std::vector<int> &block_end) {
	if (!ctx->direct) ctx->direct = new DirectSolver();
	ctx->direct->block_end = block_end;
}

// y of the direct solve, for the right-hand-side kernel to clear (false: no direct solver set up)
bool direct_vectors(admmb_ctx *ctx, double **y) {
	if (!ctx->direct || !ctx->direct->d_y.p) return false;
	*y = ctx->direct->d_y.p;
	return true;
}

void direct_fill_info(const admmb_ctx *ctx, admmb_info *out) {
	if (!ctx->direct) return;
	const DirectSolver &S = *ctx->direct;
	out->nnz_L = S.F.nnz_L;
	out->n_supernodes = S.F.nb;
	out->n_levels = S.F.nlevels;
	out->factor_bytes = (long)(S.data_doubles * sizeof(double));
	out->device_fronts = S.F.device_fronts;
}

// Cuts supernode J's panel T_J (m x w, column-major) into tiles.  Planning (descriptors + offsets, serial and cheap) is
// separated from copying (parallel over tiles): at 1 M tets the packed factor is 1.6 GB and a serial copy took 1 s.
struct TilePlan {
	SolveTile t;        // t.off relative to the start of the forward / backward array
	const double *src;  // T_J
	int m, lo, k0;      // leading dimension of T_J, first trapezoid row, first supernode column of the tile
};

static inline size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

static void plan_supernode(const SupernodalFactor &F, int J, size_t &foff, std::vector<TilePlan> &ftiles, size_t &boff,
                           std::vector<TilePlan> &btiles) {
	const int c0 = F.start[J], w = F.start[J + 1] - c0;
	const int r = F.rptr[J + 1] - F.rptr[J];
	const int m = w + r;
	const double *T = F.T.data() + F.toff[J];
	// trapezoid row ranges: the diagonal block [0,w) and the panel [w,m) are tiled separately so that no tile
	// straddles the boundary (their outputs / inputs live in different vectors)
	struct Range { int lo, hi; bool diag; };
	auto ranges = [&](int step) {
		std::vector<Range> v;
		for (int i = 0; i < w; i += step) v.push_back(Range{ i, std::min(i + step, w), true });
		for (int i = w; i < m; i += step) v.push_back(Range{ i, std::min(i + step, m), false });
		return v;
	};
	// ---- forward: rows = trapezoid rows (outputs), cols = supernode columns (inputs b_J) ----
	for (const Range &R : ranges(TILE_R)) {
		for (int k0 = 0; k0 < w; k0 += TILE_C) {
			const int k1 = std::min(k0 + TILE_C, w);
			if (R.diag && k0 >= R.hi) continue; // strictly above the diagonal: zeros
			TilePlan p;
			SolveTile &t = p.t;
			t.off = foff; t.nrows = R.hi - R.lo; t.ncols = k1 - k0; t.pad = 0;
			t.in_idx = c0 + k0;
			if (R.diag) { t.out_idx = c0 + R.lo; t.flags = (0 << TF_IN_SHIFT) | (1 << TF_OUT_SHIFT); }                       // y_J += Linv b_J
			else { t.out_idx = F.rptr[J] + (R.lo - w); t.flags = TF_OUT_LIST | TF_NEG | (0 << TF_IN_SHIFT) | (0 << TF_OUT_SHIFT); } // b_R -= G b_J
			p.src = T; p.m = m; p.lo = R.lo; p.k0 = k0;
			foff = align16(foff + (size_t)t.nrows * t.ncols);
			ftiles.push_back(p);
		}
	}
	// ---- backward: rows = supernode columns (outputs x_J), cols = trapezoid rows (inputs y_J / x_R) ----
	for (int k0 = 0; k0 < w; k0 += TILE_R) {
		const int k1 = std::min(k0 + TILE_R, w);
		for (const Range &R : ranges(TILE_C)) {
			if (R.diag && R.hi <= k0) continue; // Linv^T is upper triangular: only trapezoid rows i >= column k
			TilePlan p;
			SolveTile &t = p.t;
			t.off = boff; t.nrows = k1 - k0; t.ncols = R.hi - R.lo; t.pad = 0;
			t.out_idx = c0 + k0;
			if (R.diag) { t.in_idx = c0 + R.lo; t.flags = (1 << TF_IN_SHIFT) | (2 << TF_OUT_SHIFT); }                        // x_J += Linv^T y_J
			else { t.in_idx = F.rptr[J] + (R.lo - w); t.flags = TF_IN_LIST | TF_NEG | (2 << TF_IN_SHIFT) | (2 << TF_OUT_SHIFT); } // x_J -= G^T x_R
			p.src = T; p.m = m; p.lo = R.lo; p.k0 = k0;
			boff = align16(boff + (size_t)t.nrows * t.ncols);
			btiles.push_back(p);
		}
	}
}

// forward tile: column-major (trapezoid rows x supernode columns) block of T_J as it lies
static void fill_forward(const TilePlan &p, double *dst) {
	const SolveTile &t = p.t;
	for (int k = 0; k < t.ncols; ++k) memcpy(dst + (size_t)k * t.nrows, p.src + p.lo + (size_t)(p.k0 + k) * p.m, sizeof(double) * t.nrows);
	for (size_t i = (size_t)t.nrows * t.ncols; i < align16((size_t)t.nrows * t.ncols); ++i) dst[i] = 0.0;
}
// backward tile: the transposed block, column-major (supernode columns x trapezoid rows)
static void fill_backward(const TilePlan &p, double *dst) {
	const SolveTile &t = p.t;
	for (int i = 0; i < t.ncols; ++i)
		for (int k = 0; k < t.nrows; ++k) dst[(size_t)i * t.nrows + k] = p.src[(p.lo + i) + (size_t)(p.k0 + k) * p.m];
	for (size_t i = (size_t)t.nrows * t.ncols; i < align16((size_t)t.nrows * t.ncols); ++i) dst[i] = 0.0;
}

__global__ void k_rows_zero(int cnt, const int *__restrict__ rows, double *v) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 3 * cnt) return;
	v[3 * (size_t)rows[i / 3] + i % 3] = 0.0;
}
__global__ void k_rows_pack(int cnt, const int *__restrict__ rows, const double *__restrict__ v, double *__restrict__ buf) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 3 * cnt) return;
	buf[i] = v[3 * (size_t)rows[i / 3] + i % 3];
}
__global__ void k_rows_unpack(int cnt, const int *__restrict__ rows, const double *__restrict__ buf, double *__restrict__ v) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= 3 * cnt) return;
	v[3 * (size_t)rows[i / 3] + i % 3] = buf[i];
}

bool direct_is_sharded(const admmb_ctx *ctx) { return ctx->direct && ctx->direct->sharded; }

int direct_setup(admmb_ctx *ctx) {
	if (!ctx->direct) ADMMB_FAIL(ctx, ADMMB_E_STATE, "no dissection blocks (finalize order)");
	DirectSolver &S = *ctx->direct;
	std::string err;
	const bool verbose = getenv("ADMMB_FACTOR_VERBOSE") != nullptr;
	auto tick = [] { return std::chrono::steady_clock::now(); };
	auto since = [](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count(); };
	auto ts0 = tick();
	{
		// large fronts are factored on the device (front_gpu.cu); everything else, and everything when the device
		// backend is unavailable, on the host cores
		std::unique_ptr<FrontBackend> backend(make_device_front_backend(ctx));
		if (supernodal_factorize(ctx->n, ctx->A_ptr.data(), ctx->A_idx.data(), ctx->A_val.data(), S.block_end, S.F, err, backend.get()) != 0)
			ADMMB_FAIL(ctx, ADMMB_E_NUMERIC, "%s", err.c_str());
	}
	const SupernodalFactor &F = S.F;
	if (verbose) fprintf(stderr, "[setup] factorise %.3f s (symbolic %.3f, numeric %.3f)\n", since(ts0), F.seconds_symbolic, F.seconds_numeric);
	auto ts1 = tick();
	// tiles grouped by level: forward ascending, backward descending
	std::vector<std::vector<int> > by_level(F.nlevels);
	for (int J = 0; J < F.nb; ++J) by_level[F.level[J]].push_back(J);
	std::vector<TilePlan> fplan, bplan;
	size_t fsz = 0, bsz = 0;
	std::vector<int> f_first(F.nlevels), f_count(F.nlevels), b_first(F.nlevels), b_count(F.nlevels);
	// a mesh partitioned over ranks: shard the solve by subtrees unless the bit-reproducible solve was asked for (its
	// reduction tables assume the whole tree) or ADMMB_DIST_SOLVE=replicated
	S.sharded = ctx->dist_world > 1 && !ctx->deterministic;
	if (const char *e = getenv("ADMMB_DIST_SOLVE")) if (e[0] == 'r') S.sharded = false;
	std::vector<int> owner;
	std::vector<DirectSolver::Range> shf_sub, shf_top, shb_sub, shb_top; // per level, tile indices within fplan / bplan
	if (S.sharded) {
		owner = shard_owners(F, ctx->dist_world, &S.top_fraction);
		for (int pass = 0; pass < 2; ++pass) // own subtrees first, then the replicated top: two contiguous tile ranges per level
			for (int lv = 0; lv < F.nlevels; ++lv) {
				const int f0 = (int)fplan.size(), b0 = (int)bplan.size();
				for (int J : by_level[lv])
					if (pass == 0 ? owner[J] == ctx->dist_rank : owner[J] < 0) plan_supernode(F, J, fsz, fplan, bsz, bplan);
				const int fc = (int)fplan.size() - f0, bc = (int)bplan.size() - b0;
				if (fc) (pass == 0 ? shf_sub : shf_top).push_back(DirectSolver::Range{ f0, fc });
				if (bc) (pass == 0 ? shb_sub : shb_top).push_back(DirectSolver::Range{ b0, bc });
			}
		std::vector<int> top_rows;
		for (int J = 0; J < F.nb; ++J)
			if (owner[J] < 0) for (int c = F.start[J]; c < F.start[J + 1]; ++c) top_rows.push_back(c);
		S.n_top_rows = (int)top_rows.size();
		ADMMB_CUDA(ctx, S.d_top_rows.alloc(std::max<size_t>(top_rows.size(), 1)));
		ADMMB_CUDA(ctx, S.d_top_rows.upload(top_rows.data(), top_rows.size(), ctx->stream));
		ADMMB_CUDA(ctx, S.d_top_buf.alloc(3 * std::max<size_t>(top_rows.size(), 1)));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		// (the per-level arrays of the unsharded paths stay empty: those paths are not taken)
		std::fill(f_count.begin(), f_count.end(), 0);
		std::fill(b_count.begin(), b_count.end(), 0);
	} else
	for (int lv = 0; lv < F.nlevels; ++lv) {
		f_first[lv] = (int)fplan.size();
		b_first[lv] = (int)bplan.size();
		for (int J : by_level[lv]) plan_supernode(F, J, fsz, fplan, bsz, bplan);
		f_count[lv] = (int)fplan.size() - f_first[lv];
		b_count[lv] = (int)bplan.size() - b_first[lv];
	}
	// one array: forward data then backward data (every tile starts 128-byte aligned: bulk copies need 16 B)
	S.data_doubles = fsz + bsz;
	S.n_tiles = fplan.size() + bplan.size();
	std::unique_ptr<double[]> packed(new double[S.data_doubles + 16]);
	for (int i = 0; i < 16; ++i) packed[S.data_doubles + i] = 0.0;
#pragma omp parallel for schedule(dynamic, 64)
	for (long i = 0; i < (long)fplan.size(); ++i) fill_forward(fplan[i], packed.get() + fplan[i].t.off);
#pragma omp parallel for schedule(dynamic, 64)
	for (long i = 0; i < (long)bplan.size(); ++i) fill_backward(bplan[i], packed.get() + fsz + bplan[i].t.off);
	std::vector<SolveTile> ftiles(fplan.size()), btiles(bplan.size());
	for (size_t i = 0; i < fplan.size(); ++i) ftiles[i] = fplan[i].t;
	for (size_t i = 0; i < bplan.size(); ++i) { btiles[i] = bplan[i].t; btiles[i].off += fsz; }
	if (verbose) fprintf(stderr, "[setup] pack tiles %.3f s\n", since(ts1));
	auto ts2 = tick();
	cudaStream_t s = ctx->stream;
	ADMMB_CUDA(ctx, S.d_data.alloc(S.data_doubles + 16)); // + slack: bulk copies round their size up to 16 B
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_data.p, packed.get(), (S.data_doubles + 16) * sizeof(double), cudaMemcpyHostToDevice, s));
	std::vector<SolveTile> all(ftiles);
	all.insert(all.end(), btiles.begin(), btiles.end());
	S.det = ctx->deterministic;
	if (S.det) {
		// phases: forward levels ascending, backward levels descending (the order direct_solve walks them)
		const int nl = F.nlevels;
		std::vector<int> key, ptr(1, 0), slot;
		S.red_first.assign(2 * nl, 0);
		S.red_count.assign(2 * nl, 0);
		size_t max_rows = 1;
		std::vector<std::pair<int, int> > pairs; // (target key, partial row)
		std::vector<int> tgt_of_row;
		S.red_row_first.assign(2 * nl, 0);
		for (int ph = 0; ph < 2 * nl; ++ph) {
			const int lv = ph < nl ? ph : 2 * nl - 1 - ph;
			const int first = ph < nl ? f_first[lv] : (int)ftiles.size() + b_first[lv];
			const int cnt = ph < nl ? f_count[lv] : b_count[lv];
			pairs.clear();
			int rows = 0;
			for (int i = first; i < first + cnt; ++i) {
				SolveTile &t = all[i];
				t.pad = rows;
				const int vec = (t.flags >> TF_OUT_SHIFT) & 3;
				for (int r = 0; r < t.nrows; ++r) {
					const int go = (t.flags & TF_OUT_LIST) ? F.rows[t.out_idx + r] : t.out_idx + r;
					pairs.push_back(std::make_pair(vec * ctx->n + go, rows + r));
				}
				rows += t.nrows;
			}
			max_rows = std::max<size_t>(max_rows, rows);
			std::sort(pairs.begin(), pairs.end());
			S.red_first[ph] = (int)key.size();
			S.red_row_first[ph] = (int)tgt_of_row.size();
			tgt_of_row.resize(tgt_of_row.size() + rows);
			int *tor = tgt_of_row.data() + S.red_row_first[ph];
			for (size_t q = 0; q < pairs.size(); ++q) {
				if (q == 0 || pairs[q].first != pairs[q - 1].first) { key.push_back(pairs[q].first); ptr.push_back((int)slot.size()); }
				slot.push_back(pairs[q].second);
				ptr.back() = (int)slot.size();
				tor[pairs[q].second] = (int)key.size() - 1 - S.red_first[ph]; // target index within the phase
			}
			S.red_count[ph] = (int)key.size() - S.red_first[ph];
		}
		// ptr holds, for target j, the END of its slot range at ptr[j + 1] (ptr[0] = 0)
		ADMMB_CUDA(ctx, S.d_part.alloc(3 * max_rows));
		ADMMB_CUDA(ctx, S.d_red_key.upload(key, ctx->stream));
		ADMMB_CUDA(ctx, S.d_red_ptr.upload(ptr, ctx->stream));
		ADMMB_CUDA(ctx, S.d_red_slot.upload(slot, ctx->stream));
		ADMMB_CUDA(ctx, S.d_red_tgt.upload(tgt_of_row, ctx->stream));
		ADMMB_CUDA(ctx, S.d_red_counter.alloc(std::max<size_t>(key.size(), 1)));
		ADMMB_CUDA(ctx, S.d_red_counter.zero(ctx->stream));
		ADMMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	}
	ADMMB_CUDA(ctx, S.d_tiles.alloc(std::max<size_t>(all.size(), 1)));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_tiles.p, all.data(), all.size() * sizeof(SolveTile), cudaMemcpyHostToDevice, s));
	ADMMB_CUDA(ctx, S.d_pool.alloc(std::max<size_t>(F.rows.size(), 1)));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_pool.p, F.rows.data(), F.rows.size() * sizeof(int), cudaMemcpyHostToDevice, s));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	S.fwd_first = f_first; S.fwd_count = f_count;
	S.bwd_first.resize(F.nlevels); S.bwd_count = b_count;
	for (int lv = 0; lv < F.nlevels; ++lv) S.bwd_first[lv] = (int)ftiles.size() + b_first[lv];
	if (S.sharded) {
		// forward: levels ascending; backward: levels descending, and its tiles sit behind the forward tiles in d_tiles
		S.sh_sub_fwd = shf_sub;
		S.sh_top_fwd = shf_top;
		S.sh_top_bwd.assign(shb_top.rbegin(), shb_top.rend());
		S.sh_sub_bwd.assign(shb_sub.rbegin(), shb_sub.rend());
		for (auto &r : S.sh_top_bwd) r.first += (int)ftiles.size();
		for (auto &r : S.sh_sub_bwd) r.first += (int)ftiles.size();
		if (verbose) fprintf(stderr, "[setup] rank %d: sharded solve, %.1f %% of the factor replicated, %zu + %zu own / top forward levels, %d top rows\n",
		                     ctx->dist_rank, 100.0 * S.top_fraction, S.sh_sub_fwd.size(), S.sh_top_fwd.size(), S.n_top_rows);
	}
	// solve variant and its launch parameters
	{
		ADMMB_CUDA(ctx, S.d_y.alloc(3 * (size_t)ctx->n));
		int sms = 0;
		ADMMB_CUDA(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
		{
			const char *e = getenv("ADMMB_SOLVE_MODE");
			S.mode = 4;
			if (e && (e[0] == '3' || e[0] == '4')) S.mode = e[0] - '0';
			if (cudaFuncSetAttribute(k_solve_level_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LT_SMEM) != cudaSuccess) {
				cudaGetLastError();
				S.mode = 3;
			}
			if (S.sharded && S.mode != 4) ADMMB_FAIL(ctx, ADMMB_E_STATE, "the sharded solve of a partitioned mesh needs solve mode 4 (ADMMB_DIST_SOLVE=replicated selects the replicated solve)");
			const char *u = getenv("ADMMB_SOLVE_UNROLL");
			if (u) { const int v = atoi(u); if (v == 8 || v == 16 || v == 32) S.unroll = v; }
			int occ_pf = 0;
			if (S.unroll == 32) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pf, k_solve_level_pf<32>, TILE_R, 0);
			else if (S.unroll == 16) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pf, k_solve_level_pf<16>, TILE_R, 0);
			else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pf, k_solve_level_pf<8>, TILE_R, 0);
			S.slots = sms * (occ_pf < 1 ? 1 : occ_pf);
			if (const char *np = getenv("ADMMB_SOLVE_NO_PDL")) S.no_pdl = np[0] && np[0] != '0';
			const char *sp = getenv("ADMMB_SOLVE_SPLIT");
			if (sp) { const int v = atoi(sp); S.split = (v == 1) ? 1 : (v == 2) ? 2 : (v == 4) ? 3 : 0; }
		}
		ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	}
	if (verbose) fprintf(stderr, "[setup] upload %.3f s\n", since(ts2));
	// the host copy of the panels is no longer needed
	std::vector<double>().swap(S.F.T);
	return ADMMB_OK;
}

int direct_solve(admmb_ctx *ctx) {
	DirectSolver &S = *ctx->direct;
	cudaStream_t s = ctx->stream;
	const size_t bytes = 3 * (size_t)ctx->n * sizeof(double);
	const bool zeroed = ctx->solve_vectors_zeroed;
	ctx->solve_vectors_zeroed = false;
	if (!zeroed) {
		ADMMB_CUDA(ctx, cudaMemsetAsync(S.d_y.p, 0, S.d_y.bytes(), s));
		ADMMB_CUDA(ctx, cudaMemsetAsync(ctx->d_currx.p, 0, bytes, s));
	}
	const int nl = S.F.nlevels;
	if (S.sharded) {
		// own subtrees forward -> all-reduce of the top rows of b -> top forward / backward (every rank) -> own subtrees
		// backward -> all-reduce of x (every row has exactly one non-zero contribution: the sum is exact and identical on all ranks)
		bool chain = zeroed && ctx->use_pdl && !S.no_pdl; // may the next level launch follow its predecessor programmatically?
		auto level = [&](const DirectSolver::Range &r) -> int {
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(r.count); cfg.blockDim = dim3(LT_THREADS); cfg.dynamicSmemBytes = LT_SMEM; cfg.stream = s;
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
			attr[0].val.programmaticStreamSerializationAllowed = 1;
			cfg.attrs = attr; cfg.numAttrs = chain ? 1 : 0;
			ADMMB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_solve_level_tma, (const SolveTile *)(S.d_tiles.p + r.first), (const double *)S.d_data.p,
			                                   (const int *)S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p));
			ctx->launches++;
			chain = ctx->use_pdl && !S.no_pdl;
			return ADMMB_OK;
		};
		const int nt = S.n_top_rows, g = (3 * nt + 255) / 256;
		int rc;
		if (ctx->dist_rank > 0 && nt > 0) { k_rows_zero<<<g, 256, 0, s>>>(nt, S.d_top_rows.p, ctx->d_b.p); ctx->launches++; chain = false; }
		for (const auto &r : S.sh_sub_fwd) if ((rc = level(r))) return rc;
		if (nt > 0) {
			k_rows_pack<<<g, 256, 0, s>>>(nt, S.d_top_rows.p, ctx->d_b.p, S.d_top_buf.p);
			if ((rc = dist_allreduce_sum(ctx, S.d_top_buf.p, 3 * nt))) return rc;
			k_rows_unpack<<<g, 256, 0, s>>>(nt, S.d_top_rows.p, S.d_top_buf.p, ctx->d_b.p);
			ctx->launches += 2;
			chain = false;
		}
		for (const auto &r : S.sh_top_fwd) if ((rc = level(r))) return rc;
		for (const auto &r : S.sh_top_bwd) if ((rc = level(r))) return rc;
		for (const auto &r : S.sh_sub_bwd) if ((rc = level(r))) return rc;
		if (ctx->dist_rank > 0 && nt > 0) { k_rows_zero<<<g, 256, 0, s>>>(nt, S.d_top_rows.p, ctx->d_currx.p); ctx->launches++; }
		ADMMB_CUDA(ctx, cudaGetLastError());
		return dist_allreduce_sum(ctx, ctx->d_currx.p, 3 * ctx->n);
	}
	if (S.mode == 4 && !S.det) {
		// launch per level, tile by bulk copy; every launch after the first is programmatically dependent on the previous one
		bool first = true;
		for (int ph = 0; ph < 2 * nl; ++ph) {
			const int lv = ph < nl ? ph : 2 * nl - 1 - ph;
			const int cnt = ph < nl ? S.fwd_count[lv] : S.bwd_count[lv];
			const int off = ph < nl ? S.fwd_first[lv] : S.bwd_first[lv];
			if (cnt == 0) continue;
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(cnt); cfg.blockDim = dim3(LT_THREADS); cfg.dynamicSmemBytes = LT_SMEM; cfg.stream = s;
			cudaLaunchAttribute attr[1];
			attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
			attr[0].val.programmaticStreamSerializationAllowed = 1;
			// (the first level may follow the right-hand-side kernel programmatically too when that kernel has cleared the
			// vectors itself: there is no memset node in between)
			cfg.attrs = attr; cfg.numAttrs = ((first && !(zeroed && ctx->use_pdl)) || S.no_pdl) ? 0 : 1;
			ADMMB_CUDA(ctx, cudaLaunchKernelEx(&cfg, k_solve_level_tma, (const SolveTile *)(S.d_tiles.p + off), (const double *)S.d_data.p,
			                                   (const int *)S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p));
			ctx->launches++;
			first = false;
		}
		return ADMMB_OK;
	}
	for (int ph = 0; ph < 2 * nl; ++ph) {
		const int lv = ph < nl ? ph : 2 * nl - 1 - ph;
		const int cnt = ph < nl ? S.fwd_count[lv] : S.bwd_count[lv];
		const SolveTile *tl = S.d_tiles.p + (ph < nl ? S.fwd_first[lv] : S.bwd_first[lv]);
		if (cnt == 0) continue;
		if (S.det) {
			const int *tor = S.d_red_tgt.p + S.red_row_first[ph], *tk = S.d_red_key.p + S.red_first[ph], *tp = S.d_red_ptr.p + S.red_first[ph];
			int *cn = S.d_red_counter.p + S.red_first[ph];
			if (S.unroll == 32) k_solve_level_det<32><<<cnt, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, S.d_part.p, tor, tk, tp, S.d_red_slot.p, cn, ctx->n);
			else if (S.unroll == 16) k_solve_level_det<16><<<cnt, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, S.d_part.p, tor, tk, tp, S.d_red_slot.p, cn, ctx->n);
			else k_solve_level_det<8><<<cnt, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, S.d_part.p, tor, tk, tp, S.d_red_slot.p, cn, ctx->n);
			ctx->launches++;
			continue;
		}
		{
			// register-streaming fallback (mode 3).  Column split: as long as twice the CTAs still fit the machine at once,
			// halve the columns per CTA
			int sl = 0;
			if (S.split > 0) sl = S.split - 1;
			else while (sl < 2 && (long)cnt * (2 << sl) <= (long)S.slots) ++sl;
			const int g = cnt << sl;
			if (S.unroll == 32) k_solve_level_pf<32><<<g, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, sl);
			else if (S.unroll == 16) k_solve_level_pf<16><<<g, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, sl);
			else k_solve_level_pf<8><<<g, TILE_R, 0, s>>>(tl, S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p, sl);
		}
		ctx->launches++;
	}
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

void direct_destroy(admmb_ctx *ctx) {
	if (!ctx->direct) return;
	DirectSolver &S = *ctx->direct;
	S.d_data.free(); S.d_tiles.free(); S.d_pool.free(); S.d_y.free();
	S.d_top_rows.free(); S.d_top_buf.free();
	S.d_part.free(); S.d_red_key.free(); S.d_red_ptr.free(); S.d_red_slot.free(); S.d_red_tgt.free(); S.d_red_counter.free();
	delete ctx->direct;
	ctx->direct = nullptr;
}

} // namespace admmb
