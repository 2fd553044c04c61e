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
// transposed copy serves the backward pass, so both passes read memory linearly); one thread owns one output
// row of a tile, reads its row of the tile with unit stride across the warp and multiplies it into the three
// right-hand sides held in shared memory.  The kernel is bandwidth bound: 8 B of factor per 3 FMAs.
// Node vectors are in elimination order already (the device node order IS the nested-dissection order), so
// there is no permutation step around the solve.
#include <algorithm>
#include <chrono>

#include "common.h"
#include "direct_factor.h"

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
};

__global__ void __launch_bounds__(TILE_R) k_solve_level(const SolveTile *__restrict__ tiles, const double *__restrict__ data,
                                                        const int *__restrict__ pool, double *vb, double *vy, double *vx) {
	const SolveTile t = tiles[blockIdx.x];
	__shared__ double sv[TILE_C][3];
	double *vecs[3] = { vb, vy, vx };
	const double *vin = vecs[(t.flags >> TF_IN_SHIFT) & 3];
	double *vout = vecs[(t.flags >> TF_OUT_SHIFT) & 3];
	for (int c = threadIdx.x; c < t.ncols; c += TILE_R) {
		const int gi = (t.flags & TF_IN_LIST) ? pool[t.in_idx + c] : t.in_idx + c;
		sv[c][0] = vin[3 * (size_t)gi + 0];
		sv[c][1] = vin[3 * (size_t)gi + 1];
		sv[c][2] = vin[3 * (size_t)gi + 2];
	}
	__syncthreads();
	const int r = threadIdx.x;
	if (r >= t.nrows) return;
	const double *M = data + t.off + r;
	const int ld = t.nrows;
	double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 8
	for (int c = 0; c < t.ncols; ++c) {
		const double m = M[(size_t)c * ld];
		a0 += m * sv[c][0];
		a1 += m * sv[c][1];
		a2 += m * sv[c][2];
	}
	if (t.flags & TF_NEG) { a0 = -a0; a1 = -a1; a2 = -a2; }
	const int go = (t.flags & TF_OUT_LIST) ? pool[t.out_idx + r] : t.out_idx + r;
	atomicAdd(vout + 3 * (size_t)go + 0, a0);
	atomicAdd(vout + 3 * (size_t)go + 1, a1);
	atomicAdd(vout + 3 * (size_t)go + 2, a2);
}

void direct_set_blocks(admmb_ctx *ctx, const std::vector<int> &block_end) {
	if (!ctx->direct) ctx->direct = new DirectSolver();
	ctx->direct->block_end = block_end;
}

void direct_fill_info(const admmb_ctx *ctx, admmb_info *out) {
	if (!ctx->direct) return;
	const DirectSolver &S = *ctx->direct;
	out->nnz_L = S.F.nnz_L;
	out->n_supernodes = S.F.nb;
	out->n_levels = S.F.nlevels;
	out->factor_bytes = (long)(S.data_doubles * sizeof(double));
}

// Cuts supernode J's panel T_J (m x w, column-major) into tiles and appends them to the level lists.
static void pack_supernode(const SupernodalFactor &F, int J, std::vector<double> &fdata, std::vector<SolveTile> &ftiles,
                           std::vector<double> &bdata, std::vector<SolveTile> &btiles) {
	const int c0 = F.start[J], w = F.start[J + 1] - c0;
	const int r = F.rptr[J + 1] - F.rptr[J];
	const int m = w + r;
	const double *T = F.T.data() + F.toff[J];
	(void)r;
	auto align = [](std::vector<double> &v) { while (v.size() % 16) v.push_back(0.0); };
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
			align(fdata);
			SolveTile t;
			t.off = fdata.size(); t.nrows = R.hi - R.lo; t.ncols = k1 - k0; t.pad = 0;
			t.in_idx = c0 + k0;
			if (R.diag) { t.out_idx = c0 + R.lo; t.flags = (0 << TF_IN_SHIFT) | (1 << TF_OUT_SHIFT); }                       // y_J += Linv b_J
			else { t.out_idx = F.rptr[J] + (R.lo - w); t.flags = TF_OUT_LIST | TF_NEG | (0 << TF_IN_SHIFT) | (0 << TF_OUT_SHIFT); } // b_R -= G b_J
			for (int k = k0; k < k1; ++k)
				for (int i = R.lo; i < R.hi; ++i) fdata.push_back(T[i + (size_t)k * m]);
			ftiles.push_back(t);
		}
	}
	// ---- backward: rows = supernode columns (outputs x_J), cols = trapezoid rows (inputs y_J / x_R) ----
	for (int k0 = 0; k0 < w; k0 += TILE_R) {
		const int k1 = std::min(k0 + TILE_R, w);
		for (const Range &R : ranges(TILE_C)) {
			if (R.diag && R.hi <= k0) continue; // Linv^T is upper triangular: only trapezoid rows i >= column k
			align(bdata);
			SolveTile t;
			t.off = bdata.size(); t.nrows = k1 - k0; t.ncols = R.hi - R.lo; t.pad = 0;
			t.out_idx = c0 + k0;
			if (R.diag) { t.in_idx = c0 + R.lo; t.flags = (1 << TF_IN_SHIFT) | (2 << TF_OUT_SHIFT); }                        // x_J += Linv^T y_J
			else { t.in_idx = F.rptr[J] + (R.lo - w); t.flags = TF_IN_LIST | TF_NEG | (2 << TF_IN_SHIFT) | (2 << TF_OUT_SHIFT); } // x_J -= G^T x_R
			for (int i = R.lo; i < R.hi; ++i)
				for (int k = k0; k < k1; ++k) bdata.push_back(T[i + (size_t)k * m]);
			btiles.push_back(t);
		}
	}
}

int direct_setup(admmb_ctx *ctx) {
	if (!ctx->direct) ADMMB_FAIL(ctx, ADMMB_E_STATE, "no dissection blocks (finalize order)");
	DirectSolver &S = *ctx->direct;
	std::string err;
	if (supernodal_factorize(ctx->n, ctx->A_ptr.data(), ctx->A_idx.data(), ctx->A_val.data(), S.block_end, S.F, err) != 0)
		ADMMB_FAIL(ctx, ADMMB_E_NUMERIC, "%s", err.c_str());
	const SupernodalFactor &F = S.F;
	// tiles grouped by level: forward ascending, backward descending
	std::vector<std::vector<int> > by_level(F.nlevels);
	for (int J = 0; J < F.nb; ++J) by_level[F.level[J]].push_back(J);
	std::vector<double> fdata, bdata;
	std::vector<SolveTile> ftiles, btiles;
	fdata.reserve(F.T.size() + F.T.size() / 8);
	bdata.reserve(F.T.size() + F.T.size() / 8);
	std::vector<int> f_first(F.nlevels), f_count(F.nlevels), b_first(F.nlevels), b_count(F.nlevels);
	for (int lv = 0; lv < F.nlevels; ++lv) {
		f_first[lv] = (int)ftiles.size();
		b_first[lv] = (int)btiles.size();
		for (int J : by_level[lv]) pack_supernode(F, J, fdata, ftiles, bdata, btiles);
		f_count[lv] = (int)ftiles.size() - f_first[lv];
		b_count[lv] = (int)btiles.size() - b_first[lv];
	}
	// one device array: forward data then backward data; one tile array: forward tiles then backward tiles
	const size_t fsz = fdata.size();
	for (SolveTile &t : btiles) t.off += fsz;
	S.data_doubles = fdata.size() + bdata.size();
	S.n_tiles = ftiles.size() + btiles.size();
	cudaStream_t s = ctx->stream;
	ADMMB_CUDA(ctx, S.d_data.alloc(std::max<size_t>(S.data_doubles, 1)));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_data.p, fdata.data(), fdata.size() * sizeof(double), cudaMemcpyHostToDevice, s));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_data.p + fsz, bdata.data(), bdata.size() * sizeof(double), cudaMemcpyHostToDevice, s));
	std::vector<SolveTile> all(ftiles);
	all.insert(all.end(), btiles.begin(), btiles.end());
	ADMMB_CUDA(ctx, S.d_tiles.alloc(std::max<size_t>(all.size(), 1)));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_tiles.p, all.data(), all.size() * sizeof(SolveTile), cudaMemcpyHostToDevice, s));
	ADMMB_CUDA(ctx, S.d_pool.alloc(std::max<size_t>(F.rows.size(), 1)));
	ADMMB_CUDA(ctx, cudaMemcpyAsync(S.d_pool.p, F.rows.data(), F.rows.size() * sizeof(int), cudaMemcpyHostToDevice, s));
	ADMMB_CUDA(ctx, S.d_y.alloc(3 * (size_t)ctx->n));
	ADMMB_CUDA(ctx, cudaStreamSynchronize(s));
	S.fwd_first = f_first; S.fwd_count = f_count;
	S.bwd_first.resize(F.nlevels); S.bwd_count = b_count;
	for (int lv = 0; lv < F.nlevels; ++lv) S.bwd_first[lv] = (int)ftiles.size() + b_first[lv];
	// the host copy of the panels is no longer needed
	std::vector<double>().swap(S.F.T);
	return ADMMB_OK;
}

int direct_solve(admmb_ctx *ctx) {
	DirectSolver &S = *ctx->direct;
	cudaStream_t s = ctx->stream;
	const size_t bytes = 3 * (size_t)ctx->n * sizeof(double);
	ADMMB_CUDA(ctx, cudaMemsetAsync(S.d_y.p, 0, bytes, s));
	ADMMB_CUDA(ctx, cudaMemsetAsync(ctx->d_currx.p, 0, bytes, s));
	const int nl = S.F.nlevels;
	for (int lv = 0; lv < nl; ++lv) {
		if (S.fwd_count[lv] == 0) continue;
		k_solve_level<<<S.fwd_count[lv], TILE_R, 0, s>>>(S.d_tiles.p + S.fwd_first[lv], S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p);
		ctx->launches++;
	}
	for (int lv = nl - 1; lv >= 0; --lv) {
		if (S.bwd_count[lv] == 0) continue;
		k_solve_level<<<S.bwd_count[lv], TILE_R, 0, s>>>(S.d_tiles.p + S.bwd_first[lv], S.d_data.p, S.d_pool.p, ctx->d_b.p, S.d_y.p, ctx->d_currx.p);
		ctx->launches++;
	}
	ADMMB_CUDA(ctx, cudaGetLastError());
	return ADMMB_OK;
}

void direct_destroy(admmb_ctx *ctx) {
	if (!ctx->direct) return;
	DirectSolver &S = *ctx->direct;
	S.d_data.free(); S.d_tiles.free(); S.d_pool.free(); S.d_y.free();
	delete ctx->direct;
	ctx->direct = nullptr;
}

} // namespace admmb
