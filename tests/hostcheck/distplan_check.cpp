// tests/hostcheck/distplan_check.cpp -- TEST-ONLY: the product's host-side plans for ONE mesh partitioned over several ranks
// (csrc/dist_plan.cpp: subtree-to-rank mapping of the sharded direct solve, neighbour-only halo of the partitioned PCG rows),
// exercised on the CPU for any world size.
//  * dp_shard_solve: factorises like factorcheck.cpp, asks shard_owners for the mapping and then EMULATES the sharded solve
//    with plain loops -- every rank forward over its own subtrees, the rows of the replicated top summed over the ranks (the
//    all-reduce), the top forward / backward by everyone, every rank backward over its own subtrees -- checking on the way the
//    property the scheme rests on (a supernode's row structure only touches its owner's subtree or the top).
//  * dp_halo: one rank's halo plan (send / receive lists, remapped column slots) for a CSR pattern.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../admm-elastic-sca_b200/csrc/common.h"
#include "../../admm-elastic-sca_b200/csrc/direct_factor.h"
#include "../../admm-elastic-sca_b200/csrc/dist_plan.h"

using namespace admmb;

namespace {
void forward_block(const SupernodalFactor &F, int J, std::vector<double> &rb, std::vector<double> &y) {
	const int c0 = F.start[J], w = F.start[J + 1] - c0, r = F.rptr[J + 1] - F.rptr[J], m = w + r;
	const double *T = F.T.data() + F.toff[J];
	const int *R = F.rows.data() + F.rptr[J];
	for (int i = 0; i < m; ++i)
		for (int k = 0; k < 3; ++k) {
			double s = 0.0;
			for (int c = 0; c < w; ++c) s += T[i + (size_t)c * m] * rb[3 * (size_t)(c0 + c) + k];
			if (i < w) y[3 * (size_t)(c0 + i) + k] = s; else rb[3 * (size_t)R[i - w] + k] -= s;
		}
}
void backward_block(const SupernodalFactor &F, int J, const std::vector<double> &y, std::vector<double> &xs) {
	const int c0 = F.start[J], w = F.start[J + 1] - c0, r = F.rptr[J + 1] - F.rptr[J], m = w + r;
	const double *T = F.T.data() + F.toff[J];
	const int *R = F.rows.data() + F.rptr[J];
	for (int c = 0; c < w; ++c)
		for (int k = 0; k < 3; ++k) {
			double s = 0.0;
			for (int i = c; i < w; ++i) s += T[i + (size_t)c * m] * y[3 * (size_t)(c0 + i) + k];
			for (int i = w; i < m; ++i) s -= T[i + (size_t)c * m] * xs[3 * (size_t)R[i - w] + k];
			xs[3 * (size_t)(c0 + c) + k] = s;
		}
}
} // namespace

extern "C" {

// A: CSR (user order, full symmetric, sorted columns); xyz: node coordinates; b, x: n x 3 row-major.
// owner_out / parent_out: capacity `cap` supernodes.  stats: [0] supernodes, [1] top fraction of the factor, [2] heaviest
// rank's share, [3] lightest rank's share, [4] violations of the locality property (must be 0).
int dp_shard_solve(int n, const int *Ap, const int *Ai, const double *Ax, const double *xyz, int leaf, int world, const double *b,
                   double *x, int *owner_out, int *parent_out, int cap, double *stats) {
	std::vector<int> gi, gptr(n + 1, 0);
	for (int i = 0; i < n; ++i) {
		for (int p = Ap[i]; p < Ap[i + 1]; ++p) if (Ai[p] != i) gi.push_back(Ai[p]);
		gptr[i + 1] = (int)gi.size();
	}
	std::vector<int> perm, blocks;
	compute_node_order(n, xyz, gptr, gi, leaf, perm, blocks);
	std::vector<int> iperm(n);
	for (int i = 0; i < n; ++i) iperm[perm[i]] = i;
	std::vector<int> Bp(n + 1, 0), Bi;
	std::vector<double> Bx;
	for (int i = 0; i < n; ++i) {
		std::vector<std::pair<int, double> > row;
		for (int p = Ap[perm[i]]; p < Ap[perm[i] + 1]; ++p) row.push_back(std::make_pair(iperm[Ai[p]], Ax[p]));
		std::sort(row.begin(), row.end());
		for (auto &e : row) { Bi.push_back(e.first); Bx.push_back(e.second); }
		Bp[i + 1] = (int)Bi.size();
	}
	SupernodalFactor F;
	std::string err;
	if (supernodal_factorize(n, Bp.data(), Bi.data(), Bx.data(), blocks, F, err) != 0) return -1;
	if (F.nb > cap) return -2;
	double top_fraction = 0.0;
	const std::vector<int> owner = shard_owners(F, world, &top_fraction);
	for (int J = 0; J < F.nb; ++J) { owner_out[J] = owner[J]; parent_out[J] = F.parent[J]; }
	std::vector<int> col_owner(n, -1);
	std::vector<double> load(world, 0.0);
	double total = 0.0;
	for (int J = 0; J < F.nb; ++J) {
		for (int c = F.start[J]; c < F.start[J + 1]; ++c) col_owner[c] = owner[J];
		const double w = F.start[J + 1] - F.start[J], r = F.rptr[J + 1] - F.rptr[J];
		total += (w + r) * w;
		if (owner[J] >= 0) load[owner[J]] += (w + r) * w;
	}
	long violations = 0;
	for (int J = 0; J < F.nb; ++J)
		for (int q = F.rptr[J]; q < F.rptr[J + 1]; ++q) {
			const int o = col_owner[F.rows[q]];
			if (owner[J] < 0 ? (o != -1) : (o != -1 && o != owner[J])) ++violations;
		}
	std::vector<double> pb(3 * (size_t)n);
	for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) pb[3 * (size_t)i + k] = b[3 * (size_t)perm[i] + k];
	// forward, every rank on its own subtrees; the top rows start at zero and collect the rank's updates
	std::vector<double> y(3 * (size_t)n, 0.0), xs(3 * (size_t)n, 0.0), top_sum(3 * (size_t)n, 0.0);
	for (int r = 0; r < world; ++r) {
		std::vector<double> rb(3 * (size_t)n, 0.0), yr(3 * (size_t)n, 0.0);
		for (int c = 0; c < n; ++c) if (col_owner[c] == r) for (int k = 0; k < 3; ++k) rb[3 * (size_t)c + k] = pb[3 * (size_t)c + k];
		for (int J = 0; J < F.nb; ++J) if (owner[J] == r) forward_block(F, J, rb, yr);
		for (int c = 0; c < n; ++c)
			for (int k = 0; k < 3; ++k) {
				if (col_owner[c] == r) y[3 * (size_t)c + k] = yr[3 * (size_t)c + k];
				else if (col_owner[c] == -1) top_sum[3 * (size_t)c + k] += rb[3 * (size_t)c + k]; // the all-reduce
			}
	}
	// the replicated top: right-hand side = b + the summed updates; forward and backward by every rank alike
	{
		std::vector<double> rb(3 * (size_t)n, 0.0);
		for (int c = 0; c < n; ++c) if (col_owner[c] == -1) for (int k = 0; k < 3; ++k) rb[3 * (size_t)c + k] = pb[3 * (size_t)c + k] + top_sum[3 * (size_t)c + k];
		for (int J = 0; J < F.nb; ++J) if (owner[J] < 0) forward_block(F, J, rb, y);
		for (int J = F.nb - 1; J >= 0; --J) if (owner[J] < 0) backward_block(F, J, y, xs);
	}
	// backward, every rank on its own subtrees: reads only its own columns and the top's
	for (int r = 0; r < world; ++r) {
		std::vector<double> xr(3 * (size_t)n, 0.0);
		for (int c = 0; c < n; ++c) if (col_owner[c] == -1) for (int k = 0; k < 3; ++k) xr[3 * (size_t)c + k] = xs[3 * (size_t)c + k];
		for (int J = F.nb - 1; J >= 0; --J) if (owner[J] == r) backward_block(F, J, y, xr);
		for (int c = 0; c < n; ++c) if (col_owner[c] == r) for (int k = 0; k < 3; ++k) xs[3 * (size_t)c + k] = xr[3 * (size_t)c + k];
	}
	for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) x[3 * (size_t)perm[i] + k] = xs[3 * (size_t)i + k];
	stats[0] = F.nb;
	stats[1] = top_fraction;
	stats[2] = total > 0 ? *std::max_element(load.begin(), load.end()) / total : 0.0;
	stats[3] = total > 0 ? *std::min_element(load.begin(), load.end()) / total : 0.0;
	stats[4] = (double)violations;
	return 0;
}

// One rank's halo plan for the CSR pattern (Ap, Ai) with `chunk` consecutive rows per rank.  send_cnt / recv_cnt: world
// entries; send_idx / recv_idx: concatenated lists by peer (capacity cap); cols: for every entry of the rank's own rows, in
// row order, the index the SpMV reads -- the global column if owned, npad + slot otherwise.  Returns the number of entries.
long dp_halo(int n, const int *Ap, const int *Ai, int chunk, int world, int rank, int *send_cnt, int *recv_cnt, int *send_idx,
             int *recv_idx, int cap, int *cols, long cols_cap) {
	const int r0 = std::min(n, rank * chunk), r1 = std::min(n, r0 + chunk);
	HaloPlan H;
	plan_halo(Ap, Ai, chunk, world, r0, r1, H);
	if (H.send_total > cap || H.recv_total > cap) return -1;
	const int npad = chunk * world;
	for (int q = 0; q < world; ++q) { send_cnt[q] = H.send_cnt[q]; recv_cnt[q] = H.recv_cnt[q]; }
	std::copy(H.send_idx.begin(), H.send_idx.end(), send_idx);
	for (int q = 0; q < world; ++q) std::copy(H.recv[q].begin(), H.recv[q].end(), recv_idx + H.recv_off[q]);
	long e = 0;
	for (int i = r0; i < r1; ++i)
		for (int p = Ap[i]; p < Ap[i + 1]; ++p, ++e) {
			if (e >= cols_cap) return -1;
			const int j = Ai[p];
			cols[e] = (j >= r0 && j < r1) ? j : npad + H.slot_of(j, chunk);
		}
	return e;
}
}
