// ordering.cpp -- node and element orderings chosen for the device, computed once in admmb_finalize().
//
// Nodes: geometric nested dissection (recursive coordinate bisection with vertex separators) on the node
// graph of A = M + dt^2 D^T W^2 D.  The same order is (i) the fill-reducing elimination order of the
// Cholesky factor -- the reference leaves this to Eigen's AMD (SimplicialCholesky.h:245); any ordering
// gives the same x up to rounding -- (ii) the memory order of x/v/b on the device, so the triangular
// solves need no permutation pass, and (iii) a spatially coherent order (leaf sub-domains are compact
// boxes) for the local step's vertex gathers.  Nested dissection is used instead of minimum degree because
// its elimination tree is short and bushy: O(log n) levels of independent supernodes for the
// level-scheduled device solve, where AMD on the same meshes yields thousands of levels.
//
// Elements: Morton order of the rest-shape centroid, so that a thread block's elements share vertices.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>

#include "common.h"

namespace admmb {

namespace {

struct ND {
	int n;
	const double *x;
	const std::vector<int> &ap, &ai;
	int leaf;
	std::vector<int> order;      // internal -> user
	std::vector<int> block_end;  // end offset (in `order`) of every block, in elimination order
	std::vector<int> inset;      // call id a node currently belongs to
	std::vector<char> side;
	int next_id = 0;

	ND(int n_, const double *x_, const std::vector<int> &ap_, const std::vector<int> &ai_, int leaf_)
	    : n(n_), x(x_), ap(ap_), ai(ai_), leaf(leaf_), inset(n_, -1), side(n_, 0) {
		order.reserve(n_);
	}

	void emit(const std::vector<int> &S) {
		if (S.empty()) return;
		order.insert(order.end(), S.begin(), S.end());
		block_end.push_back((int)order.size());
	}

	void rec(std::vector<int> &S) {
		if ((int)S.size() <= leaf) { emit(S); return; }
		const int id = next_id++;
		double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
		for (int v : S) {
			inset[v] = id;
			for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], x[3 * v + k]); hi[k] = std::max(hi[k], x[3 * v + k]); }
		}
		int axis = 0;
		for (int k = 1; k < 3; ++k) if (hi[k] - lo[k] > hi[axis] - lo[axis]) axis = k;
		const size_t half = S.size() / 2;
		std::nth_element(S.begin(), S.begin() + half, S.end(), [&](int a, int b) {
			const double xa = x[3 * a + axis], xb = x[3 * b + axis];
			return xa < xb || (xa == xb && a < b);
		});
		for (size_t i = 0; i < S.size(); ++i) side[S[i]] = (i < half) ? 1 : 2;
		// two candidate vertex separators: boundary of the upper half, or of the lower half
		std::vector<int> sepA, sepB;
		for (size_t i = 0; i < S.size(); ++i) {
			const int v = S[i];
			const char other = (side[v] == 1) ? 2 : 1;
			bool touches = false;
			for (int p = ap[v]; p < ap[v + 1] && !touches; ++p) {
				const int u = ai[p];
				touches = (inset[u] == id && side[u] == other);
			}
			if (touches) { if (side[v] == 2) sepA.push_back(v); else sepB.push_back(v); }
		}
		const bool useA = sepA.size() <= sepB.size();
		std::vector<int> &sep = useA ? sepA : sepB;
		for (int v : sep) side[v] = 3;
		std::vector<int> L, R;
		L.reserve(half); R.reserve(S.size() - half);
		for (int v : S) { if (side[v] == 1) L.push_back(v); else if (side[v] == 2) R.push_back(v); }
		if (L.empty() || R.empty()) {
			// the cut removed a whole side (tiny or clique-like set): no further dissection
			emit(S);
			return;
		}
		std::vector<int> sepKeep(sep);
		S.clear(); S.shrink_to_fit();
		rec(L);
		rec(R);
		emit(sepKeep);
	}
};

inline uint64_t spread21(uint64_t v) {
	v &= 0x1fffffULL;
	v = (v | v << 32) & 0x1f00000000ffffULL;
	v = (v | v << 16) & 0x1f0000ff0000ffULL;
	v = (v | v << 8) & 0x100f00f00f00f00fULL;
	v = (v | v << 4) & 0x10c30c30c30c30c3ULL;
	v = (v | v << 2) & 0x1249249249249249ULL;
	return v;
}

} // namespace

// perm: internal -> user.  sep_tree: end offsets of the dissection blocks (leaves and separators) in
// elimination order; the direct solver uses them as its supernode partition.
void compute_node_order(int n, const double *x3n, const std::vector<int> &adj_ptr, const std::vector<int> &adj_idx,
                        int leaf_size, std::vector<int> &perm, std::vector<int> &sep_tree) {
	ND nd(n, x3n, adj_ptr, adj_idx, leaf_size);
	std::vector<int> all(n);
	std::iota(all.begin(), all.end(), 0);
	nd.rec(all);
	perm.swap(nd.order);
	sep_tree.swap(nd.block_end);
}

// Sort a batch's elements by the Morton code of their rest centroid (internal position -> user element).
void morton_order_elements(const admmb_ctx *ctx, Batch &b) {
	b.perm.resize(b.count);
	std::iota(b.perm.begin(), b.perm.end(), 0);
	if (b.count <= 1 || b.type == BT_COLLISION) return;
	double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
	for (int i = 0; i < ctx->n; ++i)
		for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], ctx->h_x0[3 * i + k]); hi[k] = std::max(hi[k], ctx->h_x0[3 * i + k]); }
	double ext = 0.0;
	for (int k = 0; k < 3; ++k) ext = std::max(ext, hi[k] - lo[k]);
	if (!(ext > 0.0)) ext = 1.0;
	std::vector<uint64_t> code(b.count);
	for (int e = 0; e < b.count; ++e) {
		double c[3] = { 0, 0, 0 };
		for (int j = 0; j < b.nv; ++j)
			for (int k = 0; k < 3; ++k) c[k] += ctx->h_x0[3 * b.idx[(size_t)e * b.nv + j] + k];
		uint64_t m = 0;
		for (int k = 0; k < 3; ++k) {
			double t = (c[k] / b.nv - lo[k]) / ext;
			t = std::min(std::max(t, 0.0), 1.0);
			m |= spread21((uint64_t)(t * 2097151.0)) << k;
		}
		code[e] = m;
	}
	std::stable_sort(b.perm.begin(), b.perm.end(), [&](int a, int c) { return code[a] < code[c]; });
}

} // namespace admmb
