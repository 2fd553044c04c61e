// tests/hostcheck/factorcheck.cpp -- TEST-ONLY: runs the product's host setup code (nested-dissection ordering,
// csrc/ordering.cpp, and the supernodal factorisation, csrc/direct_factor.cpp) on the CPU and applies the
// resulting panels T_J = [inv(L_JJ); L_RJ inv(L_JJ)] with plain loops, so the factor can be validated against
// scipy on a box without a GPU.  The device solve (direct_solve.cu) applies the same panels tile by tile.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../admm-elastic-sca_b200/csrc/common.h"
#include "../../admm-elastic-sca_b200/csrc/direct_factor.h"

using namespace admmb;

extern "C" {

// A: CSR (user order, full symmetric, sorted columns); xyz: node coordinates; b, x: n x 3 row-major.
// info: [0]=n_supernodes [1]=n_levels [2]=nnz_L [3]=max supernode width
int fc_solve(int n, const int *Ap, const int *Ai, const double *Ax, const double *xyz, int leaf, const double *b, double *x,
             long *info, double *seconds) {
	std::vector<int> gp(Ap, Ap + n + 1), gi;
	// graph = pattern without the diagonal
	std::vector<int> gptr(n + 1, 0);
	for (int i = 0; i < n; ++i) {
		for (int p = Ap[i]; p < Ap[i + 1]; ++p) if (Ai[p] != i) gi.push_back(Ai[p]);
		gptr[i + 1] = (int)gi.size();
	}
	std::vector<int> perm, blocks;
	compute_node_order(n, xyz, gptr, gi, leaf, perm, blocks);
	std::vector<int> iperm(n);
	for (int i = 0; i < n; ++i) iperm[perm[i]] = i;
	// permuted CSR with sorted columns
	std::vector<int> Bp(n + 1, 0), Bi;
	std::vector<double> Bx;
	for (int i = 0; i < n; ++i) {
		const int u = perm[i];
		std::vector<std::pair<int, double> > row;
		for (int p = Ap[u]; p < Ap[u + 1]; ++p) row.push_back(std::make_pair(iperm[Ai[p]], Ax[p]));
		std::sort(row.begin(), row.end());
		for (auto &e : row) { Bi.push_back(e.first); Bx.push_back(e.second); }
		Bp[i + 1] = (int)Bi.size();
	}
	SupernodalFactor F;
	std::string err;
	if (supernodal_factorize(n, Bp.data(), Bi.data(), Bx.data(), blocks, F, err) != 0) return -1;
	if (seconds) { seconds[0] = F.seconds_symbolic; seconds[1] = F.seconds_numeric; }
	int maxw = 0;
	for (int J = 0; J < F.nb; ++J) maxw = std::max(maxw, F.start[J + 1] - F.start[J]);
	if (info) { info[0] = F.nb; info[1] = F.nlevels; info[2] = F.nnz_L; info[3] = maxw; }
	// forward / backward with the panels, level order = block order (children precede parents)
	std::vector<double> rb(3 * (size_t)n), y(3 * (size_t)n, 0.0), xs(3 * (size_t)n, 0.0);
	for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) rb[3 * (size_t)i + k] = b[3 * (size_t)perm[i] + k];
	for (int J = 0; J < F.nb; ++J) {
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
	for (int J = F.nb - 1; J >= 0; --J) {
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
	for (int i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) x[3 * (size_t)perm[i] + k] = xs[3 * (size_t)i + k];
	return 0;
}
}
