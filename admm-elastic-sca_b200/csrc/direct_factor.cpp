// direct_factor.cpp -- host multifrontal supernodal Cholesky on the nested-dissection partition, producing the
// inverse-multifrontal panels T_J = [inv(L_JJ); L_RJ inv(L_JJ)] consumed by the device solve.  Setup path only
// (System::initialize() / recompute_weights() in the reference, System.cpp:138-140,159-179).
//
// Parallelism: fronts of one elimination-tree level are independent (OpenMP over fronts while a level is wide);
// near the root, where a level holds a few large fronts, the dense kernels parallelise internally instead.
#include "direct_factor.h"

#include <omp.h>
#include <memory>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>

namespace admmb {

namespace {

typedef double v4d __attribute__((vector_size(32)));
inline v4d loadu(const double *p) { v4d r; std::memcpy(&r, p, 32); return r; }
inline void storeu(double *p, v4d v) { std::memcpy(p, &v, 32); }

// C(MxN) -= A(MxK) * B(KxN).  A, C column-major; B(p,c) = B[p*bps + c*bcs] (covers B and B^T operands).
void gemm_sub(int M, int N, int K, const double *A, long lda, const double *B, long bps, long bcs, double *C, long ldc) {
	if (M <= 0 || N <= 0 || K <= 0) return;
	int j = 0;
	for (; j + 4 <= N; j += 4) {
		int i = 0;
		for (; i + 8 <= M; i += 8) {
			v4d c00 = { 0, 0, 0, 0 }, c01 = c00, c02 = c00, c03 = c00, c10 = c00, c11 = c00, c12 = c00, c13 = c00;
			const double *a = A + i;
			const double *b = B + (long)j * bcs;
			for (int p = 0; p < K; ++p) {
				const v4d a0 = loadu(a), a1 = loadu(a + 4);
				const double b0 = b[0], b1 = b[bcs], b2 = b[2 * bcs], b3 = b[3 * bcs];
				c00 += a0 * b0; c10 += a1 * b0;
				c01 += a0 * b1; c11 += a1 * b1;
				c02 += a0 * b2; c12 += a1 * b2;
				c03 += a0 * b3; c13 += a1 * b3;
				a += lda;
				b += bps;
			}
			double *c = C + i + (long)j * ldc;
			storeu(c, loadu(c) - c00); storeu(c + 4, loadu(c + 4) - c10);
			c += ldc; storeu(c, loadu(c) - c01); storeu(c + 4, loadu(c + 4) - c11);
			c += ldc; storeu(c, loadu(c) - c02); storeu(c + 4, loadu(c + 4) - c12);
			c += ldc; storeu(c, loadu(c) - c03); storeu(c + 4, loadu(c + 4) - c13);
		}
		for (; i < M; ++i) {
			double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
			const double *b = B + (long)j * bcs;
			for (int p = 0; p < K; ++p) {
				const double av = A[i + (long)p * lda];
				s0 += av * b[0]; s1 += av * b[bcs]; s2 += av * b[2 * bcs]; s3 += av * b[3 * bcs];
				b += bps;
			}
			C[i + (long)j * ldc] -= s0; C[i + (long)(j + 1) * ldc] -= s1; C[i + (long)(j + 2) * ldc] -= s2; C[i + (long)(j + 3) * ldc] -= s3;
		}
	}
	for (; j < N; ++j) {
		for (int p = 0; p < K; ++p) {
			const double bv = B[(long)p * bps + (long)j * bcs];
			if (bv == 0.0) continue;
			const double *a = A + (long)p * lda;
			double *c = C + (long)j * ldc;
			for (int i = 0; i < M; ++i) c[i] -= a[i] * bv;
		}
	}
}

const int NB = 64;    // panel width
const int RB = 256;   // row block of the parallel tile decomposition

// In-place partial Cholesky of the leading w columns of the symmetric m x m front F (lower triangle used):
// on exit F11 = L11, F21 = L21 and F22 = F22 - L21 L21^T (Schur complement).  Returns false if not SPD.
bool partial_chol(int m, int w, double *F, long ld, bool par) {
	for (int k0 = 0; k0 < w; k0 += NB) {
		const int kb = std::min(NB, w - k0);
		// panel: unblocked right-looking on columns k0..k0+kb, all rows below
		for (int j = 0; j < kb; ++j) {
			double *cj = F + (long)(k0 + j) * ld;
			const double d = cj[k0 + j];
			if (!(d > 0.0)) return false;
			const double r = std::sqrt(d), inv = 1.0 / r;
			cj[k0 + j] = r;
			for (int i = k0 + j + 1; i < m; ++i) cj[i] *= inv;
			for (int c = j + 1; c < kb; ++c) {
				double *cc = F + (long)(k0 + c) * ld;
				const double f = cj[k0 + c];
				if (f == 0.0) continue;
				for (int i = k0 + c; i < m; ++i) cc[i] -= cj[i] * f;
			}
		}
		// trailing update: C(i,j) -= sum_p P(i,p) P(j,p), i >= j >= k0+kb, P = panel columns
		const int t0 = k0 + kb;
		if (t0 >= m) break;
		const double *P = F + (long)k0 * ld;
		const int ncb = (m - t0 + NB - 1) / NB;
		if (par && (long)(m - t0) * (m - t0) > 200000) {
			// flattened (column block, row block) tiles, lower part only
			std::vector<std::pair<int, int> > tiles;
			for (int cb = 0; cb < ncb; ++cb) {
				const int j0 = t0 + cb * NB;
				for (int i0 = j0; i0 < m; i0 += RB) tiles.push_back(std::make_pair(j0, i0));
			}
#pragma omp parallel for schedule(dynamic, 1)
			for (long t = 0; t < (long)tiles.size(); ++t) {
				const int j0 = tiles[t].first, i0 = tiles[t].second;
				const int nj = std::min(NB, m - j0), mi = std::min(RB, m - i0);
				gemm_sub(mi, nj, kb, P + i0, ld, P + j0, ld, 1, F + i0 + (long)j0 * ld, ld);
			}
		} else {
			for (int cb = 0; cb < ncb; ++cb) {
				const int j0 = t0 + cb * NB;
				const int nj = std::min(NB, m - j0);
				gemm_sub(m - j0, nj, kb, P + j0, ld, P + j0, ld, 1, F + j0 + (long)j0 * ld, ld);
			}
		}
	}
	return true;
}

// X = inv(L) for the lower-triangular w x w block L (ld ldl); X written to (ldx), strict upper part zeroed.
void tri_inverse(int w, const double *L, long ldl, double *X, long ldx, bool par) {
	for (int c = 0; c < w; ++c)
		for (int i = 0; i < c; ++i) X[i + (long)c * ldx] = 0.0;
	std::vector<double> Dinv((size_t)NB * NB);
	for (int i0 = 0; i0 < w; i0 += NB) {
		const int ib = std::min(NB, w - i0);
		// Dinv = inv(L_ii) (small, unblocked): solve L_ii Y = I column by column
		for (int c = 0; c < ib; ++c) {
			for (int i = 0; i < ib; ++i) Dinv[i + (size_t)c * NB] = 0.0;
			Dinv[c + (size_t)c * NB] = 1.0 / L[(i0 + c) + (long)(i0 + c) * ldl];
			for (int i = c + 1; i < ib; ++i) {
				double s = 0.0;
				for (int p = c; p < i; ++p) s += L[(i0 + i) + (long)(i0 + p) * ldl] * Dinv[p + (size_t)c * NB];
				Dinv[i + (size_t)c * NB] = -s / L[(i0 + i) + (long)(i0 + i) * ldl];
			}
		}
		// W = L[iblk, 0:i0] * X[0:i0, 0:i0]  (X lower triangular), computed into X[iblk, 0:i0] as -W, then
		// X[iblk, 0:i0] = Dinv * (-W)
		const int ncb = (i0 + NB - 1) / NB;
#pragma omp parallel for schedule(dynamic, 1) if (par && i0 > 256)
		for (int cb = 0; cb < ncb; ++cb) {
			const int c0 = cb * NB, nc = std::min(NB, i0 - c0);
			double *dst = X + i0 + (long)c0 * ldx;
			for (int c = 0; c < nc; ++c)
				for (int i = 0; i < ib; ++i) dst[i + (long)c * ldx] = 0.0;
			// rows p >= c0 of X's column block are the only nonzeros
			gemm_sub(ib, nc, i0 - c0, L + i0 + (long)c0 * ldl, ldl, X + c0 + (long)c0 * ldx, 1, ldx, dst, ldx);
			// dst now holds -W; apply Dinv (lower triangular) from the left, in place, bottom-up
			for (int c = 0; c < nc; ++c) {
				double *col = dst + (long)c * ldx;
				for (int i = ib - 1; i >= 0; --i) {
					double s = 0.0;
					for (int p = 0; p <= i; ++p) s += Dinv[i + (size_t)p * NB] * col[p];
					col[i] = s;
				}
			}
		}
		for (int c = 0; c < ib; ++c)
			for (int i = c; i < ib; ++i) X[(i0 + i) + (long)(i0 + c) * ldx] = Dinv[i + (size_t)c * NB];
	}
}

// G(r x w) = L21(r x w) * X(w x w lower).  G is written (not accumulated).
void panel_times_inverse(int r, int w, const double *L21, long ldl, const double *X, long ldx, double *G, long ldg, bool par) {
	if (r <= 0) return;
	const int ncb = (w + NB - 1) / NB, nrb = (r + RB - 1) / RB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1) if (par && (long)r * w > 100000)
	for (int cb = 0; cb < ncb; ++cb)
		for (int rb = 0; rb < nrb; ++rb) {
			const int c0 = cb * NB, nc = std::min(NB, w - c0);
			const int r0 = rb * RB, mr = std::min(RB, r - r0);
			double *dst = G + r0 + (long)c0 * ldg;
			for (int c = 0; c < nc; ++c)
				for (int i = 0; i < mr; ++i) dst[i + (long)c * ldg] = 0.0;
			gemm_sub(mr, nc, w - c0, L21 + r0 + (long)c0 * ldl, ldl, X + c0 + (long)c0 * ldx, 1, ldx, dst, ldg);
			for (int c = 0; c < nc; ++c)
				for (int i = 0; i < mr; ++i) dst[i + (long)c * ldg] = -dst[i + (long)c * ldg];
		}
}

} // namespace

int supernodal_factorize(int n, const int *Ap, const int *Ai, const double *Ax, const std::vector<int> &block_end,
                         SupernodalFactor &F, std::string &err, FrontBackend *backend) {
	auto t0 = std::chrono::steady_clock::now();
	const int nb = (int)block_end.size();
	F = SupernodalFactor();
	F.n = n; F.nb = nb;
	F.start.assign(nb + 1, 0);
	for (int b = 0; b < nb; ++b) F.start[b + 1] = block_end[b];
	if (nb == 0 || F.start[nb] != n) { err = "dissection blocks do not cover the matrix"; return -1; }
	std::vector<int> col2blk(n);
	for (int b = 0; b < nb; ++b)
		for (int c = F.start[b]; c < F.start[b + 1]; ++c) col2blk[c] = b;

	// ---- symbolic: row structure of every supernode, elimination tree, levels -------------------------------
	std::vector<std::vector<int> > st(nb), children(nb);
	F.parent.assign(nb, -1);
	F.level.assign(nb, 0);
	{
		std::vector<int> mark(n, -1);
		for (int J = 0; J < nb; ++J) {
			const int c1 = F.start[J + 1];
			std::vector<int> &s = st[J];
			for (int c = F.start[J]; c < c1; ++c)
				for (int p = Ap[c]; p < Ap[c + 1]; ++p) {
					const int i = Ai[p];
					if (i >= c1 && mark[i] != J) { mark[i] = J; s.push_back(i); }
				}
			for (int K : children[J])
				for (int i : st[K])
					if (i >= c1 && mark[i] != J) { mark[i] = J; s.push_back(i); }
			std::sort(s.begin(), s.end());
			if (!s.empty()) {
				const int P = col2blk[s[0]];
				F.parent[J] = P;
				children[P].push_back(J);
			}
			int lv = 0;
			for (int K : children[J]) lv = std::max(lv, F.level[K] + 1);
			F.level[J] = lv;
		}
	}
	F.rptr.assign(nb + 1, 0);
	F.toff.assign(nb + 1, 0);
	F.nnz_L = 0;
	for (int J = 0; J < nb; ++J) {
		const long w = F.start[J + 1] - F.start[J], r = (long)st[J].size();
		F.rptr[J + 1] = F.rptr[J] + (int)r;
		F.toff[J + 1] = F.toff[J] + (size_t)((w + r) * w);
		F.nnz_L += w * (w + 1) / 2 + r * w;
		F.nlevels = std::max(F.nlevels, F.level[J] + 1);
	}
	F.rows.resize(F.rptr[nb]);
	for (int J = 0; J < nb; ++J) std::copy(st[J].begin(), st[J].end(), F.rows.begin() + F.rptr[J]);
	F.T.assign(F.toff[nb], 0.0);
	F.seconds_symbolic = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

	// ---- numeric: multifrontal, level by level ----------------------------------------------------------------
	auto t1 = std::chrono::steady_clock::now();
	std::vector<std::vector<int> > by_level(F.nlevels);
	for (int J = 0; J < nb; ++J) by_level[F.level[J]].push_back(J);
	std::vector<std::unique_ptr<double[]> > update(nb); // Schur complements waiting for their parent (|R| x |R|, lower; not zero-filled)
	const int nthreads = omp_get_max_threads();
	std::vector<std::vector<int> > relpos_t(nthreads, std::vector<int>(n, -1));
	bool failed = false;
	int fail_col = -1;

	const bool verbose = getenv("ADMMB_FACTOR_VERBOSE") != nullptr;
	double t_phase[4] = { 0, 0, 0, 0 }; // assemble + extend-add, partial Cholesky, inverse, panel x inverse (serial levels only)
	auto now = [] { return std::chrono::steady_clock::now(); };
	auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
	auto process = [&](int J, bool par, std::vector<int> &relpos) {
		auto p0 = now();
		const int c0 = F.start[J], w = F.start[J + 1] - c0;
		const int *R = F.rows.data() + F.rptr[J];
		const int r = F.rptr[J + 1] - F.rptr[J];
		const int m = w + r;
		const long ld = m;
		// the front: on the host heap, or -- for the large fronts that go to the device backend -- in its page-locked buffer
		const bool dev = backend && par && m >= backend->min_front();
		std::unique_ptr<double[]> Fr_heap;
		double *Fr = dev ? backend->front_buffer((size_t)m * m) : nullptr;
		const bool on_dev = Fr != nullptr;
		if (!on_dev) { Fr_heap.reset(new double[(size_t)m * m]); Fr = Fr_heap.get(); } // zeroed below, in parallel for large fronts
#pragma omp parallel for schedule(static) if (par && m > 512)
		for (int c = 0; c < m; ++c) std::memset(Fr + (size_t)c * m, 0, sizeof(double) * (size_t)m);
		for (int c = 0; c < w; ++c) relpos[c0 + c] = c;
		for (int k = 0; k < r; ++k) relpos[R[k]] = w + k;
		// assemble the original entries of the supernode's columns (lower part)
		for (int c = 0; c < w; ++c) {
			const int j = c0 + c;
			for (int p = Ap[j]; p < Ap[j + 1]; ++p) {
				const int i = Ai[p];
				if (i >= j) Fr[relpos[i] + (long)c * ld] += Ax[p];
			}
		}
		// extend-add the children's Schur complements (distinct destination columns: parallel over the child's columns)
		for (int K : children[J]) {
			const int *RK = F.rows.data() + F.rptr[K];
			const int rk = F.rptr[K + 1] - F.rptr[K];
			const double *U = update[K].get();
			std::vector<int> loc(rk);
			for (int a = 0; a < rk; ++a) loc[a] = relpos[RK[a]];
#pragma omp parallel for schedule(dynamic, 16) if (par && rk > 256)
			for (int b = 0; b < rk; ++b) {
				double *dst = Fr + (long)loc[b] * ld;
				const double *src = U + (size_t)b * rk;
				for (int a = b; a < rk; ++a) dst[loc[a]] += src[a];
			}
			update[K].reset();
		}
		auto p1 = now();
		double *T = F.T.data() + F.toff[J];
		const bool want_U = r > 0 && F.parent[J] >= 0;
		if (on_dev) {
			if (want_U) update[J].reset(new double[(size_t)r * r]);
			const int rc = backend->factor_front(m, w, Fr, T, want_U ? update[J].get() : nullptr);
			if (rc == 0) {
				F.device_fronts++;
				t_phase[0] += secs(p0, p1); t_phase[1] += secs(p1, now());
				return;
			}
			if (rc > 0) {
#pragma omp critical
				{ failed = true; fail_col = c0; }
				return;
			}
			// backend failure: the assembled front is intact, continue on the host
		}
		if (!partial_chol(m, w, Fr, ld, par)) {
#pragma omp critical
			{ failed = true; fail_col = c0; }
			return;
		}
		auto p2 = now();
		tri_inverse(w, Fr, ld, T, ld, par);
		auto p3 = now();
		panel_times_inverse(r, w, Fr + w, ld, T, ld, T + w, ld, par);
		if (par) { t_phase[0] += secs(p0, p1); t_phase[1] += secs(p1, p2); t_phase[2] += secs(p2, p3); t_phase[3] += secs(p3, now()); }
		// note: panel_times_inverse leaves +L21*X in G (it negates the -product gemm_sub produced)
		if (want_U) {
			update[J].reset(new double[(size_t)r * r]);
			double *U = update[J].get();
#pragma omp parallel for schedule(static) if (par && r > 512)
			for (int b = 0; b < r; ++b)
				std::memcpy(U + (size_t)b * r + b, Fr + (w + b) + (long)(w + b) * ld, sizeof(double) * (r - b));
		}
	};

	if (backend) {
		size_t mf = 0, mp = 0;
		for (int J = 0; J < nb; ++J) {
			const size_t w = F.start[J + 1] - F.start[J], m = w + (F.rptr[J + 1] - F.rptr[J]);
			if ((int)m >= backend->min_front()) { mf = std::max(mf, m * m); mp = std::max(mp, m * w); }
		}
		if (mf) backend->reserve(mf, mp);
	}
	for (int lv = 0; lv < F.nlevels && !failed; ++lv) {
		const std::vector<int> &L = by_level[lv];
		auto tl0 = now();
		// large fronts one at a time (parallel inside the front, or on the device); the rest of the level tree-parallel
		std::vector<int> big, small;
		for (int J : L) {
			const int m = (F.start[J + 1] - F.start[J]) + (F.rptr[J + 1] - F.rptr[J]);
			((backend && m >= backend->min_front()) ? big : small).push_back(J);
		}
		for (int J : big) {
			if (failed) break;
			process(J, true, relpos_t[0]);
		}
		if ((int)small.size() >= 2 * nthreads) {
#pragma omp parallel for schedule(dynamic, 1)
			for (long t = 0; t < (long)small.size(); ++t) {
				if (failed) continue;
				process(small[t], false, relpos_t[omp_get_thread_num()]);
			}
		} else {
			for (int J : small) {
				if (failed) break;
				process(J, true, relpos_t[0]);
			}
		}
		if (verbose) {
			long wmax = 0, mmax = 0;
			for (int J : L) { wmax = std::max<long>(wmax, F.start[J + 1] - F.start[J]); mmax = std::max<long>(mmax, F.start[J + 1] - F.start[J] + F.rptr[J + 1] - F.rptr[J]); }
			fprintf(stderr, "[factor] level %2d: %6zu supernodes (max width %ld, max front %ld), %zu large + %zu %s: %.3f s\n", lv, L.size(), wmax, mmax,
			        big.size(), small.size(), (int)small.size() >= 2 * nthreads ? "tree-parallel " : "front-parallel", secs(tl0, now()));
		}
	}
	if (verbose) fprintf(stderr, "[factor] one-at-a-time fronts: assemble %.3f  cholesky (or whole device front) %.3f  inverse %.3f  panel*inverse %.3f s; %d fronts on the device\n", t_phase[0], t_phase[1], t_phase[2], t_phase[3], F.device_fronts);
	F.seconds_numeric = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
	if (failed) {
		char buf[160];
		snprintf(buf, sizeof(buf), "system matrix is not positive definite (pivot failure in the supernode starting at column %d)", fail_col);
		err = buf;
		return -1;
	}
	return 0;
}

} // namespace admmb
