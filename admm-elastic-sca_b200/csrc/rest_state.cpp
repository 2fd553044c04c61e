// rest_state.cpp -- host-side restatement of the reference's Force::initialize() bodies: rest shapes,
// volumes / areas and the (float-rounded) ADMM weights.  Runs once in admmb_finalize().  Citations are to
// /root/reference/deps/admm-elastic-sca/src/system (A/src/system).
#include <algorithm>
#include <cmath>

#include "common.h"

namespace admmb {

namespace {

struct V3 {
	double v[3];
	double &operator[](int i) { return v[i]; }
	double operator[](int i) const { return v[i]; }
};
inline V3 sub(const V3 &a, const V3 &b) { return V3{ { a[0] - b[0], a[1] - b[1], a[2] - b[2] } }; }
inline V3 cross(const V3 &a, const V3 &b) {
	return V3{ { a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0] } };
}
// fixed-size Eigen reductions add as a0 + (a1 + a2) (Eigen/src/Core/Redux.h, redux_novec_unroller)
inline double dot(const V3 &a, const V3 &b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }
inline double norm(const V3 &a) { return std::sqrt(dot(a, a)); }
inline V3 node(const std::vector<double> &x, int i) { return V3{ { x[3 * i], x[3 * i + 1], x[3 * i + 2] } }; }

// Eigen 3.2.5 Matrix3d::inverse (Eigen/src/LU/Inverse.h:118-160), m and r column-major
void inverse3(const double *m, double *r) {
#define M_(i, j) m[3 * (j) + (i)]
#define COF(i, j) (M_((i + 1) % 3, (j + 1) % 3) * M_((i + 2) % 3, (j + 2) % 3) - M_((i + 1) % 3, (j + 2) % 3) * M_((i + 2) % 3, (j + 1) % 3))
	const double c0 = COF(0, 0), c1 = COF(1, 0), c2 = COF(2, 0);
	const double det = c0 * M_(0, 0) + (c1 * M_(1, 0) + c2 * M_(2, 0));
	const double invdet = 1.0 / det;
#define R_(i, j) r[3 * (j) + (i)]
	R_(0, 0) = c0 * invdet; R_(0, 1) = c1 * invdet; R_(0, 2) = c2 * invdet;
	R_(1, 0) = COF(0, 1) * invdet; R_(1, 1) = COF(1, 1) * invdet; R_(1, 2) = COF(2, 1) * invdet;
	R_(2, 0) = COF(0, 2) * invdet; R_(2, 1) = COF(1, 2) * invdet; R_(2, 2) = COF(2, 2) * invdet;
#undef R_
#undef COF
#undef M_
}

} // namespace

int compute_rest_state(admmb_ctx *ctx, Batch &b) {
	const std::vector<double> &x = ctx->h_x0;
	const int n = ctx->n;
	for (size_t i = 0; i < b.idx.size(); ++i)
		if (b.idx[i] < 0 || b.idx[i] >= n) ADMMB_FAIL(ctx, ADMMB_E_ARG, "force node index %d out of range [0,%d)", b.idx[i], n);

	b.S.assign((size_t)b.count * b.nsel, 0.0);
	b.kk.assign(b.count, 0.0);
	b.aux.assign((size_t)b.count * b.naux, 0.0);
	const bool keep_w = (b.w.size() == (size_t)b.count); // weights already overridden by the caller
	if (!keep_w) b.w.assign(b.count, 0.0);

	switch (b.type) {
	case BT_TETS:
		for (int e = 0; e < b.count; ++e) {
			// helper::init_tet_force  TetForce.cpp:28-57
			const int *id = &b.idx[4 * e];
			const V3 v0 = node(x, id[0]), v1 = node(x, id[1]), v2 = node(x, id[2]), v3 = node(x, id[3]);
			double edges[9], inv[9];
			const V3 e0 = sub(v1, v0), e1 = sub(v2, v0), e2 = sub(v3, v0);
			for (int r = 0; r < 3; ++r) { edges[r] = e0[r]; edges[3 + r] = e1[r]; edges[6 + r] = e2[r]; }
			inverse3(edges, inv);
			// B = D * Xg^-1, D = [-1 -1 -1; I]: B(c,r) stored at S[e*12 + c*3 + r]
			double *B = &b.S[(size_t)e * 12];
			for (int r = 0; r < 3; ++r) {
				const double a0 = inv[3 * r + 0], a1 = inv[3 * r + 1], a2 = inv[3 * r + 2]; // Xinv(k,r)
				B[0 * 3 + r] = (-1.0 * a0 + -1.0 * a1) + -1.0 * a2;
				B[1 * 3 + r] = a0;
				B[2 * 3 + r] = a1;
				B[3 * 3 + r] = a2;
			}
			const double volume = std::fabs(dot(sub(v0, v3), cross(sub(v1, v3), sub(v2, v3)))) / 6.0;
			double stiff;
			if (b.kind == ADMMB_TET_NEOHOOKEAN || b.kind == ADMMB_TET_STVK) {
				stiff = std::min(b.p0, b.p1); // TetForce.cpp:306
				b.kk[e] = stiff;              // prox penalty k (NHProx/StVKProx k)
			} else {
				stiff = b.p0;
				b.kk[e] = stiff * volume;     // TetForce.cpp:147,205
			}
			// weight = sqrtf(stiffness)*sqrtf(volume)  (float arithmetic, TetForce.cpp:116,162,307)
			if (!keep_w) b.w[e] = (double)(sqrtf((float)stiff) * sqrtf((float)volume));
		}
		break;
	case BT_TRIS:
		for (int e = 0; e < b.count; ++e) {
			// LimitedTriangleStrain::initialize  TriangleForce.cpp:29-63 (FungTriangle :170-203 is the same shape code)
			const int *id = &b.idx[3 * e];
			const V3 x1 = node(x, id[0]), x2 = node(x, id[1]), x3 = node(x, id[2]);
			const V3 e12 = sub(x2, x1), e13 = sub(x3, x1);
			const double l12 = norm(e12);
			const V3 n1 = V3{ { e12[0] / l12, e12[1] / l12, e12[2] / l12 } };
			const double pr = dot(e13, n1);
			const V3 t = V3{ { e13[0] - pr * n1[0], e13[1] - pr * n1[1], e13[2] - pr * n1[2] } };
			const double lt = norm(t);
			const V3 n2 = V3{ { t[0] / lt, t[1] / lt, t[2] / lt } };
			// Xg = basis^T * edges (2x2)
			auto d3 = [](const V3 &a, const V3 &c) { return (a[0] * c[0] + a[1] * c[1]) + a[2] * c[2]; };
			const double g00 = d3(n1, e12), g01 = d3(n1, e13), g10 = d3(n2, e12), g11 = d3(n2, e13);
			const double det = g00 * g11 - g10 * g01;
			const double invdet = 1.0 / det;
			const double i00 = g11 * invdet, i10 = -g10 * invdet, i01 = -g01 * invdet, i11 = g00 * invdet;
			// B = D * Xg^-1 (3x2), D = [-1 -1; 1 0; 0 1]; B(j,c) stored at S[e*6 + j*2 + c]
			double *B = &b.S[(size_t)e * 6];
			B[0] = -1.0 * i00 + -1.0 * i10; B[1] = -1.0 * i01 + -1.0 * i11;
			B[2] = i00; B[3] = i01;
			B[4] = i10; B[5] = i11;
			const double area = std::fabs(det / 2.0);
			if (b.kind == ADMMB_TRI_FUNG) {
				if (!keep_w) b.w[e] = std::sqrt(b.p0) * std::sqrt(area); // TriangleForce.cpp:196
				b.kk[e] = b.p0;
			} else {
				if (!keep_w) b.w[e] = (double)(sqrtf((float)b.p0) * sqrtf((float)area)); // :62
				b.kk[e] = b.p0 * area;                                                   // :97
			}
		}
		break;
	case BT_SPRINGS:
		for (int e = 0; e < b.count; ++e) {
			// Spring::initialize  Force.cpp:29-38
			const V3 disp = sub(node(x, b.idx[2 * e]), node(x, b.idx[2 * e + 1]));
			b.aux[e] = norm(disp);
			b.kk[e] = b.stiffness[e];
			if (!keep_w) b.w[e] = std::sqrt(b.stiffness[e]);
		}
		break;
	case BT_BENDS:
		for (int e = 0; e < b.count; ++e) {
			// BendForce::initialize  BendForce.cpp:26-55
			const int *id = &b.idx[4 * e];
			const V3 x0 = node(x, id[0]), x1 = node(x, id[1]), x2 = node(x, id[2]), x3 = node(x, id[3]);
			const V3 xA = sub(x0, x2), xB = sub(x1, x2), xC = V3{ { 0, 0, 0 } }, xD = sub(x3, x2);
			const double area1 = 0.5 * norm(cross(xA, xD));
			const double area2 = 0.5 * norm(cross(xD, xB));
			const double hA = 2.0 * area1 / norm(xD);
			const double hB = 2.0 * area2 / norm(xD);
			const V3 nC = cross(sub(xC, xB), sub(xC, xA));
			const V3 nD = cross(sub(xD, xA), sub(xD, xB));
			double *al = &b.aux[(size_t)e * 4];
			al[0] = hB / (hA + hB);
			al[1] = hA / (hA + hB);
			al[2] = -norm(nD) / (norm(nC) + norm(nD));
			al[3] = -norm(nC) / (norm(nC) + norm(nD));
			b.kk[e] = b.p0;
			if (!keep_w) b.w[e] = std::sqrt(b.p0);
		}
		break;
	case BT_STATIC_ANCHORS:
		for (int e = 0; e < b.count; ++e) {
			// StaticAnchor::initialize  AnchorForce.cpp:29-33; default weight 1000.f (AnchorForce.hpp:57-60)
			for (int j = 0; j < 3; ++j) b.aux[(size_t)e * 3 + j] = x[3 * b.idx[e] + j];
			if (!keep_w) b.w[e] = (b.anchor_weight > 0.0) ? b.anchor_weight : 1000.0;
		}
		break;
	case BT_MOVING_ANCHORS:
		// aux (control point positions) was filled by admmb_add_moving_anchors; keep it
		b.aux = b.stiffness; // staged there by add_moving_anchors
		for (int e = 0; e < b.count; ++e)
			if (!keep_w) b.w[e] = (b.anchor_weight > 0.0) ? b.anchor_weight : 1000.0;
		break;
	case BT_COLLISION:
		for (int e = 0; e < b.count; ++e)
			if (!keep_w) b.w[e] = b.anchor_weight; // CollisionForce.hpp:30 (default 32)
		break;
	default:
		ADMMB_FAIL(ctx, ADMMB_E_ARG, "unknown batch type %d", b.type);
	}
	// Corner order.  The reference forms Dx = m_D * curr_x with Eigen's column-major sparse product
	// (System.cpp:54, SparseDenseProduct.h:190-209): every row accumulates its terms in ascending COLUMN, i.e.
	// ascending node index.  D_i x is symmetric in the corners, so we store each tet / triangle with its corners
	// sorted by node index; the kernels' left-to-right sums then round exactly like the reference's Dx.
	if (b.type == BT_TETS || b.type == BT_TRIS) {
		const int nv = b.nv, per = b.nsel / b.nv;
		for (int e = 0; e < b.count; ++e) {
			int *id = &b.idx[(size_t)e * nv];
			double *S = &b.S[(size_t)e * b.nsel];
			for (int a = 1; a < nv; ++a)          // insertion sort of <= 4 corners, rows of S move with them
				for (int c = a; c > 0 && id[c - 1] > id[c]; --c) {
					std::swap(id[c - 1], id[c]);
					for (int k = 0; k < per; ++k) std::swap(S[(c - 1) * per + k], S[c * per + k]);
				}
		}
	}
	return ADMMB_OK;
}

void wind_wavefronts(int n, int ntris, const int *tris3, std::vector<int> &level_ptr, std::vector<int> &order) {
	std::vector<int> node_level(n, 0), level(ntris);
	int depth = 0;
	for (int t = 0; t < ntris; ++t) {
		int l = 0;
		for (int c = 0; c < 3; ++c) l = std::max(l, node_level[tris3[3 * (size_t)t + c]]);
		level[t] = l;
		for (int c = 0; c < 3; ++c) node_level[tris3[3 * (size_t)t + c]] = l + 1;
		depth = std::max(depth, l + 1);
	}
	level_ptr.assign(depth + 1, 0);
	for (int t = 0; t < ntris; ++t) level_ptr[level[t] + 1]++;
	for (int l = 0; l < depth; ++l) level_ptr[l + 1] += level_ptr[l];
	std::vector<int> fill(level_ptr.begin(), level_ptr.end() - 1);
	order.resize(ntris);
	for (int t = 0; t < ntris; ++t) order[fill[level[t]]++] = t;
}

} // namespace admmb
