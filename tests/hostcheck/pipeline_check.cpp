// tests/hostcheck/pipeline_check.cpp -- TEST-ONLY: one force batch's local step on the CPU through the product's
// own code: csrc/rest_state.cpp (Force::initialize restatement) + csrc/local_bodies.h (the per-force body the
// CUDA kernels execute).  Lets the "not gpu" test suite replay the golden reference dumps without a device.
#include <cstring>
#include <numeric>
#include <vector>

#include "../../admm-elastic-sca_b200/csrc/common.h"
#include "../../admm-elastic-sca_b200/csrc/local_bodies.h"

using namespace admmb;

namespace {
template <class T>
void to_soa(const std::vector<T> &src, int ncomp, int cnt, std::vector<T> &dst) {
	dst.resize((size_t)cnt * ncomp);
	for (int k = 0; k < ncomp; ++k)
		for (int p = 0; p < cnt; ++p) dst[(size_t)k * cnt + p] = src[(size_t)p * ncomp + k];
}
} // namespace

extern "C" int hc_batch_local(int type, int kind, int n, const double *x_rest, int count, const int *idx, const double *stiffness,
                              double p0, double p1, double p2, int maxit, int flag, double anchor_weight, const double *anchor_pos,
                              const int *active, int nshapes, const int *shape_kind, const double *shape_params, double dt,
                              const double *x_cur, const double *u_in, double *state, double *z_out, double *u_out, double *P_out,
                              double *w_out, double *anchor_pos_out) {
	static const int NV[] = { 4, 3, 2, 4, 1, 1, 1 }, ROWS[] = { 9, 6, 3, 9, 3, 3, 3 }, NSEL[] = { 12, 6, 0, 0, 0, 0, 0 }, NAUX[] = { 0, 0, 1, 4, 3, 3, 0 };
	admmb_ctx ctx;
	ctx.n = n;
	ctx.dt = dt;
	ctx.h_x0.assign(x_rest, x_rest + 3 * (size_t)n);
	Batch b;
	b.type = type; b.kind = kind; b.count = count; b.nv = NV[type]; b.rows = ROWS[type]; b.nsel = NSEL[type]; b.naux = NAUX[type];
	const bool hyper = type == BT_TETS && (kind == ADMMB_TET_NEOHOOKEAN || kind == ADMMB_TET_STVK);
	b.nstate = hyper ? 4 : ((type == BT_TRIS && kind == ADMMB_TRI_FUNG) ? 1 : 0);
	b.p0 = p0; b.p1 = p1; b.p2 = p2; b.max_iterations = maxit; b.flag = flag; b.anchor_weight = anchor_weight;
	if (type == BT_COLLISION) { b.idx.resize(n); std::iota(b.idx.begin(), b.idx.end(), 0); }
	else b.idx.assign(idx, idx + (size_t)count * b.nv);
	if (type == BT_SPRINGS) b.stiffness.assign(stiffness, stiffness + count);
	if (type == BT_MOVING_ANCHORS) { b.stiffness.assign(anchor_pos, anchor_pos + 3 * (size_t)count); b.active.assign(active, active + count); }
	if (type == BT_COLLISION) {
		b.shape_kind.assign(shape_kind, shape_kind + nshapes);
		b.shape_params.assign(shape_params, shape_params + 4 * (size_t)nshapes);
		for (int i = 0; i < nshapes; ++i) if (shape_kind[i] == ADMMB_SHAPE_CYLINDER) b.shape_params[4 * i + 2] = 0.0;
	}
	if (compute_rest_state(&ctx, b) != 0) return -1;
	// structure-of-arrays staging exactly as upload_batch() builds it (identity node / force order)
	std::vector<int> idx_soa;
	to_soa(b.idx, b.nv, count, idx_soa);
	std::vector<double> S, aux, u, z((size_t)b.rows * count), st, wdt2(count), P((size_t)count * b.nv * 3);
	if (b.nsel) to_soa(b.S, b.nsel, count, S);
	if (b.naux) to_soa(b.aux, b.naux, count, aux);
	std::vector<double> u_aos(u_in, u_in + (size_t)b.rows * count);
	to_soa(u_aos, b.rows, count, u);
	if (b.nstate) { std::vector<double> s_aos(state, state + (size_t)b.nstate * count); to_soa(s_aos, b.nstate, count, st); }
	for (int e = 0; e < count; ++e) wdt2[e] = dt * dt * b.w[e] * b.w[e];
	std::vector<int> its(count, 0);
	LocalArgs a;
	memset(&a, 0, sizeof(a));
	a.count = count; a.idx = idx_soa.data(); a.S = S.data(); a.w = b.w.data(); a.wdt2 = wdt2.data(); a.kk = b.kk.data();
	a.aux = aux.data(); a.u = u.data(); a.z = z.data(); a.state = st.data(); a.its = its.data();
	a.active = (type == BT_MOVING_ANCHORS) ? b.active.data() : nullptr;
	a.x = x_cur; a.P = P.data(); a.p0 = p0; a.p1 = p1; a.p2 = p2; a.kprox = std::min(p0, p1); a.max_iterations = maxit; a.flag = flag;
	a.shape_kind = b.shape_kind.data(); a.shape_params = b.shape_params.data(); a.nshapes = nshapes;
	for (int e = 0; e < count; ++e) {
		switch (type) {
		case BT_TETS:
			if (kind == ADMMB_TET_LINEAR_STRAIN) local_tet<ADMMB_TET_LINEAR_STRAIN, 1>(a, e);
			else if (kind == ADMMB_TET_VOLUME) local_tet<ADMMB_TET_VOLUME, 1>(a, e);
			else if (kind == ADMMB_TET_NEOHOOKEAN) { double park[18]; if (maxit <= 5) local_tet_hyper<NHModel, 5>(a, e, park, 1); else local_tet_hyper<NHModel, 10>(a, e, park, 1); }
			else { double park[18]; if (maxit <= 5) local_tet_hyper<StVKModel, 5>(a, e, park, 1); else local_tet_hyper<StVKModel, 10>(a, e, park, 1); }
			break;
		case BT_TRIS:
			if (kind == ADMMB_TRI_LIMITED_STRAIN) local_tri<ADMMB_TRI_LIMITED_STRAIN>(a, e);
			else if (kind == ADMMB_TRI_AREA) local_tri<ADMMB_TRI_AREA>(a, e);
			else local_tri<ADMMB_TRI_FUNG>(a, e);
			break;
		case BT_SPRINGS: local_spring(a, e); break;
		case BT_BENDS: local_bend(a, e); break;
		case BT_STATIC_ANCHORS:
		case BT_MOVING_ANCHORS: local_anchor(a, e); break;
		case BT_COLLISION: local_collision(a, e); break;
		}
	}
	for (int k = 0; k < b.rows; ++k)
		for (int e = 0; e < count; ++e) { z_out[(size_t)e * b.rows + k] = z[(size_t)k * count + e]; u_out[(size_t)e * b.rows + k] = u[(size_t)k * count + e]; }
	for (int k = 0; k < b.nstate; ++k)
		for (int e = 0; e < count; ++e) state[(size_t)e * b.nstate + k] = st[(size_t)k * count + e];
	if (P_out) memcpy(P_out, P.data(), P.size() * sizeof(double));
	if (w_out) memcpy(w_out, b.w.data(), count * sizeof(double));
	if (anchor_pos_out && type == BT_MOVING_ANCHORS)
		for (int k = 0; k < 3; ++k)
			for (int e = 0; e < count; ++e) anchor_pos_out[(size_t)e * 3 + k] = aux[(size_t)k * count + e];
	return 0;
}

// WindForce::project through the product's wavefront schedule (csrc/rest_state.cpp wind_wavefronts + elastic_math.h
// wind_triangle), levels walked in order and the triangles of a level in REVERSE order -- any order inside a level must
// give the serial loop's result.  serial != 0 walks the triangle list as given instead.
extern "C" int hc_wind(int n, int ntris, const int *tris3, const double *dir, double dt, const double *x, double *v, int serial) {
	std::vector<int> ptr, order;
	if (serial) { ptr = { 0, ntris }; order.resize(ntris); for (int t = 0; t < ntris; ++t) order[t] = t; }
	else admmb::wind_wavefronts(n, ntris, tris3, ptr, order);
	for (size_t l = 0; l + 1 < ptr.size(); ++l) {
		const int lo = ptr[l], hi = ptr[l + 1];
		for (int q = 0; q < hi - lo; ++q) {
			const int t = order[serial ? lo + q : hi - 1 - q];
			const size_t a = 3 * (size_t)tris3[3 * t], b = 3 * (size_t)tris3[3 * t + 1], c = 3 * (size_t)tris3[3 * t + 2];
			double f[3];
			admmb::wind_triangle(x + a, x + b, x + c, v + a, v + b, v + c, dir, dt, f);
			const size_t ids[3] = { a, b, c };
			for (int k = 0; k < 3; ++k) { v[ids[k]] += f[0]; v[ids[k] + 1] += f[1]; v[ids[k] + 2] += f[2]; }
		}
	}
	return (int)ptr.size() - 1;
}
