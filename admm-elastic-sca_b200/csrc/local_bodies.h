// local_bodies.h -- the per-force body of the ADMM local step, shared by the CUDA kernels (kernels_local.cu, one
// thread per force) and by the test-only CPU harness (tests/hostcheck), so that the "not gpu" tests execute the
// very code the device runs.  See kernels_local.cu for the layout conventions.
#pragma once
#include "elastic_math.h"
#include "../../include/admm_b200.h"

namespace admmb {

struct LocalArgs {
	int count;
	const int *idx;       // [nv][count]
	const double *S;      // [nsel][count]
	const double *w;      // [count]
	const double *wdt2;   // [count]  dt^2 w^2
	const double *kk;     // [count]
	double *aux;          // [naux][count]
	double *u, *z;        // [rows][count]
	double *state;        // [nstate][count]
	int *its;             // [count] or null
	int *trips;           // [count] or null: line-search trial points of the last project() (diagnostics: lane divergence)
	const int *active;    // moving anchors
	const double *x;      // [n][3]
	double *P;            // [slots][3], already offset to this batch's first slot
	double p0, p1, p2;
	double kprox;         // hyperelastic tets: the prox penalty k = min(mu, lambda), the same for every tet (TetForce.cpp:306)
	int max_iterations, flag;
	const int *shape_kind;
	const double *shape_params;
	int nshapes;
};

// ---- tets -------------------------------------------------------------------------------------------
// B(c,r) of force e and row 3r+j of D_i x = sum_c B(c,r) x_c[j] (TetForce.cpp:67-75), corners in ascending node
// order so that the left-to-right sum rounds like Eigen's column-major sparse product (rest_state.cpp).
ADMMB_HD void tet_load_Dx(const LocalArgs &a, const int e, double *B, double *Dx) {
	const int n = a.count;
#pragma unroll
	for (int k = 0; k < 12; ++k) B[k] = a.S[(size_t)k * n + e];
	double xs[4][3];
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		const int id = a.idx[(size_t)c * n + e];
#pragma unroll
		for (int j = 0; j < 3; ++j) xs[c][j] = a.x[3 * (size_t)id + j];
	}
#pragma unroll
	for (int r = 0; r < 3; ++r)
#pragma unroll
		for (int j = 0; j < 3; ++j)
			Dx[3 * r + j] = ((B[0 * 3 + r] * xs[0][j] + B[1 * 3 + r] * xs[1][j]) + B[2 * 3 + r] * xs[2][j]) + B[3 * 3 + r] * xs[3][j];
	ADMMB_FLOPS(9 * 7);
}

// u += D_i x - z, store u and z, and this force's share of the right-hand side dt^2 D_i^T W_i^2 (z - u).
ADMMB_HD void tet_finish(const LocalArgs &a, const int e, const double *B, const double *Dx, const double *u, const double *z) {
	const int n = a.count;
	double zu[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) {
		const double un = u[k] + (Dx[k] - z[k]); // ui += (Dix - zi)
		a.u[(size_t)k * n + e] = un;
		a.z[(size_t)k * n + e] = z[k];
		zu[k] = z[k] - un;                       // curr_z - curr_u
	}
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 12;
#pragma unroll
	for (int v = 0; v < 4; ++v)
#pragma unroll
		for (int j = 0; j < 3; ++j)
			P[3 * v + j] = c * ((B[v * 3 + 0] * zu[j] + B[v * 3 + 1] * zu[3 + j]) + B[v * 3 + 2] * zu[6 + j]);
	ADMMB_FLOPS(9 * 3 + 12 * 6);
}

// LinearTetStrain / TetVolume: closed-form projection, everything stays in registers.
template <int KIND, int MH>
ADMMB_HD void local_tet(const LocalArgs &a, const int e) {
	const int n = a.count;
	double B[12], Dx[9], u[9], q[9], z[9];
	tet_load_Dx(a, e, B, Dx);
#pragma unroll
	for (int k = 0; k < 9; ++k) { u[k] = a.u[(size_t)k * n + e]; q[k] = Dx[k] + u[k]; }
	if (KIND == ADMMB_TET_LINEAR_STRAIN) arap_tet_z(q, a.kk[e], a.w[e], z);
	else volume_tet_z(q, a.kk[e], a.w[e], a.p1, a.p2, z);
	tet_finish(a, e, B, Dx, u, z);
}

// HyperElasticTet (TetForce.cpp:320-364).  The optimiser on the three singular values is where the time goes
// (in steady state the reference's line search burns its full 20 + 2 evaluations per tet), and it needs few live
// values: Sigma_0, the iterate, the L-BFGS history.  Everything else is kept OUT of registers while it runs --
// U and V are parked in `park` (shared memory on the device: park[k * ps], k < 18), and D_i x, u and B are simply
// re-read / re-formed afterwards (bit-identical: same operations on the same data) -- which takes the kernel
// from 168 to ~100 registers and doubles the resident warps that hide the FP64 latency.
template <class Model, int MH>
ADMMB_HD void local_tet_hyper(const LocalArgs &a, const int e, double *park, const int ps) {
	const int n = a.count;
	ProxParams P;
	P.mu = a.p0; P.lambda = a.p1; P.k = a.kprox;
	{
		double B[12], q[9], U[9], V[9];
		tet_load_Dx(a, e, B, q);
#pragma unroll
		for (int k = 0; k < 9; ++k) q[k] = q[k] + a.u[(size_t)k * n + e];
		ADMMB_FLOPS(9);
		oriented_svd3(q, U, P.s0, V);
#pragma unroll
		for (int k = 0; k < 9; ++k) { park[k * ps] = U[k]; park[(9 + k) * ps] = V[k]; }
	}
	double x2[3] = { a.state[e], a.state[(size_t)n + e], a.state[2 * (size_t)n + e] };
	double ih = a.state[3 * (size_t)n + e];
	// initial guess needs positive entries; collapsed-node test case (TetForce.cpp:339-347)
	if (x2[2] < 0.0) { x2[2] *= -1.0; }
	else if (fabs(x2[0]) < 1.e-3 && fabs(x2[1]) < 1.e-3 && fabs(x2[2]) < 1.e-3) { x2[0] = 1.e-3; x2[1] = 1.e-3; x2[2] = 1.e-3; }
	int trips = 0;
	const int its = lbfgs_minimize<Model, ProxParams, 3, MH>(P, x2, a.max_iterations, 1e-8, ih, &trips);
	if (a.trips) a.trips[e] = trips;
	a.state[e] = x2[0]; a.state[(size_t)n + e] = x2[1]; a.state[2 * (size_t)n + e] = x2[2]; a.state[3 * (size_t)n + e] = ih;
	if (a.its) a.its[e] = its;

	double z[9];
	{
		double U[9], V[9];
#pragma unroll
		for (int k = 0; k < 9; ++k) { U[k] = park[k * ps]; V[k] = park[(9 + k) * ps]; }
		usvt3(U, x2, V, z);
	}
	double B[12], Dx[9], u[9];
	tet_load_Dx(a, e, B, Dx);
	ADMMB_FLOPS(-9 * 7); // D_i x is re-formed here only to keep it out of registers during the optimiser: counted once
#pragma unroll
	for (int k = 0; k < 9; ++k) u[k] = a.u[(size_t)k * n + e];
	tet_finish(a, e, B, Dx, u, z);
}

// ---- triangles --------------------------------------------------------------------------------------
template <int KIND>
ADMMB_HD void local_tri(const LocalArgs &a, const int e) {
	const int n = a.count;
	double B[6], Dx[6], u[6], q[6], z[6]; // B(j,c) at [j*2+c]
#pragma unroll
	for (int k = 0; k < 6; ++k) B[k] = a.S[(size_t)k * n + e];
	{
		double xs[3][3];
#pragma unroll
		for (int c = 0; c < 3; ++c) {
			const int id = a.idx[(size_t)c * n + e];
#pragma unroll
			for (int j = 0; j < 3; ++j) xs[c][j] = a.x[3 * (size_t)id + j];
		}
		// rows i and 3+i: sum_j B(j,0) x_j[i], sum_j B(j,1) x_j[i]   (TriangleForce.cpp:65-76)
#pragma unroll
		for (int c = 0; c < 2; ++c)
#pragma unroll
			for (int i = 0; i < 3; ++i)
				Dx[3 * c + i] = (B[0 * 2 + c] * xs[0][i] + B[1 * 2 + c] * xs[1][i]) + B[2 * 2 + c] * xs[2][i];
	}
#pragma unroll
	for (int k = 0; k < 6; ++k) { u[k] = a.u[(size_t)k * n + e]; q[k] = Dx[k] + u[k]; }

	if (KIND == ADMMB_TRI_LIMITED_STRAIN) tri_strain_z(q, a.kk[e], a.w[e], a.p1, a.p2, a.flag != 0, z);
	else if (KIND == ADMMB_TRI_AREA) tri_area_z(q, a.kk[e], a.w[e], a.p1, a.p2, a.flag, z);
	else {
		double ih = a.state[e];
		fung_tri_z(q, a.p0, &ih, z);
		a.state[e] = ih;
	}
	double zu[6];
#pragma unroll
	for (int k = 0; k < 6; ++k) {
		const double un = u[k] + (Dx[k] - z[k]);
		a.u[(size_t)k * n + e] = un;
		a.z[(size_t)k * n + e] = z[k];
		zu[k] = z[k] - un;
	}
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 9;
#pragma unroll
	for (int v = 0; v < 3; ++v)
#pragma unroll
		for (int j = 0; j < 3; ++j) P[3 * v + j] = c * (B[v * 2 + 0] * zu[j] + B[v * 2 + 1] * zu[3 + j]);
}

// ---- springs ----------------------------------------------------------------------------------------
ADMMB_HD void local_spring(const LocalArgs &a, const int e) {
	const int n = a.count;
	const int i0 = a.idx[e], i1 = a.idx[(size_t)n + e];
	double Dx[3], u[3], q[3], z[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		Dx[j] = a.x[3 * (size_t)i0 + j] - a.x[3 * (size_t)i1 + j]; // rows: +1 at idx0, -1 at idx1 (Force.cpp:44-47)
		u[j] = a.u[(size_t)j * n + e];
		q[j] = Dx[j] + u[j];
	}
	spring_z(q, a.kk[e], a.w[e], a.aux[e], z);
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 6;
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		const double un = u[j] + (Dx[j] - z[j]);
		a.u[(size_t)j * n + e] = un;
		a.z[(size_t)j * n + e] = z[j];
		const double f = c * (z[j] - un);
		P[j] = f;
		P[3 + j] = -f;
	}
}

// ---- bending hinges ---------------------------------------------------------------------------------
ADMMB_HD void local_bend(const LocalArgs &a, const int e) {
	const int n = a.count;
	double xs[4][3];
#pragma unroll
	for (int c = 0; c < 4; ++c) {
		const int id = a.idx[(size_t)c * n + e];
#pragma unroll
		for (int j = 0; j < 3; ++j) xs[c][j] = a.x[3 * (size_t)id + j];
	}
	double Dx[9], u[9], q[9], z[9], al[4];
#pragma unroll
	for (int j = 0; j < 3; ++j) { // rows (x0-x2), (x3-x2), (x1-x2)   BendForce.cpp:89-128
		Dx[j] = xs[0][j] - xs[2][j];
		Dx[3 + j] = xs[3][j] - xs[2][j];
		Dx[6 + j] = xs[1][j] - xs[2][j];
	}
#pragma unroll
	for (int k = 0; k < 9; ++k) { u[k] = a.u[(size_t)k * n + e]; q[k] = Dx[k] + u[k]; }
#pragma unroll
	for (int k = 0; k < 4; ++k) al[k] = a.aux[(size_t)k * n + e];
	bend_z(q, a.kk[e], a.w[e], al, z);
	double zu[9];
#pragma unroll
	for (int k = 0; k < 9; ++k) {
		const double un = u[k] + (Dx[k] - z[k]);
		a.u[(size_t)k * n + e] = un;
		a.z[(size_t)k * n + e] = z[k];
		zu[k] = z[k] - un;
	}
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 12;
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		P[0 * 3 + j] = c * zu[j];                                  // node idx0: +row0
		P[1 * 3 + j] = c * zu[6 + j];                              // node idx1: +row2
		P[2 * 3 + j] = c * ((-zu[j] - zu[3 + j]) - zu[6 + j]);     // node idx2: -row0 -row1 -row2
		P[3 * 3 + j] = c * zu[3 + j];                              // node idx3: +row1
	}
}

// ---- anchors (StaticAnchor / MovingAnchor) ---------------------------------------------------------
ADMMB_HD void local_anchor(const LocalArgs &a, const int e) {
	const int n = a.count;
	const int id = a.idx[e];
	const bool act = a.active ? (a.active[e] != 0) : true;
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 3;
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		const double Dx = a.x[3 * (size_t)id + j];
		const double u = a.u[(size_t)j * n + e];
		double z;
		if (act) {
			z = a.aux[(size_t)j * n + e];       // zi = pos (AnchorForce.cpp:46-55, :76-77)
		} else {
			z = Dx + u;                          // zi = Dix + ui; point->pos = Dx  (AnchorForce.cpp:79-83)
			a.aux[(size_t)j * n + e] = Dx;
		}
		const double un = u + (Dx - z);
		a.u[(size_t)j * n + e] = un;
		a.z[(size_t)j * n + e] = z;
		P[j] = c * (z - un);
	}
}

// ---- collision (one "force" spanning all nodes; one thread per node) --------------------------------
ADMMB_HD void local_collision(const LocalArgs &a, const int e) {
	const int n = a.count;
	double Dx[3], u[3], p[3];
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		Dx[j] = a.x[3 * (size_t)e + j];
		u[j] = a.u[(size_t)j * n + e];
		p[j] = Dx[j] + u[j];
	}
	collide_point(p, a.shape_params, a.shape_kind, a.nshapes);
	const double c = a.wdt2[e];
	double *P = a.P + (size_t)e * 3;
#pragma unroll
	for (int j = 0; j < 3; ++j) {
		const double un = u[j] + (Dx[j] - p[j]);
		a.u[(size_t)j * n + e] = un;
		a.z[(size_t)j * n + e] = p[j];
		P[j] = c * (p[j] - un);
	}
}

} // namespace admmb
