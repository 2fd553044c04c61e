// tests/hostcheck/hostcheck.cpp -- TEST-ONLY harness: compiles csrc/elastic_math.h as plain C++ so the
// exact per-element arithmetic of the CUDA kernels can be checked on a CPU-only box against the
// reference.  Not linked into the product library; the product has no CPU path.
#include "../../admm-elastic-sca_b200/csrc/elastic_math.h"
using namespace admmb;

extern "C" {

// kind: 0 ARAP (p0 = stiffness*volume, w), 1 NH, 2 StVK (p0=mu,p1=lambda,k=min), 3 volume (p0=k, p1=lmin, p2=lmax)
void hc_tet_z(int kind, int n, const double *q, const double *w, const double *kk, double p0, double p1, double p2,
              int maxit, double *state, double *z, int *its) {
	for (int e = 0; e < n; ++e) {
		const double *qe = q + 9 * e;
		double *ze = z + 9 * e;
		int it = 0;
		switch (kind) {
		case 0: arap_tet_z(qe, kk[e], w[e], ze); break;
		case 1: it = (maxit <= 5) ? hyperelastic_tet_z<NHModel, 5>(qe, p0, p1, dmin(p0, p1), maxit, state + 4 * e, ze)
		                          : hyperelastic_tet_z<NHModel, 10>(qe, p0, p1, dmin(p0, p1), maxit, state + 4 * e, ze); break;
		case 2: it = (maxit <= 5) ? hyperelastic_tet_z<StVKModel, 5>(qe, p0, p1, dmin(p0, p1), maxit, state + 4 * e, ze)
		                          : hyperelastic_tet_z<StVKModel, 10>(qe, p0, p1, dmin(p0, p1), maxit, state + 4 * e, ze); break;
		case 3: volume_tet_z(qe, kk[e], w[e], p1, p2, ze); break;
		}
		if (its) its[e] = it;
	}
}

// kind: 0 strain (flag = strain_limiting), 1 area (flag = iters), 2 fung (p0 = mu, state = init_hess per tri)
void hc_tri_z(int kind, int n, const double *q, const double *w, const double *kk, double lmin, double lmax, int flag,
              double mu, double *state, double *z) {
	for (int e = 0; e < n; ++e) {
		const double *qe = q + 6 * e;
		double *ze = z + 6 * e;
		switch (kind) {
		case 0: tri_strain_z(qe, kk[e], w[e], lmin, lmax, flag != 0, ze); break;
		case 1: tri_area_z(qe, kk[e], w[e], lmin, lmax, flag, ze); break;
		case 2: fung_tri_z(qe, mu, state + e, ze); break;
		}
	}
}

void hc_spring_z(int n, const double *q, const double *stiffness, const double *w, const double *rest, double *z) {
	for (int e = 0; e < n; ++e) spring_z(q + 3 * e, stiffness[e], w[e], rest[e], z + 3 * e);
}
void hc_bend_z(int n, const double *q, double stiffness, double w, const double *alpha, double *z) {
	for (int e = 0; e < n; ++e) bend_z(q + 9 * e, stiffness, w, alpha + 4 * e, z + 9 * e);
}
void hc_collide(int n, double *p, const double *shapes, const int *kinds, int ns) {
	for (int i = 0; i < n; ++i) collide_point(p + 3 * i, shapes, kinds, ns);
}
void hc_svd3(const double *F, double *U, double *S, double *V, int oriented) {
	if (oriented) oriented_svd3(F, U, S, V); else jacobi_svd3(F, U, S, V);
}
void hc_svd32(const double *F, double *Ut, double *S, double *V) { svd32(F, Ut, S, V); }
// the libm clones against the libm of this box, element by element; returns the number of results that differ in any bit
long hc_exp_mismatches(long n, const double *x) {
	long bad = 0;
	for (long i = 0; i < n; ++i) { const double a = glibc_exp(x[i]), b = exp(x[i]); if (memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) ++bad; }
	return bad;
}
long hc_log_mismatches(long n, const double *x) {
	long bad = 0;
	for (long i = 0; i < n; ++i) { const double a = glibc_log(x[i]), b = log(x[i]); if (memcmp(&a, &b, 8) != 0 && !(a != a && b != b)) ++bad; }
	return bad;
}

}
