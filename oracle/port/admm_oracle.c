/* oracle/port/admm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C restatement of the reference's solver path for SMALL systems (dense Cholesky for the global step):
 * admm::System::initialize()/step() and every Force::project, each function citing the reference file:line it
 * follows (A/ = /root/reference/deps/admm-elastic-sca).  It exists so that a checker is available where the
 * unmodified reference (oracle/_ref, built from /root/reference) is not; it is pinned against the reference's
 * golden dumps by tests/test_oracle_port.py: replayed from the reference's own inputs, the local step of every golden
 * iteration (z, u, optimiser state; every force class) is BIT-EXACT; free-running,
 * x agrees to rounding where the reference is reproducible and within its own sensitivity where it is not.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.  Nothing here is shared with the product's csrc/ (written separately on purpose).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define FLT_MAX_D ((double)FLT_MAX)
enum { F_TET = 0, F_TRI, F_SPRING, F_BEND, F_SANCHOR, F_MANCHOR, F_COLLISION };

typedef struct {
	int type, kind, rows, nv;
	int idx[4];
	double B[12];       /* tets: B(c,r) at [3c+r]; tris: B(j,c) at [2j+c] */
	double w, k;        /* ADMM weight, blend constant / prox penalty */
	double p0, p1, p2;  /* material */
	int maxit, flag;
	double aux[4];      /* spring: rest length; bend: alpha; anchors: pos */
	int active;
	double prox[3], init_hess; /* HyperElasticTet::last_prox_result, lbfgssolver settings_.init_hess */
	int last_iters;
	long row;           /* first live row */
} force_t;

typedef struct oracle_sys {
	int n;
	double dt;
	double *x0, *m;     /* rest positions 3n, per-node mass n */
	force_t *f;
	int nf, cap;
	int nshapes; int *shape_kind; double *shape_par; double coll_w; int has_coll; long coll_row;
	double grav[8][3]; int ngrav;
	int *wind_tris; int wind_nt; double wind_dir[3]; int wind_after_gravity;
	long rows;
	double *L;          /* dense Cholesky factor of A_n (n x n, lower, row-major) */
	double *u, *z;
} oracle_sys;

/* ---------------------------------------------------------------------------------------------------------
 * Eigen 3.2.5 JacobiSVD<Matrix3d> (A/deps/Eigen3/Eigen/src/SVD/JacobiSVD.h:824-930), matrices row-major m[r][c]
 * --------------------------------------------------------------------------------------------------------- */
static double hyp(double x, double y) { /* Eigen/src/Core/MathFunctions.h:284-302 */
	double p = fmax(fabs(x), fabs(y)), q = fmin(fabs(x), fabs(y)), t;
	if (p == 0.0) return 0.0;
	t = q / p;
	return p * sqrt(1.0 + t * t);
}
static void rot_rows(double a[3][3], int p, int q, double c, double s) { /* Jacobi.h:300-420, x' = c x + s y */
	int j;
	if (c == 1.0 && s == 0.0) return;
	for (j = 0; j < 3; ++j) { double x = a[p][j], y = a[q][j]; a[p][j] = c * x + s * y; a[q][j] = -s * x + c * y; }
}
static void rot_cols(double a[3][3], int p, int q, double c, double s) {
	int i;
	if (c == 1.0 && s == 0.0) return;
	for (i = 0; i < 3; ++i) { double x = a[i][p], y = a[i][q]; a[i][p] = c * x + s * y; a[i][q] = -s * x + c * y; }
}
static void jacobi_svd3(double F[3][3], double U[3][3], double S[3], double V[3][3]) {
	double W[3][3], scale = 0.0;
	int i, j, p, q, finished = 0, guard = 0;
	for (i = 0; i < 3; ++i) for (j = 0; j < 3; ++j) if (fabs(F[i][j]) > scale) scale = fabs(F[i][j]);
	if (scale == 0.0) scale = 1.0;
	for (i = 0; i < 3; ++i) for (j = 0; j < 3; ++j) { W[i][j] = F[i][j] / scale; U[i][j] = V[i][j] = (i == j); }
	while (!finished && guard++ < 64) {
		finished = 1;
		for (p = 1; p < 3; ++p) for (q = 0; q < p; ++q) {
			double thr = fmax(2.0 * 4.9406564584124654e-324, 2.0 * DBL_EPSILON * fmax(fabs(W[p][p]), fabs(W[q][q])));
			double m00, m01, m10, m11, t, d, c1, s1, cr, sr, cl, sl, x, y;
			if (!(fabs(W[p][q]) > thr || fabs(W[q][p]) > thr)) continue;
			finished = 0;
			/* real_2x2_jacobi_svd (JacobiSVD.h:414-441) */
			m00 = W[p][p]; m01 = W[p][q]; m10 = W[q][p]; m11 = W[q][q];
			t = m00 + m11; d = m10 - m01;
			if (t == 0.0) { c1 = 0.0; s1 = d > 0.0 ? 1.0 : -1.0; }
			else { double h = hyp(t, d); c1 = fabs(t) / h; s1 = d / h; if (t < 0.0) s1 = -s1; }
			if (!(c1 == 1.0 && s1 == 0.0)) {
				x = m00; y = m10; m00 = c1 * x + s1 * y; m10 = -s1 * x + c1 * y;
				x = m01; y = m11; m01 = c1 * x + s1 * y; m11 = -s1 * x + c1 * y;
			}
			/* makeJacobi (Jacobi.h:83-113) on [m00 m01; . m11] */
			if (m01 == 0.0) { cr = 1.0; sr = 0.0; }
			else {
				double tau = (m00 - m11) / (2.0 * fabs(m01)), w = sqrt(tau * tau + 1.0), tt, n;
				tt = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
				n = 1.0 / sqrt(tt * tt + 1.0);
				sr = -(tt > 0.0 ? 1.0 : -1.0) * (m01 / fabs(m01)) * fabs(tt) * n;
				cr = n;
			}
			/* j_left = rot1 * j_right^T (Jacobi.h:51-56) */
			cl = c1 * cr - s1 * (-sr);
			sl = c1 * (-sr) + s1 * cr;
			rot_rows(W, p, q, cl, sl);      /* applyOnTheLeft(p,q,j_left)            :888 */
			rot_cols(U, p, q, cl, sl);      /* U.applyOnTheRight(p,q,j_left^T)       :889 */
			rot_cols(W, p, q, cr, -sr);     /* applyOnTheRight(p,q,j_right)          :891 */
			rot_cols(V, p, q, cr, -sr);     /*                                        :892 */
		}
	}
	for (i = 0; i < 3; ++i) { /* :899-906 */
		double a = fabs(W[i][i]);
		S[i] = a;
		if (a != 0.0) { double f = W[i][i] / a; for (j = 0; j < 3; ++j) U[j][i] *= f; }
	}
	for (i = 0; i < 3; ++i) { /* :908-927 selection sort, first maximum wins */
		int pos = i;
		double mx = S[i];
		for (j = i + 1; j < 3; ++j) if (S[j] > mx) { mx = S[j]; pos = j; }
		if (mx == 0.0) break;
		if (pos != i) {
			double t = S[i]; S[i] = S[pos]; S[pos] = t;
			for (j = 0; j < 3; ++j) { t = U[j][i]; U[j][i] = U[j][pos]; U[j][pos] = t; t = V[j][i]; V[j][i] = V[j][pos]; V[j][pos] = t; }
		}
	}
	for (i = 0; i < 3; ++i) S[i] *= scale;
}
static double det3(double m[3][3]) { /* Eigen/src/LU/Determinant.h bruteforce_det3_helper */
	return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
	       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}
/* 9-vector (column-major F, TetForce.cpp:328) <-> row-major 3x3 */
static void vec_to_mat(const double *q, double F[3][3]) { int r, c; for (c = 0; c < 3; ++c) for (r = 0; r < 3; ++r) F[r][c] = q[3 * c + r]; }
static void usvt(double U[3][3], const double *s, double V[3][3], double *out) { /* U diag(s) V^T -> column-major 9-vector */
	int r, c;
	for (c = 0; c < 3; ++c) for (r = 0; r < 3; ++r)
		out[3 * c + r] = (U[r][0] * s[0]) * V[c][0] + (U[r][1] * s[1]) * V[c][1] + (U[r][2] * s[2]) * V[c][2];
}

/* ---------------------------------------------------------------------------------------------------------
 * Prox objectives (TetForce.cpp:216-297) and the optimiser (cppoptlib lbfgssolver.h:43-144, morethuente.h:25-308)
 * --------------------------------------------------------------------------------------------------------- */
typedef struct { int model; double mu, lambda, k, s0[3]; } prob_t;

static double obj_value(const prob_t *P, const double *x) {
	double d0 = x[0] - P->s0[0], d1 = x[1] - P->s0[1], d2 = x[2] - P->s0[2];
	if (P->model == 3) { /* FungProx::value TriangleForce.cpp:120-146 (b = 1); two unknowns, x[2] is an inert 0 */
		double s3, I1, t1, t2, r0;
		if (x[0] <= 0.0 || x[1] <= 0.0) return FLT_MAX_D;
		s3 = 1.0 / (x[0] * x[1]);
		I1 = x[0] * x[0] + x[1] * x[1] + s3 * s3;
		t1 = P->mu / (1.0 * 2.0);
		t2 = exp(1.0 * (I1 - 3.0)) - 1.0;
		r0 = isfinite(t2) ? (t1 * t2) : FLT_MAX_D;
		return r0 + (P->k * 0.5) * (d0 * d0 + d1 * d1);
	}
	if (x[0] < 0.0 || x[1] < 0.0 || x[2] < 0.0) return FLT_MAX_D;
	if (P->model == 1) { /* NHProx::value :228-233 */
		double det = x[0] * x[1] * x[2], I1 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2], l = log(det * det);
		double r = 0.5 * P->mu * (I1 - l - 3.0) + 0.125 * P->lambda * l * l;
		return 1.0 * r + (P->k * 0.5) * (d0 * d0 + d1 * d1 + d2 * d2);
	} else { /* StVKProx::value :269-287 */
		double a = 0.5 * (x[0] * x[0] - 1.0), b = 0.5 * (x[1] * x[1] - 1.0), c = 0.5 * (x[2] * x[2] - 1.0);
		double tr = a + b + c;
		double r = P->mu * (a * a + (b * b + c * c)) + (P->lambda * 0.5 * (tr * tr));
		return r + (P->k * 0.5) * (d0 * d0 + (d1 * d1 + d2 * d2));
	}
}
static void obj_grad(const prob_t *P, const double *x, double *g) {
	int i;
	if (P->model == 3) { /* FungProx::gradient TriangleForce.cpp:148-168 */
		const double minval = 1.17549435082228751e-38; /* numeric_limits<float>::min() */
		double sig3, I1, t1;
		g[2] = 0.0;
		if (fabs(x[0]) < minval || fabs(x[1]) < minval) { g[0] = g[1] = 1.0 * FLT_MAX_D; return; }
		sig3 = 1.0 / (x[0] * x[1]);
		I1 = (x[0] * x[0] + x[1] * x[1] + sig3 * sig3);
		t1 = 0.5 * P->mu * exp(1.0 * (I1 - 3.0));
		g[0] = t1 * (2.0 * x[0] - 2.0 / (x[0] * x[0] * x[0] * x[1] * x[1])) + P->k * (x[0] - P->s0[0]);
		g[1] = t1 * (2.0 * x[1] - 2.0 / (x[1] * x[1] * x[1] * x[0] * x[0])) + P->k * (x[1] - P->s0[1]);
		return;
	}
	if (P->model == 1) { /* NHProx::gradient :235-243 */
		double det = x[0] * x[1] * x[2];
		if (det <= 0.0) { g[0] = g[1] = g[2] = FLT_MAX_D; return; }
		{
			double ll = P->lambda * log(det);
			for (i = 0; i < 3; ++i) { double inv = 1.0 / x[i]; g[i] = 1.0 * (P->mu * (x[i] - inv) + ll * inv) + P->k * (x[i] - P->s0[i]); }
		}
	} else { /* StVKProx::gradient :289-297 */
		double c2 = 0.5 * P->lambda * ((x[0] * x[0] + x[1] * x[1] + x[2] * x[2]) - 3.0);
		for (i = 0; i < 3; ++i) g[i] = P->mu * x[i] * (x[i] * x[i] - 1.0) + c2 * x[i] + P->k * (x[i] - P->s0[i]);
	}
}
static double dotn(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static double inf3(const double *a) { return fmax(fmax(fabs(a[0]), fabs(a[1])), fabs(a[2])); }
static double mn(double a, double b) { return b < a ? b : a; }
static double mx(double a, double b) { return a < b ? b : a; }

/* morethuente.h:169-308 */
static int cstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy, double *stp, double fp, double dp,
                 int *brackt, double stpmin, double stpmax, int *info) {
	int bound = 0;
	double sgnd, stpf = 0, stpc = 0, stpq = 0, theta, s, gamma, p, q, r;
	*info = 0;
	if ((*brackt & ((*stp <= mn(*stx, *sty)) | (*stp >= mx(*stx, *sty)))) | (*dx * (*stp - *stx) >= 0.0) | (stpmax < stpmin)) return -1;
	sgnd = dp * (*dx / fabs(*dx));
	if (fp > *fx) {
		*info = 1; bound = 1;
		theta = 3. * (*fx - fp) / (*stp - *stx) + *dx + dp;
		s = mx(theta, mx(*dx, dp));
		gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
		if (*stp < *stx) gamma = -gamma;
		p = (gamma - *dx) + theta; q = ((gamma - *dx) + gamma) + dp; r = p / q;
		stpc = *stx + r * (*stp - *stx);
		stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.) * (*stp - *stx);
		stpf = (fabs(stpc - *stx) < fabs(stpq - *stx)) ? stpc : stpc + (stpq - stpc) / 2;
		*brackt = 1;
	} else if (sgnd < 0.0) {
		*info = 2; bound = 0;
		theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
		s = mx(theta, mx(*dx, dp));
		gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
		if (*stp > *stx) gamma = -gamma;
		p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + *dx; r = p / q;
		stpc = *stp + r * (*stx - *stp);
		stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
		stpf = (fabs(stpc - *stp) > fabs(stpq - *stp)) ? stpc : stpq;
		*brackt = 1;
	} else if (fabs(dp) < fabs(*dx)) {
		*info = 3; bound = 1;
		theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
		s = mx(theta, mx(*dx, dp));
		gamma = s * sqrt(mx(0., (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
		if (*stp > *stx) gamma = -gamma;
		p = (gamma - dp) + theta; q = (gamma + (*dx - dp)) + gamma; r = p / q;
		if ((r < 0.0) & (gamma != 0.0)) stpc = *stp + r * (*stx - *stp);
		else if (*stp > *stx) stpc = stpmax;
		else stpc = stpmin;
		stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
		if (*brackt) stpf = (fabs(*stp - stpc) < fabs(*stp - stpq)) ? stpc : stpq;
		else stpf = (fabs(*stp - stpc) > fabs(*stp - stpq)) ? stpc : stpq;
	} else {
		*info = 4; bound = 0;
		if (*brackt) {
			theta = 3 * (fp - *fy) / (*sty - *stp) + *dy + dp;
			s = mx(theta, mx(*dy, dp));
			gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
			if (*stp > *sty) gamma = -gamma;
			p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + *dy; r = p / q;
			stpf = *stp + r * (*sty - *stp);
		} else if (*stp > *stx) stpf = stpmax;
		else stpf = stpmin;
	}
	if (fp > *fx) { *sty = *stp; *fy = fp; *dy = dp; }
	else {
		if (sgnd < 0.0) { *sty = *stx; *fy = *fx; *dy = *dx; }
		*stx = *stp; *fx = fp; *dx = dp;
	}
	stpf = mn(stpmax, stpf);
	stpf = mx(stpmin, stpf);
	*stp = stpf;
	if (*brackt & bound) {
		if (*sty > *stx) *stp = mn(*stx + 0.66 * (*sty - *stx), *stp);
		else *stp = mx(*stx + 0.66 * (*sty - *stx), *stp);
	}
	return 0;
}
/* morethuente.h:25-167 (linesearch + cvsrch); returns the step */
static double linesearch(const prob_t *P, const double *x0, const double *s, double alpha_init) {
	const double xtol = 1e-15, ftol = 1e-4, gtol = 1e-2, stpmin = 1e-15, stpmax = 1e15, xtrapf = 4;
	const int maxfev = 20;
	double stp = alpha_init, f = obj_value(P, x0), g[3], x[3], dginit, finit, dgtest, width, width1;
	double stx = 0.0, fx, dgx, sty = 0.0, fy, dgy, stmin, stmax;
	int info = 0, infoc = 1, nfev = 0, brackt = 0, stage1 = 1, i;
	obj_grad(P, x0, g);
	dginit = dotn(g, s);
	if (dginit >= 0.0) return stp;
	finit = f; dgtest = ftol * dginit; width = stpmax - stpmin; width1 = 2 * width;
	fx = fy = finit; dgx = dgy = dginit;
	for (;;) {
		double dg, ftest1;
		if (brackt) { stmin = mn(stx, sty); stmax = mx(stx, sty); }
		else { stmin = stx; stmax = stp + xtrapf * (stp - stx); }
		stp = mx(stp, stpmin);
		stp = mn(stp, stpmax);
		if ((brackt && ((stp <= stmin) | (stp >= stmax))) | (nfev >= maxfev - 1) | (infoc == 0) | (brackt & (stmax - stmin <= xtol * stmax))) stp = stx;
		for (i = 0; i < 3; ++i) x[i] = x0[i] + stp * s[i];
		f = obj_value(P, x);
		obj_grad(P, x, g);
		nfev++;
		dg = dotn(g, s);
		ftest1 = finit + stp * dgtest;
		if ((brackt & ((stp <= stmin) | (stp >= stmax))) | (infoc == 0)) info = 6;
		if ((stp == stpmax) & (f <= ftest1) & (dg <= dgtest)) info = 5;
		if ((stp == stpmin) & ((f > ftest1) | (dg >= dgtest))) info = 4;
		if (nfev >= maxfev) info = 3;
		if (brackt & (stmax - stmin <= xtol * stmax)) info = 2;
		if ((f <= ftest1) & (fabs(dg) <= gtol * (-dginit))) info = 1;
		if (info != 0) return stp;
		if (stage1 & (f <= ftest1) & (dg >= mn(ftol, gtol) * dginit)) stage1 = 0;
		if (stage1 & (f <= fx) & (f > ftest1)) {
			double fm = f - stp * dgtest, fxm = fx - stx * dgtest, fym = fy - sty * dgtest;
			double dgm = dg - dgtest, dgxm = dgx - dgtest, dgym = dgy - dgtest;
			cstep(&stx, &fxm, &dgxm, &sty, &fym, &dgym, &stp, fm, dgm, &brackt, stmin, stmax, &infoc);
			fx = fxm + stx * dgtest; fy = fym + sty * dgtest; dgx = dgxm + dgtest; dgy = dgym + dgtest;
		} else {
			cstep(&stx, &fx, &dgx, &sty, &fy, &dgy, &stp, f, dg, &brackt, stmin, stmax, &infoc);
		}
		if (brackt) {
			if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
			width1 = width;
			width = fabs(sty - stx);
		}
	}
}
/* lbfgssolver.h:43-144 */
static int lbfgs(const prob_t *P, double *x0, int maxIter, double gradTol, double *init_hess) {
	int m = maxIter < 10 ? maxIter : 10, k, i, j, glob = 0, maxiter = maxIter;
	double s[10][3], y[10][3], alpha[10], rho[10], grad[3], q[3], gold[3], xold[3], gamma_k = *init_hess, alpha_init, new_hess = 1.0;
	memset(s, 0, sizeof(s)); memset(y, 0, sizeof(y)); memset(alpha, 0, sizeof(alpha)); memset(rho, 0, sizeof(rho));
	obj_grad(P, x0, grad);
	alpha_init = mn(1.0, 1.0 / inf3(grad));
	for (k = 0; k < maxiter; k++) {
		int iter;
		double dir, rate, nq[3], st[3], yt[3], dx2;
		for (j = 0; j < 3; ++j) { xold[j] = x0[j]; gold[j] = grad[j]; q[j] = grad[j]; }
		glob++;
		iter = m < k ? m : k;
		for (i = iter - 1; i >= 0; --i) {
			rho[i] = 1.0 / dotn(s[i], y[i]);
			alpha[i] = rho[i] * dotn(s[i], q);
			for (j = 0; j < 3; ++j) q[j] = q[j] - alpha[i] * y[i][j];
		}
		for (j = 0; j < 3; ++j) q[j] = gamma_k * q[j];
		for (i = 0; i < iter; ++i) {
			double beta = rho[i] * dotn(q, y[i]);
			for (j = 0; j < 3; ++j) q[j] = q[j] + (alpha[i] - beta) * s[i][j];
		}
		dir = dotn(q, grad);
		if (dir < 1e-4) {
			for (j = 0; j < 3; ++j) q[j] = grad[j];
			maxiter -= k;
			k = 0;
			alpha_init = mn(1.0, 1.0 / inf3(grad));
		}
		for (j = 0; j < 3; ++j) nq[j] = -q[j];
		rate = linesearch(P, x0, nq, alpha_init);
		dx2 = 0.0;
		for (j = 0; j < 3; ++j) { double d; x0[j] = x0[j] - rate * q[j]; d = xold[j] - x0[j]; dx2 = (j == 0) ? d * d : dx2 + d * d; }
		if (dx2 < 1e-8) break;
		obj_grad(P, x0, grad);
		if (inf3(grad) < gradTol) { new_hess = gamma_k; break; }
		for (j = 0; j < 3; ++j) { st[j] = x0[j] - xold[j]; yt[j] = grad[j] - gold[j]; }
		if (k < m) { for (j = 0; j < 3; ++j) { s[k][j] = st[j]; y[k][j] = yt[j]; } }
		else {
			for (i = 0; i < m - 1; ++i) for (j = 0; j < 3; ++j) { s[i][j] = s[i + 1][j]; y[i][j] = y[i + 1][j]; }
			for (j = 0; j < 3; ++j) { s[m - 1][j] = st[j]; y[m - 1][j] = yt[j]; }
		}
		gamma_k = dotn(st, yt) / dotn(yt, yt);
		alpha_init = 1.0;
	}
	*init_hess = new_hess;
	return glob;
}

/* ---------------------------------------------------------------------------------------------------------
 * Force::project bodies: q = D_i x + u_i in, z out
 * --------------------------------------------------------------------------------------------------------- */
static void blend(const double *p, const double *q, double k, double w, int rows, double *z) { /* (k p + w^2 q)/(w^2 + k) */
	int i;
	for (i = 0; i < rows; ++i) z[i] = (k * p[i] + w * w * q[i]) / (w * w + k);
}
static void project_tet(force_t *f, const double *q, double *z) {
	double F[3][3], U[3][3], V[3][3], S[3], p[9];
	vec_to_mat(q, F);
	jacobi_svd3(F, U, S, V);
	if (f->kind == 0) { /* LinearTetStrain::project TetForce.cpp:127-153 */
		S[0] = S[1] = S[2] = 1.0;
		if (det3(F) < 0.0) S[2] = -1.0;
		usvt(U, S, V, p);
		blend(p, q, f->k, f->w, 9, z);
	} else if (f->kind == 3) { /* TetVolume::project :173-210 */
		double S0[3] = { S[0], S[1], S[2] }, d[3] = { 0, 0, 0 };
		int i;
		for (i = 0; i < 4; i++) {
			double detS = S[0] * S[1] * S[2], fv = detS - mn(mx(detS, f->p1), f->p2);
			double g[3] = { S[1] * S[2], S[0] * S[2], S[0] * S[1] };
			double gd = g[0] * d[0] + (g[1] * d[1] + g[2] * d[2]), gg = g[0] * g[0] + (g[1] * g[1] + g[2] * g[2]);
			double c = -((fv - gd) / gg);
			d[0] = c * g[0]; d[1] = c * g[1]; d[2] = c * g[2];
			S[0] = S0[0] + d[0]; S[1] = S0[1] + d[1]; S[2] = S0[2] + d[2];
		}
		if (det3(F) < 0.0) S[2] = -1.0;
		usvt(U, S, V, p);
		blend(p, q, f->k, f->w, 9, z);
	} else { /* HyperElasticTet::project :320-364 with helper::oriented_svd :80-102 */
		prob_t P;
		double Vt[3][3], x2[3];
		int i, j;
		if (det3(U) < 0.0) { for (i = 0; i < 3; ++i) U[i][2] = -U[i][2]; S[2] *= -1.0; }
		for (i = 0; i < 3; ++i) for (j = 0; j < 3; ++j) Vt[i][j] = V[j][i];
		if (det3(Vt) < 0.0) { for (i = 0; i < 3; ++i) V[i][2] = -V[i][2]; S[2] *= -1.0; }
		P.model = f->kind; P.mu = f->p0; P.lambda = f->p1; P.k = f->k;
		for (i = 0; i < 3; ++i) { P.s0[i] = S[i]; x2[i] = f->prox[i]; }
		if (x2[2] < 0.0) x2[2] *= -1.0;
		else if (fabs(x2[0]) < 1.e-3 && fabs(x2[1]) < 1.e-3 && fabs(x2[2]) < 1.e-3) x2[0] = x2[1] = x2[2] = 1.e-3;
		f->last_iters = lbfgs(&P, x2, f->maxit, 1e-8, &f->init_hess);
		for (i = 0; i < 3; ++i) f->prox[i] = x2[i];
		usvt(U, x2, V, z);
	}
}
/* 3x2: Eigen 3.2.5 JacobiSVD<Matrix<double,3,2>>(F, ComputeFullU | ComputeFullV), step by step: scaling
 * (JacobiSVD.h:839-846), column-pivoting Householder QR preconditioner (QR/ColPivHouseholderQR.h:430-508 with
 * Householder/Householder.h:60-130; U = householderQ, HouseholderSequence.h:236-277; V = column permutation), the 2x2
 * Jacobi step on R (JacobiSVD.h:414-441, Jacobi.h:83-113), sign fix, sort, unscale (:899-929).  U here is its first two
 * columns (all the triangle forces use).  q: column-major 3x2. */
static void svd32(const double *q, double U[3][2], double S[2], double V[2][2]) {
	double A[6], scale = 0.0, n0, n1, tau0, beta0, e00, e01, tau1, beta1, e10, Q[3][3], W[2][2], tmp;
	int i, r, c, swapped, guard = 0;
	for (i = 0; i < 6; ++i) if (fabs(q[i]) > scale) scale = fabs(q[i]);
	if (scale == 0.0) scale = 1.0;
	for (i = 0; i < 6; ++i) A[i] = q[i] / scale;
	n0 = A[0] * A[0] + (A[1] * A[1] + A[2] * A[2]);     /* fixed-size column norms: a0 + (a1 + a2) */
	n1 = A[3] * A[3] + (A[4] * A[4] + A[5] * A[5]);
	swapped = n1 > n0;                                  /* first maximum wins */
	if (swapped) for (r = 0; r < 3; ++r) { double t = A[r]; A[r] = A[3 + r]; A[3 + r] = t; }
	{ /* k = 0 */
		double c0 = A[0], tail = A[1] * A[1] + A[2] * A[2];
		if (tail == 0.0) { tau0 = 0.0; beta0 = c0; e00 = e01 = 0.0; }
		else { beta0 = sqrt(c0 * c0 + tail); if (c0 >= 0.0) beta0 = -beta0; e00 = A[1] / (c0 - beta0); e01 = A[2] / (c0 - beta0); tau0 = (beta0 - c0) / beta0; }
		tmp = e00 * A[4] + e01 * A[5]; tmp += A[3];
		A[3] -= tau0 * tmp; A[4] -= (tau0 * e00) * tmp; A[5] -= (tau0 * e01) * tmp;
	}
	{ /* k = 1 */
		double c0 = A[4], tail = A[5] * A[5];
		if (tail == 0.0) { tau1 = 0.0; beta1 = c0; e10 = 0.0; }
		else { beta1 = sqrt(c0 * c0 + tail); if (c0 >= 0.0) beta1 = -beta1; e10 = A[5] / (c0 - beta1); tau1 = (beta1 - c0) / beta1; }
	}
	for (r = 0; r < 3; ++r) for (c = 0; c < 3; ++c) Q[r][c] = (r == c);
	for (c = 1; c < 3; ++c) { tmp = e10 * Q[2][c]; tmp += Q[1][c]; Q[1][c] -= tau1 * tmp; Q[2][c] -= (tau1 * e10) * tmp; }
	for (c = 0; c < 3; ++c) { tmp = e00 * Q[1][c] + e01 * Q[2][c]; tmp += Q[0][c]; Q[0][c] -= tau0 * tmp; Q[1][c] -= (tau0 * e00) * tmp; Q[2][c] -= (tau0 * e01) * tmp; }
	W[0][0] = beta0; W[1][0] = 0.0; W[0][1] = A[3]; W[1][1] = beta1;
	V[0][0] = swapped ? 0.0 : 1.0; V[1][0] = swapped ? 1.0 : 0.0; V[0][1] = swapped ? 1.0 : 0.0; V[1][1] = swapped ? 0.0 : 1.0;
	while (guard++ < 64) { /* pair (p, q) = (1, 0) */
		double thr = fmax(2.0 * 4.9406564584124654e-324, 2.0 * DBL_EPSILON * fmax(fabs(W[1][1]), fabs(W[0][0])));
		double m00, m01, m10, m11, t, d, c1, s1, cr, sr, cl, sl, x, y;
		if (!(fabs(W[1][0]) > thr || fabs(W[0][1]) > thr)) break;
		m00 = W[1][1]; m01 = W[1][0]; m10 = W[0][1]; m11 = W[0][0];
		t = m00 + m11; d = m10 - m01;
		if (t == 0.0) { c1 = 0.0; s1 = d > 0.0 ? 1.0 : -1.0; }
		else { double h = hyp(t, d); c1 = fabs(t) / h; s1 = d / h; if (t < 0.0) s1 = -s1; }
		if (!(c1 == 1.0 && s1 == 0.0)) {
			x = m00; y = m10; m00 = c1 * x + s1 * y; m10 = -s1 * x + c1 * y;
			x = m01; y = m11; m01 = c1 * x + s1 * y; m11 = -s1 * x + c1 * y;
		}
		if (m01 == 0.0) { cr = 1.0; sr = 0.0; }
		else {
			double tau = (m00 - m11) / (2.0 * fabs(m01)), w = sqrt(tau * tau + 1.0), tt, n;
			tt = (tau > 0.0) ? 1.0 / (tau + w) : 1.0 / (tau - w);
			n = 1.0 / sqrt(tt * tt + 1.0);
			sr = -(tt > 0.0 ? 1.0 : -1.0) * (m01 / fabs(m01)) * fabs(tt) * n;
			cr = n;
		}
		cl = c1 * cr - s1 * (-sr);
		sl = c1 * (-sr) + s1 * cr;
		if (!(cl == 1.0 && sl == 0.0)) { /* rows 1, 0 of W; columns 1, 0 of Q */
			for (c = 0; c < 2; ++c) { x = W[1][c]; y = W[0][c]; W[1][c] = cl * x + sl * y; W[0][c] = -sl * x + cl * y; }
			for (r = 0; r < 3; ++r) { x = Q[r][1]; y = Q[r][0]; Q[r][1] = cl * x + sl * y; Q[r][0] = -sl * x + cl * y; }
		}
		if (!(cr == 1.0 && -sr == 0.0)) { /* columns 1, 0 of W and V with (c, s) = (cr, -sr) */
			for (r = 0; r < 2; ++r) { x = W[r][1]; y = W[r][0]; W[r][1] = cr * x + (-sr) * y; W[r][0] = -(-sr) * x + cr * y; }
			for (r = 0; r < 2; ++r) { x = V[r][1]; y = V[r][0]; V[r][1] = cr * x + (-sr) * y; V[r][0] = -(-sr) * x + cr * y; }
		}
	}
	for (i = 0; i < 2; ++i) {
		double a = fabs(W[i][i]);
		S[i] = a;
		if (a != 0.0) { double f = W[i][i] / a; for (r = 0; r < 3; ++r) Q[r][i] *= f; }
	}
	if (S[1] > S[0]) {
		double t = S[0]; S[0] = S[1]; S[1] = t;
		for (r = 0; r < 3; ++r) { t = Q[r][0]; Q[r][0] = Q[r][1]; Q[r][1] = t; }
		for (r = 0; r < 2; ++r) { t = V[r][0]; V[r][0] = V[r][1]; V[r][1] = t; }
	}
	S[0] *= scale; S[1] *= scale;
	for (r = 0; r < 3; ++r) { U[r][0] = Q[r][0]; U[r][1] = Q[r][1]; }
}
static void project_tri(force_t *f, const double *q, double *z) {
	double U[3][2], S[2], V[2][2], p[6];
	int r, c;
	svd32(q, U, S, V);
	if (f->kind == 0) { /* LimitedTriangleStrain::project TriangleForce.cpp:79-113 */
		for (c = 0; c < 2; ++c) for (r = 0; r < 3; ++r) p[3 * c + r] = U[r][0] * V[c][0] + U[r][1] * V[c][1];
		blend(p, q, f->k, f->w, 6, z);
		if (f->flag) {
			double l0 = sqrt(z[0] * z[0] + (z[1] * z[1] + z[2] * z[2])), l1 = sqrt(z[3] * z[3] + (z[4] * z[4] + z[5] * z[5]));
			double f0 = (double)fmaxf((float)l0, 1e-6f), f1 = (double)fmaxf((float)l1, 1e-6f), sc;
			if (l0 < f->p1) { sc = f->p1 / f0; z[0] *= sc; z[1] *= sc; z[2] *= sc; }
			if (l1 < f->p1) { sc = f->p1 / f1; z[3] *= sc; z[4] *= sc; z[5] *= sc; }
			if (l0 > f->p2) { sc = f->p2 / f0; z[0] *= sc; z[1] *= sc; z[2] *= sc; }
			if (l1 > f->p2) { sc = f->p2 / f1; z[3] *= sc; z[4] *= sc; z[5] *= sc; }
		}
	} else if (f->kind == 2) { /* FungTriangle::project TriangleForce.cpp:217-248: L-BFGS on the two singular values */
		prob_t P;
		double x2[3] = { S[0], S[1], 0.0 };
		P.model = 3; P.mu = f->p0; P.lambda = 0.0; P.k = f->p0; P.s0[0] = S[0]; P.s0[1] = S[1]; P.s0[2] = 0.0;
		f->last_iters = lbfgs(&P, x2, 10, 1e-6, &f->init_hess);
		for (c = 0; c < 2; ++c) for (r = 0; r < 3; ++r) z[3 * c + r] = (U[r][0] * x2[0]) * V[c][0] + (U[r][1] * x2[1]) * V[c][1];
	} else { /* TriArea::project :257-295 */
		double S0[2] = { S[0], S[1] }, d[2] = { 0, 0 };
		int i;
		for (i = 0; i < f->flag; ++i) {
			double v = S[0] * S[1], cl = v < f->p2 ? v : f->p2, fv, g0 = S[1], g1 = S[0], cc;
			cl = cl > f->p1 ? cl : f->p1;
			fv = v - cl;
			cc = -((fv - (g0 * d[0] + g1 * d[1])) / (g0 * g0 + g1 * g1));
			d[0] = cc * g0; d[1] = cc * g1;
			S[0] = S0[0] + d[0]; S[1] = S0[1] + d[1];
		}
		for (c = 0; c < 2; ++c) for (r = 0; r < 3; ++r) p[3 * c + r] = (U[r][0] * S[0]) * V[c][0] + (U[r][1] * S[1]) * V[c][1];
		blend(p, q, f->k, f->w, 6, z);
	}
}
static void project_spring(force_t *f, const double *q, double *z) { /* Spring::project Force.cpp:52-71 */
	double nrm = sqrt(q[0] * q[0] + (q[1] * q[1] + q[2] * q[2])), inv = 1.0 / (f->w * f->w + f->k);
	int i;
	for (i = 0; i < 3; ++i) { double nd = (nrm <= 0.0) ? 0.0 : q[i] / nrm; z[i] = inv * (f->k * (f->aux[0] * nd) + f->w * f->w * q[i]); }
}
static void project_bend(force_t *f, const double *q, double *z) { /* BendForce::project BendForce.cpp:134-161 */
	const double *al = f->aux;
	double den = al[0] * al[0] + al[3] * al[3] + al[1] * al[1], p[9], inv = 1.0 / (f->w * f->w + f->k);
	int j, i;
	for (j = 0; j < 3; ++j) {
		double lam = 2.0 * (al[0] * q[j] + al[3] * q[3 + j] + al[1] * q[6 + j]) / den;
		p[j] = q[j] - 0.5 * al[0] * lam; p[3 + j] = q[3 + j] - 0.5 * al[3] * lam; p[6 + j] = q[6 + j] - 0.5 * al[1] * lam;
	}
	for (i = 0; i < 9; ++i) z[i] = inv * (f->k * p[i] + f->w * f->w * q[i]);
}
static void collide(const oracle_sys *S, double *p) { /* CollisionForce::handleCollisions CollisionForce.cpp:53-70 */
	int j;
	for (j = 0; j < S->nshapes; ++j) {
		const double *s = S->shape_par + 4 * j;
		if (S->shape_kind[j] == 0) { /* CollisionSphere.hpp:47-62 */
			double d0 = p[0] - s[0], d1 = p[1] - s[1], d2 = p[2] - s[2], n = sqrt(d0 * d0 + (d1 * d1 + d2 * d2));
			if (s[3] - n > 0) { p[0] = s[0] + s[3] * (d0 / n); p[1] = s[1] + s[3] * (d1 / n); p[2] = s[2] + s[3] * (d2 / n); }
		} else if (S->shape_kind[j] == 1) { /* CollisionCylinder.hpp:45-65 (axis || z, centre z = 0) */
			double d0 = p[0] - s[0], d1 = p[1] - s[1], n = sqrt(d0 * d0 + (d1 * d1 + 0.0));
			if (s[3] - n > 0) { double pz = p[2]; p[0] = s[0] + s[3] * (d0 / n) + 0.0; p[1] = s[1] + s[3] * (d1 / n) + 0.0; p[2] = 0.0 + s[3] * (0.0 / n) + pz; }
		} else if (s[1] - p[1] > 0) p[1] = s[1]; /* CollisionFloor.hpp:47-55 */
	}
}

/* ---------------------------------------------------------------------------------------------------------
 * System
 * --------------------------------------------------------------------------------------------------------- */
oracle_sys *oracle_create(int n, const double *x3n, const double *m_node, double dt) {
	oracle_sys *S = (oracle_sys *)calloc(1, sizeof(oracle_sys));
	S->n = n; S->dt = dt;
	S->x0 = (double *)malloc(sizeof(double) * 3 * n); memcpy(S->x0, x3n, sizeof(double) * 3 * n);
	S->m = (double *)malloc(sizeof(double) * n); memcpy(S->m, m_node, sizeof(double) * n);
	return S;
}
void oracle_destroy(oracle_sys *S) {
	if (!S) return;
	free(S->x0); free(S->m); free(S->f); free(S->shape_kind); free(S->shape_par); free(S->wind_tris); free(S->L); free(S->u); free(S->z);
	free(S);
}
static force_t *new_force(oracle_sys *S) {
	if (S->nf == S->cap) { S->cap = S->cap ? 2 * S->cap : 64; S->f = (force_t *)realloc(S->f, sizeof(force_t) * S->cap); }
	memset(&S->f[S->nf], 0, sizeof(force_t));
	return &S->f[S->nf++];
}
static void sub3(const double *a, const double *b, double *o) { o[0] = a[0] - b[0]; o[1] = a[1] - b[1]; o[2] = a[2] - b[2]; }
static void cross3(const double *a, const double *b, double *o) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; }
static double fdot(const double *a, const double *b) { return a[0] * b[0] + (a[1] * b[1] + a[2] * b[2]); }

int oracle_add_tets(oracle_sys *S, int kind, int count, const int *idx, double p0, double p1, double p2, int maxit) {
	int e, r, c;
	for (e = 0; e < count; ++e) { /* helper::init_tet_force TetForce.cpp:28-57 */
		force_t *f = new_force(S);
		const double *v[4];
		double E[3][3], inv[3][3], e0[3], e1[3], e2[3], a[3], b[3], cc[3], cr[3], det, vol, stiff;
		f->type = F_TET; f->kind = kind; f->rows = 9; f->nv = 4; f->p0 = p0; f->p1 = p1; f->p2 = p2; f->maxit = maxit;
		for (c = 0; c < 4; ++c) { f->idx[c] = idx[4 * e + c]; v[c] = S->x0 + 3 * f->idx[c]; }
		sub3(v[1], v[0], e0); sub3(v[2], v[0], e1); sub3(v[3], v[0], e2);
		for (r = 0; r < 3; ++r) { E[r][0] = e0[r]; E[r][1] = e1[r]; E[r][2] = e2[r]; }
		/* Matrix3d::inverse, cofactor form (Eigen/src/LU/Inverse.h:118-160) */
		{
			double c00 = E[1][1] * E[2][2] - E[1][2] * E[2][1], c10 = E[2][1] * E[0][2] - E[2][2] * E[0][1], c20 = E[0][1] * E[1][2] - E[0][2] * E[1][1];
			double id;
			det = c00 * E[0][0] + (c10 * E[1][0] + c20 * E[2][0]);
			id = 1.0 / det;
			inv[0][0] = c00 * id; inv[0][1] = c10 * id; inv[0][2] = c20 * id;
			inv[1][0] = (E[1][2] * E[2][0] - E[1][0] * E[2][2]) * id; inv[1][1] = (E[2][2] * E[0][0] - E[2][0] * E[0][2]) * id; inv[1][2] = (E[0][2] * E[1][0] - E[0][0] * E[1][2]) * id;
			inv[2][0] = (E[1][0] * E[2][1] - E[1][1] * E[2][0]) * id; inv[2][1] = (E[2][0] * E[0][1] - E[2][1] * E[0][0]) * id; inv[2][2] = (E[0][0] * E[1][1] - E[0][1] * E[1][0]) * id;
		}
		for (r = 0; r < 3; ++r) { /* B = D * Xg^-1, D = [-1 -1 -1; I] */
			f->B[0 * 3 + r] = (-1.0 * inv[0][r] + -1.0 * inv[1][r]) + -1.0 * inv[2][r];
			f->B[1 * 3 + r] = inv[0][r]; f->B[2 * 3 + r] = inv[1][r]; f->B[3 * 3 + r] = inv[2][r];
		}
		sub3(v[0], v[3], a); sub3(v[1], v[3], b); sub3(v[2], v[3], cc); cross3(b, cc, cr);
		vol = fabs(fdot(a, cr)) / 6.0;
		stiff = (kind == 1 || kind == 2) ? mn(p0, p1) : p0;       /* TetForce.cpp:306 / :116,162 */
		f->w = (double)(sqrtf((float)stiff) * sqrtf((float)vol));
		f->k = (kind == 1 || kind == 2) ? stiff : stiff * vol;
		f->prox[0] = f->prox[1] = f->prox[2] = 1.0; f->init_hess = 1.0;
	}
	return 0;
}
int oracle_add_tris(oracle_sys *S, int kind, int count, const int *idx, double stiffness, double lmin, double lmax, int flag) {
	int e, c;
	for (e = 0; e < count; ++e) { /* LimitedTriangleStrain::initialize TriangleForce.cpp:29-63 */
		force_t *f = new_force(S);
		const double *x1, *x2, *x3;
		double e12[3], e13[3], n1[3], t[3], n2[3], l, pr, g00, g01, g10, g11, det, id, i00, i10, i01, i11, area;
		f->type = F_TRI; f->kind = kind; f->rows = 6; f->nv = 3; f->p0 = stiffness; f->p1 = lmin; f->p2 = lmax; f->flag = flag;
		for (c = 0; c < 3; ++c) f->idx[c] = idx[3 * e + c];
		x1 = S->x0 + 3 * f->idx[0]; x2 = S->x0 + 3 * f->idx[1]; x3 = S->x0 + 3 * f->idx[2];
		sub3(x2, x1, e12); sub3(x3, x1, e13);
		l = sqrt(fdot(e12, e12)); for (c = 0; c < 3; ++c) n1[c] = e12[c] / l;
		pr = fdot(e13, n1); for (c = 0; c < 3; ++c) t[c] = e13[c] - pr * n1[c];
		l = sqrt(fdot(t, t)); for (c = 0; c < 3; ++c) n2[c] = t[c] / l;
		g00 = dotn(n1, e12); g01 = dotn(n1, e13); g10 = dotn(n2, e12); g11 = dotn(n2, e13);
		det = g00 * g11 - g10 * g01; id = 1.0 / det;
		i00 = g11 * id; i10 = -g10 * id; i01 = -g01 * id; i11 = g00 * id;
		f->B[0] = -1.0 * i00 + -1.0 * i10; f->B[1] = -1.0 * i01 + -1.0 * i11; f->B[2] = i00; f->B[3] = i01; f->B[4] = i10; f->B[5] = i11;
		area = fabs(det / 2.0);
		f->w = (double)(sqrtf((float)stiffness) * sqrtf((float)area));
		f->k = stiffness * area;
		if (kind == 2) { f->w = sqrt(stiffness) * sqrt(area); f->k = stiffness; f->init_hess = 1.0; } /* FungTriangle::initialize :190-196 */
	}
	return 0;
}
int oracle_add_springs(oracle_sys *S, int count, const int *idx, const double *stiffness) {
	int e;
	for (e = 0; e < count; ++e) { /* Spring::initialize Force.cpp:29-38 */
		force_t *f = new_force(S);
		double d[3];
		f->type = F_SPRING; f->rows = 3; f->nv = 2; f->idx[0] = idx[2 * e]; f->idx[1] = idx[2 * e + 1];
		sub3(S->x0 + 3 * f->idx[0], S->x0 + 3 * f->idx[1], d);
		f->aux[0] = sqrt(fdot(d, d)); f->k = stiffness[e]; f->w = sqrt(stiffness[e]);
	}
	return 0;
}
int oracle_add_bends(oracle_sys *S, int count, const int *idx, double stiffness) {
	int e, c;
	for (e = 0; e < count; ++e) { /* BendForce::initialize BendForce.cpp:26-55 */
		force_t *f = new_force(S);
		double xA[3], xB[3], xD[3], zero[3] = { 0, 0, 0 }, t0[3], t1[3], nC[3], nD[3], cr[3], a1, a2, hA, hB, lD, lC, lDn;
		f->type = F_BEND; f->rows = 9; f->nv = 4;
		for (c = 0; c < 4; ++c) f->idx[c] = idx[4 * e + c];
		sub3(S->x0 + 3 * f->idx[0], S->x0 + 3 * f->idx[2], xA); sub3(S->x0 + 3 * f->idx[1], S->x0 + 3 * f->idx[2], xB); sub3(S->x0 + 3 * f->idx[3], S->x0 + 3 * f->idx[2], xD);
		cross3(xA, xD, cr); a1 = 0.5 * sqrt(fdot(cr, cr));
		cross3(xD, xB, cr); a2 = 0.5 * sqrt(fdot(cr, cr));
		lD = sqrt(fdot(xD, xD)); hA = 2.0 * a1 / lD; hB = 2.0 * a2 / lD;
		sub3(zero, xB, t0); sub3(zero, xA, t1); cross3(t0, t1, nC);
		sub3(xD, xA, t0); sub3(xD, xB, t1); cross3(t0, t1, nD);
		lC = sqrt(fdot(nC, nC)); lDn = sqrt(fdot(nD, nD));
		f->aux[0] = hB / (hA + hB); f->aux[1] = hA / (hA + hB); f->aux[2] = -lDn / (lC + lDn); f->aux[3] = -lC / (lC + lDn);
		f->k = stiffness; f->w = sqrt(stiffness);
	}
	return 0;
}
int oracle_add_anchors(oracle_sys *S, int moving, int count, const int *idx, const double *pos, double weight) {
	int e, j;
	for (e = 0; e < count; ++e) { /* AnchorForce.hpp:55-106, AnchorForce.cpp:29-43 */
		force_t *f = new_force(S);
		f->type = moving ? F_MANCHOR : F_SANCHOR; f->rows = 3; f->nv = 1; f->idx[0] = idx[e]; f->active = 1;
		for (j = 0; j < 3; ++j) f->aux[j] = moving ? pos[3 * e + j] : S->x0[3 * idx[e] + j];
		f->w = weight > 0.0 ? weight : 1000.0;
	}
	return S->nf - count; /* index of the first anchor force */
}
int oracle_set_anchor(oracle_sys *S, int force, const double *pos, int active) {
	if (pos) memcpy(S->f[force].aux, pos, 3 * sizeof(double));
	if (active >= 0) S->f[force].active = active;
	return 0;
}
void oracle_get_anchor(oracle_sys *S, int force, double *pos) { memcpy(pos, S->f[force].aux, 3 * sizeof(double)); }
void oracle_set_weight(oracle_sys *S, int force, double w) { S->f[force].w = w; }
int oracle_add_collision(oracle_sys *S, int nshapes, const int *kind, const double *par4, double weight) {
	int i;
	S->nshapes = nshapes; S->coll_w = weight; S->has_coll = 1;
	S->shape_kind = (int *)malloc(sizeof(int) * (nshapes + 1)); S->shape_par = (double *)malloc(sizeof(double) * 4 * (nshapes + 1));
	memcpy(S->shape_kind, kind, sizeof(int) * nshapes); memcpy(S->shape_par, par4, sizeof(double) * 4 * nshapes);
	for (i = 0; i < nshapes; ++i) if (kind[i] == 1) S->shape_par[4 * i + 2] = 0.0;
	return 0;
}
void oracle_add_gravity(oracle_sys *S, const double *d) { memcpy(S->grav[S->ngrav++], d, 3 * sizeof(double)); if (S->wind_tris) S->wind_after_gravity = 0; }
void oracle_add_wind(oracle_sys *S, int nt, const int *tris, const double *dir) {
	S->wind_tris = (int *)malloc(sizeof(int) * 3 * nt); memcpy(S->wind_tris, tris, sizeof(int) * 3 * nt);
	S->wind_nt = nt; memcpy(S->wind_dir, dir, 3 * sizeof(double)); S->wind_after_gravity = 1;
}

/* scalar selector of one force: sel[r][c], returns #scalar rows (rows/3) */
static int selector(const force_t *f, double sel[3][4]) {
	int r, c;
	memset(sel, 0, sizeof(double) * 12);
	switch (f->type) {
	case F_TET: for (r = 0; r < 3; ++r) for (c = 0; c < 4; ++c) sel[r][c] = f->B[3 * c + r]; return 3;    /* TetForce.cpp:59-77 */
	case F_TRI: for (r = 0; r < 2; ++r) for (c = 0; c < 3; ++c) sel[r][c] = f->B[2 * c + r]; return 2;    /* TriangleForce.cpp:65-76 */
	case F_SPRING: sel[0][0] = 1.0; sel[0][1] = -1.0; return 1;                                             /* Force.cpp:40-50 */
	case F_BEND: sel[0][0] = 1; sel[0][2] = -1; sel[1][3] = 1; sel[1][2] = -1; sel[2][1] = 1; sel[2][2] = -1; return 3; /* BendForce.cpp:74-131 */
	default: sel[0][0] = 1.0; return 1;                                                                       /* AnchorForce.cpp:36-43 */
	}
}
/* System::initialize / recompute_weights (System.cpp:98-179) on the scalar n x n matrix (A = A_n (x) I_3) */
int oracle_initialize(oracle_sys *S) {
	int n = S->n, i, j, k, e, a, b, r;
	double dt2 = S->dt * S->dt, *A = (double *)calloc((size_t)n * n, sizeof(double));
	long row = 0;
	for (i = 0; i < n; ++i) A[(size_t)i * n + i] = S->m[i];
	for (e = 0; e < S->nf; ++e) {
		force_t *f = &S->f[e];
		double sel[3][4], c = dt2 * f->w * f->w;
		int nr = selector(f, sel);
		f->row = row; row += f->rows;
		for (a = 0; a < f->nv; ++a) for (b = 0; b < f->nv; ++b) {
			double s = 0.0;
			for (r = 0; r < nr; ++r) s += sel[r][a] * sel[r][b];
			A[(size_t)f->idx[a] * n + f->idx[b]] += c * s;
		}
	}
	if (S->has_coll) { S->coll_row = row; row += 3 * n; for (i = 0; i < n; ++i) A[(size_t)i * n + i] += dt2 * S->coll_w * S->coll_w; }
	if (row != S->rows) { free(S->u); free(S->z); S->u = (double *)calloc(row, sizeof(double)); S->z = (double *)calloc(row, sizeof(double)); S->rows = row; }
	/* dense Cholesky A = L L^T (the reference factors with Eigen::SimplicialLDLT, System.cpp:140) */
	for (j = 0; j < n; ++j) {
		double d = A[(size_t)j * n + j];
		for (k = 0; k < j; ++k) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k];
		if (!(d > 0.0)) { free(A); return -1; }
		d = sqrt(d);
		A[(size_t)j * n + j] = d;
		for (i = j + 1; i < n; ++i) {
			double s = A[(size_t)i * n + j];
			for (k = 0; k < j; ++k) s -= A[(size_t)i * n + k] * A[(size_t)j * n + k];
			A[(size_t)i * n + j] = s / d;
		}
	}
	free(S->L);
	S->L = A;
	return 0;
}
long oracle_rows(const oracle_sys *S) { return S->rows; }
int oracle_num_hyper(const oracle_sys *S) { int e, c = 0; for (e = 0; e < S->nf; ++e) if (S->f[e].type == F_TET && (S->f[e].kind == 1 || S->f[e].kind == 2)) ++c; return c; }
void oracle_get_prox(const oracle_sys *S, double *out4, int *iters) {
	int e, c = 0;
	for (e = 0; e < S->nf; ++e) if (S->f[e].type == F_TET && (S->f[e].kind == 1 || S->f[e].kind == 2)) {
		if (out4) { memcpy(out4 + 4 * c, S->f[e].prox, 3 * sizeof(double)); out4[4 * c + 3] = S->f[e].init_hess; }
		if (iters) iters[c] = S->f[e].last_iters;
		++c;
	}
}

static void wind(const oracle_sys *S, const double *x, double *v) { /* WindForce::project ExplicitForce.cpp:42-98, serial order */
	int t, j, k;
	for (t = 0; t < S->wind_nt; ++t) {
		int id[3] = { 3 * S->wind_tris[3 * t], 3 * S->wind_tris[3 * t + 1], 3 * S->wind_tris[3 * t + 2] };
		double cv[3], vr[3], a[3], b[3], nrm[3], len, un[3], area, vn, fo[3];
		for (j = 0; j < 3; ++j) { cv[j] = (v[id[0] + j] + v[id[1] + j] + v[id[2] + j]) / 3.0; vr[j] = cv[j] - S->wind_dir[j]; }
		sub3(x + id[1], x + id[0], a); sub3(x + id[2], x + id[0], b); cross3(a, b, nrm);
		len = sqrt(fdot(nrm, nrm));
		for (j = 0; j < 3; ++j) un[j] = nrm[j] / len;
		area = 0.5 * len; vn = fdot(un, vr);
		for (j = 0; j < 3; ++j) { fo[j] = -1000.0 * area * vn * fabs(vn) * un[j]; fo[j] *= 0.33; fo[j] *= S->dt; }
		for (k = 0; k < 3; ++k) for (j = 0; j < 3; ++j) v[id[k] + j] += fo[j];
	}
}

/* The local step of one ADMM iteration (System.cpp:54-58) and, if b != NULL, this iteration's share of the right-hand
 * side (System.cpp:61).  Dx sums a row's terms in ascending node index: Eigen's column-major m_D * curr_x visits the
 * columns in that order (SparseDenseProduct.h:190-209). */
static void local_pass(oracle_sys *S, const double *cx, double *b) {
	int n = S->n, e, i, j, k, c, r;
	double dt2 = S->dt * S->dt;
	for (e = 0; e < S->nf; ++e) {
		force_t *f = &S->f[e];
		double sel[3][4], Dx[9], q[9], z[9], *u = S->u + f->row, cw = dt2 * f->w * f->w;
		int nr = selector(f, sel), ord[4], a, t;
		for (c = 0; c < f->nv; ++c) ord[c] = c;
		for (a = 1; a < f->nv; ++a) { t = ord[a]; c = a - 1; while (c >= 0 && f->idx[ord[c]] > f->idx[t]) { ord[c + 1] = ord[c]; --c; } ord[c + 1] = t; }
		for (r = 0; r < nr; ++r) for (j = 0; j < 3; ++j) {
			double s = 0.0;
			int first = 1;
			for (a = 0; a < f->nv; ++a) {
				c = ord[a];
				if (sel[r][c] == 0.0) continue;           /* structural zeros are not stored in m_D */
				if (first) { s = sel[r][c] * cx[3 * f->idx[c] + j]; first = 0; }
				else s += sel[r][c] * cx[3 * f->idx[c] + j];
			}
			Dx[3 * r + j] = s;
		}
		for (k = 0; k < f->rows; ++k) q[k] = Dx[k] + u[k];
		switch (f->type) {
		case F_TET: project_tet(f, q, z); break;
		case F_TRI: project_tri(f, q, z); break;
		case F_SPRING: project_spring(f, q, z); break;
		case F_BEND: project_bend(f, q, z); break;
		case F_SANCHOR: memcpy(z, f->aux, 3 * sizeof(double)); break;                 /* AnchorForce.cpp:46-55 */
		default: /* MovingAnchor::project AnchorForce.cpp:71-89 */
			if (f->active) memcpy(z, f->aux, 3 * sizeof(double));
			else { for (k = 0; k < 3; ++k) { z[k] = q[k]; f->aux[k] = Dx[k]; } }
		}
		for (k = 0; k < f->rows; ++k) { u[k] = u[k] + (Dx[k] - z[k]); S->z[f->row + k] = z[k]; }
		if (b) for (r = 0; r < nr; ++r) for (j = 0; j < 3; ++j) { /* b += dt^2 D^T W^2 (z - u), System.cpp:61 */
			double zu = z[3 * r + j] - u[3 * r + j];
			for (c = 0; c < f->nv; ++c) b[3 * f->idx[c] + j] += cw * sel[r][c] * zu;
		}
	}
	if (S->has_coll) { /* CollisionForce::project CollisionForce.cpp:36-46 */
		double cw = dt2 * S->coll_w * S->coll_w;
		for (i = 0; i < n; ++i) {
			double *u = S->u + S->coll_row + 3 * i, p[3];
			for (j = 0; j < 3; ++j) p[j] = cx[3 * i + j] + u[j];
			collide(S, p);
			for (j = 0; j < 3; ++j) { u[j] = u[j] + (cx[3 * i + j] - p[j]); S->z[S->coll_row + 3 * i + j] = p[j]; if (b) b[3 * i + j] += cw * (p[j] - u[j]); }
		}
	}
}

/* Teacher-forced half iteration for the parity tests: u (and the hyperelastic optimiser state, 4 per tet, may be NULL) are
 * set, the local step runs on curr_x = x, and z, u and the new optimiser state are returned. */
int oracle_local_step(oracle_sys *S, const double *x, const double *u_in, const double *prox_in, double *z_out, double *u_out, double *prox_out) {
	int e, c = 0;
	memcpy(S->u, u_in, sizeof(double) * S->rows);
	if (prox_in) for (e = 0; e < S->nf; ++e) if (S->f[e].type == F_TET && (S->f[e].kind == 1 || S->f[e].kind == 2)) {
		memcpy(S->f[e].prox, prox_in + 4 * c, 3 * sizeof(double)); S->f[e].init_hess = prox_in[4 * c + 3]; ++c;
	}
	local_pass(S, x, 0);
	memcpy(z_out, S->z, sizeof(double) * S->rows);
	memcpy(u_out, S->u, sizeof(double) * S->rows);
	if (prox_out) oracle_get_prox(S, prox_out, 0);
	return 0;
}

/* System::step (System.cpp:26-75).  x_it/z_it/u_it (optional): per-iteration dumps in the layout of oracle/ref_shim.cpp */
int oracle_step(oracle_sys *S, int iters, double *x, double *v, double *x_it, double *z_it, double *u_it, double *prox_it) {
	int n = S->n, n3 = 3 * n, it, i, j, k, g;
	double dt = S->dt, *xbar = (double *)malloc(sizeof(double) * n3), *cx = (double *)malloc(sizeof(double) * n3), *b = (double *)malloc(sizeof(double) * n3);
	int nh = oracle_num_hyper(S);
	if (S->wind_tris && !S->wind_after_gravity) wind(S, x, v);
	for (g = 0; g < S->ngrav; ++g) for (i = 0; i < n; ++i) for (j = 0; j < 3; ++j) v[3 * i + j] += (dt * S->grav[g][j]);
	if (S->wind_tris && S->wind_after_gravity) wind(S, x, v);
	for (i = 0; i < n3; ++i) { xbar[i] = x[i] + dt * v[i]; cx[i] = xbar[i]; }
	for (it = 0; it < iters; ++it) {
		if (x_it) memcpy(x_it + (size_t)it * n3, cx, sizeof(double) * n3);
		for (i = 0; i < n3; ++i) b[i] = S->m[i / 3] * xbar[i];
		local_pass(S, cx, b);
		if (z_it) memcpy(z_it + (size_t)it * S->rows, S->z, sizeof(double) * S->rows);
		if (u_it) memcpy(u_it + (size_t)it * S->rows, S->u, sizeof(double) * S->rows);
		if (prox_it && nh) oracle_get_prox(S, prox_it + (size_t)it * nh * 4, 0);
		/* curr_x = A^-1 b: forward / backward substitution with the dense factor, three columns */
		for (j = 0; j < 3; ++j) {
			for (i = 0; i < n; ++i) { double s = b[3 * i + j]; for (k = 0; k < i; ++k) s -= S->L[(size_t)i * n + k] * cx[3 * k + j]; cx[3 * i + j] = s / S->L[(size_t)i * n + i]; }
			for (i = n - 1; i >= 0; --i) { double s = cx[3 * i + j]; for (k = i + 1; k < n; ++k) s -= S->L[(size_t)k * n + i] * cx[3 * k + j]; cx[3 * i + j] = s / S->L[(size_t)i * n + i]; }
		}
	}
	for (i = 0; i < n3; ++i) { v[i] = (cx[i] - x[i]) * (1.0 / dt); x[i] = cx[i]; }
	free(xbar); free(cx); free(b);
	return 0;
}
