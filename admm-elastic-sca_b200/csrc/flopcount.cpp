// flopcount.cpp -- ALGORITHMIC FP64 flop counter of the hyperelastic local step ("counted by an instrumented
// host restatement on the actual run", because the L-BFGS / line-search trip counts are data dependent).
//
// Built into its own library (libadmm_b200_flopcount.so), never into libadmm_b200.so: it compiles the very per-force
// body the kernels execute (csrc/local_bodies.h + csrc/elastic_math.h) for the host with ADMMB_COUNT_FLOPS, which turns
// the ADMMB_FLOPS(n) annotations of those headers into a counter.  Convention: one flop per add, subtract, multiply,
// divide, square root and logarithm the REFERENCE algorithm performs (TetForce.cpp:320-364, JacobiSVD.h:824-930,
// lbfgssolver.h:43-144, morethuente.h:25-308); comparisons, selections, sign flips and the expansion of a division /
// logarithm into several machine instructions count nothing.  bench.py feeds it a random sample of the tets of the run
// it has just timed (positions, duals and optimiser state downloaded from the device) and scales to the mesh.
// It measures work, it does not produce results anyone uses: z / u computed here are discarded.
#define ADMMB_COUNT_FLOPS 1
#define ADMMB_COUNT_EVALS 1
#include <cstring>
#include <vector>

#include "common.h"
#include "local_bodies.h"

using namespace admmb;

// kind: ADMMB_TET_NEOHOOKEAN / ADMMB_TET_STVK.  x_rest, x_cur: [n][3]; idx: [count][4] node ids; u: [count][9];
// state: [count][4] (last_prox_result + init_hess) entering the iteration.  out3 = { flops, objective evaluations,
// L-BFGS iterations } summed over the `count` tets for ONE ADMM iteration.
extern "C" int admmb_flopcount_hyper_tets(int kind, int n, const double *x_rest, int count, const int *idx, double mu, double lambda,
                                          int maxit, double dt, const double *x_cur, const double *u, const double *state, double *out3) {
	admmb_ctx ctx;
	ctx.n = n;
	ctx.dt = dt;
	ctx.h_x0.assign(x_rest, x_rest + 3 * (size_t)n);
	Batch b;
	b.type = BT_TETS; b.kind = kind; b.count = count; b.nv = 4; b.rows = 9; b.nsel = 12; b.naux = 0; b.nstate = 4;
	b.p0 = mu; b.p1 = lambda; b.max_iterations = maxit;
	b.idx.assign(idx, idx + 4 * (size_t)count);
	if (compute_rest_state(&ctx, b) != 0) return -1;
	double total = 0.0, evals = 0.0, its_total = 0.0;
	std::vector<double> P(12);
	for (int e = 0; e < count; ++e) {
		// one force at a time: with count == 1 the structure-of-arrays layout is the plain array
		int id[4] = { 0, 1, 2, 3 };
		double xs[12], uu[9], zz[9], st[4], wdt2 = dt * dt * b.w[e] * b.w[e];
		for (int c = 0; c < 4; ++c)
			for (int j = 0; j < 3; ++j) xs[3 * c + j] = x_cur[3 * (size_t)idx[4 * (size_t)e + c] + j];
		memcpy(uu, u + 9 * (size_t)e, sizeof(uu));
		memcpy(st, state + 4 * (size_t)e, sizeof(st));
		int its = 0;
		LocalArgs a;
		memset(&a, 0, sizeof(a));
		a.count = 1; a.idx = id; a.S = &b.S[12 * (size_t)e]; a.w = &b.w[e]; a.wdt2 = &wdt2; a.kk = &b.kk[e];
		a.u = uu; a.z = zz; a.state = st; a.its = &its; a.x = xs; a.P = P.data();
		a.p0 = mu; a.p1 = lambda; a.kprox = std::min(mu, lambda); a.max_iterations = maxit;
		g_flops = 0.0;
		g_eval_count = 0;
		double park[18];
		if (kind == ADMMB_TET_NEOHOOKEAN) { if (maxit <= 5) local_tet_hyper<NHModel, 5>(a, 0, park, 1); else local_tet_hyper<NHModel, 10>(a, 0, park, 1); }
		else { if (maxit <= 5) local_tet_hyper<StVKModel, 5>(a, 0, park, 1); else local_tet_hyper<StVKModel, 10>(a, 0, park, 1); }
		total += g_flops;
		evals += (double)g_eval_count;
		its_total += its;
	}
	out3[0] = total; out3[1] = evals; out3[2] = its_total;
	return 0;
}
