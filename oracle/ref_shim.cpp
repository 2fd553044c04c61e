// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A thin extern "C" wrapper that drives the UNMODIFIED reference solver
// (admm::System and its Force subclasses, compiled from the sources where they
// lie under /root/reference by oracle/Makefile) so that Python tests, the golden
// generator and bench.py's cpu_baseline / --impl reference legs can call it via
// ctypes.  Nothing from the reference is copied here: this file only includes
// the reference's public headers and calls its public API:
//   admm::System::{add_nodes,initialize,step}   A/src/system/System.hpp:55-66
//   admm::Force subclasses' constructors         A/src/system/*Force.hpp
// Per-iteration dumps are obtained WITHOUT restating System::step(): a probe
// Force with D_i = I and weight 0 is appended as the last force; its project()
// is called once per ADMM iteration by the unmodified loop (System.cpp:57-58),
// where Dx[global_idx..] is exactly curr_x of that iteration and u/z hold the
// other forces' results (the dump is taken with one OpenMP thread so that all
// other forces have finished).  Weight-0 rows contribute exact zeros to
// A = M + dt^2 D^T W^2 D and to the right-hand side.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
// arm may load the library built from this file.

#include "System.hpp"
#include "TetForce.hpp"
#include "TriangleForce.hpp"
#include "BendForce.hpp"
#include "AnchorForce.hpp"
#include "CollisionForce.hpp"
#include "CollisionSphere.hpp"
#include "CollisionCylinder.hpp"
#include "CollisionFloor.hpp"
#include "ExplicitForce.hpp"

#include <chrono>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace admm;

namespace {

struct Shim;

// Probe force: D_i = I (3n rows), weight 0.  See header comment.
class ProbeForce : public Force {
public:
	Shim *owner;
	int rows;
	explicit ProbeForce(Shim *o) : owner(o), rows(0) { weight = 0.0; }
	void get_selector(const Eigen::VectorXd &x, std::vector<Eigen::Triplet<double> > &triplets,
	                  std::vector<double> &weights) {
		global_idx = weights.size();
		rows = x.size();
		for (int i = 0; i < rows; ++i) {
			triplets.push_back(Eigen::Triplet<double>(i + global_idx, i, 1.0));
			weights.push_back(0.0);
		}
	}
	void project(double dt, const Eigen::VectorXd &Dx, Eigen::VectorXd &u, Eigen::VectorXd &z) const;
};

// Subclass only to read the protected ADMM vectors after a step.
struct ProbeSystem : public System {
	const Eigen::VectorXd &get_u() const { return curr_u; }
	const Eigen::VectorXd &get_z() const { return curr_z; }
	Eigen::VectorXd &mut_u() { return curr_u; }
	long D_rows() const { return m_D.rows(); }
	long D_nnz() const { return m_D.nonZeros(); }
	long L_nnz() { return solver.matrixL().nestedExpression().nonZeros(); }
};

struct Shim {
	ProbeSystem sys;
	std::vector<std::shared_ptr<ControlPoint> > cps;
	std::vector<std::shared_ptr<CollisionShape> > shapes;
	std::shared_ptr<ProbeForce> probe;
	// live-row map: for each force (excluding the probe), [start,count) in the reference layout
	std::vector<long> live_start, live_count;
	long live_rows;
	// dump targets (set by ref_step_dump)
	double *dump_x, *dump_z, *dump_u, *dump_prox;
	int n_hyper;
	mutable int dump_iter;
	Shim() : live_rows(0), dump_x(0), dump_z(0), dump_u(0), dump_prox(0), n_hyper(0), dump_iter(0) {}

	static int rows_of(const Force *f, long n3) {
		if (dynamic_cast<const Spring *>(f)) return 3;
		if (dynamic_cast<const LinearTetStrain *>(f)) return 9;
		if (dynamic_cast<const TetVolume *>(f)) return 9;
		if (dynamic_cast<const HyperElasticTet *>(f)) return 9;
		if (dynamic_cast<const LimitedTriangleStrain *>(f)) return 6; // includes TriArea
		if (dynamic_cast<const FungTriangle *>(f)) return 6;
		if (dynamic_cast<const BendForce *>(f)) return 9;
		if (dynamic_cast<const StaticAnchor *>(f)) return 3;
		if (dynamic_cast<const MovingAnchor *>(f)) return 3;
		if (dynamic_cast<const CollisionForce *>(f)) return (int)n3;
		return 0;
	}
	void build_live_map() {
		live_start.clear(); live_count.clear(); live_rows = 0;
		for (size_t i = 0; i < sys.forces.size(); ++i) {
			const Force *f = sys.forces[i].get();
			if (f == probe.get()) continue;
			int r = rows_of(f, sys.m_x.size());
			live_start.push_back(f->global_idx);
			live_count.push_back(r);
			live_rows += r;
		}
	}
	void compact(const Eigen::VectorXd &full, double *out) const {
		long o = 0;
		for (size_t i = 0; i < live_start.size(); ++i) {
			std::memcpy(out + o, full.data() + live_start[i], sizeof(double) * live_count[i]);
			o += live_count[i];
		}
	}
	void expand(const double *in, Eigen::VectorXd &full) const {
		long o = 0;
		for (size_t i = 0; i < live_start.size(); ++i) {
			std::memcpy(full.data() + live_start[i], in + o, sizeof(double) * live_count[i]);
			o += live_count[i];
		}
	}
};

void ProbeForce::project(double, const Eigen::VectorXd &Dx, Eigen::VectorXd &u, Eigen::VectorXd &z) const {
	// keep our own rows inert: z = Dx + u  =>  u unchanged (stays 0)
	for (int i = 0; i < rows; ++i) { z[global_idx + i] = Dx[global_idx + i] + u[global_idx + i]; }
	if (!owner->dump_x) return;
	const int it = owner->dump_iter++;
	const long n3 = rows;
	std::memcpy(owner->dump_x + (long)it * n3, Dx.data() + global_idx, sizeof(double) * n3);
	if (owner->dump_z) owner->compact(z, owner->dump_z + (long)it * owner->live_rows);
	if (owner->dump_u) owner->compact(u, owner->dump_u + (long)it * owner->live_rows);
	if (owner->dump_prox) {
		double *o = owner->dump_prox + (long)it * owner->n_hyper * 4;
		for (size_t i = 0; i < owner->sys.forces.size(); ++i) {
			const HyperElasticTet *t = dynamic_cast<const HyperElasticTet *>(owner->sys.forces[i].get());
			if (!t) continue;
			for (int j = 0; j < 3; ++j) o[j] = t->last_prox_result[j];
			o[3] = t->solver->settings_.init_hess;
			o += 4;
		}
	}
}

} // namespace

extern "C" {

void *ref_create(double dt, int iters, int verbose) {
	Shim *s = new Shim();
	s->sys.settings.timestep_s = dt;
	s->sys.settings.admm_iters = iters;
	s->sys.settings.verbose = verbose;
	return s;
}
void ref_destroy(void *h) { delete (Shim *)h; }

int ref_add_nodes(void *h, const double *x, const double *m, int n3) {
	Shim *s = (Shim *)h;
	Eigen::VectorXd xv = Eigen::Map<const Eigen::VectorXd>(x, n3);
	Eigen::VectorXd mv = Eigen::Map<const Eigen::VectorXd>(m, n3);
	return s->sys.add_nodes(xv, mv);
}

// kind: 0 LinearTetStrain(p0=stiffness)  1 HyperElasticTet "nh"(p0=mu,p1=lambda,maxit)
//       2 HyperElasticTet "stvk"         3 TetVolume(p0=stiffness,p1=min,p2=max)
int ref_add_tets(void *h, int kind, int T, const int *idx, double p0, double p1, double p2, int maxit) {
	Shim *s = (Shim *)h;
	for (int t = 0; t < T; ++t) {
		const int *p = idx + 4 * t;
		std::shared_ptr<Force> f;
		switch (kind) {
		case 0: f.reset(new LinearTetStrain(p[0], p[1], p[2], p[3], p0)); break;
		case 1: f.reset(new HyperElasticTet(p[0], p[1], p[2], p[3], p0, p1, maxit, "nh")); break;
		case 2: f.reset(new HyperElasticTet(p[0], p[1], p[2], p[3], p0, p1, maxit, "stvk")); break;
		case 3: f.reset(new TetVolume(p[0], p[1], p[2], p[3], p0, p1, p2)); break;
		default: return -1;
		}
		s->sys.forces.push_back(f);
	}
	return 0;
}

// kind: 0 LimitedTriangleStrain(stiffness,min,max,flag=strain_limiting) 1 TriArea(flag=iters) 2 FungTriangle(stiffness=mu)
int ref_add_tris(void *h, int kind, int T, const int *idx, double stiffness, double lmin, double lmax, int flag) {
	Shim *s = (Shim *)h;
	for (int t = 0; t < T; ++t) {
		const int *p = idx + 3 * t;
		std::shared_ptr<Force> f;
		switch (kind) {
		case 0: f.reset(new LimitedTriangleStrain(p[0], p[1], p[2], stiffness, lmin, lmax, flag != 0)); break;
		case 1: f.reset(new TriArea(p[0], p[1], p[2], stiffness, flag, lmin, lmax)); break;
		case 2: f.reset(new FungTriangle(p[0], p[1], p[2], stiffness, lmin, lmax)); break;
		default: return -1;
		}
		s->sys.forces.push_back(f);
	}
	return 0;
}

int ref_add_springs(void *h, int S, const int *idx, const double *stiffness) {
	Shim *s = (Shim *)h;
	for (int i = 0; i < S; ++i)
		s->sys.forces.push_back(std::shared_ptr<Force>(new Spring(idx[2 * i], idx[2 * i + 1], stiffness[i])));
	return 0;
}

int ref_add_bends(void *h, int H, const int *idx, double stiffness) {
	Shim *s = (Shim *)h;
	for (int i = 0; i < H; ++i) {
		const int *p = idx + 4 * i;
		s->sys.forces.push_back(std::shared_ptr<Force>(new BendForce(p[0], p[1], p[2], p[3], stiffness)));
	}
	return 0;
}

int ref_add_static_anchors(void *h, int A, const int *idx, double weight) {
	Shim *s = (Shim *)h;
	for (int i = 0; i < A; ++i)
		s->sys.forces.push_back(std::shared_ptr<Force>(new StaticAnchor(idx[i], weight)));
	return 0;
}

// returns the index of the first control point created
int ref_add_moving_anchors(void *h, int A, const int *idx, const double *pos, double weight) {
	Shim *s = (Shim *)h;
	int first = (int)s->cps.size();
	for (int i = 0; i < A; ++i) {
		std::shared_ptr<ControlPoint> cp(new ControlPoint(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2])));
		s->cps.push_back(cp);
		s->sys.forces.push_back(std::shared_ptr<Force>(new MovingAnchor(idx[i], cp, weight)));
	}
	return first;
}
int ref_set_control_point(void *h, int cp, const double *pos, int active) {
	Shim *s = (Shim *)h;
	if (cp < 0 || cp >= (int)s->cps.size()) return -1;
	if (pos) s->cps[cp]->pos = Eigen::Vector3d(pos[0], pos[1], pos[2]);
	s->cps[cp]->active = active != 0;
	return 0;
}
int ref_get_control_point(void *h, int cp, double *pos) {
	Shim *s = (Shim *)h;
	if (cp < 0 || cp >= (int)s->cps.size()) return -1;
	for (int j = 0; j < 3; ++j) pos[j] = s->cps[cp]->pos[j];
	return s->cps[cp]->active ? 1 : 0;
}
int ref_set_moving_anchor_weight(void *h, int cp, double w) {
	Shim *s = (Shim *)h;
	if (cp < 0 || cp >= (int)s->cps.size()) return -1;
	s->cps[cp]->anchorForce->weight = w;
	return 0;
}
void ref_recompute_weights(void *h) { ((Shim *)h)->sys.recompute_weights(); }

// shape kind: 0 sphere(center,radius) 1 cylinder(center,radius) 2 floor(center)
int ref_add_collision(void *h, int S, const int *kind, const double *params4, double weight) {
	Shim *s = (Shim *)h;
	for (int i = 0; i < S; ++i) {
		Eigen::Vector3d c(params4[4 * i], params4[4 * i + 1], params4[4 * i + 2]);
		double r = params4[4 * i + 3];
		std::shared_ptr<CollisionShape> sh;
		switch (kind[i]) {
		case 0: sh.reset(new CollisionSphere(c, r)); break;
		case 1: sh.reset(new CollisionCylinder(c, Eigen::Vector3d(1, 1, 1), r)); break;
		case 2: sh.reset(new CollisionFloor(c)); break;
		default: return -1;
		}
		s->shapes.push_back(sh);
	}
	s->sys.forces.push_back(std::shared_ptr<Force>(new CollisionForce(s->shapes, weight)));
	return 0;
}

int ref_add_explicit(void *h, const double *dir) {
	Shim *s = (Shim *)h;
	s->sys.explicit_forces.push_back(std::shared_ptr<ExplicitForce>(
	    new ExplicitForce(Eigen::Vector3d(dir[0], dir[1], dir[2]))));
	return 0;
}
int ref_add_wind(void *h, int ntris, const int *tris, const double *dir) {
	Shim *s = (Shim *)h;
	std::vector<int> t(tris, tris + 3 * ntris);
	std::shared_ptr<WindForce> w(new WindForce(t));
	w->direction = Eigen::Vector3d(dir[0], dir[1], dir[2]);
	s->sys.explicit_forces.push_back(w);
	return (int)s->sys.explicit_forces.size() - 1;
}
int ref_set_explicit_direction(void *h, int which, const double *dir) {
	Shim *s = (Shim *)h;
	if (which < 0 || which >= (int)s->sys.explicit_forces.size()) return -1;
	s->sys.explicit_forces[which]->direction = Eigen::Vector3d(dir[0], dir[1], dir[2]);
	return 0;
}

// with_probe != 0 appends the probe force (needed for ref_step_dump).
int ref_initialize(void *h, int with_probe) {
	Shim *s = (Shim *)h;
	if (with_probe) {
		s->probe.reset(new ProbeForce(s));
		s->sys.forces.push_back(s->probe);
	}
	if (!s->sys.initialize()) return -1;
	s->build_live_map();
	return 0;
}

long ref_num_dof(void *h) { return ((Shim *)h)->sys.m_x.size(); }
long ref_live_rows(void *h) { return ((Shim *)h)->live_rows; }
long ref_D_rows(void *h) { return ((Shim *)h)->sys.D_rows(); }
long ref_D_nnz(void *h) { return ((Shim *)h)->sys.D_nnz(); }
long ref_L_nnz(void *h) { return ((Shim *)h)->sys.L_nnz(); }
double ref_elapsed(void *h) { return ((Shim *)h)->sys.elapsed_s; }
void ref_set_iters(void *h, int it) { ((Shim *)h)->sys.settings.admm_iters = it; }

void ref_get_x(void *h, double *x) { Shim *s = (Shim *)h; std::memcpy(x, s->sys.m_x.data(), sizeof(double) * s->sys.m_x.size()); }
void ref_set_x(void *h, const double *x) { Shim *s = (Shim *)h; std::memcpy(s->sys.m_x.data(), x, sizeof(double) * s->sys.m_x.size()); }
void ref_get_v(void *h, double *v) { Shim *s = (Shim *)h; std::memcpy(v, s->sys.m_v.data(), sizeof(double) * s->sys.m_v.size()); }
void ref_set_v(void *h, const double *v) { Shim *s = (Shim *)h; std::memcpy(s->sys.m_v.data(), v, sizeof(double) * s->sys.m_v.size()); }
void ref_get_z(void *h, double *z) { Shim *s = (Shim *)h; s->compact(s->sys.get_z(), z); }
void ref_get_u(void *h, double *u) { Shim *s = (Shim *)h; s->compact(s->sys.get_u(), u); }
void ref_set_u(void *h, const double *u) { Shim *s = (Shim *)h; s->expand(u, s->sys.mut_u()); }

// weights per live row (W diagonal) in compact order, and per-force weight
void ref_get_force_weights(void *h, double *w) {
	Shim *s = (Shim *)h; long o = 0;
	for (size_t i = 0; i < s->sys.forces.size(); ++i) {
		if (s->sys.forces[i].get() == s->probe.get()) continue;
		w[o++] = s->sys.forces[i]->weight;
	}
}

// Hyperelastic per-tet optimiser state: last_prox_result (3) + init_hess (1), for each
// HyperElasticTet in force order.  Returns the number of such tets.
int ref_get_prox_state(void *h, double *out4) {
	Shim *s = (Shim *)h; int c = 0;
	for (size_t i = 0; i < s->sys.forces.size(); ++i) {
		HyperElasticTet *t = dynamic_cast<HyperElasticTet *>(s->sys.forces[i].get());
		if (!t) continue;
		if (out4) {
			for (int j = 0; j < 3; ++j) out4[4 * c + j] = t->last_prox_result[j];
			out4[4 * c + 3] = t->solver->settings_.init_hess;
		}
		++c;
	}
	return c;
}
// Setter for the same state (teacher-forced replays that start in the middle of a run).
int ref_set_prox_state(void *h, const double *in4) {
	Shim *s = (Shim *)h; int c = 0;
	for (size_t i = 0; i < s->sys.forces.size(); ++i) {
		HyperElasticTet *t = dynamic_cast<HyperElasticTet *>(s->sys.forces[i].get());
		if (!t) continue;
		for (int j = 0; j < 3; ++j) t->last_prox_result[j] = in4[4 * c + j];
		t->solver->settings_.init_hess = in4[4 * c + 3];
		++c;
	}
	return c;
}
// L-BFGS outer-iteration counts of the last project() per hyperelastic tet
int ref_get_prox_iters(void *h, int *out) {
	Shim *s = (Shim *)h; int c = 0;
	for (size_t i = 0; i < s->sys.forces.size(); ++i) {
		HyperElasticTet *t = dynamic_cast<HyperElasticTet *>(s->sys.forces[i].get());
		if (!t) continue;
		if (out) out[c] = t->solver->n_iters;
		++c;
	}
	return c;
}

int ref_step(void *h) { return ((Shim *)h)->sys.step() ? 0 : -1; }

// One unmodified step() with per-iteration dumps: x_it[it] = curr_x entering iteration it
// (x_it[0] = x_bar), z_it/u_it[it] = compact z,u after the local step of iteration it.
// prox_it[it] (may be NULL) = {last_prox_result[3], init_hess} of every HyperElasticTet after that local step.
// Final x is read with ref_get_x.  Requires ref_initialize(h, 1).  Runs single-threaded.
int ref_step_dump(void *h, double *x_it, double *z_it, double *u_it, double *prox_it) {
	Shim *s = (Shim *)h;
	if (!s->probe) return -2;
	s->dump_prox = prox_it;
	s->n_hyper = 0;
	for (size_t i = 0; i < s->sys.forces.size(); ++i)
		if (dynamic_cast<HyperElasticTet *>(s->sys.forces[i].get())) s->n_hyper++;
#ifdef _OPENMP
	int old = omp_get_max_threads();
	omp_set_num_threads(1);
#endif
	s->dump_x = x_it; s->dump_z = z_it; s->dump_u = u_it; s->dump_iter = 0;
	bool ok = s->sys.step();
	s->dump_x = s->dump_z = s->dump_u = s->dump_prox = 0;
#ifdef _OPENMP
	omp_set_num_threads(old);
#endif
	return ok ? 0 : -1;
}

// Times `frames` unmodified step() calls; returns seconds.
double ref_step_timed(void *h, int frames) {
	Shim *s = (Shim *)h;
	auto t0 = std::chrono::steady_clock::now();
	for (int f = 0; f < frames; ++f) s->sys.step();
	auto t1 = std::chrono::steady_clock::now();
	return std::chrono::duration<double>(t1 - t0).count();
}

int ref_omp_threads() {
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}
void ref_set_omp_threads(int n) {
#ifdef _OPENMP
	omp_set_num_threads(n);
#else
	(void)n;
#endif
}

} // extern "C"
