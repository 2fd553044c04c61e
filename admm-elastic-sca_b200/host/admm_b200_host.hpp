// admm_b200_host.hpp -- the reference's C++ solver surface (namespace admm: System, Force and its subclasses,
// ExplicitForce, WindForce, ControlPoint, the collision shapes) re-implemented as a thin host layer over the
// C ABI of libadmm_b200.so.  Same class names, constructors, public members and error behaviour as
//   A/src/system/{System,Force,TetForce,TriangleForce,BendForce,AnchorForce,CollisionForce,ExplicitForce}.hpp
//   A/src/collision/Collision{Shape,Sphere,Cylinder,Floor}.hpp        (A/ = deps/admm-elastic-sca)
// so that code written against the reference (its samples, src/SimContext.cpp, src/ForceBuilder.cpp) compiles
// unchanged with this directory first on the include path and runs its solver path on the GPU.
//
// What differs by design:
//  * A Force here only DESCRIBES a constraint (indices + material); System::initialize() flattens runs of equal
//    forces into batches and hands them to admmb_add_*; there is no per-force project() on the host and a Force
//    subclass unknown to the device is a hard error (no CPU fallback).
//  * Force::weight / rest-shape members are filled in by System::initialize() from the values the device
//    library computed (same formulas as the reference's Force::initialize bodies, csrc/rest_state.cpp).
//  * The reference's own explicit forces (ExplicitForce over all nodes or a subset, WindForce) run ON THE DEVICE
//    (admmb_set_gravity / admmb_add_explicit_subset / admmb_add_wind; their `direction` is re-read every step, as
//    windyflag.cpp:150 changes it per frame).  A user-defined subclass of ExplicitForce cannot run there: if the list
//    holds one -- or was edited after initialize() -- EVERY explicit force runs on the host in list order through its own
//    project(), exactly as in the reference, and the device copies are switched off (no reordering, no double application).
//
// Needs Eigen (the reference vendors 3.2.5 under A/deps/Eigen3; any 3.x works): only VectorXd / Vector3d /
// Vector4d / Matrix types appear in the interface.
#ifndef ADMM_B200_HOST_HPP
#define ADMM_B200_HOST_HPP 1

#include <Eigen/Dense>
#include <Eigen/Sparse> // reference headers expose Eigen::Triplet in Force::get_selector's signature

#include <cmath>
#include <cstdio>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

#include "../../include/admm_b200.h"

namespace admm {

// ---- constraint descriptions --------------------------------------------------------------------------
class Force {
public:
	int global_idx;  // first row of this force in the (compact) z/u export, set by System::initialize()
	double weight;   // ADMM weight w_i; computed by initialize(), may be changed + System::recompute_weights()
	Force() : global_idx(-1), weight(0.0) {}
	virtual ~Force() {}
	virtual void set_eps(double) {}

	// device-side description (not part of the reference API)
	enum B200Class { B_TET, B_TRI, B_SPRING, B_BEND, B_STATIC_ANCHOR, B_MOVING_ANCHOR, B_COLLISION };
	virtual B200Class b200_class() const = 0;
	virtual bool b200_same_batch(const Force &) const { return true; } // same class assumed by the caller
	virtual int b200_rows() const = 0;
	virtual bool b200_is_batch() const { return false; } // ForceBatch below: ONE object for a whole run of elements
};

class Spring : public Force {
public:
	Spring(int idx0_, int idx1_, double stiffness_) : idx0(idx0_), idx1(idx1_), stiffness(stiffness_), rest_length(0.0) {}
	int idx0, idx1;
	double stiffness, rest_length;
	B200Class b200_class() const { return B_SPRING; }
	int b200_rows() const { return 3; }
};

class TetBase : public Force {
public:
	int idx[4];
	double volume;
	B200Class b200_class() const { return B_TET; }
	int b200_rows() const { return 9; }
	virtual int b200_kind() const = 0;
	virtual void b200_params(double &p0, double &p1, double &p2, int &maxit) const = 0;
	bool b200_same_batch(const Force &o) const {
		const TetBase *t = dynamic_cast<const TetBase *>(&o);
		if (!t || t->b200_kind() != b200_kind()) return false;
		double a0, a1, a2, b0, b1, b2; int am, bm;
		b200_params(a0, a1, a2, am); t->b200_params(b0, b1, b2, bm);
		return a0 == b0 && a1 == b1 && a2 == b2 && am == bm;
	}
protected:
	TetBase(int i0, int i1, int i2, int i3) : volume(0.0) { idx[0] = i0; idx[1] = i1; idx[2] = i2; idx[3] = i3; }
};

class LinearTetStrain : public TetBase {
public:
	LinearTetStrain(int i0, int i1, int i2, int i3, double stiffness_, double weight_scale_ = 1.f)
	    : TetBase(i0, i1, i2, i3), stiffness(stiffness_), weight_scale(weight_scale_) {}
	double stiffness, weight_scale;
	int b200_kind() const { return ADMMB_TET_LINEAR_STRAIN; }
	void b200_params(double &p0, double &p1, double &p2, int &m) const { p0 = stiffness; p1 = p2 = 0.0; m = 0; }
};

class TetVolume : public TetBase {
public:
	TetVolume(int i0, int i1, int i2, int i3, double stiffness_, double limit_min_, double limit_max_)
	    : TetBase(i0, i1, i2, i3), stiffness(stiffness_), rest_volume(0.0), limit_min(limit_min_), limit_max(limit_max_) {}
	double stiffness, rest_volume, limit_min, limit_max;
	int b200_kind() const { return ADMMB_TET_VOLUME; }
	void b200_params(double &p0, double &p1, double &p2, int &m) const { p0 = stiffness; p1 = limit_min; p2 = limit_max; m = 0; }
};

class HyperElasticTet : public TetBase {
public:
	HyperElasticTet(int i0, int i1, int i2, int i3, double mu_, double lambda_, int max_iterations_, std::string type_)
	    : TetBase(i0, i1, i2, i3), type(0), max_iterations(max_iterations_), mu(mu_), lambda(lambda_) {
		if (type_ == "stvk" || type_ == "1") type = 1;
	}
	int type, max_iterations;
	double mu, lambda;
	int b200_kind() const { return type == 1 ? ADMMB_TET_STVK : ADMMB_TET_NEOHOOKEAN; }
	void b200_params(double &p0, double &p1, double &p2, int &m) const { p0 = mu; p1 = lambda; p2 = 0.0; m = max_iterations; }
};

class LimitedTriangleStrain : public Force {
public:
	LimitedTriangleStrain(int id0_, int id1_, int id2_, double stiffness_, double limit_min_, double limit_max_, bool strain_limiting_ = true)
	    : id0(id0_), id1(id1_), id2(id2_), stiffness(stiffness_), limit_min(limit_min_), limit_max(limit_max_), area(0.0),
	      strain_limiting(strain_limiting_) {}
	int id0, id1, id2;
	double stiffness, limit_min, limit_max, area;
	bool strain_limiting;
	B200Class b200_class() const { return B_TRI; }
	int b200_rows() const { return 6; }
	virtual int b200_kind() const { return ADMMB_TRI_LIMITED_STRAIN; }
	virtual int b200_flag() const { return strain_limiting ? 1 : 0; }
	bool b200_same_batch(const Force &o) const {
		const LimitedTriangleStrain *t = dynamic_cast<const LimitedTriangleStrain *>(&o);
		return t && t->b200_kind() == b200_kind() && t->stiffness == stiffness && t->limit_min == limit_min && t->limit_max == limit_max &&
		       t->b200_flag() == b200_flag();
	}
};

class TriArea : public LimitedTriangleStrain {
public:
	TriArea(int id0_, int id1_, int id2_, double stiffness_, int iters_, double limit_min_, double limit_max_)
	    : LimitedTriangleStrain(id0_, id1_, id2_, stiffness_, limit_min_, limit_max_), iters(iters_) {}
	int iters;
	int b200_kind() const { return ADMMB_TRI_AREA; }
	int b200_flag() const { return iters; }
};

class FungTriangle : public Force {
public:
	FungTriangle(int id0_, int id1_, int id2_, double mu_, double limit_min_, double limit_max_)
	    : id0(id0_), id1(id1_), id2(id2_), mu(mu_), limit_min(limit_min_), limit_max(limit_max_), area(0.0) {}
	int id0, id1, id2;
	double mu, limit_min, limit_max, area;
	B200Class b200_class() const { return B_TRI; }
	int b200_rows() const { return 6; }
	bool b200_same_batch(const Force &o) const {
		const FungTriangle *t = dynamic_cast<const FungTriangle *>(&o);
		return t && t->mu == mu;
	}
};

class BendForce : public Force {
public:
	BendForce(int i0, int i1, int i2, int i3, double stiffness_) : stiffness(stiffness_) {
		idx[0] = i0; idx[1] = i1; idx[2] = i2; idx[3] = i3;
		weight = std::sqrt(stiffness);
		alpha.setZero();
	}
	int idx[4];
	Eigen::Vector4d alpha;
	double stiffness;
	B200Class b200_class() const { return B_BEND; }
	int b200_rows() const { return 9; }
	bool b200_same_batch(const Force &o) const {
		const BendForce *t = dynamic_cast<const BendForce *>(&o);
		return t && t->stiffness == stiffness;
	}
};

// ---- SoA batches (not in the reference API) ----------------------------------------------------
// ONE Force object that describes a whole run of equal elements: the corner indices as one flat array, the material once.
// System::initialize() hands the arrays to admmb_add_* as they are -- no heap object, no virtual call and no copy per
// element, which is what the reference's ForceBuilder costs (ForceBuilder.cpp:310-313,346-349: one shared_ptr<Force> per
// tet) and what makes an 8 M-tet scene slow to load.  A batch sits in System::forces like any other force (so the order
// of the rows of z / u is the order of the list, as in the reference); `weights` holds the per-element ADMM weights after
// initialize() and may be edited before recompute_weights(), like Force::weight.  host/scene/ForceBuilderBatched.cpp builds
// these from the reference's XML vocabulary.
class ForceBatch : public Force {
public:
	std::vector<int> idx;        // corners, element-major
	std::vector<double> weights; // one per element (Force::weight of the batch object itself is unused)
	size_t count() const { return idx.size() / (size_t)b200_corners(); }
	bool b200_is_batch() const { return true; }
	int b200_rows() const { return (int)count() * b200_rows_per_element(); }
	virtual int b200_corners() const = 0;
	virtual int b200_rows_per_element() const = 0;
	virtual int b200_add(admmb_ctx *ctx) const = 0; // registers the batch, returns the device batch id (< 0: error)
};

class TetBatch : public ForceBatch { // TetForce.hpp:31-147
public:
	// kind: ADMMB_TET_LINEAR_STRAIN (p0 = stiffness), ADMMB_TET_VOLUME (stiffness, range_min, range_max),
	// ADMMB_TET_NEOHOOKEAN / ADMMB_TET_STVK (mu, lambda; max_iterations)
	TetBatch(int kind_, double p0_, double p1_ = 0.0, double p2_ = 0.0, int max_iterations_ = 0)
	    : kind(kind_), max_iterations(max_iterations_) { p[0] = p0_; p[1] = p1_; p[2] = p2_; }
	int kind, max_iterations;
	double p[3];
	B200Class b200_class() const { return B_TET; }
	int b200_corners() const { return 4; }
	int b200_rows_per_element() const { return 9; }
	int b200_add(admmb_ctx *ctx) const { return admmb_add_tets(ctx, kind, (int)count(), idx.data(), p[0], p[1], p[2], max_iterations); }
};

class TriangleBatch : public ForceBatch { // LimitedTriangleStrain, TriangleForce.hpp:31-62
public:
	TriangleBatch(double stiffness_, double limit_min_, double limit_max_, bool strain_limiting_ = true)
	    : stiffness(stiffness_), limit_min(limit_min_), limit_max(limit_max_), strain_limiting(strain_limiting_) {}
	double stiffness, limit_min, limit_max;
	bool strain_limiting;
	B200Class b200_class() const { return B_TRI; }
	int b200_corners() const { return 3; }
	int b200_rows_per_element() const { return 6; }
	int b200_add(admmb_ctx *ctx) const { return admmb_add_tris(ctx, ADMMB_TRI_LIMITED_STRAIN, (int)count(), idx.data(), stiffness, limit_min, limit_max, strain_limiting ? 1 : 0); }
};

class BendBatch : public ForceBatch { // BendForce.hpp:31-60; idx = the four hinge vertices in the reference's (Volino) order
public:
	explicit BendBatch(double stiffness_) : stiffness(stiffness_) {}
	double stiffness;
	B200Class b200_class() const { return B_BEND; }
	int b200_corners() const { return 4; }
	int b200_rows_per_element() const { return 9; }
	int b200_add(admmb_ctx *ctx) const { return admmb_add_bends(ctx, (int)count(), idx.data(), stiffness); }
};

class SpringBatch : public ForceBatch { // Force.hpp:78-95
public:
	explicit SpringBatch(double stiffness_) : stiffness(stiffness_) {}
	double stiffness;
	B200Class b200_class() const { return B_SPRING; }
	int b200_corners() const { return 2; }
	int b200_rows_per_element() const { return 3; }
	int b200_add(admmb_ctx *ctx) const {
		const std::vector<double> k(count(), stiffness);
		return admmb_add_springs(ctx, (int)count(), idx.data(), k.data());
	}
};

namespace helper {
// AnchorForce.hpp:33-47
static inline Eigen::Vector3d smooth_move(double t, double t0, double t1, Eigen::Vector3d a, Eigen::Vector3d b) {
	if (t < t0) return a;
	const double r = (t - t0) / (t1 - t0);
	if (r > 1.0) return b;
	const Eigen::Vector3d d = b - a;
	return (a + (3.0 * r * r - 2.0 * r * r * r) * d);
}
static inline Eigen::Vector3d linear_move(double t, double t0, double t1, Eigen::Vector3d a, Eigen::Vector3d b) {
	if (t < t0) return a;
	const double r = (t - t0) / (t1 - t0);
	if (r > 1.0) return b;
	const Eigen::Vector3d d = b - a;
	return (a + d); // sic: the reference does not scale by r
}
} // namespace helper

class StaticAnchor : public Force {
public:
	StaticAnchor(int idx_, double use_weight_ = -1.0) : idx(idx_) {
		weight = (use_weight_ > 0.0) ? use_weight_ : 1000.f;
		pos.setZero();
	}
	int idx;
	Eigen::Vector3d pos;
	B200Class b200_class() const { return B_STATIC_ANCHOR; }
	int b200_rows() const { return 3; }
	bool b200_same_batch(const Force &o) const { return dynamic_cast<const StaticAnchor *>(&o) != 0; }
};

class MovingAnchor;
class ControlPoint {
public:
	ControlPoint() : active(true), anchorForce(0) { pos.setZero(); }
	ControlPoint(Eigen::Vector3d pos_) : pos(pos_), active(true), anchorForce(0) {}
	Eigen::Vector3d pos;
	bool active;
	MovingAnchor *anchorForce;
};

class MovingAnchor : public Force {
public:
	MovingAnchor(int idx_, std::shared_ptr<ControlPoint> p_, double use_weight_ = -1.0) : idx(idx_), point(p_) {
		point->anchorForce = this;
		weight = (use_weight_ > 0.0) ? use_weight_ : 1000.f;
	}
	int idx;
	std::shared_ptr<ControlPoint> point;
	B200Class b200_class() const { return B_MOVING_ANCHOR; }
	int b200_rows() const { return 3; }
	bool b200_same_batch(const Force &o) const { return dynamic_cast<const MovingAnchor *>(&o) != 0; }
};

// ---- collision shapes (A/src/collision) ------------------------------------------------------------------
class CollisionShape {
public:
	CollisionShape(Eigen::Vector3d shapeCenter) { center = shapeCenter; }
	virtual ~CollisionShape() {}
	virtual double isColliding(Eigen::Vector3d pos) const = 0;   // > 0 inside
	virtual Eigen::Vector3d projectOut(const Eigen::Vector3d currPos) const = 0;
	virtual int b200_shape_kind() const = 0;
	virtual double b200_radius() const { return 0.0; }
	Eigen::Vector3d center;
};
class CollisionSphere : public CollisionShape {
public:
	CollisionSphere(Eigen::Vector3d c, double r) : CollisionShape(c), radius(r) {}
	double isColliding(Eigen::Vector3d p) const { return radius - (p - center).norm(); }
	Eigen::Vector3d projectOut(const Eigen::Vector3d p) const { const Eigen::Vector3d d = p - center; return center + radius * (d / d.norm()); }
	int b200_shape_kind() const { return ADMMB_SHAPE_SPHERE; }
	double b200_radius() const { return radius; }
	double radius;
};
class CollisionCylinder : public CollisionShape { // axis parallel to z, centre z forced to 0 (CollisionCylinder.hpp:45)
public:
	CollisionCylinder(Eigen::Vector3d c, Eigen::Vector3d, double r) : CollisionShape(Eigen::Vector3d(c[0], c[1], 0)), radius(r), length(0.0) {}
	double isColliding(Eigen::Vector3d p) const { return radius - (Eigen::Vector3d(p[0], p[1], 0) - center).norm(); }
	Eigen::Vector3d projectOut(const Eigen::Vector3d p) const {
		const Eigen::Vector3d d = Eigen::Vector3d(p[0], p[1], 0) - center;
		return center + radius * (d / d.norm()) + Eigen::Vector3d(0, 0, p[2]);
	}
	int b200_shape_kind() const { return ADMMB_SHAPE_CYLINDER; }
	double b200_radius() const { return radius; }
	double radius, length;
};
class CollisionFloor : public CollisionShape {
public:
	CollisionFloor(Eigen::Vector3d c) : CollisionShape(c), radius(0.0) {}
	double isColliding(Eigen::Vector3d p) const { return center[1] - p[1]; }
	Eigen::Vector3d projectOut(const Eigen::Vector3d p) const { return Eigen::Vector3d(p[0], center[1], p[2]); }
	int b200_shape_kind() const { return ADMMB_SHAPE_FLOOR; }
	double radius;
};

class CollisionForce : public Force {
public:
	CollisionForce(std::vector<std::shared_ptr<CollisionShape> > &collShapes, double use_weight = 32.0) : collisionShapes(collShapes), Di_rows(0), n_nodes(0) {
		weight = use_weight;
	}
	std::vector<std::shared_ptr<CollisionShape> > collisionShapes;
	int Di_rows, n_nodes;
	B200Class b200_class() const { return B_COLLISION; }
	int b200_rows() const { return Di_rows; }
	bool b200_same_batch(const Force &) const { return false; }
};

// ---- explicit forces: host side, exactly as the reference (ExplicitForce.cpp:29-98) -------------------------
class ExplicitForce {
public:
	ExplicitForce(std::vector<int> indices_ = std::vector<int>(0)) : indices(indices_) { direction.setZero(); }
	ExplicitForce(Eigen::Vector3d direction_, std::vector<int> indices_ = std::vector<int>(0)) : direction(direction_), indices(indices_) {}
	virtual ~ExplicitForce() {}
	virtual void project(double dt, Eigen::VectorXd &x, Eigen::VectorXd &v, Eigen::VectorXd &m) const {
		(void)x; (void)m;
		const bool all = indices.empty();
		const int count = all ? (int)(v.size() / 3) : (int)indices.size();
		for (int i = 0; i < count; ++i) {
			const int node = all ? i : indices[i];
			for (int j = 0; j < 3; ++j) v[node * 3 + j] += (dt * direction[j]);
		}
	}
	Eigen::Vector3d direction;
	std::vector<int> indices;
};

class WindForce : public ExplicitForce {
public:
	WindForce(std::vector<int> &tris_) : tris(tris_) { direction = Eigen::Vector3d(0, 0, 0); }
	void project(double dt, Eigen::VectorXd &x, Eigen::VectorXd &v, Eigen::VectorXd &m) const {
		(void)m;
		// serial triangle order = the reference with one OpenMP thread (its parallel version races on v)
		const int nt = (int)tris.size() / 3;
		for (int t = 0; t < nt; ++t) {
			const int a = tris[3 * t] * 3, b = tris[3 * t + 1] * 3, c = tris[3 * t + 2] * 3;
			const Eigen::Vector3d vavg = Eigen::Vector3d(v[a] + v[b] + v[c], v[a + 1] + v[b + 1] + v[c + 1], v[a + 2] + v[b + 2] + v[c + 2]) / 3.0;
			const Eigen::Vector3d rel = vavg - direction;
			const Eigen::Vector3d p0(x[a], x[a + 1], x[a + 2]), p1(x[b], x[b + 1], x[b + 2]), p2(x[c], x[c + 1], x[c + 2]);
			const Eigen::Vector3d nrm = (p1 - p0).cross(p2 - p0);
			const Eigen::Vector3d unit = nrm.normalized();
			const double area = 0.5 * nrm.norm();
			const double vn = unit.dot(rel);
			Eigen::Vector3d f = -1000.0 * area * vn * std::fabs(vn) * unit;
			f *= 0.33;
			f *= dt;
			const int ids[3] = { a, b, c };
			for (int k = 0; k < 3; ++k) { v[ids[k]] += f[0]; v[ids[k] + 1] += f[1]; v[ids[k] + 2] += f[2]; }
		}
	}
	std::vector<int> tris;
};

// ---- the solver -------------------------------------------------------------------------------------------
class System {
public:
	System() : elapsed_s(0.0), initialized(false), ctx(0) {}
	~System() { if (ctx) admmb_destroy(ctx); }

	struct Settings {
		void parse_args(int argc, char **argv) {
			for (int i = 1; i < argc - 1; ++i) {
				std::string arg(argv[i]);
				std::stringstream val(argv[i + 1]);
				if (arg == "-help") help();
				else if (arg == "-dt") val >> timestep_s;
				else if (arg == "-v") val >> verbose;
				else if (arg == "-it") val >> admm_iters;
			}
			if (argc > 0 && std::string(argv[argc - 1]) == "-help") help();
		}
		void help() { printf("\n==========================================\nArgs:\n\t-dt: time step (s)\n\t-v: verbosity (higher -> show more)\n\t-it: # admm iters\n==========================================\n"); }
		double timestep_s;
		int verbose;
		int admm_iters;
		int device;      // (new) CUDA device of this system
		int solver;      // (new) ADMMB_SOLVER_DIRECT / ADMMB_SOLVER_PCG
		bool pin_host;   // (new) page-lock the storage of m_x / m_v so step() transfers it by direct DMA
		bool deterministic; // (new) atomic-free solve: runs are bit-reproducible, like the reference's serial solve (about 10 % slower)
		bool check_finite;  // (new) step() returns false once positions are not finite (the reference's step() always returns true)
		Settings() : timestep_s(0.04), verbose(1), admm_iters(10), device(0), solver(ADMMB_SOLVER_DIRECT), pin_host(true), deterministic(false), check_finite(false) {}
	} settings;

	double elapsed_s;
	Eigen::VectorXd m_x, m_v, m_masses;
	std::vector<std::shared_ptr<ExplicitForce> > explicit_forces;
	std::vector<std::shared_ptr<Force> > forces;
	std::vector<std::function<void(admm::System *)> > pre_step_callbacks;

	int add_nodes(Eigen::VectorXd x, Eigen::VectorXd m) {
		const int old_n = (int)m_x.size(), add = (int)x.size();
		m_x.conservativeResize(old_n + add);
		m_v.conservativeResize(old_n + add);
		m_masses.conservativeResize(old_n + add);
		for (int i = 0; i < add; ++i) { m_x[old_n + i] = x[i]; m_v[old_n + i] = 0.0; m_masses[old_n + i] = m[i]; }
		return (old_n + add) / 3;
	}

	bool initialize();
	bool step();
	void recompute_weights();

	// The names of the later admm-elastic releases (`admm::Solver`, add_force): thin aliases of the members above.
	void add_force(std::shared_ptr<Force> f) { forces.push_back(f); }
	void add_explicit_force(std::shared_ptr<ExplicitForce> f) { explicit_forces.push_back(f); }

	admmb_ctx *b200_context() { return ctx; } // (new) for state dumps / timing through the C ABI

protected:
	bool initialized;
	admmb_ctx *ctx;
	// storage of m_x / m_v currently page-locked in `ctx` (Settings::pin_host), keyed on (pointer, bytes).  Eigen frees the
	// old block when it reallocates, so a stale entry is unregistered BEFORE anything else touches it and never
	// dereferenced; a new context starts with nothing registered (admmb_destroy unregistered what the old one held).
	double *pinned[2] = { 0, 0 };
	long pinned_bytes[2] = { 0, 0 };
	void pin_state() {
		double *cur[2] = { m_x.data(), m_v.data() };
		const long bytes[2] = { (long)(m_x.size() * sizeof(double)), (long)(m_v.size() * sizeof(double)) };
		for (int k = 0; k < 2; ++k) {
			if (!settings.pin_host || (cur[k] == pinned[k] && bytes[k] == pinned_bytes[k])) continue;
			if (pinned[k]) admmb_unregister_host_buffer(ctx, pinned[k]);
			const bool ok = admmb_register_host_buffer(ctx, cur[k], bytes[k]) == 0; // failure: staged path
			pinned[k] = ok ? cur[k] : 0;
			pinned_bytes[k] = ok ? bytes[k] : 0;
		}
	}
	struct BatchRef { int id; Force::B200Class cls; size_t first, count; std::vector<double> pos; std::vector<int> act; bool any_inactive; bool soa; };
	std::vector<BatchRef> batches;
	// explicit forces as registered on the device at initialize(): the object, its device id and the direction last sent
	struct ExplicitRef { const ExplicitForce *obj; int id; double dir[3]; };
	std::vector<ExplicitRef> device_explicit;
	bool explicit_on_host = false;
	bool explicit_list_unchanged() const {
		if (device_explicit.size() != explicit_forces.size()) return false;
		for (size_t i = 0; i < device_explicit.size(); ++i)
			if (device_explicit[i].obj != explicit_forces[i].get()) return false;
		return true;
	}
	bool fail(const char *what) {
		std::cerr << "\n**Solver Error: " << what << ": " << admmb_last_error(ctx) << std::endl;
		return false;
	}
};

inline bool System::initialize() {
	if (settings.verbose > 0) std::cout << "Solver::initialize: " << std::endl;
	if (settings.timestep_s <= 0.0) {
		std::cerr << "\n**Solver Error: timestep set to " << settings.timestep_s << "s, changing to 0.04s." << std::endl;
		settings.timestep_s = 0.04;
	}
	if (!(m_masses.size() == m_x.size() && m_x.size() >= 3)) {
		std::cerr << "\n**Solver Error: Problem with node data!" << std::endl;
		return false;
	}
	if (m_v.size() < m_x.size()) m_v.resize(m_x.size());
	m_v.setZero();
	if (ctx) { admmb_destroy(ctx); ctx = 0; }
	pinned[0] = pinned[1] = 0; // the destroyed context unregistered its buffers
	pinned_bytes[0] = pinned_bytes[1] = 0;
	if (admmb_create(settings.device, &ctx) != ADMMB_OK) {
		std::cerr << "\n**Solver Error: " << admmb_last_error(0) << std::endl;
		return false;
	}
	const int n = (int)(m_x.size() / 3);
	if (admmb_set_nodes(ctx, n, m_x.data(), m_masses.data()) < 0) return fail("set_nodes");
	admmb_set_solver(ctx, settings.solver, 0.0, 0);
	if (settings.deterministic) admmb_set_deterministic(ctx, 1);
	if (settings.check_finite) admmb_set_check_finite(ctx, 1);

	// flatten maximal runs of forces that share class and material into batches, keeping the force order
	batches.clear();
	size_t i = 0;
	long row = 0;
	while (i < forces.size()) {
		Force *f0 = forces[i].get();
		if (f0->b200_is_batch()) { // an SoA batch goes to the device as it is
			ForceBatch *fb = static_cast<ForceBatch *>(f0);
			fb->global_idx = (int)row;
			if (fb->count() > 0) {
				if (fb->idx.size() != fb->count() * (size_t)fb->b200_corners()) { std::cerr << "\n**Solver Error: ragged force batch" << std::endl; return false; }
				const int bid = fb->b200_add(ctx);
				if (bid < 0) return fail("add forces");
				BatchRef br;
				br.id = bid; br.cls = fb->b200_class(); br.first = i; br.count = fb->count(); br.any_inactive = false; br.soa = true;
				batches.push_back(br);
				row += fb->b200_rows();
			}
			++i;
			continue;
		}
		size_t j = i + 1;
		while (j < forces.size() && !forces[j]->b200_is_batch() && forces[j]->b200_class() == f0->b200_class() && f0->b200_same_batch(*forces[j])) ++j;
		const int cnt = (int)(j - i);
		int id = -1;
		std::vector<int> idx;
		switch (f0->b200_class()) {
		case Force::B_TET: {
			for (size_t k = i; k < j; ++k) { const TetBase *t = static_cast<const TetBase *>(forces[k].get()); idx.insert(idx.end(), t->idx, t->idx + 4); }
			double p0, p1, p2; int m;
			const TetBase *t0 = static_cast<const TetBase *>(f0);
			t0->b200_params(p0, p1, p2, m);
			id = admmb_add_tets(ctx, t0->b200_kind(), cnt, idx.data(), p0, p1, p2, m);
			break;
		}
		case Force::B_TRI: {
			if (const FungTriangle *g = dynamic_cast<const FungTriangle *>(f0)) {
				for (size_t k = i; k < j; ++k) { const FungTriangle *t = static_cast<const FungTriangle *>(forces[k].get()); idx.push_back(t->id0); idx.push_back(t->id1); idx.push_back(t->id2); }
				id = admmb_add_tris(ctx, ADMMB_TRI_FUNG, cnt, idx.data(), g->mu, g->limit_min, g->limit_max, 0);
			} else {
				const LimitedTriangleStrain *t0 = static_cast<const LimitedTriangleStrain *>(f0);
				for (size_t k = i; k < j; ++k) { const LimitedTriangleStrain *t = static_cast<const LimitedTriangleStrain *>(forces[k].get()); idx.push_back(t->id0); idx.push_back(t->id1); idx.push_back(t->id2); }
				id = admmb_add_tris(ctx, t0->b200_kind(), cnt, idx.data(), t0->stiffness, t0->limit_min, t0->limit_max, t0->b200_flag());
			}
			break;
		}
		case Force::B_SPRING: {
			std::vector<double> k_(cnt);
			for (size_t k = i; k < j; ++k) { const Spring *s = static_cast<const Spring *>(forces[k].get()); idx.push_back(s->idx0); idx.push_back(s->idx1); k_[k - i] = s->stiffness; }
			id = admmb_add_springs(ctx, cnt, idx.data(), k_.data());
			break;
		}
		case Force::B_BEND: {
			for (size_t k = i; k < j; ++k) { const BendForce *b = static_cast<const BendForce *>(forces[k].get()); idx.insert(idx.end(), b->idx, b->idx + 4); }
			id = admmb_add_bends(ctx, cnt, idx.data(), static_cast<const BendForce *>(f0)->stiffness);
			break;
		}
		case Force::B_STATIC_ANCHOR: {
			for (size_t k = i; k < j; ++k) idx.push_back(static_cast<const StaticAnchor *>(forces[k].get())->idx);
			id = admmb_add_static_anchors(ctx, cnt, idx.data(), -1.0);
			break;
		}
		case Force::B_MOVING_ANCHOR: {
			std::vector<double> pos(3 * (size_t)cnt);
			for (size_t k = i; k < j; ++k) {
				const MovingAnchor *a = static_cast<const MovingAnchor *>(forces[k].get());
				idx.push_back(a->idx);
				for (int c = 0; c < 3; ++c) pos[3 * (k - i) + c] = a->point->pos[c];
			}
			id = admmb_add_moving_anchors(ctx, cnt, idx.data(), pos.data(), -1.0);
			break;
		}
		case Force::B_COLLISION: {
			CollisionForce *c = static_cast<CollisionForce *>(f0);
			c->n_nodes = n; c->Di_rows = 3 * n;
			std::vector<int> kinds; std::vector<double> par;
			for (size_t s = 0; s < c->collisionShapes.size(); ++s) {
				const CollisionShape *sh = c->collisionShapes[s].get();
				kinds.push_back(sh->b200_shape_kind());
				par.push_back(sh->center[0]); par.push_back(sh->center[1]); par.push_back(sh->center[2]); par.push_back(sh->b200_radius());
			}
			id = admmb_add_collision(ctx, (int)kinds.size(), kinds.data(), par.data(), c->weight);
			break;
		}
		}
		if (id < 0) return fail("add forces");
		// explicit weights for classes whose weight is a constructor argument / public member in the reference
		if (f0->b200_class() == Force::B_STATIC_ANCHOR || f0->b200_class() == Force::B_MOVING_ANCHOR) {
			std::vector<double> w(cnt);
			for (size_t k = i; k < j; ++k) w[k - i] = forces[k]->weight;
			admmb_set_batch_weights(ctx, id, w.data());
		}
		BatchRef br;
		br.id = id; br.cls = f0->b200_class(); br.first = i; br.count = (size_t)cnt; br.any_inactive = false; br.soa = false;
		if (br.cls == Force::B_MOVING_ANCHOR) { // what the device holds: the positions just sent, all active
			br.pos.resize(3 * (size_t)cnt);
			br.act.assign(cnt, 1);
			for (size_t k = i; k < j; ++k)
				for (int c = 0; c < 3; ++c) br.pos[3 * (k - i) + c] = static_cast<const MovingAnchor *>(forces[k].get())->point->pos[c];
		}
		batches.push_back(br);
		for (size_t k = i; k < j; ++k) { forces[k]->global_idx = (int)row; row += forces[k]->b200_rows(); }
		i = j;
	}
	// explicit forces: the reference's own classes go to the device, in list order; one unknown subclass keeps all on the host
	device_explicit.clear();
	explicit_on_host = false;
	for (size_t e = 0; e < explicit_forces.size(); ++e) {
		const ExplicitForce *f = explicit_forces[e].get();
		if (!(typeid(*f) == typeid(ExplicitForce) || typeid(*f) == typeid(WindForce))) explicit_on_host = true;
	}
	for (size_t e = 0; e < explicit_forces.size() && !explicit_on_host; ++e) {
		const ExplicitForce *f = explicit_forces[e].get();
		ExplicitRef r;
		r.obj = f;
		for (int c = 0; c < 3; ++c) r.dir[c] = f->direction[c];
		if (const WindForce *w = dynamic_cast<const WindForce *>(f)) r.id = admmb_add_wind(ctx, (int)(w->tris.size() / 3), w->tris.data(), r.dir);
		else if (f->indices.empty()) r.id = admmb_set_gravity(ctx, -1, r.dir);
		else r.id = admmb_add_explicit_subset(ctx, (int)f->indices.size(), f->indices.data(), r.dir);
		if (r.id < 0) return fail("explicit forces");
		device_explicit.push_back(r);
	}
	if (admmb_finalize(ctx, settings.timestep_s) < 0) return fail("finalize");
	// read back the weights the library computed (Force::initialize in the reference)
	for (size_t b = 0; b < batches.size(); ++b) {
		if (batches[b].soa) {
			ForceBatch *fb = static_cast<ForceBatch *>(forces[batches[b].first].get());
			fb->weights.resize(batches[b].count);
			if (admmb_get_batch_weights(ctx, batches[b].id, fb->weights.data()) < 0) return fail("get weights");
			continue;
		}
		std::vector<double> w(batches[b].cls == Force::B_COLLISION ? (size_t)n : batches[b].count);
		if (admmb_get_batch_weights(ctx, batches[b].id, w.data()) < 0) return fail("get weights");
		if (batches[b].cls == Force::B_COLLISION) forces[batches[b].first]->weight = w.empty() ? 0.0 : w[0];
		else for (size_t k = 0; k < batches[b].count; ++k) forces[batches[b].first + k]->weight = w[k];
	}
	for (size_t k = 0; k < forces.size(); ++k)
		if (StaticAnchor *a = dynamic_cast<StaticAnchor *>(forces[k].get()))
			a->pos = Eigen::Vector3d(m_x[a->idx * 3], m_x[a->idx * 3 + 1], m_x[a->idx * 3 + 2]);
	if (settings.verbose >= 1) std::cout << m_x.size() / 3 << " nodes, " << forces.size() << " forces" << std::endl;
	initialized = true;
	return true;
}

inline bool System::step() {
	for (size_t cb = 0; cb < pre_step_callbacks.size(); ++cb) pre_step_callbacks[cb](this);
	if (!initialized || !ctx) return false;
	const double dt = settings.timestep_s;
	// explicit forces (System.cpp:37-39): on the device unless the list holds a user subclass or was edited since initialize()
	if (!explicit_on_host && !explicit_list_unchanged()) {
		for (size_t e = 0; e < device_explicit.size(); ++e) admmb_enable_explicit(ctx, device_explicit[e].id, 0);
		explicit_on_host = true;
	}
	if (explicit_on_host) {
		for (size_t i = 0; i < explicit_forces.size(); ++i) explicit_forces[i]->project(dt, m_x, m_v, m_masses);
	} else {
		for (size_t e = 0; e < device_explicit.size(); ++e) { // `direction` is a public member callers change between steps
			ExplicitRef &r = device_explicit[e];
			const Eigen::Vector3d &d = r.obj->direction;
			if (d[0] == r.dir[0] && d[1] == r.dir[1] && d[2] == r.dir[2]) continue;
			for (int c = 0; c < 3; ++c) r.dir[c] = d[c];
			if (admmb_set_gravity(ctx, r.id, r.dir) < 0) return fail("explicit force direction");
		}
	}
	// control points are owned by the caller and may have moved / been released since the last step: sent when they changed
	for (size_t b = 0; b < batches.size(); ++b) {
		BatchRef &B = batches[b];
		if (B.cls != Force::B_MOVING_ANCHOR) continue;
		bool changed = false;
		B.any_inactive = false;
		for (size_t k = 0; k < B.count; ++k) {
			const MovingAnchor *a = static_cast<const MovingAnchor *>(forces[B.first + k].get());
			const int act = a->point->active ? 1 : 0;
			for (int c = 0; c < 3; ++c)
				if (B.pos[3 * k + c] != a->point->pos[c]) { B.pos[3 * k + c] = a->point->pos[c]; changed = true; }
			if (B.act[k] != act) { B.act[k] = act; changed = true; }
			B.any_inactive = B.any_inactive || !act;
		}
		if (changed && admmb_update_anchor_targets(ctx, B.id, 0, (int)B.count, B.pos.data(), B.act.data()) < 0) return fail("anchor targets");
	}
	pin_state();
	if (admmb_step(ctx, settings.admm_iters, m_x.data(), m_v.data()) < 0) return fail("step");
	// inactive control points follow the mesh (MovingAnchor::project writes point->pos, AnchorForce.cpp:82): read back only then
	for (size_t b = 0; b < batches.size(); ++b) {
		BatchRef &B = batches[b];
		if (B.cls != Force::B_MOVING_ANCHOR || !B.any_inactive) continue;
		std::vector<double> pos(3 * B.count);
		if (admmb_get_anchor_targets(ctx, B.id, 0, (int)B.count, pos.data(), 0) < 0) return fail("anchor targets");
		for (size_t k = 0; k < B.count; ++k) {
			MovingAnchor *a = static_cast<MovingAnchor *>(forces[B.first + k].get());
			if (a->point->active) continue;
			a->point->pos = Eigen::Vector3d(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]);
			for (int c = 0; c < 3; ++c) B.pos[3 * k + c] = pos[3 * k + c]; // the device already holds this value
		}
	}
	elapsed_s += dt;
	return true;
}

inline void System::recompute_weights() {
	if (!ctx) return;
	const int n = (int)(m_x.size() / 3);
	for (size_t b = 0; b < batches.size(); ++b) {
		if (batches[b].soa) {
			const ForceBatch *fb = static_cast<const ForceBatch *>(forces[batches[b].first].get());
			if (fb->weights.size() == batches[b].count) admmb_set_batch_weights(ctx, batches[b].id, fb->weights.data());
			continue;
		}
		std::vector<double> w(batches[b].cls == Force::B_COLLISION ? (size_t)n : batches[b].count);
		if (batches[b].cls == Force::B_COLLISION) std::fill(w.begin(), w.end(), forces[batches[b].first]->weight);
		else for (size_t k = 0; k < batches[b].count; ++k) w[k] = forces[batches[b].first + k]->weight;
		admmb_set_batch_weights(ctx, batches[b].id, w.data());
	}
	if (admmb_recompute_weights(ctx) < 0) fail("recompute_weights");
}

// `admm::Solver` is what the later releases of the library call this class (BASELINE.json north_star uses that name).
typedef System Solver;

} // namespace admm

#endif
