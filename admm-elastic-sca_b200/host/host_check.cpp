// host_check.cpp -- drives the C++ host layer (admm::System and the Force classes of admm_b200_host.hpp) the way
// src/ForceBuilder.cpp does -- one heap Force object per element pushed on System::forces -- from a plain-text
// scene written by tests/test_host_cpp.py, and dumps m_x after every frame.  Used by the -m gpu tests to check
// that the C++ surface produces the same numbers as the C ABI driven directly.
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "System.hpp"
#include "TetForce.hpp"
#include "TriangleForce.hpp"
#include "BendForce.hpp"
#include "AnchorForce.hpp"
#include "CollisionForce.hpp"
#include "CollisionSphere.hpp"
#include "CollisionCylinder.hpp"
#include "CollisionFloor.hpp"

using namespace admm;

// a user-defined ExplicitForce subclass (what the reference lets callers write): cannot run on the device, so its
// presence makes the host layer apply EVERY explicit force on the host in list order (argv[3] == "userforce")
class UserGravity : public ExplicitForce {
public:
	UserGravity(Eigen::Vector3d d) : ExplicitForce(d) {}
	void project(double dt, Eigen::VectorXd &x, Eigen::VectorXd &v, Eigen::VectorXd &m) const { ExplicitForce::project(dt, x, v, m); }
};

int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: host_check scene.txt out.bin [userforce|soa]\n"); return 2; }
	const bool userforce = argc > 3 && std::string(argv[3]) == "userforce";
	// "soa": runs of tets / strain triangles / bends / uniform springs go into ONE ForceBatch object each (what
	// host/scene/ForceBuilderBatched.cpp builds) instead of one Force object per element; everything else as before
	const bool soa = argc > 3 && std::string(argv[3]) == "soa";
	std::vector<std::shared_ptr<ForceBatch> > soa_batches;
	std::ifstream in(argv[1]);
	if (!in) { fprintf(stderr, "cannot open %s\n", argv[1]); return 2; }
	Solver system; // the later releases' name of admm::System (alias)
	system.settings.verbose = 0;
	int frames, n;
	in >> system.settings.timestep_s >> system.settings.admm_iters >> frames >> n;
	Eigen::VectorXd x(3 * n), m(3 * n), x_after(3 * n);
	for (int i = 0; i < n; ++i) { double mi; in >> x[3 * i] >> x[3 * i + 1] >> x[3 * i + 2] >> mi; m[3 * i] = m[3 * i + 1] = m[3 * i + 2] = mi; }
	int has_after;
	in >> has_after;
	if (has_after) for (int i = 0; i < 3 * n; ++i) in >> x_after[i];
	system.add_nodes(x, m);
	int nbatches;
	in >> nbatches;
	std::vector<std::shared_ptr<ControlPoint> > cps;
	for (int b = 0; b < nbatches; ++b) {
		std::string type;
		int kind, count, maxit, flag;
		double p0, p1, p2;
		in >> type >> kind >> count >> p0 >> p1 >> p2 >> maxit >> flag;
		if (soa && (type == "tets" || (type == "tris" && kind == 0) || type == "bends" || type == "springs")) {
			std::shared_ptr<ForceBatch> fb;
			const int corners = type == "tris" ? 3 : (type == "springs" ? 2 : 4);
			std::vector<int> idx((size_t)corners * count);
			std::vector<double> ks(count, 0.0);
			for (int e = 0; e < count; ++e) {
				for (int c = 0; c < corners; ++c) in >> idx[(size_t)corners * e + c];
				if (type == "springs") in >> ks[e];
			}
			bool uniform = true;
			for (int e = 1; e < count; ++e) uniform = uniform && ks[e] == ks[0];
			static const int tet_kinds[4] = { ADMMB_TET_LINEAR_STRAIN, ADMMB_TET_NEOHOOKEAN, ADMMB_TET_STVK, ADMMB_TET_VOLUME };
			if (type == "tets") fb.reset(new TetBatch(tet_kinds[kind], p0, p1, p2, (kind == 1 || kind == 2) ? maxit : 0));
			else if (type == "tris") fb.reset(new TriangleBatch(p0, p1, p2, flag != 0));
			else if (type == "bends") fb.reset(new BendBatch(p0));
			else if (uniform && count > 0) fb.reset(new SpringBatch(ks[0]));
			if (fb) {
				fb->idx = idx;
				soa_batches.push_back(fb);
				system.add_force(fb);
			} else {
				for (int e = 0; e < count; ++e) system.add_force(std::shared_ptr<Force>(new Spring(idx[2 * e], idx[2 * e + 1], ks[e])));
			}
			continue;
		}
		for (int e = 0; e < count; ++e) {
			std::shared_ptr<Force> f;
			int i0, i1, i2, i3;
			if (type == "tets") {
				in >> i0 >> i1 >> i2 >> i3;
				if (kind == 0) f.reset(new LinearTetStrain(i0, i1, i2, i3, p0));
				else if (kind == 1) f.reset(new HyperElasticTet(i0, i1, i2, i3, p0, p1, maxit, "nh"));
				else if (kind == 2) f.reset(new HyperElasticTet(i0, i1, i2, i3, p0, p1, maxit, "stvk"));
				else f.reset(new TetVolume(i0, i1, i2, i3, p0, p1, p2));
			} else if (type == "tris") {
				in >> i0 >> i1 >> i2;
				if (kind == 0) f.reset(new LimitedTriangleStrain(i0, i1, i2, p0, p1, p2, flag != 0));
				else if (kind == 1) f.reset(new TriArea(i0, i1, i2, p0, flag, p1, p2));
				else f.reset(new FungTriangle(i0, i1, i2, p0, p1, p2));
			} else if (type == "springs") {
				double k;
				in >> i0 >> i1 >> k;
				f.reset(new Spring(i0, i1, k));
			} else if (type == "bends") {
				in >> i0 >> i1 >> i2 >> i3;
				f.reset(new BendForce(i0, i1, i2, i3, p0));
			} else if (type == "static_anchors") {
				in >> i0;
				f.reset(new StaticAnchor(i0, p0));
			} else if (type == "moving_anchors") {
				double px, py, pz;
				in >> i0 >> px >> py >> pz;
				std::shared_ptr<ControlPoint> cp(new ControlPoint(Eigen::Vector3d(px, py, pz)));
				cps.push_back(cp);
				f.reset(new MovingAnchor(i0, cp, p0));
			} else { fprintf(stderr, "unknown batch type %s\n", type.c_str()); return 2; }
			if (e % 2) system.add_force(f); else system.forces.push_back(f); // add_force: alias of forces.push_back
		}
		if (type == "collision") { // count = number of shapes
			fprintf(stderr, "collision batches use the 'shapes' record\n");
			return 2;
		}
	}
	int nshapes;
	in >> nshapes;
	if (nshapes > 0) {
		std::vector<std::shared_ptr<CollisionShape> > shapes;
		double weight;
		in >> weight;
		for (int s = 0; s < nshapes; ++s) {
			int kind; double cx, cy, cz, r;
			in >> kind >> cx >> cy >> cz >> r;
			if (kind == 0) shapes.push_back(std::shared_ptr<CollisionShape>(new CollisionSphere(Eigen::Vector3d(cx, cy, cz), r)));
			else if (kind == 1) shapes.push_back(std::shared_ptr<CollisionShape>(new CollisionCylinder(Eigen::Vector3d(cx, cy, cz), Eigen::Vector3d(1, 1, 1), r)));
			else shapes.push_back(std::shared_ptr<CollisionShape>(new CollisionFloor(Eigen::Vector3d(cx, cy, cz))));
		}
		system.forces.push_back(std::shared_ptr<Force>(new CollisionForce(shapes, weight)));
	}
	int nexplicit;
	in >> nexplicit;
	for (int k = 0; k < nexplicit; ++k) {
		std::string type;
		double dx, dy, dz;
		in >> type >> dx >> dy >> dz;
		if (type == "gravity" && userforce) system.explicit_forces.push_back(std::shared_ptr<ExplicitForce>(new UserGravity(Eigen::Vector3d(dx, dy, dz))));
		else if (type == "gravity") system.explicit_forces.push_back(std::shared_ptr<ExplicitForce>(new ExplicitForce(Eigen::Vector3d(dx, dy, dz))));
		else {
			int nt;
			in >> nt;
			std::vector<int> tris(3 * nt);
			for (int t = 0; t < 3 * nt; ++t) in >> tris[t];
			std::shared_ptr<WindForce> w(new WindForce(tris));
			w->direction = Eigen::Vector3d(dx, dy, dz);
			system.explicit_forces.push_back(w);
		}
	}
	if (!in) { fprintf(stderr, "scene file truncated\n"); return 2; }
	if (!system.initialize()) return 1;
	if (has_after) system.m_x = x_after;
	for (size_t b = 0; b < soa_batches.size(); ++b) { // the weights initialize() computed are visible per element, like Force::weight
		if (soa_batches[b]->weights.size() != soa_batches[b]->count()) { fprintf(stderr, "batch %zu: weights not read back\n", b); return 1; }
		for (size_t e = 0; e < soa_batches[b]->count(); ++e)
			if (!(soa_batches[b]->weights[e] > 0.0)) { fprintf(stderr, "batch %zu: weight %zu = %g\n", b, e, soa_batches[b]->weights[e]); return 1; }
	}
	if (soa) system.recompute_weights(); // pushes the (unchanged) per-element weights back and refactors: same numbers
	FILE *out = fopen(argv[2], "wb");
	for (int f = 0; f < frames; ++f) {
		if (!system.step()) return 1;
		fwrite(system.m_x.data(), sizeof(double), 3 * n, out);
	}
	fclose(out);
	printf("host_check: %d nodes, %zu force objects (%zu SoA batches), %d frames, elapsed %.3f s\n", n, system.forces.size(), soa_batches.size(), frames, system.elapsed_s);
	return 0;
}
