// tests/dropin/scene_runner.cpp -- TEST INFRASTRUCTURE: the drop-in demonstration at the level of the reference's scene
// layer.  The reference's OWN src/SimContext.cpp + src/ForceBuilder.cpp + mclscene (XML loader, TetMesh, TriMesh,
// tetgen, trimesh2) are compiled UNMODIFIED from where they lie under /root/reference, but with
// admm-elastic-sca_b200/host first on the include path, so `#include "System.hpp"` etc. resolve to the B200 host layer and
// SimContext::step() runs on the GPU through libadmm_b200.so.  This file only restates, headlessly, what the GUI sample
// mains do around SimContext (the mains need GLFW): samples/*/ *.cpp setup(), cited below.
//
//   ref_scene_runner <bunnyexpand|windyflag|poordillo|plinkopony> <scene.xml> <frames> <out.bin>
//   ref_scene_runner dump-forces <scene.xml> <out.txt>
//
// out.bin: frames x 3n doubles, m_x after every SimContext::step.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <random>

#include "SimContext.hpp"
#include "CollisionCylinder.hpp"
#include "CollisionForce.hpp"

using namespace admm;

static double smooth(double t, double t0, double t1) { // helper::smooth_move, AnchorForce.hpp:33-40
	double r = (t - t0) / (t1 - t0);
	r = r < 0.0 ? 0.0 : (r > 1.0 ? 1.0 : r);
	return 3.0 * r * r - 2.0 * r * r * r;
}

// `dump-forces`: the force list the scene layer built while loading the XML (ForceBuilder::admm_build_object ->
// build_tetmesh / build_trimesh), one line per ELEMENT whether it sits in the list as its own object (src/ForceBuilder.cpp) or
// inside an SoA batch (host/scene/ForceBuilderBatched.cpp); needs no GPU.  The two runners must print the same text.
static int dump_forces(const char *xml, const char *path) {
	SimContext context;
	context.system->settings.verbose = 0;
	const auto t0 = std::chrono::steady_clock::now();
	context.load(xml);
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	FILE *o = fopen(path, "w");
	if (!o) return 2;
	size_t objects = 0, elements = 0;
	const std::vector<std::shared_ptr<Force> > &F = context.system->forces;
	for (size_t k = 0; k < F.size(); ++k, ++objects) {
		const Force *f = F[k].get();
		if (const TetBatch *b = dynamic_cast<const TetBatch *>(f)) {
			for (size_t e = 0; e < b->count(); ++e, ++elements)
				fprintf(o, "tet %d %.17g %.17g %.17g %d : %d %d %d %d\n", b->kind, b->p[0], b->p[1], b->p[2], b->max_iterations, b->idx[4 * e], b->idx[4 * e + 1], b->idx[4 * e + 2], b->idx[4 * e + 3]);
		} else if (const TetBase *t = dynamic_cast<const TetBase *>(f)) {
			double p0, p1, p2; int m;
			t->b200_params(p0, p1, p2, m);
			fprintf(o, "tet %d %.17g %.17g %.17g %d : %d %d %d %d\n", t->b200_kind(), p0, p1, p2, m, t->idx[0], t->idx[1], t->idx[2], t->idx[3]);
			++elements;
		} else if (const TriangleBatch *b = dynamic_cast<const TriangleBatch *>(f)) {
			for (size_t e = 0; e < b->count(); ++e, ++elements)
				fprintf(o, "tri %.17g %.17g %.17g %d : %d %d %d\n", b->stiffness, b->limit_min, b->limit_max, b->strain_limiting ? 1 : 0, b->idx[3 * e], b->idx[3 * e + 1], b->idx[3 * e + 2]);
		} else if (const LimitedTriangleStrain *t = dynamic_cast<const LimitedTriangleStrain *>(f)) {
			fprintf(o, "tri %.17g %.17g %.17g %d : %d %d %d\n", t->stiffness, t->limit_min, t->limit_max, t->b200_flag(), t->id0, t->id1, t->id2);
			++elements;
		} else if (const BendBatch *b = dynamic_cast<const BendBatch *>(f)) {
			for (size_t e = 0; e < b->count(); ++e, ++elements)
				fprintf(o, "bend %.17g : %d %d %d %d\n", b->stiffness, b->idx[4 * e], b->idx[4 * e + 1], b->idx[4 * e + 2], b->idx[4 * e + 3]);
		} else if (const BendForce *t = dynamic_cast<const BendForce *>(f)) {
			fprintf(o, "bend %.17g : %d %d %d %d\n", t->stiffness, t->idx[0], t->idx[1], t->idx[2], t->idx[3]);
			++elements;
		} else if (const SpringBatch *b = dynamic_cast<const SpringBatch *>(f)) {
			for (size_t e = 0; e < b->count(); ++e, ++elements) fprintf(o, "spring %.17g : %d %d\n", b->stiffness, b->idx[2 * e], b->idx[2 * e + 1]);
		} else if (const Spring *t = dynamic_cast<const Spring *>(f)) {
			fprintf(o, "spring %.17g : %d %d\n", t->stiffness, t->idx0, t->idx1);
			++elements;
		} else {
			fprintf(o, "other\n");
			++elements;
		}
	}
	fprintf(o, "nodes %ld bend_index %d\n", (long)(context.system->m_x.size() / 3), ForceBuilder::bend_index);
	fclose(o);
	printf("load %.3f s, %zu force objects, %zu elements\n", sec, objects, elements);
	return 0;
}

int main(int argc, char **argv) {
	if (argc == 4 && std::string(argv[1]) == "dump-forces") return dump_forces(argv[2], argv[3]);
	if (argc < 5) { fprintf(stderr, "usage: ref_scene_runner <scene> <scene.xml> <frames> <out.bin>\n"); return 2; }
	const std::string which = argv[1];
	const int frames = atoi(argv[3]);
	SimContext context;
	context.system->settings.verbose = 0;
	context.load(argv[2]);
	std::vector<std::shared_ptr<ControlPoint> > hand_cp;
	std::vector<Eigen::Vector3d> hand_start;
	std::vector<std::shared_ptr<Force> > hand_forces;
	if (which == "bunnyexpand") { // samples/bunnyexpand/bunnyexpand.cpp:33-63, scramble with the goldens' fixed seed
		context.initialize();
		std::mt19937 gen(12345);
		std::uniform_real_distribution<double> dis(-0.75, 0.75);
		Eigen::VectorXd &x = context.system->m_x;
		for (int i = 0; i < x.size(); i += 3) { x[i] = dis(gen); x[i + 1] = dis(gen); x[i + 2] = dis(gen); }
	} else if (which == "windyflag") { // samples/windyflag/windyflag.cpp:68-129
		trimesh::TriMesh *cloth_m = context.scene->objects_map["cloth1"]->get_TriMesh().get();
		std::vector<mcl::Param> cloth_params = context.scene->object_params["cloth1"];
		int cloth_height = 0;
		for (size_t i = 0; i < cloth_params.size(); ++i) if (cloth_params[i].tag == "length") cloth_height = cloth_params[i].as_int();
		context.system->forces.push_back(std::shared_ptr<Force>(new StaticAnchor(0)));
		context.system->forces.push_back(std::shared_ptr<Force>(new StaticAnchor(cloth_height)));
		std::vector<int> faces;
		for (size_t f = 0; f < cloth_m->faces.size(); ++f) { faces.push_back(cloth_m->faces[f][0]); faces.push_back(cloth_m->faces[f][1]); faces.push_back(cloth_m->faces[f][2]); }
		std::shared_ptr<ExplicitForce> wind(new WindForce(faces));
		wind->direction = Eigen::Vector3d(10, 0, 2);
		context.system->explicit_forces.push_back(wind);
		context.initialize();
	} else if (which == "poordillo") { // samples/poordillo/poordillo.cpp:129-163
		trimesh::TriMesh *dillo = context.scene->objects_map["dillo"]->get_TriMesh().get();
		const trimesh::point hand_c(.6, .8, .5), foot_c(-.25, -.6, -.1);
		const double rad = 0.2;
		std::vector<int> hand_ids, foot_ids;
		std::vector<std::shared_ptr<ControlPoint> > foot_cp;
		for (size_t i = 0; i < dillo->vertices.size(); ++i) {
			trimesh::point p = dillo->vertices[i];
			if (trimesh::len(p - hand_c) < rad) { hand_ids.push_back((int)i); hand_cp.push_back(std::shared_ptr<ControlPoint>(new ControlPoint(Eigen::Vector3d(p[0], p[1], p[2])))); hand_start.push_back(Eigen::Vector3d(p[0], p[1], p[2])); }
			if (trimesh::len(p - foot_c) < rad) { foot_ids.push_back((int)i); foot_cp.push_back(std::shared_ptr<ControlPoint>(new ControlPoint(Eigen::Vector3d(p[0], p[1], p[2])))); }
		}
		for (size_t i = 0; i < hand_ids.size(); ++i) { hand_forces.push_back(std::shared_ptr<Force>(new MovingAnchor(hand_ids[i], hand_cp[i]))); context.system->forces.push_back(hand_forces.back()); }
		for (size_t i = 0; i < foot_ids.size(); ++i) context.system->forces.push_back(std::shared_ptr<Force>(new MovingAnchor(foot_ids[i], foot_cp[i])));
		context.initialize();
	} else if (which == "plinkopony") { // samples/plinkopony/plinkopony.cpp:53-96
		std::vector<std::shared_ptr<CollisionShape> > shapes;
		std::unordered_map<std::string, std::vector<mcl::Param> >::iterator it = context.scene->object_params.begin();
		for (; it != context.scene->object_params.end(); ++it) {
			if (it->first[0] != 'c') continue;
			double r = 1.f;
			Eigen::Vector3d center(0, 0, 0), scale(1, 1, 1);
			for (size_t i = 0; i < it->second.size(); ++i) {
				if (it->second[i].tag == "scale_copy") { trimesh::vec v = it->second[i].as_vec3(); scale = Eigen::Vector3d(v[0], v[1], v[2]); }
				else if (it->second[i].tag == "translate_copy") { trimesh::vec v = it->second[i].as_vec3(); center = Eigen::Vector3d(v[0], v[1], v[2]); }
				else if (it->second[i].tag == "radius") r = it->second[i].as_double();
			}
			shapes.push_back(std::shared_ptr<CollisionShape>(new CollisionCylinder(center, scale, r)));
		}
		context.system->forces.push_back(std::shared_ptr<Force>(new CollisionForce(shapes)));
		context.initialize();
	} else { fprintf(stderr, "unknown scene %s\n", which.c_str()); return 2; }

	context.settings.run_realtime = false; // one System::step per SimContext::step (SimContext.cpp:198-210), whatever the XML says
	std::ofstream out(argv[4], std::ios::binary);
	const double dt = context.system->settings.timestep_s;
	for (int f = 0; f < frames; ++f) {
		if (which == "poordillo") {
			// headless stand-in for the GUI interaction: the hand is dragged 2 units in +x over 1.2 s with smooth_move
			// (poordillo.cpp:51-58) and released at frame 20 as the H key does (poordillo.cpp:196-204)
			if (f < 20) {
				const double s = smooth(f * dt, 0.0, 1.2);
				for (size_t i = 0; i < hand_cp.size(); ++i) hand_cp[i]->pos = hand_start[i] + s * Eigen::Vector3d(2.0, 0.0, 0.0);
			}
			if (f == 20) {
				for (size_t i = 0; i < hand_cp.size(); ++i) { hand_cp[i]->active = false; hand_forces[i]->weight = 0.0; }
				context.system->recompute_weights();
			}
		}
		if (!context.step(context.scene.get(), 0.f)) { fprintf(stderr, "step %d failed\n", f); return 1; }
		out.write((const char *)context.system->m_x.data(), sizeof(double) * context.system->m_x.size());
	}
	printf("%s: %ld nodes, %zu forces, %d frames on the B200 solver\n", which.c_str(), (long)context.system->m_x.size() / 3, context.system->forces.size(), frames);
	return 0;
}
