// oracle/scene_export.cpp -- TEST INFRASTRUCTURE.  Loads the reference's four shipped scenes with the reference's
// OWN scene layer (src/SimContext.cpp + src/ForceBuilder.cpp + mclscene, compiled from where they lie by
// oracle/Makefile target `scene_export`), re-states the GUI samples' setup() headlessly (the sample mains need
// GLFW), and writes the resulting admm::System -- node positions as the loader produced them (float-rounded,
// transformed), masses, every Force with its indices and material, explicit forces -- as plain text that
// tests/golden/make_golden.py turns into scene fixtures.  Nothing here reaches the product.
//
//   scene_export <bunnyexpand|windyflag|poordillo|plinkopony> out.txt
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <random>

#include "SimContext.hpp"
#include "CollisionCylinder.hpp"
#include "CollisionForce.hpp"

using namespace admm;

static void write_system(const char *path, SimContext &ctx, const Eigen::VectorXd *x_after) {
	System &S = *ctx.system;
	std::ofstream o(path);
	o << std::setprecision(17);
	const int n = (int)(S.m_x.size() / 3);
	o << "settings " << S.settings.timestep_s << " " << S.settings.admm_iters << "\n";
	o << "nodes " << n << "\n";
	for (int i = 0; i < n; ++i) o << S.m_x[3 * i] << " " << S.m_x[3 * i + 1] << " " << S.m_x[3 * i + 2] << " " << S.m_masses[3 * i] << "\n";
	if (x_after) {
		o << "x_after " << n << "\n";
		for (int i = 0; i < 3 * n; ++i) o << (*x_after)[i] << (i % 3 == 2 ? "\n" : " ");
	}
	o << "forces " << S.forces.size() << "\n";
	for (size_t i = 0; i < S.forces.size(); ++i) {
		Force *f = S.forces[i].get();
		if (LinearTetStrain *t = dynamic_cast<LinearTetStrain *>(f)) o << "tet 0 " << t->idx[0] << " " << t->idx[1] << " " << t->idx[2] << " " << t->idx[3] << " " << t->stiffness << " 0 0 0\n";
		else if (HyperElasticTet *t = dynamic_cast<HyperElasticTet *>(f)) o << "tet " << (t->type == 1 ? 2 : 1) << " " << t->idx[0] << " " << t->idx[1] << " " << t->idx[2] << " " << t->idx[3] << " " << t->mu << " " << t->lambda << " 0 " << t->solver->settings_.maxIter << "\n";
		else if (TetVolume *t = dynamic_cast<TetVolume *>(f)) o << "tet 3 " << t->idx[0] << " " << t->idx[1] << " " << t->idx[2] << " " << t->idx[3] << " " << t->stiffness << " " << t->limit_min << " " << t->limit_max << " 0\n";
		else if (TriArea *t = dynamic_cast<TriArea *>(f)) o << "tri 1 " << t->id0 << " " << t->id1 << " " << t->id2 << " " << t->stiffness << " " << t->limit_min << " " << t->limit_max << " " << t->iters << "\n";
		else if (LimitedTriangleStrain *t = dynamic_cast<LimitedTriangleStrain *>(f)) o << "tri 0 " << t->id0 << " " << t->id1 << " " << t->id2 << " " << t->stiffness << " " << t->limit_min << " " << t->limit_max << " " << (t->strain_limiting ? 1 : 0) << "\n";
		else if (BendForce *b = dynamic_cast<BendForce *>(f)) o << "bend " << b->idx[0] << " " << b->idx[1] << " " << b->idx[2] << " " << b->idx[3] << " " << b->stiffness << "\n";
		else if (Spring *s = dynamic_cast<Spring *>(f)) o << "spring " << s->idx0 << " " << s->idx1 << " " << s->stiffness << "\n";
		else if (StaticAnchor *a = dynamic_cast<StaticAnchor *>(f)) o << "sanchor " << a->idx << " " << a->weight << "\n";
		else if (MovingAnchor *a = dynamic_cast<MovingAnchor *>(f)) o << "manchor " << a->idx << " " << a->weight << " " << a->point->pos[0] << " " << a->point->pos[1] << " " << a->point->pos[2] << "\n";
		else if (CollisionForce *c = dynamic_cast<CollisionForce *>(f)) {
			o << "collision " << c->weight << " " << c->collisionShapes.size() << "\n";
			for (size_t s = 0; s < c->collisionShapes.size(); ++s) {
				CollisionCylinder *cy = dynamic_cast<CollisionCylinder *>(c->collisionShapes[s].get());
				o << "shape 1 " << cy->center[0] << " " << cy->center[1] << " " << cy->center[2] << " " << cy->radius << "\n";
			}
		} else o << "unknown\n";
	}
	o << "explicit " << S.explicit_forces.size() << "\n";
	for (size_t i = 0; i < S.explicit_forces.size(); ++i) {
		ExplicitForce *e = S.explicit_forces[i].get();
		if (WindForce *w = dynamic_cast<WindForce *>(e)) {
			o << "wind " << w->direction[0] << " " << w->direction[1] << " " << w->direction[2] << " " << w->tris.size() / 3 << "\n";
			for (size_t t = 0; t < w->tris.size(); ++t) o << w->tris[t] << (t % 3 == 2 ? "\n" : " ");
		} else o << "gravity " << e->direction[0] << " " << e->direction[1] << " " << e->direction[2] << "\n";
	}
}

int main(int argc, char **argv) {
	if (argc < 3) { fprintf(stderr, "usage: scene_export <scene> out.txt\n"); return 2; }
	const std::string which = argv[1];
	const std::string root = SRC_ROOT_DIR;
	SimContext context;
	context.system->settings.verbose = 0;
	if (which == "bunnyexpand") { // samples/bunnyexpand/bunnyexpand.cpp:33-63, scramble with a FIXED seed
		context.load(root + "/samples/bunnyexpand/bunnyexpand.xml");
		context.initialize();
		std::mt19937 gen(12345);
		std::uniform_real_distribution<double> dis(-0.75, 0.75);
		Eigen::VectorXd xa = context.system->m_x;
		for (int i = 0; i < xa.size(); i += 3) { xa[i] = dis(gen); xa[i + 1] = dis(gen); xa[i + 2] = dis(gen); }
		write_system(argv[2], context, &xa);
	} else if (which == "windyflag") { // samples/windyflag/windyflag.cpp:68-129
		context.load(root + "/samples/windyflag/cloth.xml");
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
		write_system(argv[2], context, 0);
	} else if (which == "poordillo") { // samples/poordillo/poordillo.cpp:129-163
		context.load(root + "/samples/poordillo/poordillo.xml");
		trimesh::TriMesh *dillo = context.scene->objects_map["dillo"]->get_TriMesh().get();
		const trimesh::point hand_c(.6, .8, .5), foot_c(-.25, -.6, -.1);
		const double rad = 0.2;
		std::vector<int> hand_ids, foot_ids;
		std::vector<std::shared_ptr<ControlPoint> > hand_cp, foot_cp;
		for (size_t i = 0; i < dillo->vertices.size(); ++i) {
			trimesh::point p = dillo->vertices[i];
			if (trimesh::len(p - hand_c) < rad) { hand_ids.push_back((int)i); hand_cp.push_back(std::shared_ptr<ControlPoint>(new ControlPoint(Eigen::Vector3d(p[0], p[1], p[2])))); }
			if (trimesh::len(p - foot_c) < rad) { foot_ids.push_back((int)i); foot_cp.push_back(std::shared_ptr<ControlPoint>(new ControlPoint(Eigen::Vector3d(p[0], p[1], p[2])))); }
		}
		for (size_t i = 0; i < hand_ids.size(); ++i) context.system->forces.push_back(std::shared_ptr<Force>(new MovingAnchor(hand_ids[i], hand_cp[i])));
		for (size_t i = 0; i < foot_ids.size(); ++i) context.system->forces.push_back(std::shared_ptr<Force>(new MovingAnchor(foot_ids[i], foot_cp[i])));
		context.initialize();
		write_system(argv[2], context, 0);
		printf("poordillo: %zu hand anchors, %zu foot anchors\n", hand_ids.size(), foot_ids.size());
	} else if (which == "plinkopony") { // samples/plinkopony/plinkopony.cpp:53-96
		context.load(root + "/samples/plinkopony/plinko.xml");
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
		write_system(argv[2], context, 0);
	} else { fprintf(stderr, "unknown scene %s\n", which.c_str()); return 2; }
	printf("%s: %ld nodes, %zu forces\n", which.c_str(), (long)context.system->m_x.size() / 3, context.system->forces.size());
	return 0;
}
