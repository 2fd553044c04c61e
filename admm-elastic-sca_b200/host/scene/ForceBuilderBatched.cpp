// ForceBuilderBatched.cpp -- batched (SoA) definitions of the reference scene layer's force factory.
//
// A maintainer of the reference swaps THIS file for src/ForceBuilder.cpp in the build of the scene layer (SimContext +
// ForceBuilder + mclscene); src/ForceBuilder.hpp -- the class declaration, admm_build_object, the mass lumping -- stays
// the reference's own and is included from where it lies.  Same static API, same XML vocabulary, same error behaviour
// (reference: src/ForceBuilder.hpp:50-68, src/ForceBuilder.cpp:76-446), but
//   * one admm::TetBatch / TriangleBatch / BendBatch / SpringBatch per (object, force) instead of one heap-allocated Force
//     per element (ForceBuilder.cpp:126-129,168-170,251-255,310-313,346-349,362-365,427-430): the corner indices go into one
//     flat array that System::initialize() hands to admmb_add_* unchanged;
//   * hinges are deduplicated through a hash set of their sorted vertex quadruple instead of a linear scan that sorts
//     every stored signature again for every candidate (isUniqueHinge, ForceBuilder.cpp:55-74: O(H^2 log) for H hinges);
//     the hinges, their vertex order and their order in the list are the reference's.
// The element order inside a batch is the reference's loop order, so the rows of z / u, the weights and the trajectories
// are identical to a scene built by src/ForceBuilder.cpp (tests/test_dropin_scene_layer_gpu.py runs both).
#include "ForceBuilder.hpp"

#include <array>
#include <unordered_set>

using namespace admm;

namespace {

struct QuadHash {
	size_t operator()(const std::array<int, 4> &q) const {
		unsigned long long h = 0x9E3779B97F4A7C15ull;
		for (int k = 0; k < 4; ++k) { h ^= (unsigned long long)(unsigned)q[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xFF51AFD7ED558CCDull; }
		return (size_t)(h ^ (h >> 32));
	}
};
struct PairHash {
	size_t operator()(const std::pair<int, int> &p) const {
		unsigned long long h = ((unsigned long long)(unsigned)p.first << 32) | (unsigned)p.second;
		h *= 0xFF51AFD7ED558CCDull;
		return (size_t)(h ^ (h >> 29));
	}
};

bool need_param(mcl::Component &force, const char *name) {
	if (force.exists(name)) return true;
	std::cerr << "\n**ForceBuilder Error: force \"" << force.name << "\" needs a " << name << " parameter." << std::endl;
	return false;
}

// the vertex of face `other` that face `f` does not have (the fourth hinge vertex); the reference ends the process when
// the two faces do not share exactly two vertices (ForceBuilder.cpp:24-52)
int opposite_vertex(const trimesh::TriMesh &mesh, int other, int f) {
	int shared = 0, lone = -1;
	for (int i = 0; i < 3; ++i) {
		const int v = mesh.faces[other].v[i];
		const bool in_f = (v == mesh.faces[f].v[0] || v == mesh.faces[f].v[1] || v == mesh.faces[f].v[2]);
		if (in_f) ++shared;
		else if (lone < 0) lone = v;
	}
	if (shared != 2) {
		std::cout << "Error in getUniqueVert: two input faces do not share 2 verts!\n";
		exit(0);
	}
	return lone;
}

} // namespace

bool ForceBuilder::build_trimesh(std::shared_ptr<trimesh::TriMesh> mesh, mcl::Component &force,
                                 std::vector<std::shared_ptr<Force> > *sys_forces, int idx_offset) {
	const std::string force_type = mcl::parse::to_lower(force.type);
	mesh->need_faces();
	const int nf = (int)mesh->faces.size();
	if (nf == 0) return true; // the reference looks at the force only inside its loop over the faces

	if (force_type == "lineartrianglestrain" || force_type == "trianglestrain") {
		trimesh::vec2 limit(0.f, 9999999.f);
		if (force.exists("limit")) limit = force["limit"].as_vec2();
		if (!need_param(force, "stiffness")) return false;
		std::shared_ptr<TriangleBatch> batch(new TriangleBatch(force["stiffness"].as_double(), limit[0], limit[1]));
		batch->idx.resize(3 * (size_t)nf);
		for (int f = 0; f < nf; ++f)
			for (int c = 0; c < 3; ++c) batch->idx[3 * (size_t)f + c] = mesh->faces[f].v[c] + idx_offset;
		sys_forces->push_back(batch);
		return true;
	}

	if (force_type == "bend") {
		if (!need_param(force, "stiffness")) return false;
		if (mesh->across_edge.empty()) mesh->need_across_edge();
		std::shared_ptr<BendBatch> batch(new BendBatch(force["stiffness"].as_double()));
		std::unordered_set<std::array<int, 4>, QuadHash> seen;
		seen.reserve(2 * (size_t)nf);
		// corner i of face f and the face across the edge opposite to it: hinge vertices in the reference's order
		// (p_i, the other face's lone vertex, then the shared edge as listed in ForceBuilder.cpp:163-211)
		static const int edge_order[3][2] = { { 2, 1 }, { 0, 2 }, { 1, 0 } };
		for (int f = 0; f < nf; ++f) {
			const trimesh::TriMesh::Face &face = mesh->faces[f];
			for (int i = 0; i < 3; ++i) {
				const int other = mesh->across_edge[f][i];
				if (other < 0) continue;
				const std::array<int, 4> hv = { { face.v[i] + idx_offset, opposite_vertex(*mesh, other, f) + idx_offset,
				                                  face.v[edge_order[i][0]] + idx_offset, face.v[edge_order[i][1]] + idx_offset } };
				std::array<int, 4> key = hv;
				std::sort(key.begin(), key.end());
				if (!seen.insert(key).second) continue;
				batch->idx.insert(batch->idx.end(), hv.begin(), hv.end());
				bend_index += 1;
			}
		}
		sys_forces->push_back(batch);
		return true;
	}

	if (force_type == "spring") {
		trimesh::vec2 limit(-1.f, -1.f);
		if (force.exists("limit")) limit = force["limit"].as_vec2();
		if (!need_param(force, "stiffness")) return false;
		if (limit[0] >= 0.f) {
			std::cout << "TODO: ForceBuilder::build_trimesh with limited springs" << std::endl;
			return false;
		}
		std::shared_ptr<SpringBatch> batch(new SpringBatch(force["stiffness"].as_double()));
		std::unordered_set<std::pair<int, int>, PairHash> seen;
		seen.reserve(3 * (size_t)nf);
		static const int ends[3][2] = { { 0, 1 }, { 0, 2 }, { 1, 2 } };
		for (int f = 0; f < nf; ++f)
			for (int e = 0; e < 3; ++e) {
				const int a = mesh->faces[f].v[ends[e][0]] + idx_offset, b = mesh->faces[f].v[ends[e][1]] + idx_offset;
				if (!seen.insert(std::make_pair(std::min(a, b), std::max(a, b))).second) continue;
				batch->idx.push_back(a); // end points in the order of the face that introduced the edge
				batch->idx.push_back(b);
			}
		sys_forces->push_back(batch);
		return true;
	}

	if (force_type != "constforce") {
		std::cout << "TODO: ForceBuilder::build_trimesh with force: " << force_type << std::endl;
		return false;
	}
	return true;
}

bool ForceBuilder::build_tetmesh(std::shared_ptr<mcl::TetMesh> mesh, mcl::Component &force,
                                 std::vector<std::shared_ptr<Force> > *sys_forces, int idx_offset) {
	const std::string force_type = mcl::parse::to_lower(force.type);
	const size_t nt = mesh->tets.size();
	if (nt == 0) return true;

	std::shared_ptr<TetBatch> batch;
	if (force_type == "lineartetstrain") {
		if (!need_param(force, "stiffness")) return false;
		// weight_scale: the reference parses it and passes it on, but its use is commented out (TetForce.cpp:116), so it has no effect
		batch.reset(new TetBatch(ADMMB_TET_LINEAR_STRAIN, force["stiffness"].as_double()));
	} else if (force_type == "neohookeantet" || force_type == "stvktet") {
		assert(force.exists("mu"));
		assert(force.exists("lambda"));
		int max_iters = 10;
		if (force.exists("max_iterations")) max_iters = force["max_iterations"].as_int();
		batch.reset(new TetBatch(force_type == "stvktet" ? ADMMB_TET_STVK : ADMMB_TET_NEOHOOKEAN, force["mu"].as_double(), force["lambda"].as_double(), 0.0, max_iters));
	} else if (force_type == "volpres") {
		if (!need_param(force, "stiffness") || !need_param(force, "range_min") || !need_param(force, "range_max")) return false;
		batch.reset(new TetBatch(ADMMB_TET_VOLUME, force["stiffness"].as_double(), force["range_min"].as_double(), force["range_max"].as_double()));
	} else if (force_type != "constforce") {
		std::cout << "TODO: ForceBuilder::build_tetmesh with force: " << force_type << std::endl;
		return false;
	} else {
		return true;
	}
	batch->idx.resize(4 * nt);
	for (size_t t = 0; t < nt; ++t)
		for (int c = 0; c < 4; ++c) batch->idx[4 * t + c] = mesh->tets[t].v[c] + idx_offset;
	sys_forces->push_back(batch);
	return true;
}

// static members (ForceBuilder.hpp:62-68)
std::shared_ptr<admm::System> ForceBuilder::system;
std::unordered_map<std::string, mcl::Component> *ForceBuilder::force_param_map;
int ForceBuilder::index_offset;
int ForceBuilder::num_objects;
int ForceBuilder::bend_index = 0;
std::unordered_map<int, std::pair<int, int> > *ForceBuilder::system_to_scene_map;
