// grayshift_b200/host/capi.cpp -- C entry points of the host layer (include/grayshift_host.h).
#include <cstring>
#include <stdexcept>

#include "../../include/grayshift_host.h"
#include "grayshift.hpp"

using namespace grayshift;

struct gsh_scene {
	HostScene scene;
	std::unique_ptr<FlatScene> flat;
	gs_camera camera{};
};

namespace {
thread_local std::string g_err;
template <class F> int guard(F&& f) {
	try { f(); return 0; }
	catch (const std::exception& e) { g_err = e.what(); return 1; }
	catch (...) { g_err = "unknown error"; return 1; }
}
}  // namespace

extern "C" {

const char* gsh_last_error(void) { return g_err.c_str(); }

gsh_scene* gsh_scene_build(const char* name, int width, int spp, int max_depth, uint64_t scene_seed, int include_media, int n_spheres,
                           const char* asset_dir) {
	gsh_scene* out = nullptr;
	guard([&] {
		SceneOptions o;
		o.width = width; o.spp = spp; o.max_depth = max_depth; o.scene_seed = scene_seed; o.include_media = include_media != 0;
		o.n_spheres = n_spheres;
		if (asset_dir && *asset_dir) o.asset_dir = asset_dir;
		auto s = std::make_unique<gsh_scene>();
		s->scene = build_scene(name ? name : "", o);
		out = s.release();
	});
	return out;
}

gsh_scene* gsh_scene_load(const char* graph_path, int include_media) {
	gsh_scene* out = nullptr;
	guard([&] {
		auto s = std::make_unique<gsh_scene>();
		s->scene = load_graph(graph_path ? graph_path : "", include_media != 0);
		out = s.release();
	});
	return out;
}

int gsh_scene_dump(const gsh_scene* scene, const char* graph_path) {
	return guard([&] {
		if (!scene || !graph_path) throw std::runtime_error("null argument");
		dump_graph(scene->scene, graph_path);
	});
}

void gsh_scene_free(gsh_scene* scene) { delete scene; }

int gsh_scene_flatten(gsh_scene* scene, int max_leaf_size, int reference_tree) {
	return guard([&] {
		if (!scene) throw std::runtime_error("null scene");
		FlattenOptions o;
		if (max_leaf_size > 0) o.max_leaf_size = max_leaf_size;
		o.reference_tree = reference_tree != 0;
		scene->flat = flatten(scene->scene.world, o);
		scene->camera = scene->scene.camera.to_abi(*scene->flat);
	});
}

const gs_scene_desc* gsh_scene_desc(gsh_scene* scene) {
	if (!scene) { g_err = "null scene"; return nullptr; }
	if (!scene->flat && gsh_scene_flatten(scene, 0, 0) != 0) return nullptr;
	return &scene->flat->desc;
}

int gsh_scene_camera(gsh_scene* scene, gs_camera* camera_out, gs_sample_settings* settings_out) {
	return guard([&] {
		if (!scene) throw std::runtime_error("null scene");
		if (!scene->flat) { if (gsh_scene_flatten(scene, 0, 0) != 0) throw std::runtime_error(g_err); }
		if (camera_out) {
			*camera_out = scene->camera;
			// image size / depth may have been changed through gsh_scene_set_sampling
			camera_out->max_depth = scene->scene.camera.max_depth;
		}
		if (settings_out) *settings_out = scene->scene.camera.settings_abi();
	});
}

int gsh_scene_set_sampling(gsh_scene* scene, double confidence, double tolerance, uint32_t batch_size, uint32_t max_samples, uint32_t max_depth) {
	return guard([&] {
		if (!scene) throw std::runtime_error("null scene");
		scene->scene.camera.sample_settings = SampleSettings{confidence, tolerance, batch_size, max_samples};
		if (max_depth) scene->scene.camera.max_depth = max_depth;
	});
}

int gsh_scene_render_ppm(gsh_scene* scene, const char* path, int device, uint64_t seed, gs_stats* stats) {
	return guard([&] {
		if (!scene || !path) throw std::runtime_error("null argument");
		std::FILE* f = std::fopen(path, "w");
		if (!f) throw std::runtime_error(std::string("cannot create ") + path);
		try { scene->scene.camera.render(scene->scene.world, f, device, seed, stats); }
		catch (...) { std::fclose(f); throw; }
		std::fclose(f);
	});
}

int gsh_bvh_build(const double* bounds, size_t n, int max_leaf_size, gs_bvh_node* nodes_out, size_t* n_nodes_out, uint32_t* order_out) {
	return guard([&] {
		if (!bounds || !nodes_out || !n_nodes_out || !order_out) throw std::runtime_error("null argument");
		std::vector<double> b(bounds, bounds + n * 6);
		std::vector<gs_bvh_node> nodes;
		std::vector<uint32_t> order;
		build_bvh_sah(b, max_leaf_size > 0 ? max_leaf_size : 4, nodes, order);
		std::memcpy(nodes_out, nodes.data(), nodes.size() * sizeof(gs_bvh_node));
		std::memcpy(order_out, order.data(), order.size() * sizeof(uint32_t));
		*n_nodes_out = nodes.size();
	});
}

void gsh_perlin_table(uint32_t seed, uint8_t* out256) { perlin_permutation_table(seed, out256); }

double gsh_scene_rng(uint64_t seed, int k) {
	seed_scene_rng(seed);
	double v = 0;
	for (int i = 0; i <= k; ++i) v = random_f64();
	return v;
}

}  // extern "C"
