// oracle/main.cpp -- command-line front end of the oracle (mirrors the reference's `main`, main.rs:30-57,
// which renders a compile-time-selected scene to image.ppm).  TEST INFRASTRUCTURE / CPU BASELINE ONLY.
//
//   oracle_render --scene c4 [--width W] [--spp N] [--depth D] [--threads T] [--rng wyrand|philox]
//                 [--seed S] [--scene-seed S] [--assets DIR] [--out image.ppm] [--graph dump.gsg]
// Prints one JSON line with paths/s and Mrays/s.
#include "oracle.hpp"

#include <cstdlib>
#include <thread>

using namespace orc;

int main(int argc, char** argv) {
	std::string scene = "scene11", out, graph, load;
	SceneOptions opt;
	int threads = 0;
	uint64_t seed = 1;
	RngMode mode = RngMode::WYRAND;
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
		if (a == "--scene") scene = next();
		else if (a == "--load") load = next();
		else if (a == "--width") opt.width = std::atoi(next().c_str());
		else if (a == "--spp") opt.spp = std::atoi(next().c_str());
		else if (a == "--depth") opt.max_depth = std::atoi(next().c_str());
		else if (a == "--spheres") opt.n_spheres = std::atoi(next().c_str());
		else if (a == "--threads") threads = std::atoi(next().c_str());
		else if (a == "--seed") seed = std::strtoull(next().c_str(), nullptr, 10);
		else if (a == "--scene-seed") opt.scene_seed = std::strtoull(next().c_str(), nullptr, 10);
		else if (a == "--assets") opt.asset_dir = next();
		else if (a == "--rng") mode = next() == "philox" ? RngMode::PHILOX : RngMode::WYRAND;
		else if (a == "--out") out = next();
		else if (a == "--graph") graph = next();
		else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
	}
	try {
		auto s = load.empty() ? build_scene(scene, opt) : load_graph(load);
		if (!graph.empty()) dump_graph(*s, graph, "");
		if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
		RenderResult r = render(s->camera, *s->root, seed, mode, threads);
		if (!out.empty()) write_ppm(out, s->camera.image_width, s->camera.image_height, r.rgb);
		std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"threads\": %d, \"seconds\": %.6f, \"paths\": %llu, \"rays\": %llu, "
		            "\"box_tests\": %llu, \"prim_tests\": %llu, \"tex_fetches\": %llu, \"paths_per_s\": %.1f, \"mrays_per_s\": %.4f}\n",
		            scene.c_str(), s->camera.image_width, s->camera.image_height, threads, r.seconds,
		            (unsigned long long)r.counters.paths, (unsigned long long)r.counters.rays, (unsigned long long)r.counters.box_tests,
		            (unsigned long long)r.counters.prim_tests, (unsigned long long)r.counters.tex_fetches,
		            (double)r.counters.paths / r.seconds, (double)r.counters.rays / r.seconds / 1e6);
	} catch (const std::exception& e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
