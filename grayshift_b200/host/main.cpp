// grayshift_b200/host/main.cpp -- command-line front end mirroring the reference's `main` (main.rs:30-57):
// pick a scene, render it, write image.ppm.  The reference selects the scene with a compile-time
// constant; here it is an argument.
//
//   grayshift [--scene scene11|c4|...] [--load graph.gsg] [--width W] [--spp N] [--depth D] [--seed S]
//             [--scene-seed S] [--assets DIR] [--device N] [--out image.ppm] [--graph dump.gsg]
#include <cstdlib>
#include <cstring>
#include <stdexcept>

#include "grayshift.hpp"

using namespace grayshift;

int main(int argc, char** argv) {
	std::string scene = "scene11", out = "image.ppm", graph, load;  // `const SCENE: u8 = 11`, main.rs:40; "image.ppm", main.rs:37
	SceneOptions opt;
	uint64_t seed = 1;
	int device = 0;
	for (int i = 1; i < argc; ++i) {
		std::string a = argv[i];
		auto next = [&]() -> std::string { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
		if (a == "--scene") scene = next();
		else if (a == "--load") load = next();
		else if (a == "--width") opt.width = std::atoi(next().c_str());
		else if (a == "--spp") opt.spp = std::atoi(next().c_str());
		else if (a == "--depth") opt.max_depth = std::atoi(next().c_str());
		else if (a == "--spheres") opt.n_spheres = std::atoi(next().c_str());
		else if (a == "--seed") seed = std::strtoull(next().c_str(), nullptr, 10);
		else if (a == "--scene-seed") opt.scene_seed = std::strtoull(next().c_str(), nullptr, 10);
		else if (a == "--assets") opt.asset_dir = next();
		else if (a == "--device") device = std::atoi(next().c_str());
		else if (a == "--out") out = next();
		else if (a == "--graph") graph = next();
		else { std::fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
	}
	try {
		HostScene s = load.empty() ? build_scene(scene, opt) : load_graph(load);
		if (!graph.empty()) dump_graph(s, graph);
		std::FILE* f = std::fopen(out.c_str(), "w");
		if (!f) throw std::runtime_error("cannot create " + out);
		gs_stats st{};
		s.camera.render(s.world, f, device, seed, &st);
		std::fclose(f);
		std::printf("{\"scene\": \"%s\", \"width\": %d, \"height\": %d, \"paths\": %llu, \"rays\": %llu, \"kernel_ms\": %.3f, \"total_ms\": %.3f, "
		            "\"paths_per_s\": %.1f, \"mrays_per_s\": %.3f}\n",
		            s.name.c_str(), s.camera.image_width, s.camera.image_height, (unsigned long long)st.paths, (unsigned long long)st.rays,
		            st.kernel_ms, st.total_ms, st.kernel_ms > 0 ? st.paths / (st.kernel_ms * 1e-3) : 0.0,
		            st.kernel_ms > 0 ? st.rays / (st.kernel_ms * 1e-3) / 1e6 : 0.0);
		std::fprintf(stderr, "Done.\n");  // info!("Done."), main.rs:787
	} catch (const std::exception& e) {
		std::fprintf(stderr, "error: %s\n", e.what());
		return 1;
	}
	return 0;
}
