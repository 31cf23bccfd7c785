// oracle/capi.cpp -- C entry points of the oracle for the Python tests (ctypes) and bench.py's
// cpu_baseline / --impl reference legs.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).
#include "oracle.hpp"

#include <stdexcept>

using namespace orc;

namespace {
thread_local std::string g_err;
template <class F> int guard(F&& f) {
	try { f(); return 0; }
	catch (const std::exception& e) { g_err = e.what(); return 1; }
	catch (...) { g_err = "unknown error"; return 1; }
}
}  // namespace

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

void* orc_scene_build(const char* name, int width, int spp, int max_depth, uint64_t scene_seed, int include_media,
                      int n_spheres, const char* asset_dir) {
	Scene* out = nullptr;
	guard([&] {
		SceneOptions o;
		o.width = width; o.spp = spp; o.max_depth = max_depth; o.scene_seed = scene_seed;
		o.include_media = include_media != 0; o.n_spheres = n_spheres;
		if (asset_dir && *asset_dir) o.asset_dir = asset_dir;
		out = build_scene(name, o).release();
	});
	return out;
}

void* orc_scene_load(const char* path, int include_media) {
	Scene* out = nullptr;
	guard([&] { out = load_graph(path, include_media != 0).release(); });
	return out;
}

int orc_scene_dump(void* scene, const char* path) {
	return guard([&] { dump_graph(*static_cast<Scene*>(scene), path, ""); });
}

void orc_scene_free(void* scene) { delete static_cast<Scene*>(scene); }

// info[0..5] = width, height, prim_count, medium_count, max_depth, batch_size
int orc_scene_info(void* scene, int64_t* info) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		info[0] = s->camera.image_width; info[1] = s->camera.image_height; info[2] = s->prim_count;
		info[3] = s->medium_count; info[4] = s->camera.max_depth; info[5] = s->camera.sample_settings.batch_size;
	});
}

int orc_scene_set_sampling(void* scene, double confidence, double tolerance, uint32_t batch, uint32_t max_samples, uint32_t max_depth) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		s->camera.sample_settings = SampleSettings{confidence, tolerance, batch, max_samples};
		if (max_depth) s->camera.max_depth = max_depth;
	});
}

// counters[0..4] = rays, box_tests, prim_tests, tex_fetches, paths.  box/prim/tex need the ORC_COUNTERS build.
int orc_render(void* scene, uint64_t seed, int rng_mode, int threads, uint32_t sample_begin, int pixel_begin, int pixel_end,
               double* out_rgb, uint32_t* out_counts, uint64_t* counters, double* seconds) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		RenderResult r = render(s->camera, *s->root, seed, rng_mode == 0 ? RngMode::PHILOX : RngMode::WYRAND, threads,
		                        sample_begin, pixel_begin, pixel_end);
		if (out_rgb) std::memcpy(out_rgb, r.rgb.data(), r.rgb.size() * sizeof(double));
		if (out_counts) std::memcpy(out_counts, r.counts.data(), r.counts.size() * sizeof(uint32_t));
		if (counters) {
			counters[0] = r.counters.rays; counters[1] = r.counters.box_tests; counters[2] = r.counters.prim_tests;
			counters[3] = r.counters.tex_fetches; counters[4] = r.counters.paths;
		}
		if (seconds) *seconds = r.seconds;
	});
}

// rays: n x 7 doubles (origin, direction, time).  prim_id = -1 on a miss.  extra (optional): n x 8 =
// position(3), normal(3, flipped against the ray), u, v.
int orc_trace_closest(void* scene, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t, double* extra) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		tls_rng.mode = RngMode::WYRAND;
		for (size_t k = 0; k < n; ++k) {
			const double* r = rays + k * 7;
			Ray ray(Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]), r[6]);
			auto rec = s->root->hit(ray, Interval(tmin, tmax));
			prim_id[k] = rec ? rec->prim_id : -1;
			t[k] = rec ? rec->t : 0.0;
			if (extra) {
				double* e = extra + k * 8;
				if (rec) {
					e[0] = rec->position.x; e[1] = rec->position.y; e[2] = rec->position.z;
					e[3] = rec->normal.x; e[4] = rec->normal.y; e[5] = rec->normal.z; e[6] = rec->u; e[7] = rec->v;
				} else for (int i = 0; i < 8; ++i) e[i] = 0.0;
			}
		}
	});
}

// Camera rays with the shared Philox scheme: rays_out = (pixel_end - pixel_begin) x 7 for sample index `sample`.
int orc_camera_rays(void* scene, uint64_t seed, uint32_t sample, int pixel_begin, int pixel_end, double* rays_out) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		const Camera& c = s->camera;
		tls_rng.mode = RngMode::PHILOX;
		for (int n = pixel_begin; n < pixel_end; ++n) {
			tls_rng.start_path(seed, (uint32_t)n, sample);
			Ray r = c.get_ray(n % c.image_width, n / c.image_width);
			double* o = rays_out + (size_t)(n - pixel_begin) * 7;
			o[0] = r.origin.x; o[1] = r.origin.y; o[2] = r.origin.z;
			o[3] = r.direction.x; o[4] = r.direction.y; o[5] = r.direction.z; o[6] = r.time;
		}
	});
}

// Radiance of individual paths under the shared Philox scheme: out = n x 3 for (pixel[k], sample[k]).
int orc_path_radiance(void* scene, uint64_t seed, const uint32_t* pixel, const uint32_t* sample, size_t n, double* out) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		const Camera& c = s->camera;
		tls_rng.mode = RngMode::PHILOX;
		for (size_t k = 0; k < n; ++k) {
			tls_rng.start_path(seed, pixel[k], sample[k]);
			Ray r = c.get_ray((int)(pixel[k] % c.image_width), (int)(pixel[k] / c.image_width));
			Vec3 col = c.ray_color(r, c.max_depth, *s->root);
			out[k * 3] = col.x; out[k * 3 + 1] = col.y; out[k * 3 + 2] = col.z;
		}
	});
}

// ---- unit-level hooks for known-answer tests ------------------------------------------------
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { philox4x32_10(ctr, key, out); }
void orc_perlin_table(uint32_t seed, uint8_t* out) { perlin_permutation_table(seed, out); }
double orc_perlin_get(double x, double y, double z) { static Perlin p(0); return p.get(x, y, z); }
void orc_xorshift_words(const uint8_t* seed16, size_t n, uint32_t* out) { xorshift_words(seed16, n, out); }
// returns the index, or -1 when the word stream ran out; *used = words consumed
int64_t orc_gen_index(const uint32_t* words, size_t n_words, uint32_t ubound, size_t* used) {
	int64_t r = -1;
	guard([&] { r = gen_index_from_words(words, n_words, ubound, used); });
	return r;
}
double orc_reflectance(double cosine, double ri) { return Dielectric::reflectance(cosine, ri); }
double orc_luminance(double r, double g, double b) { return luminance(Vec3(r, g, b)); }
int orc_color_byte(double c) { return color_byte(c); }
double orc_scene_rng(uint64_t seed, int k) { SceneRng r(seed); double v = 0; for (int i = 0; i <= k; ++i) v = r.f64(); return v; }

// texture value by oracle texture-table id (creation order)
int orc_texture_value(void* scene, int tex, double u, double v, const double* p, double* out3) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		Vec3 c = s->textures.at(tex)->value_at(u, v, Vec3(p[0], p[1], p[2]));
		out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
	});
}
int orc_background(void* scene, const double* dir, double* out3) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		Vec3 c = s->camera.sample_background(Ray(Vec3(0, 0, 0), Vec3(dir[0], dir[1], dir[2]), 0.0));
		out3[0] = c.x; out3[1] = c.y; out3[2] = c.z;
	});
}
// n draws of random_cosine_direction / random_unit_vector / in_unit_disk under Philox(seed, pixel=0, sample=k)
int orc_sample_dirs(int which, uint64_t seed, size_t n, double* out) {
	return guard([&] {
		tls_rng.mode = RngMode::PHILOX;
		for (size_t k = 0; k < n; ++k) {
			tls_rng.start_path(seed, 0, (uint32_t)k);
			Vec3 v = which == 0 ? random_cosine_direction() : (which == 1 ? random_unit_vector() : random_vector_in_unit_disk());
			out[k * 3] = v.x; out[k * 3 + 1] = v.y; out[k * 3 + 2] = v.z;
		}
	});
}
// One scatter event of material `mat` (creation-order id): in = ray o(3) d(3) time, hit p(3) n(3, outward) u v;
// out = scattered(1/0), direction(3), attenuation(3), emitted(3).  Philox(seed, pixel, sample).
int orc_scatter(void* scene, int mat, uint64_t seed, uint32_t pixel, uint32_t sample, const double* in, double* out) {
	return guard([&] {
		auto* s = static_cast<Scene*>(scene);
		tls_rng.mode = RngMode::PHILOX;
		tls_rng.start_path(seed, pixel, sample);
		Ray ray(Vec3(in[0], in[1], in[2]), Vec3(in[3], in[4], in[5]), in[6]);
		auto m = s->materials.at(mat);
		HitRecord rec(ray, 1.0, Vec3(in[7], in[8], in[9]), Vec3(in[10], in[11], in[12]), m, in[13], in[14], 0);
		Vec3 e = m->emitted(rec.u, rec.v, rec.position);
		auto sr = m->scatter(ray, rec);
		out[0] = sr ? 1.0 : 0.0;
		Vec3 d = sr ? sr->scattered_ray.direction : Vec3(0, 0, 0), a = sr ? sr->attenuation : Vec3(0, 0, 0);
		out[1] = d.x; out[2] = d.y; out[3] = d.z; out[4] = a.x; out[5] = a.y; out[6] = a.z; out[7] = e.x; out[8] = e.y; out[9] = e.z;
	});
}
// decoded HDR pixels (for checking the RGBE restatement against cv2): returns w,h and copies into out if non-null
int orc_load_hdr(const char* path, int* w, int* h, float* out) {
	return guard([&] {
		auto im = load_hdr_rgbe(path);
		*w = im->w; *h = im->h;
		if (out) std::memcpy(out, im->rgb.data(), im->rgb.size() * sizeof(float));
	});
}
void orc_set_triangle_interval(int respect) { g_triangle_respects_interval = respect != 0; }

int orc_counters_enabled() {
#ifdef ORC_COUNTERS
	return 1;
#else
	return 0;
#endif
}

}  // extern "C"
