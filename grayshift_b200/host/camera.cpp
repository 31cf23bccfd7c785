// grayshift_b200/host/camera.cpp -- Camera::new / Camera::render of the host mirror (camera.rs:39-121) and the
// output stage (color.rs:8-33).  render() = flatten -> gs_scene_create -> gs_render -> PPM text.
#include <cmath>
#include <stdexcept>

#include "grayshift.hpp"

namespace grayshift {

static const double PI = 3.14159265358979323846;

Camera Camera::new_(double aspect_ratio, int image_width, SampleSettings sample_settings, uint32_t max_depth, double v_fov,
                    const Vec3& look_from, const Vec3& look_at, const Vec3& vup, double defocus_angle, double focus_distance,
                    Background background) {  // camera.rs:39-98
	Camera c;
	c.image_width = image_width;
	c.image_height = (int)((double)image_width / aspect_ratio);
	double theta = v_fov / 180.0 * PI;
	double h = std::tan(theta / 2.0);
	double viewport_height = 2.0 * h * focus_distance;
	double viewport_width = viewport_height * ((double)c.image_width / (double)c.image_height);
	Vec3 w = (look_from - look_at).unit();
	Vec3 u = vup.cross(w).unit();
	Vec3 v = w.cross(u);
	Vec3 viewport_u = viewport_width * u;
	Vec3 viewport_v = viewport_height * -v;
	c.pixel_delta_u = viewport_u / (double)c.image_width;
	c.pixel_delta_v = viewport_v / (double)c.image_height;
	Vec3 viewport_upper_left = look_from - focus_distance * w - viewport_u / 2.0 - viewport_v / 2.0;
	c.starting_pixel_pos = viewport_upper_left + 0.5 * (c.pixel_delta_u + c.pixel_delta_v);
	double defocus_radius = focus_distance * std::tan((defocus_angle / 2.0) / 180.0 * PI);
	c.defocus_disk_u = u * defocus_radius;
	c.defocus_disk_v = v * defocus_radius;
	c.sample_settings = sample_settings; c.max_depth = max_depth; c.center = look_from; c.background = std::move(background);
	c.defocus_angle = defocus_angle;
	c.arg_aspect = aspect_ratio; c.arg_vfov = v_fov; c.arg_focus = focus_distance;
	c.arg_from = look_from; c.arg_at = look_at; c.arg_vup = vup;
	return c;
}

gs_camera Camera::to_abi(FlatScene& flat) const {
	gs_camera g{};
	g.image_width = image_width; g.image_height = image_height; g.max_depth = max_depth;
	auto put = [](double* d, const Vec3& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; };
	put(g.center, center); put(g.starting_pixel_pos, starting_pixel_pos);
	put(g.pixel_delta_u, pixel_delta_u); put(g.pixel_delta_v, pixel_delta_v);
	g.defocus_angle = defocus_angle;
	put(g.defocus_disk_u, defocus_disk_u); put(g.defocus_disk_v, defocus_disk_v);
	if (background.is_hdri) {
		g.background_kind = GS_BG_HDRI;
		g.hdri_image = flat.image_index(background.hdri.image);
		put(g.hdri_rotation, background.hdri.rotation);
		flat.finish();  // the image table may have grown
	} else {
		g.background_kind = GS_BG_SOLID;
		put(g.background_color, background.color);
	}
	return g;
}

gs_sample_settings Camera::settings_abi() const {
	gs_sample_settings s{};
	s.confidence = sample_settings.confidence; s.tolerance = sample_settings.tolerance;
	s.batch_size = sample_settings.batch_size; s.max_samples = sample_settings.max_samples;
	return s;
}

void Camera::render_to_buffer(const HittablePtr& world, std::vector<float>& rgb, std::vector<uint32_t>* counts, int device, uint64_t seed,
                              gs_stats* stats) const {
	auto flat = flatten(world);
	gs_camera cam = to_abi(*flat);
	gs_sample_settings st = settings_abi();
	gs_scene* scene = nullptr;
	if (gs_scene_create(&flat->desc, device, &scene) != GS_OK) throw std::runtime_error(std::string("gs_scene_create: ") + gs_last_error());
	const size_t n = (size_t)image_width * image_height;
	rgb.assign(n * 3, 0.f);
	if (counts) counts->assign(n, 0u);
	int rc = gs_render(scene, &cam, &st, seed, rgb.data(), counts ? counts->data() : nullptr, stats);
	std::string err = rc ? gs_last_error() : "";
	gs_scene_destroy(scene);
	if (rc) throw std::runtime_error("gs_render: " + err);
}

void write_color(std::FILE* image, const Vec3& color) {  // color.rs:8-27
	auto gamma = [](double n) { return n > 0.0 ? std::sqrt(n) : 0.0; };
	auto byte = [&](double c) { double g = gamma(c); g = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g); return (int)(256.0 * g); };
	if (std::fprintf(image, "%d %d %d\n", byte(color.x), byte(color.y), byte(color.z)) < 0) throw std::runtime_error("write_color: write failed");
}
double luminance(const Vec3& v) { return 0.299 * v.x + 0.587 * v.y + 0.144 * v.z; }  // color.rs:31-33

void Camera::render(const HittablePtr& world, std::FILE* image_file, int device, uint64_t seed, gs_stats* stats) const {  // camera.rs:100-121
	if (!image_file) throw std::runtime_error("Camera::render: null file");
	if (std::fprintf(image_file, "P3\n%d %d\n255\n", image_width, image_height) < 0) throw std::runtime_error("Camera::render: write failed");
	std::vector<float> rgb;
	render_to_buffer(world, rgb, nullptr, device, seed, stats);
	for (size_t n = 0; n < (size_t)image_width * image_height; ++n) write_color(image_file, Vec3(rgb[n * 3], rgb[n * 3 + 1], rgb[n * 3 + 2]));
}

}  // namespace grayshift
