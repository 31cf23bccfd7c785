// oracle/scenes.cpp -- the reference's scenes (main.rs:61-888) restated on the oracle's classes,
// plus the BASELINE.json configs c1..c5.  TEST INFRASTRUCTURE ONLY (see oracle.hpp).
// Scene-construction randomness (Q16) comes from SceneRng(scene_seed) in the reference's draw order.
#include "oracle.hpp"

#include <stdexcept>

namespace orc {
namespace {

struct CamArgs {
	double aspect; int width; SampleSettings settings; uint32_t depth; double vfov;
	Vec3 from, at; double defocus; Background bg;
};

Camera make_camera(CamArgs a, const SceneOptions& opt) {
	if (opt.width > 0) a.width = opt.width;
	if (opt.max_depth > 0) a.depth = (uint32_t)opt.max_depth;
	if (opt.spp > 0) a.settings = SampleSettings{0.95, 0.0, (uint32_t)opt.spp, 0};  // fixed-spp mapping, SURVEY a3
	return Camera(a.aspect, a.width, a.settings, a.depth, a.vfov, a.from, a.at, Vec3(0.0, 1.0, 0.0), a.defocus, 10.0, a.bg);
}
Background solid_bg(const Vec3& c) { Background b; b.is_hdri = false; b.solid = c; return b; }
Background hdri_bg(Scene& s, const SceneOptions& opt, const Vec3& rot) {
	auto img = load_hdr_rgbe(opt.asset_dir + "/airport.hdr");  // main.rs:138-139
	s.images_f32.push_back(img);
	Background b; b.is_hdri = true; b.hdri.image = img; b.hdri.rotation = rot;
	return b;
}
std::shared_ptr<Texture> earth_texture(Scene& s, const SceneOptions& opt) {  // main.rs:221, 696
	auto img = load_ppm_u8(opt.asset_dir + "/earthmap.ppm");  // earthmap.jpg decoded by the host (PIL) to P6
	s.images_u8.push_back(img);
	return s.add(std::make_shared<ImageTexture>(img));
}
void finish(Scene& s, HittableList& world) { s.root = s.bvh_from_list(world); }  // `BVHNode::from_list(world)` ends every scene

const SampleSettings S_32_1000_25{0.95, 0.25, 32, 1000};

// main.rs:61-167
void bouncing_spheres(Scene& s, const SceneOptions& opt, int default_width) {
	SceneRng rng(opt.scene_seed);
	auto world = s.list();
	auto even = s.solid(Vec3(0.2, 0.3, 0.1));
	auto odd = s.solid(Vec3(0.9, 0.9, 0.9));
	auto ground_texture = s.add(std::make_shared<CheckeredTexture>(0.32, even, odd));
	auto ground_material = s.lambertian(ground_texture);
	world->add(s.sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, ground_material));
	for (int a = -11; a < 11; ++a) {
		for (int b = -11; b < 11; ++b) {
			double cx = (double)a + 0.9 * rng.f64();
			double cz = (double)b + 0.9 * rng.f64();
			Vec3 center(cx, 0.2, cz);
			if ((center - Vec3(4.0, 0.2, 0.0)).length() > 0.9) {
				double material_choice = rng.f64();
				std::shared_ptr<Material> material;
				if (material_choice < 0.8) {
					double a0 = rng.range(0, 1), a1 = rng.range(0, 1), a2 = rng.range(0, 1);
					double b0 = rng.range(0, 1), b1 = rng.range(0, 1), b2 = rng.range(0, 1);
					material = s.lambertian_color(Vec3(a0, a1, a2) * Vec3(b0, b1, b2));
				} else if (material_choice < 0.95) {
					double a0 = rng.range(0.5, 1), a1 = rng.range(0.5, 1), a2 = rng.range(0.5, 1);
					double fuzz = rng.range(0.0, 0.5);
					material = s.metal(Vec3(a0, a1, a2), fuzz);
				} else {
					material = s.dielectric(1.5);
				}
				(void)rng.f64();  // center_end (main.rs:105): drawn, unused
				world->add(s.sphere(center, 0.2, material));
			}
		}
	}
	world->add(s.sphere(Vec3(4.0, 1.0, 0.0), 1.0, s.metal(Vec3(0.7, 0.6, 0.5), 0.0)));
	world->add(s.sphere(Vec3(0.0, 1.0, 0.0), 1.0, s.dielectric(1.5)));
	world->add(s.sphere(Vec3(-4.0, 1.0, 0.0), 1.0, s.metal(Vec3(0.7, 0.6, 0.5), 0.0)));
	s.camera = make_camera({16.0 / 9.0, default_width, {0.95, 0.05, 64, 200}, 50, 20.0, Vec3(13.0, 2.0, 3.0), Vec3(0, 0, 0), 0.6,
	                        hdri_bg(s, opt, Vec3(0.0, -90.0, 90.0))}, opt);
	finish(s, *world);
}

// main.rs:169-215
void checkered_spheres(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	auto tex = s.add(std::make_shared<CheckeredTexture>(0.32, s.solid(Vec3(0.2, 0.3, 0.1)), s.solid(Vec3(0.9, 0.9, 0.9))));
	auto mat = s.lambertian(tex);
	world->add(s.sphere(Vec3(0.0, -10.0, 0.0), 10.0, mat));
	world->add(s.sphere(Vec3(0.0, 10.0, 0.0), 10.0, mat));
	s.camera = make_camera({16.0 / 9.0, 400, S_32_1000_25, 50, 20.0, Vec3(13.0, 2.0, 3.0), Vec3(0, 0, 0), 0.0, solid_bg(Vec3(0.7, 0.8, 1.0))}, opt);
	finish(s, *world);
}

// main.rs:217-253 (c2 swaps in the HDR background of main.rs:805-828 and width 1280)
void earth(Scene& s, const SceneOptions& opt, bool config2) {
	auto world = s.list();
	auto mat = s.lambertian(earth_texture(s, opt));
	world->add(s.sphere(Vec3(0.0, 0.0, 0.0), 2.0, mat));
	Background bg = config2 ? hdri_bg(s, opt, Vec3(PI / 2.0, PI, 0.0)) : solid_bg(Vec3(0.7, 0.8, 1.0));
	s.camera = make_camera({16.0 / 9.0, config2 ? 1280 : 400, S_32_1000_25, 50, 20.0, Vec3(0.0, 0.0, 12.0), Vec3(0, 0, 0), 0.0, bg}, opt);
	finish(s, *world);
}

// main.rs:255-297
void perlin_spheres(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	auto mat = s.lambertian(s.add(std::make_shared<NoiseTexture>(4.0)));
	world->add(s.sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, mat));
	world->add(s.sphere(Vec3(0.0, 2.0, 0.0), 2.0, mat));
	s.camera = make_camera({16.0 / 9.0, 400, S_32_1000_25, 50, 20.0, Vec3(13.0, 2.0, 3.0), Vec3(0, 0, 0), 0.0, solid_bg(Vec3(0.7, 0.8, 1.0))}, opt);
	finish(s, *world);
}

// main.rs:299-364
void quads(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	auto left_red = s.lambertian_color(Vec3(1.0, 0.2, 0.2));
	auto back_green = s.lambertian_color(Vec3(0.2, 1.0, 0.2));
	auto right_blue = s.lambertian_color(Vec3(0.2, 0.2, 1.0));
	auto upper_orange = s.lambertian_color(Vec3(1.0, 0.5, 0.0));
	auto lower_teal = s.lambertian_color(Vec3(0.2, 0.8, 0.8));
	world->add(s.quad(Vec3(-3.0, -2.0, 5.0), Vec3(0.0, 0.0, -4.0), Vec3(0.0, 4.0, 0.0), left_red));
	world->add(s.quad(Vec3(-2.0, -2.0, 0.0), Vec3(4.0, 0.0, 0.0), Vec3(0.0, 4.0, 0.0), back_green));
	world->add(s.quad(Vec3(3.0, -2.0, 1.0), Vec3(0.0, 0.0, 4.0), Vec3(0.0, 4.0, 0.0), right_blue));
	world->add(s.quad(Vec3(-2.0, 3.0, 1.0), Vec3(4.0, 0.0, 0.0), Vec3(0.0, 0.0, 4.0), upper_orange));
	world->add(s.quad(Vec3(-2.0, -3.0, 5.0), Vec3(4.0, 0.0, 0.0), Vec3(0.0, 0.0, -4.0), lower_teal));
	s.camera = make_camera({1.0, 400, S_32_1000_25, 50, 80.0, Vec3(0.0, 0.0, 9.0), Vec3(0, 0, 0), 0.0, solid_bg(Vec3(0.7, 0.8, 1.0))}, opt);
	finish(s, *world);
}

// main.rs:366-419
void simple_light(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	auto mat = s.lambertian(s.add(std::make_shared<NoiseTexture>(4.0)));
	world->add(s.sphere(Vec3(0.0, -1000.0, 0.0), 1000.0, mat));
	world->add(s.sphere(Vec3(0.0, 2.0, 0.0), 2.0, mat));
	auto light = s.light_color(Vec3(4.0, 4.0, 4.0));
	world->add(s.quad(Vec3(3.0, 1.0, -2.0), Vec3(2.0, 0.0, 0.0), Vec3(0.0, 2.0, 0.0), light));
	s.camera = make_camera({16.0 / 9.0, 1000, S_32_1000_25, 50, 20.0, Vec3(26.0, 3.0, 6.0), Vec3(0.0, 2.0, 0.0), 0.0, solid_bg(Vec3(0, 0, 0))}, opt);
	finish(s, *world);
}

// shared by cornell_box (main.rs:421-517) and cornell_smoke (main.rs:519-624)
void cornell(Scene& s, const SceneOptions& opt, bool smoke) {
	auto world = s.list();
	auto red = s.lambertian_color(Vec3(0.65, 0.05, 0.05));
	auto white = s.lambertian_color(Vec3(0.73, 0.73, 0.73));
	auto green = s.lambertian_color(Vec3(0.12, 0.45, 0.15));
	auto light = s.light_color(smoke ? Vec3(7.0, 7.0, 7.0) : Vec3(15.0, 15.0, 15.0));
	if (smoke) world->add(s.quad(Vec3(113.0, 554.0, 127.0), Vec3(330.0, 0.0, 0.0), Vec3(0.0, 0.0, 305.0), light));
	else world->add(s.quad(Vec3(343.0, 554.0, 332.0), Vec3(-130.0, 0.0, 0.0), Vec3(0.0, 0.0, -105.0), light));
	world->add(s.quad(Vec3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), green));
	world->add(s.quad(Vec3(0.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), Vec3(0.0, 0.0, 555.0), red));
	world->add(s.quad(Vec3(0.0, 0.0, 0.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 0.0, 555.0), white));
	world->add(s.quad(Vec3(555.0, 555.0, 555.0), Vec3(-555.0, 0.0, 0.0), Vec3(0.0, 0.0, -555.0), white));
	world->add(s.quad(Vec3(0.0, 0.0, 555.0), Vec3(555.0, 0.0, 0.0), Vec3(0.0, 555.0, 0.0), white));
	auto box1 = s.translate(s.rotate_y(s.cube(Vec3(0, 0, 0), Vec3(165.0, 330.0, 165.0), white), 15.0), Vec3(265.0, 0.0, 295.0));
	if (smoke) { if (opt.include_media) world->add(s.medium(box1, 0.01, s.isotropic_color(Vec3(0.0, 0.0, 0.0)))); }
	else world->add(box1);
	auto box2 = s.translate(s.rotate_y(s.cube(Vec3(0, 0, 0), Vec3(165.0, 165.0, 165.0), white), -18.0), Vec3(130.0, 0.0, 65.0));
	if (smoke) { if (opt.include_media) world->add(s.medium(box2, 0.01, s.isotropic_color(Vec3(1.0, 1.0, 1.0)))); }
	else world->add(box2);
	SampleSettings st{0.95, smoke ? 0.25 : 0.5, 32, 1000};
	s.camera = make_camera({1.0, 600, st, 50, 40.0, Vec3(278.0, 278.0, -800.0), Vec3(278.0, 278.0, 0.0), 0.0, solid_bg(Vec3(0, 0, 0))}, opt);
	finish(s, *world);
}

// main.rs:626-789
void final_scene(Scene& s, const SceneOptions& opt, int image_width, uint32_t max_depth) {
	SceneRng rng(opt.scene_seed);
	auto world = s.list();
	auto ground = s.lambertian_color(Vec3(0.48, 0.83, 0.53));
	const int boxes_per_side = 20;
	auto boxes = s.list();
	for (int i = 0; i < boxes_per_side; ++i) {
		for (int j = 0; j < boxes_per_side; ++j) {
			double w = 100.0;
			double x0 = -1000.0 + (double)i * w, y0 = 0.0, z0 = -1000.0 + (double)j * w;
			double x1 = x0 + w, y1 = rng.range(1.0, 101.0), z1 = z0 + w;
			boxes->add(s.cube(Vec3(x0, y0, z0), Vec3(x1, y1, z1), ground));
		}
	}
	world->add(s.bvh_from_list(*boxes));
	world->add(s.quad(Vec3(123.0, 554.0, 147.0), Vec3(300.0, 0.0, 0.0), Vec3(0.0, 0.0, 265.0), s.light_color(Vec3(7.0, 7.0, 7.0))));
	Vec3 mc1(400.0, 400.0, 200.0);
	Vec3 mc2 = mc1 + Vec3(30.0, 0.0, 0.0);
	world->add(s.moving_sphere(mc1, mc2, 50.0, s.lambertian_color(Vec3(0.7, 0.3, 0.1))));
	world->add(s.sphere(Vec3(260.0, 150.0, 45.0), 50.0, s.dielectric(1.5)));
	world->add(s.sphere(Vec3(0.0, 150.0, 145.0), 50.0, s.metal(Vec3(0.8, 0.8, 0.9), 1.0)));
	world->add(s.sphere(Vec3(400.0, 200.0, 400.0), 100.0, s.lambertian(earth_texture(s, opt))));
	world->add(s.sphere(Vec3(220.0, 280.0, 300.0), 80.0, s.lambertian(s.add(std::make_shared<NoiseTexture>(0.2)))));
	auto fog_material = s.dielectric(1.5);  // only ever a boundary material (Q11)
	auto foggy_boundary = s.sphere(Vec3(360.0, 150.0, 145.0), 70.0, fog_material);
	auto foggy_material = s.lambertian_color(Vec3(0.2, 0.4, 0.9));
	if (opt.include_media) world->add(s.medium(foggy_boundary, 0.2, foggy_material));
	auto world_fog_boundary = s.sphere(Vec3(0, 0, 0), 5000.0, fog_material);
	auto world_fog_material = s.lambertian_color(Vec3(1.0, 1.0, 1.0));
	if (opt.include_media) world->add(s.medium(world_fog_boundary, 0.0001, world_fog_material));
	auto ball_material = s.lambertian_color(Vec3(0.73, 0.73, 0.73));
	auto balls = s.list();
	for (int i = 0; i < 1000; ++i) {
		double a = rng.range(0.0, 165.0), b = rng.range(0.0, 165.0), c = rng.range(0.0, 165.0);
		balls->add(s.sphere(Vec3(a, b, c), 10.0, ball_material));
	}
	world->add(s.translate(s.rotate_y(s.bvh_from_list(*balls), 15.0), Vec3(-100.0, 270.0, 395.0)));
	s.camera = make_camera({1.0, image_width, S_32_1000_25, max_depth, 40.0, Vec3(478.0, 278.0, -600.0), Vec3(278.0, 278.0, 0.0), 0.0,
	                        solid_bg(Vec3(0, 0, 0))}, opt);
	finish(s, *world);
}

// main.rs:792-834
void hdri_scene(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	world->add(s.sphere(Vec3(4.0, 1.0, 0.0), 1.0, s.metal(Vec3(0.7, 0.6, 0.5), 0.0)));
	s.camera = make_camera({16.0 / 9.0, 600, {0.95, 0.05, 64, 200}, 50, 20.0, Vec3(13.0, 2.0, 5.0), Vec3(0, 0, 0), 0.6,
	                        hdri_bg(s, opt, Vec3(PI / 2.0, PI, 0.0))}, opt);
	finish(s, *world);
}

// main.rs:836-888
void triangles(Scene& s, const SceneOptions& opt) {
	auto world = s.list();
	auto left_red = s.lambertian_color(Vec3(1.0, 0.2, 0.2));
	auto back_green = s.lambertian_color(Vec3(0.2, 1.0, 0.2));
	(void)s.lambertian_color(Vec3(0.2, 0.2, 1.0));  // right_blue: created, unused
	auto upper_orange = s.lambertian_color(Vec3(1.0, 0.5, 0.0));
	(void)s.lambertian_color(Vec3(0.2, 0.8, 0.8));  // lower_teal: created, unused
	world->add(s.triangle(Vec3(-2.0, 2.0, 0.0), Vec3(-2.0, -2.0, 0.0), Vec3(-2.0, -2.0, 4.0), left_red));
	world->add(s.triangle(Vec3(-2.0, 2.0, 0.0), Vec3(2.0, -2.0, 0.0), Vec3(-2.0, -2.0, 0.0), back_green));
	world->add(s.triangle(Vec3(-2.0, -2.0, 4.0), Vec3(-2.0, -2.0, 0.0), Vec3(2.0, -2.0, 0.0), upper_orange));
	s.camera = make_camera({1.0, 400, S_32_1000_25, 50, 80.0, Vec3(0.0, 0.0, 9.0), Vec3(0, 0, 0), 0.0, solid_bg(Vec3(0.7, 0.8, 1.0))}, opt);
	finish(s, *world);
}

// C5 (no reference counterpart): n stationary spheres r=0.25, centres U[0,100)^3,
// materials drawn as main.rs:85-103, SOLID(0.7,0.8,1.0) background.
void sphere_cloud(Scene& s, const SceneOptions& opt) {
	SceneRng rng(opt.scene_seed);
	int n = opt.n_spheres > 0 ? opt.n_spheres : 1000000;
	auto world = s.list();
	for (int i = 0; i < n; ++i) {
		double cx = rng.range(0.0, 100.0), cy = rng.range(0.0, 100.0), cz = rng.range(0.0, 100.0);
		double choice = rng.f64();
		std::shared_ptr<Material> material;
		if (choice < 0.8) {
			double a0 = rng.range(0, 1), a1 = rng.range(0, 1), a2 = rng.range(0, 1);
			double b0 = rng.range(0, 1), b1 = rng.range(0, 1), b2 = rng.range(0, 1);
			material = s.lambertian_color(Vec3(a0, a1, a2) * Vec3(b0, b1, b2));
		} else if (choice < 0.95) {
			double a0 = rng.range(0.5, 1), a1 = rng.range(0.5, 1), a2 = rng.range(0.5, 1);
			double fuzz = rng.range(0.0, 0.5);
			material = s.metal(Vec3(a0, a1, a2), fuzz);
		} else {
			material = s.dielectric(1.5);
		}
		world->add(s.sphere(Vec3(cx, cy, cz), 0.25, material));
	}
	s.camera = make_camera({16.0 / 9.0, 3840, {0.95, 0.0, 4096, 0}, 50, 40.0, Vec3(180.0, 120.0, 180.0), Vec3(50.0, 50.0, 50.0), 0.0,
	                        solid_bg(Vec3(0.7, 0.8, 1.0))}, opt);
	finish(s, *world);
}

}  // namespace

std::unique_ptr<Scene> build_scene(const std::string& name, const SceneOptions& opt_in) {
	auto s = std::make_unique<Scene>();
	s->name = name;
	SceneOptions opt = opt_in;
	auto fixed = [&](int spp) { if (opt.spp == 0) opt.spp = spp; };
	if (name == "c1") { fixed(100); bouncing_spheres(*s, opt, 400); }
	else if (name == "c2") { fixed(1000); earth(*s, opt, true); }
	else if (name == "c3") { fixed(1000); cornell(*s, opt, false); }
	else if (name == "c4") { fixed(10000); final_scene(*s, opt, 800, 40); }
	else if (name == "c5") { fixed(4096); sphere_cloud(*s, opt); }
	else if (name == "scene0") bouncing_spheres(*s, opt, 600);
	else if (name == "scene1") checkered_spheres(*s, opt);
	else if (name == "scene2") earth(*s, opt, false);
	else if (name == "scene3") perlin_spheres(*s, opt);
	else if (name == "scene4") quads(*s, opt);
	else if (name == "scene5") simple_light(*s, opt);
	else if (name == "scene6") cornell(*s, opt, false);
	else if (name == "scene7") cornell(*s, opt, true);
	else if (name == "scene8") final_scene(*s, opt, 400, 50);
	else if (name == "scene9") final_scene(*s, opt, 800, 40);
	else if (name == "scene10") triangles(*s, opt);
	else if (name == "scene11") hdri_scene(*s, opt);
	else throw std::runtime_error("unknown scene '" + name + "'");
	return s;
}

}  // namespace orc
