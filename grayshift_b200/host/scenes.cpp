// grayshift_b200/host/scenes.cpp -- the reference's scenes (main.rs:61-888) written against the host mirror
// API, statement for statement, plus the five BASELINE.json configurations.  Rust evaluates operands
// and arguments left to right; C++ does not promise that, so every RNG draw is sequenced explicitly.
#include <stdexcept>

#include "grayshift.hpp"

namespace grayshift {
namespace {

const double PI = 3.14159265358979323846;

struct CamArgs {
	double aspect; int width; SampleSettings settings; uint32_t depth; double vfov;
	Vec3 from, at; double defocus; Background bg;
};
Camera make_camera(CamArgs a, const SceneOptions& opt) {
	if (opt.width > 0) a.width = opt.width;
	if (opt.max_depth > 0) a.depth = (uint32_t)opt.max_depth;
	if (opt.spp > 0) a.settings = SampleSettings{0.95, 0.0, (uint32_t)opt.spp, 0};  // fixed-spp mapping
	return Camera::new_(a.aspect, a.width, a.settings, a.depth, a.vfov, a.from, a.at, Vec3::new_(0.0, 1.0, 0.0), a.defocus, 10.0, a.bg);
}
Background hdri_background(const SceneOptions& opt, const Vec3& rotation) {
	HDRI h;
	h.image = load_hdr(opt.asset_dir + "/airport.hdr");  // File::open("airport.hdr") + radiant::load, main.rs:138-139
	h.rotation = rotation;
	return Background::HDRI_(h);
}
const SampleSettings S_32_1000_25{0.95, 0.25, 32, 1000};

// main.rs:61-167
HostScene bouncing_spheres(const SceneOptions& opt, int default_width) {
	auto world = HittableList::new_();
	auto ground_texture = CheckeredTexture::from_colors(0.32, Vec3::new_(0.2, 0.3, 0.1), Vec3::new_(0.9, 0.9, 0.9));
	auto ground_material = Lambertian::from_texture(ground_texture);
	world->add(Sphere::new_stationary(Vec3::new_(0.0, -1000.0, 0.0), 1000.0, ground_material));
	for (int a = -11; a < 11; ++a) {
		for (int b = -11; b < 11; ++b) {
			double cx = (double)a + 0.9 * random_f64();
			double cz = (double)b + 0.9 * random_f64();
			Vec3 center = Vec3::new_(cx, 0.2, cz);
			if ((center - Vec3::new_(4.0, 0.2, 0.0)).length() > 0.9) {
				double material_choice = random_f64();
				std::shared_ptr<Material> material;
				if (material_choice < 0.8) {
					Vec3 c0 = random_vector(0.0, 1.0);
					Vec3 c1 = random_vector(0.0, 1.0);
					material = Lambertian::from_color(c0 * c1);
				} else if (material_choice < 0.95) {
					Vec3 albedo = random_vector(0.5, 1.0);
					double fuzz = random_f64(0.0, 0.5);
					material = Metal::new_(albedo, fuzz);
				} else {
					material = Dielectric::new_(1.5);
				}
				Vec3 center_end = center + Vec3::new_(0.0, random_f64(), 0.0);  // main.rs:105: drawn, unused
				(void)center_end;
				world->add(Sphere::new_stationary(center, 0.2, material));
			}
		}
	}
	world->add(Sphere::new_stationary(Vec3::new_(4.0, 1.0, 0.0), 1.0, Metal::new_(Vec3::new_(0.7, 0.6, 0.5), 0.0)));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 1.0, 0.0), 1.0, Dielectric::new_(1.5)));
	world->add(Sphere::new_stationary(Vec3::new_(-4.0, 1.0, 0.0), 1.0, Metal::new_(Vec3::new_(0.7, 0.6, 0.5), 0.0)));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, default_width, {0.95, 0.05, 64, 200}, 50, 20.0, Vec3::new_(13.0, 2.0, 3.0), Vec3::ZERO, 0.6,
	                        hdri_background(opt, Vec3::new_(0.0, -90.0, 90.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:169-215
HostScene checkered_spheres(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto checkered_texture = CheckeredTexture::from_colors(0.32, Vec3::new_(0.2, 0.3, 0.1), Vec3::new_(0.9, 0.9, 0.9));
	auto checkered_material = Lambertian::from_texture(checkered_texture);
	world->add(Sphere::new_stationary(Vec3::new_(0.0, -10.0, 0.0), 10.0, checkered_material));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 10.0, 0.0), 10.0, checkered_material));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, 400, S_32_1000_25, 50, 20.0, Vec3::new_(13.0, 2.0, 3.0), Vec3::ZERO, 0.0,
	                        Background::SOLID(Vec3::new_(0.7, 0.8, 1.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:217-253; config c2 = the same sphere at 1280 wide under the HDR background of main.rs:805-828
HostScene earth(const SceneOptions& opt, bool config2) {
	auto world = HittableList::new_();
	auto earth_texture = ImageTexture::new_(opt.asset_dir + "/earthmap.jpg");
	auto earth_material = Lambertian::from_texture(earth_texture);
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 0.0, 0.0), 2.0, earth_material));
	Background bg = config2 ? hdri_background(opt, Vec3::new_(PI / 2.0, PI, 0.0)) : Background::SOLID(Vec3::new_(0.7, 0.8, 1.0));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, config2 ? 1280 : 400, S_32_1000_25, 50, 20.0, Vec3::new_(0.0, 0.0, 12.0), Vec3::ZERO, 0.0, bg}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:255-297
HostScene perlin_spheres(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto perlin_material = Lambertian::from_texture(NoiseTexture::new_(4.0));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, -1000.0, 0.0), 1000.0, perlin_material));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 2.0, 0.0), 2.0, perlin_material));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, 400, S_32_1000_25, 50, 20.0, Vec3::new_(13.0, 2.0, 3.0), Vec3::ZERO, 0.0,
	                        Background::SOLID(Vec3::new_(0.7, 0.8, 1.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:299-364
HostScene quads(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto left_red = Lambertian::from_color(Vec3::new_(1.0, 0.2, 0.2));
	auto back_green = Lambertian::from_color(Vec3::new_(0.2, 1.0, 0.2));
	auto right_blue = Lambertian::from_color(Vec3::new_(0.2, 0.2, 1.0));
	auto upper_orange = Lambertian::from_color(Vec3::new_(1.0, 0.5, 0.0));
	auto lower_teal = Lambertian::from_color(Vec3::new_(0.2, 0.8, 0.8));
	world->add(Quad::new_(Vec3::new_(-3.0, -2.0, 5.0), Vec3::new_(0.0, 0.0, -4.0), Vec3::new_(0.0, 4.0, 0.0), left_red));
	world->add(Quad::new_(Vec3::new_(-2.0, -2.0, 0.0), Vec3::new_(4.0, 0.0, 0.0), Vec3::new_(0.0, 4.0, 0.0), back_green));
	world->add(Quad::new_(Vec3::new_(3.0, -2.0, 1.0), Vec3::new_(0.0, 0.0, 4.0), Vec3::new_(0.0, 4.0, 0.0), right_blue));
	world->add(Quad::new_(Vec3::new_(-2.0, 3.0, 1.0), Vec3::new_(4.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, 4.0), upper_orange));
	world->add(Quad::new_(Vec3::new_(-2.0, -3.0, 5.0), Vec3::new_(4.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, -4.0), lower_teal));
	HostScene s;
	s.camera = make_camera({1.0, 400, S_32_1000_25, 50, 80.0, Vec3::new_(0.0, 0.0, 9.0), Vec3::ZERO, 0.0,
	                        Background::SOLID(Vec3::new_(0.7, 0.8, 1.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:366-419
HostScene simple_light(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto noise_material = Lambertian::from_texture(NoiseTexture::new_(4.0));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, -1000.0, 0.0), 1000.0, noise_material));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 2.0, 0.0), 2.0, noise_material));
	auto light_material = DiffuseLight::from_color(Vec3::new_(4.0, 4.0, 4.0));
	world->add(Quad::new_(Vec3::new_(3.0, 1.0, -2.0), Vec3::new_(2.0, 0.0, 0.0), Vec3::new_(0.0, 2.0, 0.0), light_material));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, 1000, S_32_1000_25, 50, 20.0, Vec3::new_(26.0, 3.0, 6.0), Vec3::new_(0.0, 2.0, 0.0), 0.0,
	                        Background::SOLID(Vec3::new_(0.0, 0.0, 0.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:421-517 (cornell_box) and main.rs:519-624 (cornell_smoke)
HostScene cornell(const SceneOptions& opt, bool smoke) {
	auto world = HittableList::new_();
	auto red_material = Lambertian::from_color(Vec3::new_(0.65, 0.05, 0.05));
	auto white_material = Lambertian::from_color(Vec3::new_(0.73, 0.73, 0.73));
	auto green_material = Lambertian::from_color(Vec3::new_(0.12, 0.45, 0.15));
	auto light_material = DiffuseLight::from_color(smoke ? Vec3::new_(7.0, 7.0, 7.0) : Vec3::new_(15.0, 15.0, 15.0));
	if (smoke) world->add(Quad::new_(Vec3::new_(113.0, 554.0, 127.0), Vec3::new_(330.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, 305.0), light_material));
	else world->add(Quad::new_(Vec3::new_(343.0, 554.0, 332.0), Vec3::new_(-130.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, -105.0), light_material));
	world->add(Quad::new_(Vec3::new_(555.0, 0.0, 0.0), Vec3::new_(0.0, 555.0, 0.0), Vec3::new_(0.0, 0.0, 555.0), green_material));
	world->add(Quad::new_(Vec3::new_(0.0, 0.0, 0.0), Vec3::new_(0.0, 555.0, 0.0), Vec3::new_(0.0, 0.0, 555.0), red_material));
	world->add(Quad::new_(Vec3::new_(0.0, 0.0, 0.0), Vec3::new_(555.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, 555.0), white_material));
	world->add(Quad::new_(Vec3::new_(555.0, 555.0, 555.0), Vec3::new_(-555.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, -555.0), white_material));
	world->add(Quad::new_(Vec3::new_(0.0, 0.0, 555.0), Vec3::new_(555.0, 0.0, 0.0), Vec3::new_(0.0, 555.0, 0.0), white_material));
	HittablePtr box1 = Quad::cube(Vec3::ZERO, Vec3::new_(165.0, 330.0, 165.0), white_material);
	auto rotated_box1 = RotateY::new_(box1, 15.0);
	auto box1_final = Translate::new_(rotated_box1, Vec3::new_(265.0, 0.0, 295.0));
	if (smoke) { if (opt.include_media) world->add(ConstantMedium::from_isotropic_color(box1_final, 0.01, Vec3::new_(0.0, 0.0, 0.0))); }
	else world->add(box1_final);
	HittablePtr box2 = Quad::cube(Vec3::ZERO, Vec3::new_(165.0, 165.0, 165.0), white_material);
	auto rotated_box2 = RotateY::new_(box2, -18.0);
	auto box2_final = Translate::new_(rotated_box2, Vec3::new_(130.0, 0.0, 65.0));
	if (smoke) { if (opt.include_media) world->add(ConstantMedium::from_isotropic_color(box2_final, 0.01, Vec3::new_(1.0, 1.0, 1.0))); }
	else world->add(box2_final);
	HostScene s;
	s.camera = make_camera({1.0, 600, {0.95, smoke ? 0.25 : 0.5, 32, 1000}, 50, 40.0, Vec3::new_(278.0, 278.0, -800.0), Vec3::new_(278.0, 278.0, 0.0), 0.0,
	                        Background::SOLID(Vec3::new_(0.0, 0.0, 0.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:626-789
HostScene final_scene(const SceneOptions& opt, int image_width, uint32_t max_depth) {
	auto world = HittableList::new_();
	auto ground_material = Lambertian::from_color(Vec3::new_(0.48, 0.83, 0.53));
	const int boxes_per_side = 20;
	auto boxes = HittableList::new_();
	for (int i = 0; i < boxes_per_side; ++i) {
		for (int j = 0; j < boxes_per_side; ++j) {
			double w = 100.0;
			double x0 = -1000.0 + (double)i * w;
			double y0 = 0.0;
			double z0 = -1000.0 + (double)j * w;
			double x1 = x0 + w;
			double y1 = random_f64(1.0, 101.0);
			double z1 = z0 + w;
			boxes->add(Quad::cube(Vec3::new_(x0, y0, z0), Vec3::new_(x1, y1, z1), ground_material));
		}
	}
	world->add(BVHNode::from_list(boxes));
	auto light_material = DiffuseLight::from_color(Vec3::new_(7.0, 7.0, 7.0));
	world->add(Quad::new_(Vec3::new_(123.0, 554.0, 147.0), Vec3::new_(300.0, 0.0, 0.0), Vec3::new_(0.0, 0.0, 265.0), light_material));
	Vec3 moving_center_1 = Vec3::new_(400.0, 400.0, 200.0);
	Vec3 moving_center_2 = moving_center_1 + Vec3::new_(30.0, 0.0, 0.0);
	auto moving_material = Lambertian::from_color(Vec3::new_(0.7, 0.3, 0.1));
	world->add(Sphere::new_moving(moving_center_1, moving_center_2, 50.0, moving_material));
	world->add(Sphere::new_stationary(Vec3::new_(260.0, 150.0, 45.0), 50.0, Dielectric::new_(1.5)));
	world->add(Sphere::new_stationary(Vec3::new_(0.0, 150.0, 145.0), 50.0, Metal::new_(Vec3::new_(0.8, 0.8, 0.9), 1.0)));
	auto earth_texture = ImageTexture::new_(opt.asset_dir + "/earthmap.jpg");
	world->add(Sphere::new_stationary(Vec3::new_(400.0, 200.0, 400.0), 100.0, Lambertian::from_texture(earth_texture)));
	auto noise_texture = NoiseTexture::new_(0.2);
	world->add(Sphere::new_stationary(Vec3::new_(220.0, 280.0, 300.0), 80.0, Lambertian::from_texture(noise_texture)));
	auto fog_material = Dielectric::new_(1.5);
	auto foggy_sphere_boundary = Sphere::new_stationary(Vec3::new_(360.0, 150.0, 145.0), 70.0, fog_material);
	auto foggy_sphere_material = Lambertian::from_color(Vec3::new_(0.2, 0.4, 0.9));
	if (opt.include_media) world->add(ConstantMedium::new_(foggy_sphere_boundary, 0.2, foggy_sphere_material));
	auto world_fog_boundary = Sphere::new_stationary(Vec3::ZERO, 5000.0, fog_material);
	auto world_fog_material = Lambertian::from_color(Vec3::new_(1.0, 1.0, 1.0));
	if (opt.include_media) world->add(ConstantMedium::new_(world_fog_boundary, 0.0001, world_fog_material));
	auto ball_material = Lambertian::from_color(Vec3::new_(0.73, 0.73, 0.73));
	auto balls = HittableList::new_();
	const int ns = 1000;
	for (int i = 0; i < ns; ++i) balls->add(Sphere::new_stationary(random_vector(0.0, 165.0), 10.0, ball_material));
	world->add(Translate::new_(RotateY::new_(BVHNode::from_list(balls), 15.0), Vec3::new_(-100.0, 270.0, 395.0)));
	HostScene s;
	s.camera = make_camera({1.0, image_width, S_32_1000_25, max_depth, 40.0, Vec3::new_(478.0, 278.0, -600.0), Vec3::new_(278.0, 278.0, 0.0), 0.0,
	                        Background::SOLID(Vec3::new_(0.0, 0.0, 0.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:792-834
HostScene hdri(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto material = Metal::new_(Vec3::new_(0.7, 0.6, 0.5), 0.0);
	world->add(Sphere::new_stationary(Vec3::new_(4.0, 1.0, 0.0), 1.0, material));
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, 600, {0.95, 0.05, 64, 200}, 50, 20.0, Vec3::new_(13.0, 2.0, 5.0), Vec3::ZERO, 0.6,
	                        hdri_background(opt, Vec3::new_(PI / 2.0, PI, 0.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// main.rs:836-888
HostScene triangles(const SceneOptions& opt) {
	auto world = HittableList::new_();
	auto left_red = Lambertian::from_color(Vec3::new_(1.0, 0.2, 0.2));
	auto back_green = Lambertian::from_color(Vec3::new_(0.2, 1.0, 0.2));
	auto right_blue = Lambertian::from_color(Vec3::new_(0.2, 0.2, 1.0));   // created, unused (main.rs:842)
	auto upper_orange = Lambertian::from_color(Vec3::new_(1.0, 0.5, 0.0));
	auto lower_teal = Lambertian::from_color(Vec3::new_(0.2, 0.8, 0.8));   // created, unused (main.rs:844)
	(void)right_blue; (void)lower_teal;
	world->add(Triangle::new_(Vec3::new_(-2.0, 2.0, 0.0), Vec3::new_(-2.0, -2.0, 0.0), Vec3::new_(-2.0, -2.0, 4.0), left_red));
	world->add(Triangle::new_(Vec3::new_(-2.0, 2.0, 0.0), Vec3::new_(2.0, -2.0, 0.0), Vec3::new_(-2.0, -2.0, 0.0), back_green));
	world->add(Triangle::new_(Vec3::new_(-2.0, -2.0, 4.0), Vec3::new_(-2.0, -2.0, 0.0), Vec3::new_(2.0, -2.0, 0.0), upper_orange));
	HostScene s;
	s.camera = make_camera({1.0, 400, S_32_1000_25, 50, 80.0, Vec3::new_(0.0, 0.0, 9.0), Vec3::ZERO, 0.0,
	                        Background::SOLID(Vec3::new_(0.7, 0.8, 1.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

// Config c5 (no reference counterpart): n spheres r = 0.25 in [0,100)^3, materials as main.rs:85-103.
HostScene sphere_cloud(const SceneOptions& opt) {
	auto world = HittableList::new_();
	const int n = opt.n_spheres > 0 ? opt.n_spheres : 1000000;
	for (int i = 0; i < n; ++i) {
		Vec3 center = random_vector(0.0, 100.0);
		double material_choice = random_f64();
		std::shared_ptr<Material> material;
		if (material_choice < 0.8) {
			Vec3 c0 = random_vector(0.0, 1.0);
			Vec3 c1 = random_vector(0.0, 1.0);
			material = Lambertian::from_color(c0 * c1);
		} else if (material_choice < 0.95) {
			Vec3 albedo = random_vector(0.5, 1.0);
			double fuzz = random_f64(0.0, 0.5);
			material = Metal::new_(albedo, fuzz);
		} else {
			material = Dielectric::new_(1.5);
		}
		world->add(Sphere::new_stationary(center, 0.25, material));
	}
	HostScene s;
	s.camera = make_camera({16.0 / 9.0, 3840, {0.95, 0.0, 4096, 0}, 50, 40.0, Vec3::new_(180.0, 120.0, 180.0), Vec3::new_(50.0, 50.0, 50.0), 0.0,
	                        Background::SOLID(Vec3::new_(0.7, 0.8, 1.0))}, opt);
	s.world = BVHNode::from_list(world);
	return s;
}

}  // namespace

HostScene build_scene(const std::string& name, const SceneOptions& opt_in) {
	SceneOptions opt = opt_in;
	seed_scene_rng(opt.scene_seed);
	reset_scene_ids();
	auto fixed = [&](int spp) { if (opt.spp == 0) opt.spp = spp; };
	HostScene s;
	if (name == "c1") { fixed(100); s = bouncing_spheres(opt, 400); }
	else if (name == "c2") { fixed(1000); s = earth(opt, true); }
	else if (name == "c3") { fixed(1000); s = cornell(opt, false); }
	else if (name == "c4") { fixed(10000); s = final_scene(opt, 800, 40); }
	else if (name == "c5") { fixed(4096); s = sphere_cloud(opt); }
	else if (name == "scene0") s = bouncing_spheres(opt, 600);
	else if (name == "scene1") s = checkered_spheres(opt);
	else if (name == "scene2") s = earth(opt, false);
	else if (name == "scene3") s = perlin_spheres(opt);
	else if (name == "scene4") s = quads(opt);
	else if (name == "scene5") s = simple_light(opt);
	else if (name == "scene6") s = cornell(opt, false);
	else if (name == "scene7") s = cornell(opt, true);
	else if (name == "scene8") s = final_scene(opt, 400, 50);
	else if (name == "scene9") s = final_scene(opt, 800, 40);
	else if (name == "scene11") s = hdri(opt);
	else if (name == "scene10") s = triangles(opt);
	else throw std::runtime_error("unknown scene '" + name + "'");
	s.name = name;
	return s;
}

}  // namespace grayshift
