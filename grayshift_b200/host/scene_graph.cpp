// grayshift_b200/host/scene_graph.cpp -- constructors of the host-side scene graph (see grayshift.hpp).
// Bounding boxes follow the reference formulas exactly because they determine the order in which the
// reference's BVH visits primitives (BVH.rs:44-55), which this host reproduces for tie-breaking.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <stdexcept>

#include "grayshift.hpp"

namespace grayshift {

const Vec3 Vec3::ZERO = Vec3(0.0, 0.0, 0.0);
double Vec3::length() const { return std::sqrt(length_squared()); }
Vec3 Vec3::unit() const { return *this / length(); }

Interval::Interval() : min(DBL_MAX), max(-DBL_MAX) {}  // EMPTY (interval.rs:10; f64::MIN is -MAX)
Interval Interval::from_interval_pair(const Interval& a, const Interval& b) {  // interval.rs:18-23
	return Interval(a.min <= b.min ? a.min : b.min, a.max >= b.max ? a.max : b.max);
}
Interval Interval::expand(double delta) const {  // interval.rs:41-44
	double padding = delta / 2.0;
	return Interval(min - padding, max + padding);
}

AABB AABB::from_corners(const Vec3& a, const Vec3& b) {  // AABB.rs:24-48
	AABB r(a.x <= b.x ? Interval(a.x, b.x) : Interval(b.x, a.x), a.y <= b.y ? Interval(a.y, b.y) : Interval(b.y, a.y),
	       a.z <= b.z ? Interval(a.z, b.z) : Interval(b.z, a.z));
	const double delta = 0.0001;  // pad_to_minimums, AABB.rs:123-128
	if (r.x.size() < delta) r.x = r.x.expand(delta);
	if (r.y.size() < delta) r.y = r.y.expand(delta);
	if (r.z.size() < delta) r.z = r.z.expand(delta);
	return r;
}
AABB AABB::from_AABB_pair(const AABB& a, const AABB& b) {  // AABB.rs:50-56
	return AABB(Interval::from_interval_pair(a.x, b.x), Interval::from_interval_pair(a.y, b.y), Interval::from_interval_pair(a.z, b.z));
}
int AABB::longest_axis() const {  // AABB.rs:115-121
	if (x.size() > y.size()) return x.size() > z.size() ? 0 : 2;
	return y.size() > z.size() ? 1 : 2;
}
AABB AABB::operator+(const Vec3& o) const {  // AABB.rs:144-153
	return AABB(Interval(x.min + o.x, x.max + o.x), Interval(y.min + o.y, y.max + o.y), Interval(z.min + o.z, z.max + o.z));
}

// ---- scene RNG + ids --------------------------------------------------------------------------------
namespace {
thread_local uint64_t g_rng_state = 1;
thread_local int g_next_prim_id = 0;
thread_local int g_next_medium = 0;
}  // namespace
void seed_scene_rng(uint64_t seed) { g_rng_state = seed; }
double random_f64() {  // SplitMix64 -> 53-bit uniform
	g_rng_state += 0x9E3779B97F4A7C15ULL;
	uint64_t z = g_rng_state;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	z ^= z >> 31;
	return (double)(z >> 11) * (1.0 / 9007199254740992.0);
}
double random_f64(double min, double max) { return random_f64() * (max - min) + min; }
Vec3 random_vector(double min, double max) {
	double a = random_f64(min, max), b = random_f64(min, max), c = random_f64(min, max);
	return Vec3(a, b, c);
}
void reset_scene_ids() { g_next_prim_id = 0; g_next_medium = 0; }

// ---- textures -----------------------------------------------------------------------------------------
std::shared_ptr<Texture> SolidColorTexture::new_(const Vec3& albedo) {
	auto t = std::make_shared<Texture>();
	t->kind = GS_TEX_SOLID; t->albedo = albedo;
	return t;
}
std::shared_ptr<Texture> SolidColorTexture::from_rgb(double r, double g, double b) { return new_(Vec3(r, g, b)); }
std::shared_ptr<Texture> CheckeredTexture::new_(double scale, std::shared_ptr<Texture> even, std::shared_ptr<Texture> odd) {
	auto t = std::make_shared<Texture>();
	t->kind = GS_TEX_CHECKER; t->scale = scale; t->even_texture = std::move(even); t->odd_texture = std::move(odd);
	return t;
}
std::shared_ptr<Texture> CheckeredTexture::from_colors(double scale, const Vec3& even, const Vec3& odd) {
	auto e = SolidColorTexture::new_(even);
	auto o = SolidColorTexture::new_(odd);
	return new_(scale, e, o);
}
std::shared_ptr<Texture> ImageTexture::new_(const std::string& filepath) { return from_image(load_image_u8(filepath)); }
std::shared_ptr<Texture> ImageTexture::from_image(std::shared_ptr<ImageData> img) {
	if (!img || img->format != GS_IMG_U8_RGB) throw std::runtime_error("ImageTexture needs an 8-bit RGB image");
	auto t = std::make_shared<Texture>();
	t->kind = GS_TEX_IMAGE; t->image = std::move(img);
	return t;
}
std::shared_ptr<Texture> NoiseTexture::new_(double scale) {
	auto t = std::make_shared<Texture>();
	t->kind = GS_TEX_NOISE; t->scale = scale;
	return t;
}

// noise 0.9.0 PermutationTable::new(seed) (third-party; published algorithm, unpinned -- DESIGN.md):
// XorShiftRng seeded from [1,0,0,0, seed_le x3], rand 0.8 Fisher-Yates shuffle of 0..=255.
void perlin_permutation_table(uint32_t seed, uint8_t out[256]) {
	uint32_t x = 1, y = seed, z = seed, w = seed;
	if (!(x | y | z | w)) { x = 0x193a6754u; y = 0xa8a7d469u; z = 0x97830e05u; w = 0x113ba7bbu; }
	auto next = [&]() {
		uint32_t t = x ^ (x << 11);
		x = y; y = z; z = w;
		w = w ^ (w >> 19) ^ (t ^ (t >> 8));
		return w;
	};
	for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
	for (uint32_t i = 255; i >= 1; --i) {
		uint32_t range = i + 1;
		uint32_t zone = (range << __builtin_clz(range)) - 1u;
		uint32_t j;
		for (;;) {
			uint64_t m = (uint64_t)next() * range;
			if ((uint32_t)m <= zone) { j = (uint32_t)(m >> 32); break; }
		}
		std::swap(out[i], out[j]);
	}
}

// ---- materials ------------------------------------------------------------------------------------------
static std::shared_ptr<Material> make_material(int kind, std::shared_ptr<Texture> t, const Vec3& albedo, double param) {
	auto m = std::make_shared<Material>();
	m->kind = kind; m->texture = std::move(t); m->albedo = albedo; m->param = param;
	return m;
}
std::shared_ptr<Material> Lambertian::from_color(const Vec3& albedo) { return make_material(GS_MAT_LAMBERTIAN, SolidColorTexture::new_(albedo), Vec3(), 0); }
std::shared_ptr<Material> Lambertian::from_texture(std::shared_ptr<Texture> t) { return make_material(GS_MAT_LAMBERTIAN, std::move(t), Vec3(), 0); }
std::shared_ptr<Material> Metal::new_(const Vec3& albedo, double fuzz) { return make_material(GS_MAT_METAL, nullptr, albedo, fuzz); }
std::shared_ptr<Material> Dielectric::new_(double ri) { return make_material(GS_MAT_DIELECTRIC, nullptr, Vec3(), ri); }
std::shared_ptr<Material> DiffuseLight::new_(std::shared_ptr<Texture> t) { return make_material(GS_MAT_DIFFUSE_LIGHT, std::move(t), Vec3(), 0); }
std::shared_ptr<Material> DiffuseLight::from_color(const Vec3& c) { return new_(SolidColorTexture::new_(c)); }
std::shared_ptr<Material> Isotropic::new_(std::shared_ptr<Texture> t) { return make_material(GS_MAT_ISOTROPIC, std::move(t), Vec3(), 0); }
std::shared_ptr<Material> Isotropic::from_color(const Vec3& c) { return new_(SolidColorTexture::new_(c)); }

// ---- hittables ------------------------------------------------------------------------------------------
std::shared_ptr<Sphere> Sphere::new_stationary(const Vec3& center, double radius, std::shared_ptr<Material> m) {  // sphere.rs:21-33
	auto s = std::make_shared<Sphere>();
	Vec3 r_vec(radius, radius, radius);
	s->bbox = AABB::from_corners(center - r_vec, center + r_vec);
	s->center_start = center; s->center_end = center; s->radius = radius; s->material = std::move(m);
	s->prim_id = g_next_prim_id++;
	return s;
}
std::shared_ptr<Sphere> Sphere::new_moving(const Vec3& c1, const Vec3& c2, double radius, std::shared_ptr<Material> m) {  // sphere.rs:35-49
	auto s = std::make_shared<Sphere>();
	Vec3 r_vec(radius, radius, radius);
	s->bbox = AABB::from_AABB_pair(AABB::from_corners(c1 - r_vec, c1 + r_vec), AABB::from_corners(c2 - r_vec, c2 + r_vec));
	s->center_start = c1; s->center_path = c2 - c1; s->center_end = c2; s->is_moving = true; s->radius = radius; s->material = std::move(m);
	s->prim_id = g_next_prim_id++;
	return s;
}

void HittableList::add(HittablePtr object) {  // hittable.rs:63-66
	bbox = AABB::from_AABB_pair(bbox, object->bounding_box());
	objects.push_back(std::move(object));
}

std::shared_ptr<Quad> Quad::new_(const Vec3& q, const Vec3& u, const Vec3& v, std::shared_ptr<Material> m) {  // quad.rs:25-38
	auto s = std::make_shared<Quad>();
	s->bbox = AABB::from_AABB_pair(AABB::from_corners(q, q + u + v), AABB::from_corners(q + u, q + v));
	s->q = q; s->u = u; s->v = v; s->material = std::move(m);
	s->prim_id = g_next_prim_id++;
	return s;
}
std::shared_ptr<HittableList> Quad::cube(const Vec3& a, const Vec3& b, std::shared_ptr<Material> material) {  // quad.rs:54-80
	auto sides = HittableList::new_();
	Vec3 mn(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z));
	Vec3 mx(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z));
	Vec3 dx(mx.x - mn.x, 0.0, 0.0), dy(0.0, mx.y - mn.y, 0.0), dz(0.0, 0.0, mx.z - mn.z);
	sides->add(Quad::new_(Vec3(mn.x, mn.y, mx.z), dx, dy, material));
	sides->add(Quad::new_(Vec3(mx.x, mn.y, mx.z), -dz, dy, material));
	sides->add(Quad::new_(Vec3(mx.x, mn.y, mn.z), -dx, dy, material));
	sides->add(Quad::new_(Vec3(mn.x, mn.y, mn.z), dz, dy, material));
	sides->add(Quad::new_(Vec3(mn.x, mx.y, mx.z), dx, -dz, material));
	sides->add(Quad::new_(Vec3(mn.x, mn.y, mn.z), dx, dz, material));
	return sides;
}

std::shared_ptr<Triangle> Triangle::new_(const Vec3& a, const Vec3& b, const Vec3& c, std::shared_ptr<Material> m) {  // triangle.rs:20-28
	auto t = std::make_shared<Triangle>();
	t->bbox = AABB::from_AABB_pair(AABB::from_corners(a, b), AABB::from_corners(a, c));
	t->a = a; t->b = b; t->c = c; t->material = std::move(m);
	t->prim_id = g_next_prim_id++;
	return t;
}

std::shared_ptr<BVHNode> BVHNode::construct_tree(std::vector<HittablePtr> objects) {  // BVH.rs:18-65
	if (objects.empty()) throw std::runtime_error("BVHNode::from_list: empty list (the reference recurses forever)");
	auto node = std::make_shared<BVHNode>();
	if (objects.size() == 1) {
		node->bbox = objects[0]->bounding_box();
		node->left = objects[0];
		return node;
	}
	if (objects.size() == 2) {
		node->bbox = AABB::from_AABB_pair(objects[0]->bounding_box(), objects[1]->bounding_box());
		node->left = objects[0]; node->right = objects[1];
		return node;
	}
	AABB bbox;
	for (const auto& o : objects) bbox = AABB::from_AABB_pair(bbox, o->bounding_box());
	int axis = bbox.longest_axis();
	for (const auto& o : objects) {
		double m = o->bounding_box()[axis].min;
		if (m != m) throw std::runtime_error("BVHNode::from_list: NaN bounding box (the reference panics, BVH.rs:54)");
	}
	std::stable_sort(objects.begin(), objects.end(), [axis](const HittablePtr& a, const HittablePtr& b) {
		return a->bounding_box()[axis].min < b->bounding_box()[axis].min;
	});
	size_t middle = objects.size() / 2;
	std::vector<HittablePtr> right_objs(objects.begin() + middle, objects.end());
	objects.resize(middle);
	node->left = construct_tree(std::move(objects));
	node->right = construct_tree(std::move(right_objs));
	node->bbox = bbox;
	return node;
}
std::shared_ptr<BVHNode> BVHNode::from_list(const HittableList& list) {  // BVH.rs:15-17
	auto root = construct_tree(list.objects);
	root->source = list.objects;
	return root;
}

std::shared_ptr<Translate> Translate::new_(HittablePtr object, const Vec3& offset) {  // hittable.rs:100-103
	auto t = std::make_shared<Translate>();
	t->bbox = object->bounding_box() + offset;
	t->object = std::move(object); t->offset = offset;
	return t;
}

std::shared_ptr<RotateY> RotateY::new_(HittablePtr object, double angle) {  // hittable.rs:135-176
	auto r = std::make_shared<RotateY>();
	double radians = angle / 180.0 * 3.14159265358979323846;  // deg_to_rad, util.rs:62-64
	r->angle = angle; r->sin_theta = std::sin(radians); r->cos_theta = std::cos(radians);
	AABB b = object->bounding_box();
	Vec3 mn(DBL_MAX, DBL_MAX, DBL_MAX), mx(-DBL_MAX, -DBL_MAX, -DBL_MAX);
	for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
		double x = (double)i * b.x.max + (double)(1 - i) * b.x.min;
		double y = (double)j * b.y.max + (double)(1 - j) * b.y.min;
		double z = (double)k * b.z.max + (double)(1 - k) * b.z.min;
		double new_x = r->cos_theta * x + r->sin_theta * z;
		double new_z = -r->sin_theta * x + r->cos_theta * z;
		mn.x = std::fmin(mn.x, new_x); mx.x = std::fmax(mx.x, new_x);
		mn.y = std::fmin(mn.y, y);     mx.y = std::fmax(mx.y, y);
		mn.z = std::fmin(mn.z, new_z); mx.z = std::fmax(mx.z, new_z);
	}
	r->bbox = AABB::from_corners(mn, mx);
	r->object = std::move(object);
	return r;
}

std::shared_ptr<ConstantMedium> ConstantMedium::new_(HittablePtr boundary, double density, std::shared_ptr<Material> phase) {  // volume.rs:17-21
	auto m = std::make_shared<ConstantMedium>();
	m->boundary = std::move(boundary); m->density = density; m->phase_function = std::move(phase);
	m->medium_index = g_next_medium++;
	return m;
}
std::shared_ptr<ConstantMedium> ConstantMedium::from_isotropic_color(HittablePtr boundary, double density, const Vec3& color) {  // volume.rs:23-28
	return new_(std::move(boundary), density, Isotropic::from_color(color));
}

}  // namespace grayshift
