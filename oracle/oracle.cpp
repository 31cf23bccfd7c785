// oracle/oracle.cpp -- see oracle.hpp.  TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (see header).
// Each function cites the reference file:line (relative to /root/reference) it restates.
#include "oracle.hpp"

#include <chrono>
#include <fstream>
#include <mutex>
#include <stdexcept>
#include <thread>

namespace orc {

thread_local Counters tls_counters;
thread_local Rng tls_rng;
bool g_triangle_respects_interval = false;

// =========================================================================================
// Third-party: noise 0.9.0 Perlin (call sites texture.rs:3,99,104,113-117).  UNPINNED.
// Published algorithm: PermutationTable::new(seed) seeds rand_xorshift::XorShiftRng from the
// 16 bytes [1,0,0,0, seed_le x3] and Fisher-Yates-shuffles 0..=255 with rand 0.8's
// SliceRandom::shuffle (gen_index -> UniformInt<u32>::sample_single, widening-multiply zone
// rejection); hash = perm[perm[perm[x&255]^(y&255)]^(z&255)]; gradient = the 12 edge vectors
// picked by hash&15 with cases 12..15 repeating (1,1,0), (-1,1,0), (0,-1,1), (0,-1,-1) in the order
// of Perlin's "Improving Noise" paper, i.e. 0|12, 1|13, 9|14, 11|15 as core::perlin::perlin_3d's
// gradient_dot_v writes them (NOT the 12, 14 / 13 swap of Perlin's Java bit trick); quintic fade;
// result * 2/sqrt(3) clamped to [-1,1].
// Pinned pieces: XorShift against rand_xorshift's own test_xorshift_true_values (tests/test_oracle_kat.py);
// gen_index's arithmetic against the first draw of rand 0.8's value_stability_slice.  The shuffle as a
// whole, the gradient aliasing above and the table it yields stay unpinned.
// =========================================================================================
struct XorShift {
	uint32_t x, y, z, w;
	uint32_t next() {
		uint32_t t = x ^ (x << 11);
		x = y; y = z; z = w;
		w = w ^ (w >> 19) ^ (t ^ (t >> 8));
		return w;
	}
};
uint32_t rand08_gen_index(XorShift& rng, uint32_t ubound) {
	// UniformInt<u32>::sample_single(0, ubound): range = ubound; zone = (range << lz) - 1
	uint32_t range = ubound;
	uint32_t zone = (range << __builtin_clz(range)) - 1u;
	for (;;) {
		uint32_t v = rng.next();
		uint64_t m = (uint64_t)v * range;
		uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
		if (lo <= zone) return hi;
	}
}

// KAT hooks (tests/test_oracle_kat.py): the generator from 16 seed bytes (rand_core's from_seed reads them as four
// little-endian words; an all-zero seed is replaced), and gen_index over a caller-supplied word stream.
void xorshift_words(const uint8_t seed[16], size_t n, uint32_t* out) {
	auto le = [&](int i) { return (uint32_t)seed[i] | ((uint32_t)seed[i + 1] << 8) | ((uint32_t)seed[i + 2] << 16) | ((uint32_t)seed[i + 3] << 24); };
	XorShift rng{le(0), le(4), le(8), le(12)};
	if (!(rng.x | rng.y | rng.z | rng.w)) rng = XorShift{0x193a6754u, 0xa8a7d469u, 0x97830e05u, 0x113ba7bbu};
	for (size_t i = 0; i < n; ++i) out[i] = rng.next();
}
uint32_t gen_index_from_words(const uint32_t* words, size_t n_words, uint32_t ubound, size_t* used) {
	struct Stream { const uint32_t* w; size_t n, k; uint32_t next() { if (k >= n) throw std::runtime_error("gen_index: word stream exhausted"); return w[k++]; } } st{words, n_words, 0};
	uint32_t range = ubound;
	uint32_t zone = (range << __builtin_clz(range)) - 1u;
	for (;;) {
		uint32_t v = st.next();
		uint64_t m = (uint64_t)v * range;
		if ((uint32_t)m <= zone) { if (used) *used = st.k; return (uint32_t)(m >> 32); }
	}
}

void perlin_permutation_table(uint32_t seed, uint8_t out[256]) {
	uint8_t real[16] = {0};
	real[0] = 1;
	for (int i = 1; i < 4; ++i) {
		real[i * 4] = (uint8_t)seed; real[i * 4 + 1] = (uint8_t)(seed >> 8);
		real[i * 4 + 2] = (uint8_t)(seed >> 16); real[i * 4 + 3] = (uint8_t)(seed >> 24);
	}
	auto le = [&](int i) { return (uint32_t)real[i] | ((uint32_t)real[i + 1] << 8) | ((uint32_t)real[i + 2] << 16) | ((uint32_t)real[i + 3] << 24); };
	XorShift rng{le(0), le(4), le(8), le(12)};
	if (!(rng.x | rng.y | rng.z | rng.w)) rng = XorShift{0x193a6754u, 0xa8a7d469u, 0x97830e05u, 0x113ba7bbu};
	for (int i = 0; i < 256; ++i) out[i] = (uint8_t)i;
	for (uint32_t i = 255; i >= 1; --i) {  // for i in (1..len).rev(): swap(i, gen_index(rng, i+1))
		uint32_t j = rand08_gen_index(rng, i + 1);
		std::swap(out[i], out[j]);
	}
}

Perlin::Perlin(uint32_t seed) { perlin_permutation_table(seed, perm); }

static inline double perlin_grad3(unsigned h, double x, double y, double z) {
	switch (h & 15u) {
		case 0: case 12: return x + y;
		case 1: case 13: return -x + y;
		case 2: return x - y;
		case 3: return -x - y;
		case 4: return x + z;
		case 5: return -x + z;
		case 6: return x - z;
		case 7: return -x - z;
		case 8: return y + z;
		case 9: case 14: return -y + z;
		case 10: return y - z;
		default: return -y - z;  // 11, 15
	}
}

double Perlin::get(double px, double py, double pz) const {
	const double SCALE_FACTOR = 1.1547005383792515;  // 2/sqrt(3)
	double fx = std::floor(px), fy = std::floor(py), fz = std::floor(pz);
	long cx = (long)fx, cy = (long)fy, cz = (long)fz;
	double dx = px - fx, dy = py - fy, dz = pz - fz;
	auto hash = [&](long x, long y, long z) -> unsigned {
		unsigned a = (unsigned)(x & 0xff);
		a = perm[a] ^ (unsigned)(y & 0xff);
		a = perm[a] ^ (unsigned)(z & 0xff);
		return perm[a];
	};
	auto g = [&](int ox, int oy, int oz) { return perlin_grad3(hash(cx + ox, cy + oy, cz + oz), dx - ox, dy - oy, dz - oz); };
	double g000 = g(0, 0, 0), g100 = g(1, 0, 0), g010 = g(0, 1, 0), g110 = g(1, 1, 0);
	double g001 = g(0, 0, 1), g101 = g(1, 0, 1), g011 = g(0, 1, 1), g111 = g(1, 1, 1);
	auto quintic = [](double t) { return t * t * t * (t * (t * 6.0 - 15.0) + 10.0); };
	double a = quintic(dx), b = quintic(dy), c = quintic(dz);
	double k0 = g000, k1 = g100 - g000, k2 = g010 - g000, k3 = g001 - g000;
	double k4 = g000 + g110 - g100 - g010, k5 = g000 + g101 - g100 - g001, k6 = g000 + g011 - g010 - g001;
	double k7 = g100 + g010 + g001 + g111 - g000 - g110 - g101 - g011;
	double unscaled = k0 + k1 * a + k2 * b + k3 * c + k4 * a * b + k5 * a * c + k6 * b * c + k7 * a * b * c;
	double scaled = unscaled * SCALE_FACTOR;
	return scaled < -1.0 ? -1.0 : (scaled > 1.0 ? 1.0 : scaled);
}

// =========================================================================================
// Textures -- texture.rs
// =========================================================================================
static inline int f64_to_i32_sat(double v) {  // Rust `as i32`: saturating, NaN -> 0
	if (std::isnan(v)) return 0;
	if (v >= 2147483647.0) return 2147483647;
	if (v <= -2147483648.0) return (-2147483647 - 1);
	return (int)v;
}
static inline uint64_t f64_to_usize_sat(double v) {  // Rust `as usize` / `as u32`: negatives and NaN -> 0
	if (!(v > 0.0)) return 0;
	if (v >= 18446744073709551615.0) return UINT64_MAX;
	return (uint64_t)v;
}

Vec3 CheckeredTexture::value_at(double u, double v, const Vec3& p) const {  // texture.rs:57-70
	int x_int = f64_to_i32_sat(std::floor(scale_inv * p.x));
	int y_int = f64_to_i32_sat(std::floor(scale_inv * p.y));
	int z_int = f64_to_i32_sat(std::floor(scale_inv * p.z));
	// Rust release build wraps on overflow; `%` keeps the sign, so -1 counts as odd.
	int sum = (int)((unsigned)x_int + (unsigned)y_int + (unsigned)z_int);
	bool is_even = (sum % 2) == 0;
	return is_even ? even_texture->value_at(u, v, p) : odd_texture->value_at(u, v, p);
}

Vec3 ImageTexture::value_at(double u, double v, const Vec3&) const {  // texture.rs:84-94
	ORC_COUNT(tex_fetches);
	double u_clamp = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
	double vc = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
	double v_clamp = 1.0 - vc;
	uint64_t i = f64_to_usize_sat(u_clamp * (double)image->w);
	uint64_t j = f64_to_usize_sat(v_clamp * (double)image->h);
	// Q13: the reference panics when i == W or j == H (u == 1 or v == 0 exactly); clamp instead.
	if (i >= (uint64_t)image->w) i = image->w - 1;
	if (j >= (uint64_t)image->h) j = image->h - 1;
	const uint8_t* px = &image->rgb[(j * image->w + i) * 3];
	return Vec3((double)px[0], (double)px[1], (double)px[2]) * (1.0 / 255.0);
}

double NoiseTexture::turbulence(const Vec3& p, uint32_t depth) const {  // texture.rs:107-123
	double accum = 0.0;
	Vec3 sample_point = p;
	double weight = 1.0;
	for (uint32_t i = 0; i < depth; ++i) {
		accum += weight * noise.get(sample_point.x, sample_point.y, sample_point.z);
		weight /= 2.0;
		sample_point *= 2.0;
	}
	return std::fabs(accum);
}
Vec3 NoiseTexture::value_at(double, double, const Vec3& p) const {  // texture.rs:127-130
	return Vec3(0.5, 0.5, 0.5) * (1.0 + std::sin(scale * p.z + 10.0 * turbulence(p, 7)));
}

// =========================================================================================
// Materials -- material.rs
// =========================================================================================
std::optional<ScatterRecord> Lambertian::scatter(const Ray& ray_in, const HitRecord& rec) const {  // material.rs:45-68
	Vec3 attenuation = texture->value_at(rec.u, rec.v, rec.position);
	OrthonormalBasis basis(rec.normal);
	Vec3 scatter_direction = basis.transform(random_cosine_direction()).unit();
	Ray scattered(rec.position, scatter_direction, ray_in.time);
	double pdf = basis.w.dot(scatter_direction) / PI;  // computed and never used (Q2)
	return ScatterRecord{scattered, attenuation, pdf};
}

std::optional<ScatterRecord> Metal::scatter(const Ray& ray_in, const HitRecord& rec) const {  // material.rs:87-102
	Vec3 reflected = ray_in.direction.reflect(rec.normal);
	Vec3 ru = reflected.unit();
	reflected = ru + fuzz * random_unit_vector();  // drawn even when fuzz == 0 (Q7)
	Ray scattered(rec.position, reflected, ray_in.time);
	if (scattered.direction.dot(rec.normal) > 0.0) return ScatterRecord{scattered, albedo, 0.0};
	return std::nullopt;
}

double Dielectric::reflectance(double cosine, double refraction_index) {  // material.rs:114-119
	double r0 = (1.0 - refraction_index) / (1.0 + refraction_index);
	r0 = r0 * r0;
	double m = 1.0 - cosine;
	return r0 + (1.0 - r0) * (m * m * m * m * m);  // f64::powi(.., 5)
}
std::optional<ScatterRecord> Dielectric::scatter(const Ray& ray_in, const HitRecord& rec) const {  // material.rs:122-148
	double ri = rec.is_front_face ? 1.0 / refraction_index : refraction_index;
	Vec3 unit_direction = ray_in.direction.unit();
	double cos_theta = std::fmin((-unit_direction).dot(rec.normal), 1.0);
	double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
	bool cannot_refract = ri * sin_theta > 1.0;
	bool fresnel_reflection = reflectance(cos_theta, ri) > rand_f64();  // always drawn (Q7)
	Vec3 direction = (cannot_refract || fresnel_reflection) ? unit_direction.reflect(rec.normal)
	                                                        : unit_direction.refract(rec.normal, ri);
	return ScatterRecord{Ray(rec.position, direction, ray_in.time), Vec3(1.0, 1.0, 1.0), 0.0};
}

std::optional<ScatterRecord> Isotropic::scatter(const Ray& ray_in, const HitRecord& rec) const {  // material.rs:186-196
	Vec3 att = texture->value_at(rec.u, rec.v, rec.position);
	Vec3 dir = random_unit_vector();
	return ScatterRecord{Ray(rec.position, dir, ray_in.time), att, 1.0 / (4.0 * PI)};
}

// =========================================================================================
// Hittables
// =========================================================================================
std::optional<HitRecord> HittableList::hit(const Ray& ray, const Interval& ray_t) const {  // hittable.rs:70-86
	double closest_hit = ray_t.max;
	std::optional<HitRecord> hit_record;
	for (const auto& object : objects) {
		if (auto rec = object->hit(ray, Interval(ray_t.min, closest_hit))) {
			closest_hit = rec->t;
			hit_record = std::move(rec);
		}
	}
	return hit_record;
}

std::shared_ptr<BVHNode> BVHNode::construct_tree(std::vector<std::shared_ptr<Hittable>> objects) {  // BVH.rs:18-65
	if (objects.empty()) throw std::runtime_error("BVHNode::construct_tree: empty list (the reference recurses forever)");
	auto node = std::make_shared<BVHNode>();
	if (objects.size() == 1) {
		node->bbox = objects[0]->bounding_box();
		node->left = objects[0];
		return node;
	}
	if (objects.size() == 2) {  // given order, no sort (BVH.rs:30-42)
		node->bbox = AABB::from_pair(objects[0]->bounding_box(), objects[1]->bounding_box());
		node->left = objects[0];
		node->right = objects[1];
		return node;
	}
	AABB bbox;  // EMPTY
	for (const auto& o : objects) bbox = AABB::from_pair(bbox, o->bounding_box());
	int axis = bbox.longest_axis();
	std::stable_sort(objects.begin(), objects.end(), [axis](const std::shared_ptr<Hittable>& a, const std::shared_ptr<Hittable>& b) {
		return a->bounding_box().axis(axis).min < b->bounding_box().axis(axis).min;  // slice::sort_by is stable
	});
	size_t middle = objects.size() / 2;
	std::vector<std::shared_ptr<Hittable>> right_objs(objects.begin() + middle, objects.end());
	objects.resize(middle);
	node->left = construct_tree(std::move(objects));
	node->right = construct_tree(std::move(right_objs));
	node->bbox = bbox;
	return node;
}

std::optional<HitRecord> BVHNode::hit(const Ray& ray, const Interval& ray_t) const {  // BVH.rs:68-90
	if (!bbox.hit(ray, ray_t)) return std::nullopt;
	if (auto hit_left = left->hit(ray, ray_t)) {
		if (right) {
			auto hit_right = right->hit(ray, Interval(ray_t.min, hit_left->t));
			if (hit_right) return hit_right;
			return hit_left;
		}
		return hit_left;
	}
	if (right) return right->hit(ray, ray_t);
	return std::nullopt;
}

std::shared_ptr<Sphere> Sphere::new_stationary(const Vec3& c, double r, std::shared_ptr<Material> m) {  // sphere.rs:21-33
	auto s = std::make_shared<Sphere>();
	Vec3 r_vec(r, r, r);
	s->bbox = AABB::from_corners(c - r_vec, c + r_vec);
	s->center_start = c; s->center_path = Vec3(0, 0, 0); s->is_moving = false; s->radius = r; s->material = std::move(m);
	return s;
}
std::shared_ptr<Sphere> Sphere::new_moving(const Vec3& c1, const Vec3& c2, double r, std::shared_ptr<Material> m) {  // sphere.rs:35-49
	auto s = std::make_shared<Sphere>();
	Vec3 r_vec(r, r, r);
	AABB b0 = AABB::from_corners(c1 - r_vec, c1 + r_vec);
	AABB b1 = AABB::from_corners(c2 - r_vec, c2 + r_vec);
	s->bbox = AABB::from_pair(b0, b1);
	s->center_start = c1; s->center_path = c2 - c1; s->center_end = c2; s->is_moving = true; s->radius = r; s->material = std::move(m);
	return s;
}
std::optional<HitRecord> Sphere::hit(const Ray& ray, const Interval& ray_t) const {  // sphere.rs:64-106
	ORC_COUNT(prim_tests);
	Vec3 center = is_moving ? center_start + ray.time * center_path : center_start;  // sphere.rs:51-53
	Vec3 oc = center - ray.origin;
	double a = ray.direction.length_squared();
	double h = ray.direction.dot(oc);
	double c = oc.length_squared() - radius * radius;
	double discriminant = h * h - a * c;
	if (discriminant < 0.0) return std::nullopt;
	double sqrt_d = std::sqrt(discriminant);
	double t = (h - sqrt_d) / a;
	if (!ray_t.surrounds(t)) {
		t = (h + sqrt_d) / a;
		if (!ray_t.surrounds(t)) return std::nullopt;
	}
	Vec3 hit_pos = ray.at(t);
	Vec3 outward_normal = (hit_pos - center) / radius;
	// sphere_uv, sphere.rs:55-60
	double theta = std::acos(-outward_normal.y);
	double phi = std::atan2(-outward_normal.z, outward_normal.x) + PI;
	return HitRecord(ray, t, hit_pos, outward_normal, material, phi / (2.0 * PI), theta / PI, prim_id);
}

Quad::Quad(const Vec3& q_, const Vec3& u_, const Vec3& v_, std::shared_ptr<Material> m)  // quad.rs:25-38
	: q(q_), u(u_), v(v_), material(std::move(m)) {
	AABB d1 = AABB::from_corners(q, q + u + v);
	AABB d2 = AABB::from_corners(q + u, q + v);
	bbox = AABB::from_pair(d1, d2);
	Vec3 n = u.cross(v);
	normal = n.unit();
	w = n / n.dot(n);
	d = normal.dot(q);  // Plane::new, plane.rs:15-18
}
std::optional<HitRecord> Quad::hit(const Ray& ray, const Interval& ray_t) const {  // quad.rs:84-109 + plane.rs:20-32
	ORC_COUNT(prim_tests);
	double denominator = normal.dot(ray.direction);
	if (std::fabs(denominator) < 1e-8) return std::nullopt;
	double t = (d - normal.dot(ray.origin)) / denominator;
	if (!ray_t.contains(t)) return std::nullopt;  // inclusive (Q9)
	Vec3 intersection = ray.at(t);
	Vec3 planar_hit = intersection - q;
	double alpha = w.dot(planar_hit.cross(v));
	double beta = w.dot(u.cross(planar_hit));
	Interval unit(0.0, 1.0);
	if (!unit.contains(alpha) || !unit.contains(beta)) return std::nullopt;
	return HitRecord(ray, t, intersection, normal, material, alpha, beta, prim_id);
}

Triangle::Triangle(const Vec3& a_, const Vec3& b_, const Vec3& c_, std::shared_ptr<Material> m)  // triangle.rs:20-28
	: a(a_), b(b_), c(c_), material(std::move(m)) {
	normal = (b - a).cross(c - a);
	bbox = AABB::from_pair(AABB::from_corners(a, b), AABB::from_corners(a, c));
}
std::optional<HitRecord> Triangle::hit(const Ray& ray, const Interval& ray_t) const {  // triangle.rs:34-68: ignores ray_t (Q14)
	ORC_COUNT(prim_tests);
	Vec3 edge_1 = c - a, edge_2 = b - a;
	Vec3 p_vec = ray.direction.cross(edge_2);
	double det = edge_1.dot(p_vec);
	if (det < 1e-8) return std::nullopt;
	Vec3 t_vec = ray.origin - a;
	double u = t_vec.dot(p_vec);
	if (u < 0.0 || u > det) return std::nullopt;
	Vec3 q_vec = t_vec.cross(edge_1);
	double v = ray.direction.dot(q_vec);
	if (v < 0.0 || u + v > det) return std::nullopt;
	double t = edge_2.dot(q_vec);
	double inv_det = 1.0 / det;
	t *= inv_det; u *= inv_det; v *= inv_det;
	if (g_triangle_respects_interval && !ray_t.surrounds(t)) return std::nullopt;  // NOT in the reference; see oracle.hpp
	return HitRecord(ray, t, ray.at(t), normal, material, u, v, prim_id);
}

std::optional<HitRecord> Translate::hit(const Ray& ray, const Interval& ray_t) const {  // hittable.rs:107-120
	Ray offset_ray(ray.origin - offset, ray.direction, ray.time);
	auto rec = object->hit(offset_ray, ray_t);
	if (rec) rec->position += offset;
	return rec;
}

RotateY::RotateY(std::shared_ptr<Hittable> o, double angle) : object(std::move(o)), angle_deg(angle) {  // hittable.rs:135-176
	double radians = deg_to_rad(angle);
	sin_theta = std::sin(radians);
	cos_theta = std::cos(radians);
	AABB b = object->bounding_box();
	Vec3 mn(F64_MAX, F64_MAX, F64_MAX), mx(F64_MIN, F64_MIN, F64_MIN);
	for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
		double x = (double)i * b.x.max + (double)(1 - i) * b.x.min;
		double y = (double)j * b.y.max + (double)(1 - j) * b.y.min;
		double z = (double)k * b.z.max + (double)(1 - k) * b.z.min;
		double new_x = cos_theta * x + sin_theta * z;
		double new_z = -sin_theta * x + cos_theta * z;
		mn.x = std::fmin(mn.x, new_x); mx.x = std::fmax(mx.x, new_x);
		mn.y = std::fmin(mn.y, y);     mx.y = std::fmax(mx.y, y);
		mn.z = std::fmin(mn.z, new_z); mx.z = std::fmax(mx.z, new_z);
	}
	bbox = AABB::from_corners(mn, mx);
}
std::optional<HitRecord> RotateY::hit(const Ray& ray, const Interval& ray_t) const {  // hittable.rs:179-211
	Vec3 origin((cos_theta * ray.origin.x) - (sin_theta * ray.origin.z), ray.origin.y,
	            (sin_theta * ray.origin.x) + (cos_theta * ray.origin.z));
	Vec3 direction((cos_theta * ray.direction.x) - (sin_theta * ray.direction.z), ray.direction.y,
	               (sin_theta * ray.direction.x) + (cos_theta * ray.direction.z));
	Ray rotated_ray(origin, direction, ray.time);
	auto rec = object->hit(rotated_ray, ray_t);
	if (rec) {
		rec->position = Vec3((cos_theta * rec->position.x) + (sin_theta * rec->position.z), rec->position.y,
		                     (-sin_theta * rec->position.x) + (cos_theta * rec->position.z));
		rec->normal = Vec3((cos_theta * rec->normal.x) + (sin_theta * rec->normal.z), rec->normal.y,
		                   (-sin_theta * rec->normal.x) + (cos_theta * rec->normal.z));
	}
	return rec;
}

std::optional<HitRecord> ConstantMedium::hit(const Ray& ray, const Interval& ray_t) const {  // volume.rs:32-63
	auto r1 = boundary->hit(ray, Interval::universe());
	if (!r1) return std::nullopt;
	auto r2 = boundary->hit(ray, Interval(r1->t + 0.0001, F64_MAX));
	if (!r2) return std::nullopt;
	double t1 = r1->t, t2 = r2->t;
	if (t1 < ray_t.min) t1 = ray_t.min;
	if (t2 > ray_t.max) t2 = ray_t.max;
	if (t1 >= t2) return std::nullopt;
	if (t1 < 0.0) t1 = 0.0;
	double ray_len = ray.direction.length();
	double dist_inside_boundary = (t2 - t1) * ray_len;
	double hit_dist = density_neg_inv * std::log(rand_f64());  // RNG inside hit (H5)
	if (hit_dist > dist_inside_boundary) return std::nullopt;
	double t = t1 + hit_dist / ray_len;
	return HitRecord(ray, t, ray.at(t), Vec3(1.0, 0.0, 0.0), phase_function, 0.0, 0.0, -2 - medium_index);
}

// =========================================================================================
// Camera -- camera.rs
// =========================================================================================
Camera::Camera(double aspect_ratio, int image_width_, SampleSettings s, uint32_t max_depth_, double v_fov,
               const Vec3& look_from, const Vec3& look_at, const Vec3& vup, double defocus_angle_,
               double focus_distance, Background background_) {  // camera.rs:39-98
	image_width = image_width_;
	image_height = (int)((double)image_width / aspect_ratio);  // truncating `as i32`
	double theta = v_fov / 180.0 * PI;
	double h = std::tan(theta / 2.0);
	double viewport_height = 2.0 * h * focus_distance;
	double viewport_width = viewport_height * ((double)image_width / (double)image_height);
	Vec3 w = (look_from - look_at).unit();
	Vec3 u = vup.cross(w).unit();
	Vec3 v = w.cross(u);
	Vec3 viewport_u = viewport_width * u;
	Vec3 viewport_v = viewport_height * -v;
	pixel_delta_u = viewport_u / (double)image_width;
	pixel_delta_v = viewport_v / (double)image_height;
	Vec3 viewport_upper_left = look_from - focus_distance * w - viewport_u / 2.0 - viewport_v / 2.0;
	starting_pixel_pos = viewport_upper_left + 0.5 * (pixel_delta_u + pixel_delta_v);
	double defocus_radius = focus_distance * std::tan(deg_to_rad(defocus_angle_ / 2.0));
	defocus_disk_u = u * defocus_radius;
	defocus_disk_v = v * defocus_radius;
	sample_settings = s; max_depth = max_depth_; center = look_from; background = std::move(background_);
	defocus_angle = defocus_angle_;
	arg_aspect = aspect_ratio; arg_vfov = v_fov; arg_focus = focus_distance;
	arg_from = look_from; arg_at = look_at; arg_vup = vup;
}

Ray Camera::get_ray(int i, int j) const {  // camera.rs:204-221; defocus_disk_sample :223-226
	double offset_x = rand_f64() - 0.5;
	double offset_y = rand_f64() - 0.5;
	Vec3 pixel_sample = starting_pixel_pos + ((double)i + offset_x) * pixel_delta_u + ((double)j + offset_y) * pixel_delta_v;
	Vec3 ray_origin;
	if (defocus_angle <= 0.0) ray_origin = center;
	else {
		Vec3 v = random_vector_in_unit_disk();
		ray_origin = center + v.x * defocus_disk_u + v.y * defocus_disk_v;
	}
	Vec3 ray_direction = pixel_sample - ray_origin;  // not normalised (Q6)
	double ray_time = rand_f64();
	return Ray(ray_origin, ray_direction, ray_time);
}

Vec3 HDRI::sample(const Vec3& direction) const {  // camera.rs:257-270
	ORC_COUNT(tex_fetches);
	Vec3 rotated = rotate_vector(direction, rotation).unit();
	double theta = std::atan2(rotated.y, rotated.x);
	double phi = std::asin(rotated.z);
	double u = 0.5 + theta / (2.0 * PI);
	double v = 0.5 - phi / PI;
	uint64_t x = f64_to_usize_sat(u * (double)image->w) % (uint64_t)image->w;
	uint64_t y = f64_to_usize_sat(v * (double)image->h) % (uint64_t)image->h;
	const float* c = &image->rgb[(y * image->w + x) * 3];
	return Vec3((double)c[0], (double)c[1], (double)c[2]);
}

Vec3 Camera::sample_background(const Ray& ray) const {  // camera.rs:228-233
	return background.is_hdri ? background.hdri.sample(ray.direction) : background.solid;
}

Vec3 Camera::ray_color(const Ray& ray, uint32_t depth, const Hittable& world) const {  // camera.rs:174-202
	if (depth <= 0) return Vec3(0, 0, 0);
	ORC_COUNT_ALWAYS(rays);
	if (auto rec = world.hit(ray, Interval(0.001, F64_MAX))) {
		Vec3 emission_color = rec->material->emitted(rec->u, rec->v, rec->position);
		if (auto sr = rec->material->scatter(ray, *rec)) {
			Vec3 scatter_color = ray_color(sr->scattered_ray, depth - 1, world);
			Vec3 color_from_scatter = scatter_color * sr->attenuation;
			return emission_color + color_from_scatter;
		}
		return emission_color;
	}
	return sample_background(ray);
}

Vec3 Camera::sample(int i, int j, const Hittable& world, uint64_t seed, uint32_t sample_begin, double* count_out) const {  // camera.rs:125-171
	Vec3 pixel_color(0, 0, 0);
	double tolerance_sq = sample_settings.tolerance * sample_settings.tolerance;
	double confidence_sq = sample_settings.confidence * sample_settings.confidence;
	double sum = 0.0, sq_sum = 0.0, sample_count = 0.0;
	uint32_t pixel = (uint32_t)(j * image_width + i);
	uint32_t sidx = sample_begin;
	for (;;) {
		sample_count += (double)sample_settings.batch_size;
		for (uint32_t k = 0; k < sample_settings.batch_size; ++k) {
			if (tls_rng.mode == RngMode::PHILOX) tls_rng.start_path(seed, pixel, sidx);
			++sidx;
			ORC_COUNT_ALWAYS(paths);
			Ray ray = get_ray(i, j);
			Vec3 sample_color = ray_color(ray, max_depth, world);
			pixel_color += sample_color;
			double lum = luminance(sample_color);
			sum += lum;
			sq_sum += lum * lum;
		}
		double mean = sum / sample_count;
		double variance_sq = 1.0 / (sample_count - 1.0) * (sq_sum - sum * sum / sample_count);
		double convergence_sq = confidence_sq * variance_sq / sample_count;
		if (convergence_sq < (mean * mean * tolerance_sq)) break;
		uint32_t sc_u32 = sample_count >= 4294967295.0 ? 4294967295u : (uint32_t)sample_count;
		if (sc_u32 > sample_settings.max_samples) break;
	}
	pixel_color /= sample_count;
	if (count_out) *count_out = sample_count;
	return pixel_color;
}

RenderResult render(const Camera& cam, const Hittable& world, uint64_t seed, RngMode mode, int threads,
                    uint32_t sample_begin, int pixel_begin, int pixel_end) {  // camera.rs:100-121
	const int W = cam.image_width, H = cam.image_height;
	const int n_pixels = W * H;
	if (pixel_end < 0 || pixel_end > n_pixels) pixel_end = n_pixels;
	if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
	RenderResult out;
	out.rgb.assign((size_t)n_pixels * 3, 0.0);
	out.counts.assign((size_t)n_pixels, 0);
	std::atomic<int> next(pixel_begin);
	const int chunk = 32;
	std::mutex mu;
	auto t0 = std::chrono::steady_clock::now();
	auto worker = [&](int tid) {
		tls_counters = Counters();
		tls_rng.mode = mode;
		tls_rng.wy = 0x9E3779B97F4A7C15ULL * (uint64_t)(tid + 1) ^ seed;
		for (;;) {
			int begin = next.fetch_add(chunk);
			if (begin >= pixel_end) break;
			int end = std::min(begin + chunk, pixel_end);
			for (int n = begin; n < end; ++n) {
				int i = n % W, j = n / W;
				double count = 0;
				Vec3 c = cam.sample(i, j, world, seed, sample_begin, &count);
				out.rgb[(size_t)n * 3 + 0] = c.x; out.rgb[(size_t)n * 3 + 1] = c.y; out.rgb[(size_t)n * 3 + 2] = c.z;
				out.counts[n] = (uint32_t)count;
			}
		}
		std::lock_guard<std::mutex> lk(mu);
		out.counters.add(tls_counters);
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < threads; ++t) pool.emplace_back(worker, t);
	worker(0);
	for (auto& t : pool) t.join();
	out.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	return out;
}

void write_ppm(const std::string& path, int w, int h, const std::vector<double>& rgb) {  // camera.rs:101-103,116-118; color.rs:8-19
	FILE* f = std::fopen(path.c_str(), "w");
	if (!f) throw std::runtime_error("cannot open " + path);
	std::fprintf(f, "P3\n%d %d\n255\n", w, h);
	for (int n = 0; n < w * h; ++n)
		std::fprintf(f, "%d %d %d\n", color_byte(rgb[n * 3]), color_byte(rgb[n * 3 + 1]), color_byte(rgb[n * 3 + 2]));
	std::fclose(f);
}

// =========================================================================================
// Image loading (host-side in the reference: image::open texture.rs:79, radiant::load main.rs:139)
// =========================================================================================
static std::vector<uint8_t> read_file(const std::string& path) {
	std::ifstream f(path, std::ios::binary);
	if (!f) throw std::runtime_error("cannot open " + path);
	return std::vector<uint8_t>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

std::shared_ptr<ImageU8> load_ppm_u8(const std::string& path) {
	auto data = read_file(path);
	size_t pos = 0;
	auto token = [&]() {
		for (;;) {
			while (pos < data.size() && std::isspace(data[pos])) ++pos;
			if (pos < data.size() && data[pos] == '#') { while (pos < data.size() && data[pos] != '\n') ++pos; continue; }
			break;
		}
		std::string s;
		while (pos < data.size() && !std::isspace(data[pos])) s.push_back((char)data[pos++]);
		return s;
	};
	if (token() != "P6") throw std::runtime_error(path + ": not a binary PPM (P6)");
	auto img = std::make_shared<ImageU8>();
	img->w = std::stoi(token()); img->h = std::stoi(token());
	if (std::stoi(token()) != 255) throw std::runtime_error(path + ": maxval must be 255");
	++pos;  // single whitespace after maxval
	size_t n = (size_t)img->w * img->h * 3;
	if (data.size() - pos < n) throw std::runtime_error(path + ": truncated");
	img->rgb.assign(data.begin() + pos, data.begin() + pos + n);
	img->source = "ppm:" + path;
	return img;
}

std::shared_ptr<ImageU8> load_raw_u8(const std::string& path, int w, int h) {
	auto data = read_file(path);
	if (data.size() != (size_t)w * h * 3) throw std::runtime_error(path + ": size mismatch");
	auto img = std::make_shared<ImageU8>();
	img->w = w; img->h = h; img->rgb = std::move(data); img->source = "raw:" + path;
	return img;
}
std::shared_ptr<ImageF32> load_raw_f32(const std::string& path, int w, int h) {
	auto data = read_file(path);
	if (data.size() != (size_t)w * h * 12) throw std::runtime_error(path + ": size mismatch");
	auto img = std::make_shared<ImageF32>();
	img->w = w; img->h = h; img->rgb.resize((size_t)w * h * 3); img->source = "raw:" + path;
	std::memcpy(img->rgb.data(), data.data(), data.size());
	return img;
}

// Third-party radiant 0.3.0 (main.rs:139,806), restated (UNPINNED): Radiance RGBE with new-style
// RLE scanlines; value = mantissa * 2^(e - 136), e == 0 -> 0.
std::shared_ptr<ImageF32> load_hdr_rgbe(const std::string& path) {
	auto data = read_file(path);
	size_t pos = 0;
	auto line = [&]() { std::string s; while (pos < data.size() && data[pos] != '\n') s.push_back((char)data[pos++]); ++pos; return s; };
	std::string magic = line();
	if (magic.rfind("#?", 0) != 0) throw std::runtime_error(path + ": not a Radiance file");
	for (;;) { std::string l = line(); if (l.empty()) break; if (pos >= data.size()) throw std::runtime_error(path + ": bad header"); }
	std::string res = line();
	int H = 0, W = 0;
	if (std::sscanf(res.c_str(), "-Y %d +X %d", &H, &W) != 2) throw std::runtime_error(path + ": unsupported orientation '" + res + "'");
	auto img = std::make_shared<ImageF32>();
	img->w = W; img->h = H; img->rgb.resize((size_t)W * H * 3); img->source = "hdr:" + path;
	std::vector<uint8_t> scan((size_t)W * 4);
	for (int y = 0; y < H; ++y) {
		if (pos + 4 > data.size()) throw std::runtime_error(path + ": truncated");
		bool rle = W >= 8 && W < 32768 && data[pos] == 2 && data[pos + 1] == 2 && ((data[pos + 2] << 8) | data[pos + 3]) == W;
		if (rle) {
			pos += 4;
			for (int ch = 0; ch < 4; ++ch) {
				int x = 0;
				while (x < W) {
					if (pos >= data.size()) throw std::runtime_error(path + ": truncated");
					int count = data[pos++];
					if (count > 128) {
						count -= 128;
						if (x + count > W || pos >= data.size()) throw std::runtime_error(path + ": bad run");
						uint8_t v = data[pos++];
						for (int k = 0; k < count; ++k) scan[(size_t)(x++) * 4 + ch] = v;
					} else {
						if (count == 0 || x + count > W || pos + count > data.size()) throw std::runtime_error(path + ": bad literal");
						for (int k = 0; k < count; ++k) scan[(size_t)(x++) * 4 + ch] = data[pos++];
					}
				}
			}
		} else {
			if (pos + (size_t)W * 4 > data.size()) throw std::runtime_error(path + ": truncated");
			std::memcpy(scan.data(), &data[pos], (size_t)W * 4);
			pos += (size_t)W * 4;
		}
		for (int x = 0; x < W; ++x) {
			const uint8_t* p = &scan[(size_t)x * 4];
			float* o = &img->rgb[((size_t)y * W + x) * 3];
			if (p[3] == 0) { o[0] = o[1] = o[2] = 0.0f; }
			else {
				float f = std::ldexp(1.0f, (int)p[3] - 136);
				o[0] = (float)p[0] * f; o[1] = (float)p[1] * f; o[2] = (float)p[2] * f;
			}
		}
	}
	return img;
}

// =========================================================================================
// Scene builders (ids)
// =========================================================================================
std::shared_ptr<Sphere> Scene::sphere(const Vec3& c, double r, std::shared_ptr<Material> m) {
	auto s = addn(Sphere::new_stationary(c, r, std::move(m)));
	s->prim_id = prim_count++;
	return s;
}
std::shared_ptr<Sphere> Scene::moving_sphere(const Vec3& c1, const Vec3& c2, double r, std::shared_ptr<Material> m) {
	auto s = addn(Sphere::new_moving(c1, c2, r, std::move(m)));
	s->prim_id = prim_count++;
	return s;
}
std::shared_ptr<Quad> Scene::quad(const Vec3& q, const Vec3& u, const Vec3& v, std::shared_ptr<Material> m) {
	auto s = addn(std::make_shared<Quad>(q, u, v, std::move(m)));
	s->prim_id = prim_count++;
	return s;
}
std::shared_ptr<Triangle> Scene::triangle(const Vec3& a, const Vec3& b, const Vec3& c, std::shared_ptr<Material> m) {
	auto s = addn(std::make_shared<Triangle>(a, b, c, std::move(m)));
	s->prim_id = prim_count++;
	return s;
}
std::shared_ptr<HittableList> Scene::cube(const Vec3& point_a, const Vec3& point_b, std::shared_ptr<Material> material) {  // quad.rs:54-80
	Vec3 mn(std::fmin(point_a.x, point_b.x), std::fmin(point_a.y, point_b.y), std::fmin(point_a.z, point_b.z));
	Vec3 mx(std::fmax(point_a.x, point_b.x), std::fmax(point_a.y, point_b.y), std::fmax(point_a.z, point_b.z));
	Vec3 dx(mx.x - mn.x, 0.0, 0.0), dy(0.0, mx.y - mn.y, 0.0), dz(0.0, 0.0, mx.z - mn.z);
	std::shared_ptr<Hittable> sides6[6] = {
		quad(Vec3(mn.x, mn.y, mx.z), dx, dy, material),
		quad(Vec3(mx.x, mn.y, mx.z), -dz, dy, material),
		quad(Vec3(mx.x, mn.y, mn.z), -dx, dy, material),
		quad(Vec3(mn.x, mn.y, mn.z), dz, dy, material),
		quad(Vec3(mn.x, mx.y, mx.z), dx, -dz, material),
		quad(Vec3(mn.x, mn.y, mn.z), dx, dz, material)};
	auto sides = list();
	for (auto& s : sides6) sides->add(s);
	return sides;
}
std::shared_ptr<BVHNode> Scene::bvh_from_list(const HittableList& l) {  // BVH.rs:15-17
	auto root = BVHNode::construct_tree(l.objects);
	for (const auto& o : l.objects) root->source_children.push_back(o->id);
	return addn(root);
}

}  // namespace orc
