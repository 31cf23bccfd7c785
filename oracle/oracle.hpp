// oracle/oracle.hpp -- CPU restatement of benjisu03/grayshift's path-tracing hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under grayshift_b200/ may include, link or call this.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
//
// PARITY UNPINNED: the reference (a Rust crate) ships no tests, no golden vectors and no
// fixtures, and cannot be compiled here (no cargo/rustc, no vendored crates).  This file
// follows the reference line by line (every function cites the reference file:line it
// restates, paths relative to /root/reference) and is pinned only against analytic
// known answers (tests/test_oracle_kat.py).  Third-party arithmetic whose source is not in
// the reference tree (noise 0.9.0 Perlin, radiant 0.3.0 RGBE decode, image 0.25 JPEG decode,
// fastrand 2.1 RNG) is restated from the published algorithms; see DESIGN.md.
//
// All arithmetic is f64 in the reference's operation order; build with -ffp-contract=off
// (rustc never contracts to FMA).  Virtual dispatch, per-node heap allocation, recursive
// unordered BVH traversal and a ref-counted material handle per accepted hit are kept on
// purpose so the cost structure resembles the Rust program: this is also the timed CPU baseline.
#pragma once
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <optional>
#include <string>
#include <vector>

namespace orc {

constexpr double PI = 3.14159265358979323846;  // std::f64::consts::PI
constexpr double F64_MAX = DBL_MAX;            // f64::MAX
constexpr double F64_MIN = -DBL_MAX;           // f64::MIN is the most negative finite (interval.rs:10-11)

// ---------------------------------------------------------------------------------------
// Counters (reference-semantics work counts; SURVEY 8(d) algorithmic bytes use these).
// ---------------------------------------------------------------------------------------
struct Counters {
	uint64_t rays = 0;       // world.hit calls from ray_color (camera.rs:177)
	uint64_t box_tests = 0;  // AABB::hit calls (AABB.rs:58)
	uint64_t prim_tests = 0; // Sphere/Quad/Triangle::hit calls
	uint64_t tex_fetches = 0;// ImageTexture::value_at + HDRI::sample texel reads
	uint64_t paths = 0;      // samples taken
	void add(const Counters& o) {
		rays += o.rays; box_tests += o.box_tests; prim_tests += o.prim_tests;
		tex_fetches += o.tex_fetches; paths += o.paths;
	}
};
extern thread_local Counters tls_counters;
// false (default) = the reference's Triangle::hit, which ignores ray_t (Q14).  true = apply t_min < t < t_max, the
// order-independent rule the device uses; tests compare the two on the reference's triangle scene.
extern bool g_triangle_respects_interval;
#ifdef ORC_COUNTERS
#define ORC_COUNT(field) (++::orc::tls_counters.field)
#else
#define ORC_COUNT(field) ((void)0)
#endif
#define ORC_COUNT_ALWAYS(field) (++::orc::tls_counters.field)  // rays/paths: one increment per ray, negligible

// ---------------------------------------------------------------------------------------
// RNG.  The reference draws from fastrand's unseeded thread-local generator; its stream
// cannot be matched.  Two modes:
//   PHILOX: counter-based Philox4x32-10 keyed by (seed), counter (pixel, sample, block, stream);
//           draw k of a path = word k&3 of block k>>2, as u32 * 2^-32.  The device uses the same
//           scheme, so camera rays and early bounces line up one to one.
//   WYRAND: thread-local Wyrand (what fastrand 2.x is), 52 mantissa bits; used for CPU timing.
// ---------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
	uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
	uint32_t k0 = key_in[0], k1 = key_in[1];
	for (int r = 0; r < 10; ++r) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0;
		uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t)p1;
		uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum class RngMode { PHILOX = 0, WYRAND = 1 };

struct Rng {
	RngMode mode = RngMode::WYRAND;
	// philox state
	uint32_t key[2] = {0, 0};
	uint32_t ctr[4] = {0, 0, 0, 0};
	uint32_t buf[4] = {0, 0, 0, 0};
	uint32_t draw = 0;
	// wyrand state
	uint64_t wy = 0x853c49e6748fea9bULL;

	void start_path(uint64_t seed, uint32_t pixel, uint32_t sample) {
		key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
		ctr[0] = pixel; ctr[1] = sample; ctr[2] = 0; ctr[3] = 0;
		draw = 0;
	}
	inline double f64() {
		if (mode == RngMode::PHILOX) {
			uint32_t w = draw & 3u;
			if (w == 0) { ctr[2] = draw >> 2; philox4x32_10(ctr, key, buf); }
			++draw;
			return (double)buf[w] * (1.0 / 4294967296.0);
		}
		// fastrand 2.x: Wyrand step, f64 = from_bits(1.0 | u64>>12) - 1.0
		wy += 0xa0761d6478bd642fULL;
		__uint128_t t = (__uint128_t)wy * (wy ^ 0xe7037ed1a0b428dbULL);
		uint64_t r = (uint64_t)(t >> 64) ^ (uint64_t)t;
		uint64_t bits = 0x3ff0000000000000ULL | (r >> 12);
		double d; std::memcpy(&d, &bits, 8);
		return d - 1.0;
	}
};
extern thread_local Rng tls_rng;
inline double rand_f64() { return tls_rng.f64(); }  // stands in for fastrand::f64()

// Scene-construction RNG (main.rs draws scene parameters from the unseeded RNG;
// here SplitMix64 so the host mirror and the oracle build the same scene from a seed).
struct SceneRng {
	uint64_t s;
	explicit SceneRng(uint64_t seed) : s(seed) {}
	double f64() {
		s += 0x9E3779B97F4A7C15ULL;
		uint64_t z = s;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
		z ^= z >> 31;
		return (double)(z >> 11) * (1.0 / 9007199254740992.0);
	}
	double range(double mn, double mx) { return f64() * (mx - mn) + mn; }  // util.rs:5-7
};

// ---------------------------------------------------------------------------------------
// Vec3 -- util/vec3.rs:6-178
// ---------------------------------------------------------------------------------------
struct Vec3 {
	double x = 0, y = 0, z = 0;
	Vec3() = default;
	Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
	double length_squared() const { return x * x + y * y + z * z; }          // vec3.rs:19-21
	double length() const { return std::sqrt(length_squared()); }            // vec3.rs:23-25
	inline Vec3 unit() const;                                                // vec3.rs:27-29
	double dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }   // vec3.rs:31-35
	Vec3 cross(const Vec3& o) const {                                        // vec3.rs:37-43
		return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x);
	}
	inline Vec3 reflect(const Vec3& n) const;                                // vec3.rs:53-55
	inline Vec3 refract(const Vec3& n, double ratio) const;                  // vec3.rs:57-62
};
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(double s, const Vec3& a) { return a * s; }            // vec3.rs:123-128
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }  // vec3.rs:150-159 (3 divides)
inline Vec3& operator+=(Vec3& a, const Vec3& b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
inline Vec3& operator*=(Vec3& a, double s) { a.x *= s; a.y *= s; a.z *= s; return a; }
inline Vec3& operator/=(Vec3& a, double s) { a.x /= s; a.y /= s; a.z /= s; return a; }
inline Vec3 Vec3::unit() const { return *this / length(); }
inline Vec3 Vec3::reflect(const Vec3& n) const { return *this - 2.0 * this->dot(n) * n; }
inline Vec3 Vec3::refract(const Vec3& n, double ratio) const {
	double cos_theta = std::fmin(n.dot(-(*this)), 1.0);
	Vec3 r_out_perp = ratio * (*this + cos_theta * n);
	Vec3 r_out_parallel = -std::sqrt(std::fabs(1.0 - r_out_perp.length_squared())) * n;
	return r_out_perp + r_out_parallel;
}

// ---------------------------------------------------------------------------------------
// Interval -- util/interval.rs:3-55
// ---------------------------------------------------------------------------------------
struct Interval {
	double min, max;
	Interval() : min(F64_MAX), max(F64_MIN) {}                               // EMPTY, interval.rs:10
	Interval(double mn, double mx) : min(mn), max(mx) {}
	static Interval empty() { return Interval(F64_MAX, F64_MIN); }
	static Interval universe() { return Interval(F64_MIN, F64_MAX); }       // interval.rs:11
	static Interval from_pair(const Interval& a, const Interval& b) {       // interval.rs:18-23
		return Interval(a.min <= b.min ? a.min : b.min, a.max >= b.max ? a.max : b.max);
	}
	double size() const { return max - min; }
	bool contains(double t) const { return min <= t && t <= max; }           // interval.rs:29-31
	bool surrounds(double t) const { return min < t && t < max; }            // interval.rs:33-35
	double clamp(double t) const { return t < min ? min : (t > max ? max : t); }
	Interval expand(double delta) const {                                    // interval.rs:41-44
		double padding = delta / 2.0;
		return Interval(min - padding, max + padding);
	}
	Interval operator+(double rhs) const { return Interval(min + rhs, max + rhs); }
};

// ---------------------------------------------------------------------------------------
// Ray -- ray.rs:3-17
// ---------------------------------------------------------------------------------------
struct Ray {
	Vec3 origin, direction;
	double time = 0;
	Ray() = default;
	Ray(const Vec3& o, const Vec3& d, double t) : origin(o), direction(d), time(t) {}
	Vec3 at(double t) const { return origin + direction * t; }
};

// ---------------------------------------------------------------------------------------
// AABB -- AABB.rs:6-153
// ---------------------------------------------------------------------------------------
struct AABB {
	Interval x, y, z;  // default = EMPTY (AABB.rs:14-18)
	AABB() = default;
	AABB(const Interval& x_, const Interval& y_, const Interval& z_) : x(x_), y(y_), z(z_) {}
	static AABB from_corners(const Vec3& a, const Vec3& b) {                 // AABB.rs:24-48
		AABB r(a.x <= b.x ? Interval(a.x, b.x) : Interval(b.x, a.x),
		       a.y <= b.y ? Interval(a.y, b.y) : Interval(b.y, a.y),
		       a.z <= b.z ? Interval(a.z, b.z) : Interval(b.z, a.z));
		r.pad_to_minimums();
		return r;
	}
	static AABB from_pair(const AABB& a, const AABB& b) {                    // AABB.rs:50-56 (no padding)
		return AABB(Interval::from_pair(a.x, b.x), Interval::from_pair(a.y, b.y), Interval::from_pair(a.z, b.z));
	}
	const Interval& axis(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	int longest_axis() const {                                               // AABB.rs:115-121 (strict >)
		if (x.size() > y.size()) return x.size() > z.size() ? 0 : 2;
		return y.size() > z.size() ? 1 : 2;
	}
	void pad_to_minimums() {                                                 // AABB.rs:123-128
		const double delta = 0.0001;
		if (x.size() < delta) x = x.expand(delta);
		if (y.size() < delta) y = y.expand(delta);
		if (z.size() < delta) z = z.expand(delta);
	}
	AABB operator+(const Vec3& o) const { return AABB(x + o.x, y + o.y, z + o.z); }  // AABB.rs:144-153
	// AABB.rs:58-113: one divide per axis per call, ordered clip, reject when max <= min.
	bool hit(const Ray& ray, Interval interval) const {
		ORC_COUNT(box_tests);
		{
			double ad_inv = 1.0 / ray.direction.x;
			double t0 = (x.min - ray.origin.x) * ad_inv;
			double t1 = (x.max - ray.origin.x) * ad_inv;
			if (t0 < t1) { if (t0 > interval.min) interval.min = t0; if (t1 < interval.max) interval.max = t1; }
			else         { if (t1 > interval.min) interval.min = t1; if (t0 < interval.max) interval.max = t0; }
			if (interval.max <= interval.min) return false;
		}
		{
			double ad_inv = 1.0 / ray.direction.y;
			double t0 = (y.min - ray.origin.y) * ad_inv;
			double t1 = (y.max - ray.origin.y) * ad_inv;
			if (t0 < t1) { if (t0 > interval.min) interval.min = t0; if (t1 < interval.max) interval.max = t1; }
			else         { if (t1 > interval.min) interval.min = t1; if (t0 < interval.max) interval.max = t0; }
			if (interval.max <= interval.min) return false;
		}
		{
			double ad_inv = 1.0 / ray.direction.z;
			double t0 = (z.min - ray.origin.z) * ad_inv;
			double t1 = (z.max - ray.origin.z) * ad_inv;
			if (t0 < t1) { if (t0 > interval.min) interval.min = t0; if (t1 < interval.max) interval.max = t1; }
			else         { if (t1 > interval.min) interval.min = t1; if (t0 < interval.max) interval.max = t0; }
			if (interval.max <= interval.min) return false;
		}
		return true;
	}
};

// ---------------------------------------------------------------------------------------
// util.rs:5-86
// ---------------------------------------------------------------------------------------
inline double random_f64(double mn, double mx) { return rand_f64() * (mx - mn) + mn; }          // util.rs:5-7
inline Vec3 random_vector(double mn, double mx) {                                               // util.rs:9-16
	double a = random_f64(mn, mx), b = random_f64(mn, mx), c = random_f64(mn, mx);
	return Vec3(a, b, c);
}
inline Vec3 random_vector_in_unit_sphere() {                                                    // util.rs:18-25
	for (;;) { Vec3 v = random_vector(-1.0, 1.0); if (v.length_squared() < 1.0) return v; }
}
inline Vec3 random_unit_vector() { return random_vector_in_unit_sphere().unit(); }             // util.rs:27-29
inline Vec3 random_vector_in_unit_disk() {                                                      // util.rs:36-46
	for (;;) {
		double a = random_f64(-1.0, 1.0), b = random_f64(-1.0, 1.0);
		Vec3 v(a, b, 0.0);
		if (v.length_squared() < 1.0) return v;
	}
}
inline Vec3 random_cosine_direction() {                                                         // util.rs:48-60 (quirk Q1)
	double r_1 = rand_f64();
	double r_2 = rand_f64();
	double phi = 2.0 * PI * r_1;
	double r_2_sqrt = std::sqrt(r_2);
	double x = std::cos(phi) * std::sqrt(r_2_sqrt);
	double y = std::sin(phi) * std::sqrt(r_2_sqrt);
	double z = std::sqrt(1.0 - r_2);
	return Vec3(x, y, z);
}
inline double deg_to_rad(double degrees) { return degrees / 180.0 * PI; }                        // util.rs:62-64
inline Vec3 rotate_vector(const Vec3& v, const Vec3& rot) {                                     // util.rs:67-86 (radians)
	double sin_x = std::sin(rot.x), cos_x = std::cos(rot.x);
	double sin_y = std::sin(rot.y), cos_y = std::cos(rot.y);
	double sin_z = std::sin(rot.z), cos_z = std::cos(rot.z);
	double x = v.x * (cos_y * cos_z) + v.y * (cos_x * sin_z + sin_x * sin_y * cos_z) + v.z * (sin_x * sin_z - cos_x * sin_y * cos_z);
	double y = v.x * (-cos_y * sin_z) + v.y * (cos_x * cos_z - sin_x * sin_y * sin_z) + v.z * (sin_x * cos_z + cos_x * sin_y * sin_z);
	double z = v.x * sin_y + v.y * (-sin_x * cos_y) + v.z * (cos_x * cos_y);
	return Vec3(x, y, z);
}

// ---------------------------------------------------------------------------------------
// ONB -- ONB.rs:3-27
// ---------------------------------------------------------------------------------------
struct OrthonormalBasis {
	Vec3 u, v, w;
	explicit OrthonormalBasis(const Vec3& normal) {
		w = normal.unit();
		Vec3 a = std::fabs(w.x) > 0.9 ? Vec3(0.0, 1.0, 0.0) : Vec3(1.0, 0.0, 0.0);
		v = w.cross(a).unit();
		u = w.cross(v);
	}
	Vec3 transform(const Vec3& p) const { return u * p.x + v * p.y + w * p.z; }
};

// ---------------------------------------------------------------------------------------
// color.rs:6-33
// ---------------------------------------------------------------------------------------
inline double luminance(const Vec3& v) { return 0.299 * v.x + 0.587 * v.y + 0.144 * v.z; }       // color.rs:31-33 (0.144 sic)
inline double linear_to_gamma(double n) { return n > 0.0 ? std::sqrt(n) : 0.0; }                // color.rs:21-27
inline int color_byte(double c) {                                                               // color.rs:8-15
	double g = linear_to_gamma(c);
	double cl = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);
	return (int)(256.0 * cl);
}

// ---------------------------------------------------------------------------------------
// Images (decoded pixels).  radiant::Image{width,height,data: Vec<RGB f32>} (camera.rs:252,269)
// and image::DynamicImage 8-bit RGB (texture.rs:73-95).
// ---------------------------------------------------------------------------------------
struct ImageU8 { int w = 0, h = 0; std::vector<uint8_t> rgb; std::string source; };
struct ImageF32 { int w = 0, h = 0; std::vector<float> rgb; std::string source; };
std::shared_ptr<ImageU8> load_ppm_u8(const std::string& path);      // P6 binary (pre-decoded JPEG)
std::shared_ptr<ImageF32> load_hdr_rgbe(const std::string& path);   // Radiance RGBE, RLE
std::shared_ptr<ImageF32> load_raw_f32(const std::string& path, int w, int h);
std::shared_ptr<ImageU8> load_raw_u8(const std::string& path, int w, int h);

// ---------------------------------------------------------------------------------------
// Perlin -- third-party `noise 0.9.0` (texture.rs:3,99,104,113-117), restated (UNPINNED).
// ---------------------------------------------------------------------------------------
struct Perlin {
	uint8_t perm[256];
	explicit Perlin(uint32_t seed = 0);       // Perlin::default() -> seed 0
	double get(double px, double py, double pz) const;
};
void perlin_permutation_table(uint32_t seed, uint8_t out[256]);
// KAT hooks for the third-party pieces behind the table (rand_xorshift 0.3, rand 0.8 gen_index); see oracle.cpp
void xorshift_words(const uint8_t seed[16], size_t n, uint32_t* out);
uint32_t gen_index_from_words(const uint32_t* words, size_t n_words, uint32_t ubound, size_t* used);

// ---------------------------------------------------------------------------------------
// Textures -- texture.rs:9-131
// ---------------------------------------------------------------------------------------
struct Texture {
	int id = -1;  // graph id (oracle bookkeeping only)
	virtual ~Texture() = default;
	virtual Vec3 value_at(double u, double v, const Vec3& p) const = 0;
};
struct SolidColorTexture : Texture {                                       // texture.rs:13-31
	Vec3 albedo;
	explicit SolidColorTexture(const Vec3& a) : albedo(a) {}
	Vec3 value_at(double, double, const Vec3&) const override { return albedo; }
};
struct CheckeredTexture : Texture {                                        // texture.rs:33-71
	double scale, scale_inv;
	std::shared_ptr<Texture> even_texture, odd_texture;
	CheckeredTexture(double s, std::shared_ptr<Texture> e, std::shared_ptr<Texture> o)
		: scale(s), scale_inv(1.0 / s), even_texture(std::move(e)), odd_texture(std::move(o)) {}
	Vec3 value_at(double u, double v, const Vec3& p) const override;
};
struct ImageTexture : Texture {                                            // texture.rs:73-95
	std::shared_ptr<ImageU8> image;
	explicit ImageTexture(std::shared_ptr<ImageU8> img) : image(std::move(img)) {}
	Vec3 value_at(double u, double v, const Vec3& p) const override;
};
struct NoiseTexture : Texture {                                            // texture.rs:97-131
	double scale;
	Perlin noise;
	explicit NoiseTexture(double s) : scale(s), noise(0) {}
	double turbulence(const Vec3& p, uint32_t depth) const;
	Vec3 value_at(double u, double v, const Vec3& p) const override;
};

// ---------------------------------------------------------------------------------------
// Materials -- material.rs:11-201
// ---------------------------------------------------------------------------------------
struct HitRecord;
struct ScatterRecord { Ray scattered_ray; Vec3 attenuation; double pdf; };  // material.rs:23-27
struct Material {
	int id = -1;
	virtual ~Material() = default;
	virtual std::optional<ScatterRecord> scatter(const Ray&, const HitRecord&) const { return std::nullopt; }  // material.rs:12-16
	virtual Vec3 emitted(double, double, const Vec3&) const { return Vec3(0, 0, 0); }                        // material.rs:18
};
struct Lambertian : Material {                                             // material.rs:29-73
	std::shared_ptr<Texture> texture;
	explicit Lambertian(std::shared_ptr<Texture> t) : texture(std::move(t)) {}
	std::optional<ScatterRecord> scatter(const Ray&, const HitRecord&) const override;
};
struct Metal : Material {                                                  // material.rs:75-103
	Vec3 albedo; double fuzz;
	Metal(const Vec3& a, double f) : albedo(a), fuzz(f) {}
	std::optional<ScatterRecord> scatter(const Ray&, const HitRecord&) const override;
};
struct Dielectric : Material {                                             // material.rs:105-149
	double refraction_index;
	explicit Dielectric(double ri) : refraction_index(ri) {}
	static double reflectance(double cosine, double refraction_index);
	std::optional<ScatterRecord> scatter(const Ray&, const HitRecord&) const override;
};
struct DiffuseLight : Material {                                           // material.rs:151-169
	std::shared_ptr<Texture> texture;
	explicit DiffuseLight(std::shared_ptr<Texture> t) : texture(std::move(t)) {}
	Vec3 emitted(double u, double v, const Vec3& p) const override { return texture->value_at(u, v, p); }
};
struct Isotropic : Material {                                              // material.rs:171-201
	std::shared_ptr<Texture> texture;
	explicit Isotropic(std::shared_ptr<Texture> t) : texture(std::move(t)) {}
	std::optional<ScatterRecord> scatter(const Ray&, const HitRecord&) const override;
};

// ---------------------------------------------------------------------------------------
// Hittables -- hittable/*.rs
// ---------------------------------------------------------------------------------------
// The reference clones an Arc<dyn Material> into every accepted candidate hit (sphere.rs:102, quad.rs:105, volume.rs:59): an
// atomic increment / decrement pair on a counter shared by every thread that touches the same material (one Lambertian on 2400
// quads in the final scene).  The timing build keeps that; ORC_NO_REFCOUNT (liboracle_norc.so) replaces the handle by a plain
// pointer to show how much of the baseline's time is that contention (bench.py's cpu_baseline.no_refcount).
#ifdef ORC_NO_REFCOUNT
struct MaterialRef {
	const Material* p;
	MaterialRef(const std::shared_ptr<Material>& m) : p(m.get()) {}
	const Material* operator->() const { return p; }
};
#else
using MaterialRef = std::shared_ptr<Material>;
#endif
struct HitRecord {                                                         // hittable.rs:15-43
	double t;
	Vec3 position, normal;
	bool is_front_face;
	MaterialRef material;  // Arc clone per accepted candidate, as the reference does
	double u, v;
	int prim_id;                         // oracle bookkeeping: construction-order primitive id (media: -2 - medium index)
	HitRecord(const Ray& ray, double t_, const Vec3& pos, const Vec3& n, MaterialRef m,
	          double u_, double v_, int pid)
		: t(t_), position(pos), material(std::move(m)), u(u_), v(v_), prim_id(pid) {
		is_front_face = ray.direction.dot(n) < 0.0;
		normal = is_front_face ? n : -n;
	}
};

struct Hittable {
	int id = -1;  // graph node id
	virtual ~Hittable() = default;
	virtual std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const = 0;  // hittable.rs:11
	virtual AABB bounding_box() const = 0;                                                  // hittable.rs:12
};

struct HittableList : Hittable {                                           // hittable.rs:45-91
	std::vector<std::shared_ptr<Hittable>> objects;
	AABB bbox;
	void add(std::shared_ptr<Hittable> o) {                                // hittable.rs:63-66
		bbox = AABB::from_pair(bbox, o->bounding_box());
		objects.push_back(std::move(o));
	}
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct BVHNode : Hittable {                                                // BVH.rs:7-94
	std::shared_ptr<Hittable> left, right;  // right may be null (BVH.rs:9)
	AABB bbox;
	std::vector<int> source_children;       // graph bookkeeping: ids of the list the root was built from
	static std::shared_ptr<BVHNode> construct_tree(std::vector<std::shared_ptr<Hittable>> objects);  // BVH.rs:18-65
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;              // BVH.rs:68-90
	AABB bounding_box() const override { return bbox; }
};

struct Sphere : Hittable {                                                 // sphere.rs:11-110
	Vec3 center_start, center_path;
	Vec3 center_end;  // oracle bookkeeping for graph dumps (new_moving's second argument)
	bool is_moving;
	double radius;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	static std::shared_ptr<Sphere> new_stationary(const Vec3& c, double r, std::shared_ptr<Material> m);
	static std::shared_ptr<Sphere> new_moving(const Vec3& c1, const Vec3& c2, double r, std::shared_ptr<Material> m);
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct Quad : Hittable {                                                   // quad.rs:12-113 + plane.rs:9-32
	Vec3 normal; double d;   // Plane
	Vec3 q, u, v, w;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	Quad(const Vec3& q_, const Vec3& u_, const Vec3& v_, std::shared_ptr<Material> m);
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct Triangle : Hittable {                                               // triangle.rs:10-73 (quirk Q14)
	Vec3 normal, a, b, c;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	Triangle(const Vec3& a_, const Vec3& b_, const Vec3& c_, std::shared_ptr<Material> m);
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct Translate : Hittable {                                              // hittable.rs:93-125
	std::shared_ptr<Hittable> object;
	Vec3 offset;
	AABB bbox;
	Translate(std::shared_ptr<Hittable> o, const Vec3& off) : object(std::move(o)), offset(off) {
		bbox = object->bounding_box() + offset;
	}
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct RotateY : Hittable {                                                // hittable.rs:127-216
	std::shared_ptr<Hittable> object;
	double angle_deg, sin_theta, cos_theta;
	AABB bbox;
	RotateY(std::shared_ptr<Hittable> o, double angle);
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return bbox; }
};

struct ConstantMedium : Hittable {                                         // volume.rs:10-67
	std::shared_ptr<Hittable> boundary;
	double density, density_neg_inv;
	std::shared_ptr<Material> phase_function;
	int medium_index = 0;
	ConstantMedium(std::shared_ptr<Hittable> b, double dens, std::shared_ptr<Material> phase)
		: boundary(std::move(b)), density(dens), density_neg_inv(-1.0 / dens), phase_function(std::move(phase)) {}
	std::optional<HitRecord> hit(const Ray& ray, const Interval& ray_t) const override;
	AABB bounding_box() const override { return boundary->bounding_box(); }
};

// ---------------------------------------------------------------------------------------
// Camera / integrator -- camera.rs:17-271
// ---------------------------------------------------------------------------------------
struct SampleSettings { double confidence = 0.95, tolerance = 0.0; uint32_t batch_size = 1, max_samples = 0; };  // camera.rs:239-244
struct HDRI {                                                              // camera.rs:251-271
	std::shared_ptr<ImageF32> image;
	Vec3 rotation;
	Vec3 sample(const Vec3& direction) const;
};
struct Background { bool is_hdri = false; Vec3 solid; HDRI hdri; };        // camera.rs:246-249

struct Camera {
	int image_width = 0, image_height = 0;
	SampleSettings sample_settings;
	uint32_t max_depth = 0;
	Vec3 center, starting_pixel_pos, pixel_delta_u, pixel_delta_v;
	Background background;
	double defocus_angle = 0;
	Vec3 defocus_disk_u, defocus_disk_v;
	// constructor arguments kept for graph dumps
	double arg_aspect = 0, arg_vfov = 0, arg_focus = 0;
	Vec3 arg_from, arg_at, arg_vup;

	Camera() = default;
	Camera(double aspect_ratio, int image_width, SampleSettings s, uint32_t max_depth, double v_fov,
	       const Vec3& look_from, const Vec3& look_at, const Vec3& vup, double defocus_angle,
	       double focus_distance, Background background);               // camera.rs:39-98
	Ray get_ray(int i, int j) const;                                       // camera.rs:204-221
	Vec3 ray_color(const Ray& ray, uint32_t depth, const Hittable& world) const;  // camera.rs:174-202
	Vec3 sample_background(const Ray& ray) const;                          // camera.rs:228-233
	// camera.rs:125-171; returns mean colour, writes the sample count.  sample_base/sample_stride
	// slice the Philox sample index so a rank g of G can take every G-th... (see render()).
	Vec3 sample(int i, int j, const Hittable& world, uint64_t seed, uint32_t sample_begin, double* count_out) const;
};

struct RenderResult {
	std::vector<double> rgb;       // W*H*3 linear mean radiance (what Camera::sample returns)
	std::vector<uint32_t> counts;  // W*H samples taken
	Counters counters;
	double seconds = 0;
};
// camera.rs:100-121 minus the PPM text: parallel loop over pixels (std::thread workers with
// dynamic chunk scheduling standing in for rayon's par_iter), ordered result.
RenderResult render(const Camera& cam, const Hittable& world, uint64_t seed, RngMode mode, int threads,
                    uint32_t sample_begin = 0, int pixel_begin = 0, int pixel_end = -1);
void write_ppm(const std::string& path, int w, int h, const std::vector<double>& rgb);  // camera.rs:101-103,116-118 + color.rs:8-19

// ---------------------------------------------------------------------------------------
// Scene = world + camera + tables (for graph dumps)
// ---------------------------------------------------------------------------------------
struct Scene {
	std::vector<std::shared_ptr<ImageU8>> images_u8;
	std::vector<std::shared_ptr<ImageF32>> images_f32;
	std::vector<std::shared_ptr<Texture>> textures;
	std::vector<std::shared_ptr<Material>> materials;
	std::vector<std::shared_ptr<Hittable>> nodes;
	std::shared_ptr<Hittable> root;
	Camera camera;
	int prim_count = 0, medium_count = 0;
	std::string name;
	// builders (register + assign ids)
	std::shared_ptr<Texture> add(std::shared_ptr<Texture> t) { t->id = (int)textures.size(); textures.push_back(t); return t; }
	std::shared_ptr<Material> add(std::shared_ptr<Material> m) { m->id = (int)materials.size(); materials.push_back(m); return m; }
	template <class T> std::shared_ptr<T> addn(std::shared_ptr<T> n) { n->id = (int)nodes.size(); nodes.push_back(n); return n; }
	std::shared_ptr<Texture> solid(const Vec3& c) { return add(std::make_shared<SolidColorTexture>(c)); }
	std::shared_ptr<Material> lambertian_color(const Vec3& c) { return add(std::make_shared<Lambertian>(solid(c))); }  // material.rs:35-37
	std::shared_ptr<Material> lambertian(std::shared_ptr<Texture> t) { return add(std::make_shared<Lambertian>(std::move(t))); }
	std::shared_ptr<Material> metal(const Vec3& a, double f) { return add(std::make_shared<Metal>(a, f)); }
	std::shared_ptr<Material> dielectric(double ir) { return add(std::make_shared<Dielectric>(ir)); }
	std::shared_ptr<Material> light_color(const Vec3& c) { return add(std::make_shared<DiffuseLight>(solid(c))); }   // material.rs:160-162
	std::shared_ptr<Material> isotropic_color(const Vec3& c) { return add(std::make_shared<Isotropic>(solid(c))); }  // material.rs:180-182
	std::shared_ptr<Sphere> sphere(const Vec3& c, double r, std::shared_ptr<Material> m);
	std::shared_ptr<Sphere> moving_sphere(const Vec3& c1, const Vec3& c2, double r, std::shared_ptr<Material> m);
	std::shared_ptr<Quad> quad(const Vec3& q, const Vec3& u, const Vec3& v, std::shared_ptr<Material> m);
	std::shared_ptr<Triangle> triangle(const Vec3& a, const Vec3& b, const Vec3& c, std::shared_ptr<Material> m);
	std::shared_ptr<HittableList> list() { return addn(std::make_shared<HittableList>()); }
	std::shared_ptr<HittableList> cube(const Vec3& a, const Vec3& b, std::shared_ptr<Material> m);  // quad.rs:54-80
	std::shared_ptr<BVHNode> bvh_from_list(const HittableList& l);                                    // BVH.rs:15-17
	std::shared_ptr<Translate> translate(std::shared_ptr<Hittable> o, const Vec3& off) { return addn(std::make_shared<Translate>(std::move(o), off)); }
	std::shared_ptr<RotateY> rotate_y(std::shared_ptr<Hittable> o, double deg) { return addn(std::make_shared<RotateY>(std::move(o), deg)); }
	std::shared_ptr<ConstantMedium> medium(std::shared_ptr<Hittable> b, double density, std::shared_ptr<Material> phase) {
		auto m = addn(std::make_shared<ConstantMedium>(std::move(b), density, std::move(phase)));
		m->medium_index = medium_count++;
		return m;
	}
};

struct SceneOptions {
	int width = 0;               // 0 = the scene's own default
	int spp = 0;                 // >0: fixed-spp mapping (batch=spp, tol=0, max=0); 0 = the scene's own adaptive settings
	int max_depth = 0;           // 0 = default
	uint64_t scene_seed = 1;
	bool include_media = true;   // false: drop ConstantMedium objects (medium-free closest-hit parity)
	int n_spheres = 0;           // C5 only: sphere count (0 = 1,000,000)
	std::string asset_dir = "assets";
};
// name: "c1".."c5" (BASELINE.json configs) or "scene0".."scene11" (main.rs:42-55)
std::unique_ptr<Scene> build_scene(const std::string& name, const SceneOptions& opt);
std::unique_ptr<Scene> load_graph(const std::string& path, bool include_media = true);
void dump_graph(const Scene& scene, const std::string& path, const std::string& blob_dir);

}  // namespace orc
