// grayshift_b200/host/grayshift.hpp -- host side above the C ABI: a C++ mirror of the reference's Rust
// scene-building API (same type names, constructor names, argument order and meaning), so the
// scenes of src/main.rs translate line by line.  The reference's host language is Rust; no Rust
// toolchain exists in this build environment, so the host is C++ here and the Rust crate shell
// under rust/ mirrors it.
//
// What stays on the host (north_star): scene construction, bounding boxes, the reference's own BVH
// ordering (used only to rank primitives for exact-tie breaking), baking of Translate/RotateY
// instances into world-space primitives, the SAH BVH build, and flattening to the arrays of
// include/grayshift.h.  Nothing here intersects a ray or shades: Camera::render hands the
// flattened scene to libgrayshift_cuda.so.
//
// `new` is a C++ keyword, so Rust's `T::new(..)` is `T::new_(..)` here; everything else keeps its name.
#pragma once
#include <cstdint>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "../../include/grayshift.h"

namespace grayshift {

// ---- util/vec3.rs:6-178 (only what scene construction needs) ------------------------------------
struct Vec3 {
	double x = 0, y = 0, z = 0;
	static const Vec3 ZERO;
	Vec3() = default;
	Vec3(double x_, double y_, double z_) : x(x_), y(y_), z(z_) {}
	static Vec3 new_(double x, double y, double z) { return Vec3(x, y, z); }
	double length_squared() const { return x * x + y * y + z * z; }
	double length() const;
	Vec3 unit() const;
	double dot(const Vec3& o) const { return x * o.x + y * o.y + z * o.z; }
	Vec3 cross(const Vec3& o) const { return Vec3(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
};
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(double s, const Vec3& a) { return a * s; }
inline Vec3 operator*(const Vec3& a, const Vec3& b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }

// ---- util/interval.rs, AABB.rs (bounding boxes only; AABB::hit lives on the device) ---------------
struct Interval {
	double min, max;
	Interval();  // EMPTY
	Interval(double mn, double mx) : min(mn), max(mx) {}
	static Interval from_interval_pair(const Interval& a, const Interval& b);
	double size() const { return max - min; }
	Interval expand(double delta) const;
};
struct AABB {
	Interval x, y, z;  // default EMPTY
	AABB() = default;
	AABB(const Interval& x_, const Interval& y_, const Interval& z_) : x(x_), y(y_), z(z_) {}
	static AABB from_corners(const Vec3& a, const Vec3& b);   // AABB.rs:24-48 (pads thin axes)
	static AABB from_AABB_pair(const AABB& a, const AABB& b);  // AABB.rs:50-56
	const Interval& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	int longest_axis() const;                                  // AABB.rs:115-121
	AABB operator+(const Vec3& o) const;
};

// ---- util.rs:5-16 with a seedable generator -----------------------------------------
// The reference draws scene parameters from fastrand's unseeded thread-local RNG; here the same
// calls draw from a SplitMix64 stream so a scene can be rebuilt identically from its seed.
void seed_scene_rng(uint64_t seed);
double random_f64();                         // fastrand::f64()
double random_f64(double min, double max);   // util.rs:5-7
Vec3 random_vector(double min, double max);  // util.rs:9-16
// Primitive ids are handed out in construction order; reset at the start of a scene.
void reset_scene_ids();

// ---- decoded images (the `image` and `radiant` crates decode on the host in the reference) -------
struct ImageData {
	int width = 0, height = 0;
	int format = GS_IMG_U8_RGB;
	std::vector<uint8_t> u8;
	std::vector<float> f32;
	std::string source;  // "ppm:<path>" / "hdr:<path>" / "raw:<path>" for graph dumps
};
std::shared_ptr<ImageData> load_image_u8(const std::string& path);   // image::open (texture.rs:79): P6 PPM, or a .jpg with a decoded .ppm sibling
std::shared_ptr<ImageData> load_hdr(const std::string& path);        // radiant::load (main.rs:139)
std::shared_ptr<ImageData> load_raw(const std::string& path, int w, int h, int format);

// ---- texture.rs ---------------------------------------------------------------------------------------
struct Texture {
	int kind = GS_TEX_SOLID;
	Vec3 albedo;
	double scale = 0;
	std::shared_ptr<Texture> even_texture, odd_texture;
	std::shared_ptr<ImageData> image;
	virtual ~Texture() = default;
};
struct SolidColorTexture : Texture {
	static std::shared_ptr<Texture> new_(const Vec3& albedo);                   // texture.rs:18-20
	static std::shared_ptr<Texture> from_rgb(double r, double g, double b);     // texture.rs:22-24
};
struct CheckeredTexture : Texture {
	static std::shared_ptr<Texture> new_(double scale, std::shared_ptr<Texture> even, std::shared_ptr<Texture> odd);  // texture.rs:40-46
	static std::shared_ptr<Texture> from_colors(double scale, const Vec3& even, const Vec3& odd);                      // texture.rs:48-54
};
struct ImageTexture : Texture {
	static std::shared_ptr<Texture> new_(const std::string& filepath);          // texture.rs:78-80 (throws on error, like the Err)
	static std::shared_ptr<Texture> from_image(std::shared_ptr<ImageData> img);
};
struct NoiseTexture : Texture {
	static std::shared_ptr<Texture> new_(double scale);                         // texture.rs:103-105
};
// PermutationTable of noise::Perlin::default() (seed 0), handed to the device through the ABI.
void perlin_permutation_table(uint32_t seed, uint8_t out[256]);

// ---- material.rs ----------------------------------------------------------------------------------------
struct Material {
	int kind = GS_MAT_LAMBERTIAN;
	std::shared_ptr<Texture> texture;
	Vec3 albedo;
	double param = 0;  // fuzz or refraction index
	virtual ~Material() = default;
};
struct Lambertian : Material {
	static std::shared_ptr<Material> from_color(const Vec3& albedo);             // material.rs:35-37
	static std::shared_ptr<Material> from_texture(std::shared_ptr<Texture> t);   // material.rs:39-41
};
struct Metal : Material { static std::shared_ptr<Material> new_(const Vec3& albedo, double fuzz); };  // material.rs:81-83
struct Dielectric : Material { static std::shared_ptr<Material> new_(double refraction_index); };     // material.rs:110-112
struct DiffuseLight : Material {
	static std::shared_ptr<Material> new_(std::shared_ptr<Texture> t);           // material.rs:156-158
	static std::shared_ptr<Material> from_color(const Vec3& color);              // material.rs:160-162
};
struct Isotropic : Material {
	static std::shared_ptr<Material> new_(std::shared_ptr<Texture> t);           // material.rs:176-178
	static std::shared_ptr<Material> from_color(const Vec3& color);              // material.rs:180-182
};

// ---- hittable/*.rs ----------------------------------------------------------------------------------------
enum class NodeKind { Sphere, Quad, Triangle, List, Bvh, Translate, RotateY, Medium };
struct Hittable {
	NodeKind node_kind;
	explicit Hittable(NodeKind k) : node_kind(k) {}
	virtual ~Hittable() = default;
	virtual AABB bounding_box() const = 0;  // hittable.rs:12
};
using HittablePtr = std::shared_ptr<Hittable>;

struct Sphere : Hittable {  // sphere.rs:11-49
	Vec3 center_start, center_path, center_end;
	bool is_moving = false;
	double radius = 0;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	Sphere() : Hittable(NodeKind::Sphere) {}
	static std::shared_ptr<Sphere> new_stationary(const Vec3& center, double radius, std::shared_ptr<Material> m);
	static std::shared_ptr<Sphere> new_moving(const Vec3& c1, const Vec3& c2, double radius, std::shared_ptr<Material> m);
	AABB bounding_box() const override { return bbox; }
};

struct HittableList : Hittable {  // hittable.rs:45-67
	std::vector<HittablePtr> objects;
	AABB bbox;
	HittableList() : Hittable(NodeKind::List) {}
	static std::shared_ptr<HittableList> new_() { return std::make_shared<HittableList>(); }
	void add(HittablePtr object);
	void clear() { objects.clear(); bbox = AABB(); }
	AABB bounding_box() const override { return bbox; }
};

struct Quad : Hittable {  // quad.rs:12-80
	Vec3 q, u, v;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	Quad() : Hittable(NodeKind::Quad) {}
	static std::shared_ptr<Quad> new_(const Vec3& q, const Vec3& u, const Vec3& v, std::shared_ptr<Material> m);
	static std::shared_ptr<HittableList> cube(const Vec3& point_a, const Vec3& point_b, std::shared_ptr<Material> m);
	AABB bounding_box() const override { return bbox; }
};

struct Triangle : Hittable {  // triangle.rs:10-28
	Vec3 a, b, c;
	std::shared_ptr<Material> material;
	AABB bbox;
	int prim_id = -1;
	Triangle() : Hittable(NodeKind::Triangle) {}
	static std::shared_ptr<Triangle> new_(const Vec3& a, const Vec3& b, const Vec3& c, std::shared_ptr<Material> m);
	AABB bounding_box() const override { return bbox; }
};

// BVHNode::from_list builds the reference's own tree (median split on the longest axis, stable sort,
// leaves of 1-2; BVH.rs:18-65).  The device never traverses it: it only fixes the order in which the
// reference would visit primitives, which decides exact-t ties (gs_quad.order).
struct BVHNode : Hittable {
	HittablePtr left, right;  // right may be null (BVH.rs:9)
	AABB bbox;
	std::vector<HittablePtr> source;  // the list this root was built from, original order (graph dumps)
	BVHNode() : Hittable(NodeKind::Bvh) {}
	static std::shared_ptr<BVHNode> from_list(const HittableList& list);
	static std::shared_ptr<BVHNode> from_list(const std::shared_ptr<HittableList>& list) { return from_list(*list); }
	static std::shared_ptr<BVHNode> construct_tree(std::vector<HittablePtr> objects);
	AABB bounding_box() const override { return bbox; }
};

struct Translate : Hittable {  // hittable.rs:93-125
	HittablePtr object;
	Vec3 offset;
	AABB bbox;
	Translate() : Hittable(NodeKind::Translate) {}
	static std::shared_ptr<Translate> new_(HittablePtr object, const Vec3& offset);
	AABB bounding_box() const override { return bbox; }
};

struct RotateY : Hittable {  // hittable.rs:127-176
	HittablePtr object;
	double angle = 0, sin_theta = 0, cos_theta = 1;
	AABB bbox;
	RotateY() : Hittable(NodeKind::RotateY) {}
	static std::shared_ptr<RotateY> new_(HittablePtr object, double angle_degrees);
	AABB bounding_box() const override { return bbox; }
};

struct ConstantMedium : Hittable {  // volume.rs:10-28
	HittablePtr boundary;
	double density = 0;
	std::shared_ptr<Material> phase_function;
	int medium_index = 0;
	ConstantMedium() : Hittable(NodeKind::Medium) {}
	static std::shared_ptr<ConstantMedium> new_(HittablePtr boundary, double density, std::shared_ptr<Material> phase_function);
	static std::shared_ptr<ConstantMedium> from_isotropic_color(HittablePtr boundary, double density, const Vec3& color);
	AABB bounding_box() const override { return boundary->bounding_box(); }
};

// ---- camera.rs ----------------------------------------------------------------------------------------------
struct SampleSettings {  // camera.rs:239-244
	double confidence = 0.95, tolerance = 0.0;
	uint32_t batch_size = 1, max_samples = 0;
};
struct HDRI {  // camera.rs:251-254
	std::shared_ptr<ImageData> image;
	Vec3 rotation;
};
struct Background {  // camera.rs:246-249
	bool is_hdri = false;
	Vec3 color;
	HDRI hdri;
	static Background SOLID(const Vec3& c) { Background b; b.color = c; return b; }
	static Background HDRI_(const HDRI& h) { Background b; b.is_hdri = true; b.hdri = h; return b; }
};

// The flattened scene: owns every array a gs_scene_desc points at.
struct FlatScene {
	std::vector<gs_sphere> spheres, boundary_spheres;
	std::vector<gs_quad> quads, boundary_quads;
	std::vector<gs_triangle> triangles;
	std::vector<gs_bvh_node> nodes;
	std::vector<uint32_t> prim_refs;
	std::vector<gs_medium> media;
	std::vector<gs_material> materials;
	std::vector<gs_texture> textures;
	std::vector<gs_image> images;
	std::vector<std::shared_ptr<ImageData>> image_data;
	uint8_t perlin_perm[256];
	bool has_noise = false;
	gs_scene_desc desc{};
	int image_index(const std::shared_ptr<ImageData>& img);  // registers on first use
	void finish();                                           // fills `desc`
};

struct FlattenOptions {
	int max_leaf_size = 1;  // measured on B200: one primitive per leaf is +8 % on C4, +4 % on C3 (the child box test doubles as a per-primitive fp32 pre-test)
	bool reference_tree = false;  // emit the reference's own median-split tree instead of the SAH tree (A/B measurements)
};
// flatten(): graph -> world-space primitives (instances baked) + reference visiting ranks + SAH BVH.
std::unique_ptr<FlatScene> flatten(const HittablePtr& world, const FlattenOptions& opt = FlattenOptions());
// SAH (binned) BVH over boxes; bounds = n x (min xyz, max xyz).  Nodes are emitted in pre-order (children
// have larger indices than their parent), leaves reference `order_out` positions.
void build_bvh_sah(const std::vector<double>& bounds, int max_leaf_size, std::vector<gs_bvh_node>& nodes_out, std::vector<uint32_t>& order_out);

class Camera {
public:
	// camera.rs:39-51, same argument order
	static Camera new_(double aspect_ratio, int image_width, SampleSettings sample_settings, uint32_t max_depth, double v_fov,
	                   const Vec3& look_from, const Vec3& look_at, const Vec3& vup, double defocus_angle, double focus_distance,
	                   Background background);
	// camera.rs:100: renders `world` and writes the PPM (P3) text to image_file.  Throws std::runtime_error
	// where the reference returns Err.  device: CUDA ordinal.
	void render(const HittablePtr& world, std::FILE* image_file, int device = 0, uint64_t seed = 1, gs_stats* stats = nullptr) const;
	// the same without the file: linear mean radiance W*H*3 and per-pixel sample counts
	void render_to_buffer(const HittablePtr& world, std::vector<float>& rgb, std::vector<uint32_t>* counts, int device = 0,
	                      uint64_t seed = 1, gs_stats* stats = nullptr) const;
	// the ABI view of this camera; registers the HDR image in `flat` when there is one
	gs_camera to_abi(FlatScene& flat) const;
	gs_sample_settings settings_abi() const;

	int image_width = 0, image_height = 0;
	SampleSettings sample_settings;
	uint32_t max_depth = 0;
	Vec3 center, starting_pixel_pos, pixel_delta_u, pixel_delta_v;
	Background background;
	double defocus_angle = 0;
	Vec3 defocus_disk_u, defocus_disk_v;
	// constructor arguments, kept for graph dumps
	double arg_aspect = 0, arg_vfov = 0, arg_focus = 0;
	Vec3 arg_from, arg_at, arg_vup;
};

void write_color(std::FILE* image, const Vec3& color);  // color.rs:8-19
double luminance(const Vec3& v);                        // color.rs:31-33

// ---- scenes (main.rs:61-888) and the BASELINE.json configs ------------------------------------------------
struct HostScene {
	HittablePtr world;  // what the scene hands to camera.render (a BVHNode root)
	Camera camera;
	std::string name;
};
struct SceneOptions {
	int width = 0, spp = 0, max_depth = 0;
	uint64_t scene_seed = 1;
	bool include_media = true;
	int n_spheres = 0;
	std::string asset_dir = "assets";
};
HostScene build_scene(const std::string& name, const SceneOptions& opt);  // "c1".."c5", "scene0".."scene11"
HostScene load_graph(const std::string& path, bool include_media = true);
void dump_graph(const HostScene& scene, const std::string& path);

}  // namespace grayshift
