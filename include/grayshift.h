/* include/grayshift.h -- C ABI of libgrayshift_cuda.so, the B200 (sm_100a) implementation of
 * grayshift's per-pixel path-tracing hot path.
 *
 * The reference (benjisu03/grayshift, a pure-Rust CPU renderer) has no FFI boundary of its own
 * (SURVEY.md F1/F3): its de-facto interface for this path is
 *     Camera::new(...)                                   src/camera.rs:39-51
 *     Camera::render(&self, world: Box<dyn Hittable>, image_file: &mut File)   src/camera.rs:100
 *     SampleSettings / Background / HDRI                  src/camera.rs:239-254
 * plus the scene constructors in the src/hittable directory, src/material.rs and src/texture.rs.
 * This header is what a `src/ffi.rs` added underneath that preserved Rust API binds
 * (INTEGRATION.md shows the binding): the host (Rust in the reference; the C++ mirror under
 * grayshift_b200/host/ here) builds the scene, bakes instance transforms, builds the BVH and
 * flattens everything into the plain arrays below; the library copies them to the GPU and runs
 * ray generation, BVH traversal, primitive/medium intersection, material scattering, texture /
 * environment lookups and accumulation as hand-written CUDA.
 *
 * Conventions: plain C, plain pointers and sizes, no ownership transfer (the library copies
 * every input during the call), every function returns 0 on success and a non-zero gs_status
 * otherwise with a thread-local message in gs_last_error().  Nothing aborts or throws across
 * this boundary.  All geometry is f64, world space, as in the reference (util/vec3.rs:6-10).
 * There is NO CPU fallback: without a CUDA device every compute entry returns GS_ERR_NO_DEVICE.
 *
 * Concurrency: handles are independent of each other, and any number of threads may work on DIFFERENT
 * gs_scene handles at once (the worker pool, the staging buffers and the error string are shared
 * safely).  One handle is not re-entrant: calls on it must come from one thread at a time, like
 * `&mut self`.  Asynchronous renders of one handle may overlap on different CUDA streams; each launch
 * takes the next of 256 work counters of the handle, so at most 256 of them may be in flight at once.
 */
#ifndef GRAYSHIFT_H
#define GRAYSHIFT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GS_ABI_VERSION 2

typedef enum gs_status {
	GS_OK = 0,
	GS_ERR_INVALID = 1,    /* malformed description (bad index, empty scene, NaN bounds ...) */
	GS_ERR_NO_DEVICE = 2,  /* no usable CUDA device */
	GS_ERR_CUDA = 3,       /* CUDA runtime error; message carries the CUDA error string */
	GS_ERR_NCCL = 4,       /* gs_render_multi: a peer-to-peer / NCCL step between the devices failed */
	GS_ERR_UNSUPPORTED = 5
} gs_status;

/* ---- primitives (replace Sphere / Quad trait objects) ------------------------------------- */

/* Sphere::new_stationary / new_moving, src/hittable/sphere.rs:21-49.  center(time) = center + time*path. */
typedef struct gs_sphere {
	double center[3];
	double radius;     /* SIGNED, as handed to the constructor: outward_normal = (p - center) / radius (sphere.rs:91), so a
	                    * negative radius turns the normal inward (the hollow-glass trick).  Zero is rejected. */
	double path[3];    /* center_path (sphere.rs:44); zero when stationary */
	int32_t material;  /* index into gs_scene_desc.materials */
	int32_t prim_id;   /* caller's id, returned by gs_trace_closest */
	int32_t moving;    /* is_moving (sphere.rs:14) */
	int32_t order;     /* rank of this primitive in the reference's visiting order (see gs_quad.order) */
	/* (cos, sin) of the RotateY angle this sphere was baked with, (1, 0) when it is not under a RotateY.  Sphere::hit
	 * computes (u, v) from the OBJECT-space normal before RotateY::hit turns the normal into world space
	 * (sphere.rs:95, hittable.rs:179-211), so image-textured spheres need the rotation undone for sphere_uv. */
	double uv_rotation[2];
} gs_sphere;

/* Quad::new, src/hittable/quad.rs:25-38.  The library derives normal, w and d exactly as
 * quad.rs:32-36 and plane.rs:15-18 do.  Quad::cube (quad.rs:54-80) is six of these. */
typedef struct gs_quad {
	double q[3], u[3], v[3];
	int32_t material;
	int32_t prim_id;
	/* Rank of this primitive in the order the reference would visit primitives (HittableList order,
	 * hittable.rs:75; BVHNode left-then-right, BVH.rs:71-79).  Only used to break EXACT ties in t the
	 * way the reference does (quads accept t <= closest, so the later one wins; interval.rs:29-31).
	 * Coplanar overlapping quads are common (adjacent Quad::cube faces, main.rs:637-656). */
	int32_t order;
	int32_t reserved;
} gs_quad;

/* Triangle::new, src/hittable/triangle.rs:20-28.  Moeller-Trumbore with the reference's conventions: edge_1 = c - a,
 * edge_2 = b - a, back faces culled (det < 1e-8), u = barycentric of c, v of b, normal = (b - a) x (c - a) NOT
 * normalised (triangle.rs:21, 55-63).  ONE DELIBERATE DIFFERENCE: the reference's Triangle::hit never
 * looks at ray_t, so a triangle visited later replaces a closer hit and even hits behind the origin count -- the
 * result then depends on the shape of the reference's BVH.  This library applies the interval (t_min < t < closest)
 * like every other primitive, which is order-independent; the two agree whenever no ray meets a triangle outside
 * its interval (true for main.rs's `triangles` scene, tests/test_oracle_kat.py). */
typedef struct gs_triangle {
	double a[3], b[3], c[3];
	int32_t material;
	int32_t prim_id;
	int32_t order;
	int32_t reserved;
} gs_triangle;

/* ---- acceleration structure (replaces BVHNode, src/hittable/BVH.rs) ------------------------
 * A binary BVH over the surface primitives, built by the host.  The closest hit does not depend
 * on the tree's shape (every box is conservative), so the host may build any valid tree; the
 * library re-lays it out for the GPU.  Node 0 is the root.  A node with ref_count > 0 is a leaf
 * over prim_refs[first_ref .. first_ref + ref_count). */
typedef struct gs_bvh_node {
	double bounds_min[3];
	double bounds_max[3];
	int32_t left, right;       /* child node indices (internal nodes) */
	int32_t first_ref, ref_count;
} gs_bvh_node;

#define GS_PRIM_SPHERE 0u
#define GS_PRIM_QUAD 1u
#define GS_PRIM_TRIANGLE 2u
#define GS_PRIM_REF(kind, index) (((uint32_t)(kind) << 30) | (uint32_t)(index))

/* ---- participating media (replace ConstantMedium, src/hittable/volume.rs:10-67) -------------
 * The boundary is a convex set of primitives taken from boundary_spheres / boundary_quads
 * (never rendered as surfaces). */
typedef struct gs_medium {
	int32_t sphere_first, sphere_count;  /* range in gs_scene_desc.boundary_spheres */
	int32_t quad_first, quad_count;      /* range in gs_scene_desc.boundary_quads */
	double density;                      /* ConstantMedium::new's `density` (volume.rs:17-21) */
	int32_t phase_material;              /* any material (main.rs:720,733 use a Lambertian) */
	int32_t medium_index;                /* gs_trace_closest never reports media */
} gs_medium;

/* ---- materials (replace src/material.rs) ------------------------------------------------------ */
typedef enum gs_material_kind {
	GS_MAT_LAMBERTIAN = 0,    /* material.rs:29-73; texture */
	GS_MAT_METAL = 1,         /* material.rs:75-103; albedo, param = fuzz */
	GS_MAT_DIELECTRIC = 2,    /* material.rs:105-149; param = refraction index */
	GS_MAT_DIFFUSE_LIGHT = 3, /* material.rs:151-169; texture, two-sided */
	GS_MAT_ISOTROPIC = 4      /* material.rs:171-201; texture */
} gs_material_kind;

typedef struct gs_material {
	int32_t kind;
	int32_t texture;   /* index into textures, or -1 */
	double albedo[3];
	double param;
} gs_material;

/* ---- textures (replace src/texture.rs) ---------------------------------------------------------- */
typedef enum gs_texture_kind {
	GS_TEX_SOLID = 0,    /* texture.rs:13-31; color */
	GS_TEX_CHECKER = 1,  /* texture.rs:33-71; scale (the constructor's `scale`, not its inverse), a = even, b = odd texture */
	GS_TEX_IMAGE = 2,    /* texture.rs:73-95; a = image index (GS_IMG_U8_RGB) */
	GS_TEX_NOISE = 3     /* texture.rs:97-131; scale; uses gs_scene_desc.perlin_perm */
} gs_texture_kind;

typedef struct gs_texture {
	int32_t kind;
	int32_t a, b;
	int32_t reserved;
	double color[3];
	double scale;
} gs_texture;

typedef enum gs_image_format {
	GS_IMG_U8_RGB = 0,   /* image::DynamicImage 8-bit RGB, row 0 = top (texture.rs:89-91) */
	GS_IMG_F32_RGB = 1   /* radiant::Image, RGB f32, row 0 = top (camera.rs:265-269) */
} gs_image_format;

/* Pixels are decoded by the host (the `image` and `radiant` crates in the reference). */
typedef struct gs_image {
	int32_t width, height;
	int32_t format;
	int32_t reserved;
	const void* pixels;  /* width*height*3 interleaved u8 or f32 */
} gs_image;

/* ---- the flattened scene ---------------------------------------------------------------------- */
typedef struct gs_scene_desc {
	uint32_t abi_version;  /* GS_ABI_VERSION */
	uint32_t reserved;
	const gs_sphere* spheres;           size_t n_spheres;
	const gs_quad* quads;               size_t n_quads;
	const gs_bvh_node* nodes;           size_t n_nodes;
	const uint32_t* prim_refs;          size_t n_prim_refs;
	const gs_sphere* boundary_spheres;  size_t n_boundary_spheres;
	const gs_quad* boundary_quads;      size_t n_boundary_quads;
	const gs_medium* media;             size_t n_media;
	const gs_material* materials;       size_t n_materials;
	const gs_texture* textures;         size_t n_textures;
	const gs_image* images;             size_t n_images;
	const uint8_t* perlin_perm;         /* 256-entry permutation table of noise::Perlin::default() (texture.rs:104); may be NULL without noise textures */
	const gs_triangle* triangles;       size_t n_triangles;
} gs_scene_desc;

/* ---- camera (the private fields Camera::new computes, src/camera.rs:17-36, 80-97) ------------- */
typedef enum gs_background_kind {
	GS_BG_SOLID = 0,  /* Background::SOLID, camera.rs:247 */
	GS_BG_HDRI = 1    /* Background::HDRI,  camera.rs:248 */
} gs_background_kind;

typedef struct gs_camera {
	int32_t image_width, image_height;
	uint32_t max_depth;
	int32_t background_kind;
	double center[3];
	double starting_pixel_pos[3];
	double pixel_delta_u[3];
	double pixel_delta_v[3];
	double defocus_angle;
	double defocus_disk_u[3];
	double defocus_disk_v[3];
	double background_color[3];   /* SOLID */
	int32_t hdri_image;           /* HDRI: index into images (GS_IMG_F32_RGB) */
	int32_t reserved;
	double hdri_rotation[3];      /* radians, util.rs:67-70 */
} gs_camera;

/* SampleSettings, src/camera.rs:239-244 (public fields, same meaning).  max_samples < batch_size
 * gives exactly batch_size samples per pixel (the loop of camera.rs:135-165 always leaves after its
 * first batch); the BASELINE.json configs use {batch_size: spp, tolerance: 0, max_samples: 0}. */
typedef struct gs_sample_settings {
	double confidence;
	double tolerance;
	uint32_t batch_size;
	uint32_t max_samples;
} gs_sample_settings;

typedef struct gs_stats {
	uint64_t paths;         /* samples taken (sum of per-pixel sample counts) */
	uint64_t rays;          /* world.hit calls from ray_color, camera.rs:177 */
	uint64_t tex_fetches;   /* image-texture and HDR texel reads; fixed-count renders count them only with "count_nodes" set */
	uint64_t node_visits;   /* BVH nodes fetched; only counted after gs_scene_set_option(scene, "count_nodes", 1) */
	uint64_t kernel_launches;
	double kernel_ms;       /* device time of the render kernels (CUDA events) */
	double total_ms;        /* wall time of the call */
	uint64_t h2d_bytes, d2h_bytes;
	double exchange_ms;     /* gs_render_multi: device time of the cross-GPU sum + resolve on the first device; 0 otherwise */
} gs_stats;

typedef struct gs_scene gs_scene;

/* ---- entry points ------------------------------------------------------------------------------- */

const char* gs_last_error(void);
int gs_abi_version(void);
int gs_device_count(void);  /* number of CUDA devices; 0 when none (never an error) */

/* Validates the description and copies it to `device` (a CUDA ordinal).  Replaces handing
 * `world: Box<dyn Hittable>` to Camera::render (camera.rs:100). */
int gs_scene_create(const gs_scene_desc* desc, int device, gs_scene** out);
void gs_scene_destroy(gs_scene* scene);

/* Camera::render (camera.rs:100-121) + Camera::sample (camera.rs:125-171) without the PPM text:
 * out_rgb = W*H*3 linear mean radiance per pixel (what `sample` returns, before write_color),
 * out_counts (optional) = W*H samples taken.  Blocking.  `seed` keys the Philox4x32-10 streams
 * (counter = pixel, sample, block). */
int gs_render(gs_scene* scene, const gs_camera* camera, const gs_sample_settings* settings, uint64_t seed,
              float* out_rgb, uint32_t* out_counts, gs_stats* stats);

/* Camera::render (camera.rs:100) on SEVERAL GPUs from one host thread -- no process group, no Python in the loop.  scenes[g] is
 * the same description created on device g (gs_scene_create once per device).  Fixed sample counts are split by sample index
 * (device g of G takes [g*spp/G, (g+1)*spp/G) of every pixel: the Philox counter carries the global sample index, so the image does
 * not depend on G); the adaptive sampler splits pixels (device g takes pixels g, g+G, ...: a pixel's batch loop stays on one device).
 * The device of scenes[0] then sums the W*H float4 accumulators WHILE it resolves them: its resolve kernel reads the peers' buffers
 * through peer-mapped pointers over NVLink / NVSwitch (cross-GPU sum + /count in one pass; where two devices cannot map each other
 * the buffer is first copied with cudaMemcpyPeer).  With the environment variable GS_MULTI_REDUCE=nccl the sum is an ncclReduce over
 * a communicator the library creates with ncclCommInitAll (libnccl is loaded on demand).  GS_ERR_NCCL reports a failure of that
 * exchange step, whichever transport it used.  stats (optional): totals over the devices, kernel_ms = the slowest device.  Blocking. */
int gs_render_multi(gs_scene* const* scenes, int n_scenes, const gs_camera* camera, const gs_sample_settings* settings,
                    uint64_t seed, float* out_rgb, uint32_t* out_counts, gs_stats* stats);

/* Multi-GPU building block (one process per GPU, samples split across ranks): adds samples
 * [sample_begin, sample_begin + sample_count) of every pixel into a caller-owned DEVICE buffer of
 * W*H float4 (sum r, sum g, sum b, count) on `cuda_stream` (a cudaStream_t, NULL = default stream).
 * Asynchronous; the caller sums the buffers of all ranks (NCCL) and then calls gs_resolve. */
int gs_render_accumulate(gs_scene* scene, const gs_camera* camera, uint64_t seed, uint32_t sample_begin,
                         uint32_t sample_count, void* d_accum, void* cuda_stream);
/* The adaptive sampler of Camera::sample (camera.rs:125-171) with the reference's batch / stop-test
 * semantics, for pixels pixel_begin, pixel_begin + pixel_stride, ... (ranks interleave pixels; every
 * pixel's whole batch loop runs on one GPU).  Leaves (sum r, sum g, sum b, count) in d_accum[pixel]; the entries
 * of these pixels must be ZERO on entry (the count doubles as the sampler's running sample_count).  Runs one launch per
 * batch over the pixels still sampling plus the stop test, and reads back a 4-byte "any left?" after batches 1, 2, 4, ...:
 * it synchronises `cuda_stream` a few times and has finished issuing work when it returns.  (GS_ADAPTIVE_ROUNDS=kernel in the
 * environment selects the fully asynchronous variant -- every round inside one cooperative launch -- which measured slower.) */
int gs_render_adaptive_accumulate(gs_scene* scene, const gs_camera* camera, const gs_sample_settings* settings,
                                  uint64_t seed, uint32_t pixel_begin, uint32_t pixel_stride, void* d_accum,
                                  void* cuda_stream);
/* Options: "count_nodes" (0/1) switches to the kernel variant that counts BVH node fetches and texel reads;
 * "wavefront" (0/1) selects the wavefront pipeline for fixed-count renders; "pool_capacity" its path pool size;
 * "wavefront_sort" (0/1, default 0) its material-sorted shading queues;
 * "tail_ahead" (0/1, default 1): when the adaptive sampler is down to a few hundred pixels, render several of a pixel's
 * batches at once and keep those up to the one whose stop test fires (same samples and counts as the sequential loop);
 * "adaptive_chains" (0..4, default 0 = by image size: 2 from 131072 pixels): the adaptive sampler deals the pixels out to this many
 * independent chains of rounds on streams of their own, so that one chain's launch fills the SMs another's leaves idle as it ends;
 * "bvh_width" (2 or 8) selects the tree the kernels traverse -- the binary tree or the eight-wide compressed tree
 * the library derives from it (default: by scene size; the other tree is built on first request). */
int gs_scene_set_option(gs_scene* scene, const char* key, int64_t value);
/* Facts about the scene as laid out on the device: "bvh_width" (2 / 8, the tree in use), "nodes2" / "nodes8" (node counts of the
 * binary tree, 0 until it has been built, and of the eight-wide topology), "wide_depth", "scene_bytes" (device bytes held). */
int gs_scene_get_info(gs_scene* scene, const char* key, int64_t* value);
/* d_accum (device, W*H float4) -> host mean radiance and counts; blocking on `cuda_stream`. */
int gs_resolve(gs_scene* scene, const void* d_accum, int32_t width, int32_t height, float* out_rgb,
               uint32_t* out_counts, void* cuda_stream);
/* The output stage on the device: pixel_color /= sample_count (camera.rs:167) fused with write_color's
 * transfer (color.rs:8-26: sqrt gamma, clamp to [0, 0.999], *256 truncated) -> W*H*3 bytes on the host,
 * i.e. the numbers the reference prints into image.ppm.  A quarter of gs_resolve's device-to-host traffic. */
int gs_resolve_rgb8(gs_scene* scene, const void* d_accum, int32_t width, int32_t height, uint8_t* out_rgb8,
                    uint32_t* out_counts, void* cuda_stream);
/* Device-side totals accumulated since the last call (synchronises the device). */
int gs_stats_read(gs_scene* scene, gs_stats* stats);

/* write_color (src/color.rs:8-26) for a whole frame: sqrt gamma, clamp to [0, 0.999], *256 -> u8. */
int gs_tonemap(const float* rgb, size_t n_pixels, uint8_t* out_rgb8);

/* Parity hook: closest hit of Hittable::hit (hittable.rs:11) over the surface primitives
 * (media excluded) for n rays of 7 doubles each (origin, direction, time), interval
 * (tmin, tmax) with the reference's per-primitive rules (spheres strict, quads inclusive;
 * interval.rs:29-35).  prim_id[k] = -1 on a miss.  f64 primitive arithmetic, compiled without
 * FMA contraction. */
int gs_trace_closest(gs_scene* scene, const double* rays, size_t n, double tmin, double tmax,
                     int32_t* prim_id, double* t);
/* The same query answered by the RENDER kernels' build of the traversal (FMA contraction, approximate-plus-Newton f64
 * division and square root): shows that the timed code picks the primitives the contraction-free build picks. */
int gs_trace_closest_fast(gs_scene* scene, const double* rays, size_t n, double tmin, double tmax,
                          int32_t* prim_id, double* t);

/* Parity hooks for the stochastic parts (one Philox path (pixel, sample) each):
 * camera rays of get_ray (camera.rs:204-221) for pixels [pixel_begin, pixel_end), 7 doubles per ray. */
int gs_camera_rays(gs_scene* scene, const gs_camera* camera, uint64_t seed, uint32_t sample,
                   int32_t pixel_begin, int32_t pixel_end, double* rays_out);
/* radiance of individual paths, 3 floats each (ray_color, camera.rs:174-202). */
int gs_path_radiance(gs_scene* scene, const gs_camera* camera, uint64_t seed, const uint32_t* pixel,
                     const uint32_t* sample, size_t n, float* out_rgb);

#ifdef __cplusplus
}
#endif
#endif /* GRAYSHIFT_H */
