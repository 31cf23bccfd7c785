/* include/grayshift_host.h -- C entry points of libgrayshift_host.so, the host layer above the C ABI of
 * grayshift.h (scene construction, instance baking, BVH build, flattening).  In the reference this
 * layer is Rust (src/main.rs scene functions + the constructors of src/hittable, src/material.rs,
 * src/texture.rs, src/camera.rs:39-98); here it is the C++ mirror under grayshift_b200/host/.  These
 * functions exist so non-C++ callers (the Python tests and bench.py through ctypes, or a Rust shell
 * that prefers not to re-implement the builder) can drive it.  No ray is traced here.
 */
#ifndef GRAYSHIFT_HOST_H
#define GRAYSHIFT_HOST_H

#include "grayshift.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gsh_scene gsh_scene;

const char* gsh_last_error(void);

/* One of the reference's scenes ("scene0".."scene11", main.rs:42-55) or a BASELINE.json config
 * ("c1".."c5").  width/spp/max_depth = 0 keep the scene's own values; spp > 0 selects the fixed-count
 * mapping SampleSettings{batch_size: spp, tolerance: 0, max_samples: 0}. */
gsh_scene* gsh_scene_build(const char* name, int width, int spp, int max_depth, uint64_t scene_seed,
                           int include_media, int n_spheres, const char* asset_dir);
/* A scene graph in the "gsgraph 1" text form (grayshift_b200/host/graph_io.cpp). */
gsh_scene* gsh_scene_load(const char* graph_path, int include_media);
int gsh_scene_dump(const gsh_scene* scene, const char* graph_path);
void gsh_scene_free(gsh_scene* scene);

/* Flattens (once; reference_tree = 1 emits the reference's median-split BVH instead of the SAH one)
 * and returns views that stay valid until gsh_scene_free / the next gsh_scene_flatten. */
int gsh_scene_flatten(gsh_scene* scene, int max_leaf_size, int reference_tree);
const gs_scene_desc* gsh_scene_desc(gsh_scene* scene);
int gsh_scene_camera(gsh_scene* scene, gs_camera* camera_out, gs_sample_settings* settings_out);
int gsh_scene_set_sampling(gsh_scene* scene, double confidence, double tolerance, uint32_t batch_size,
                           uint32_t max_samples, uint32_t max_depth);

/* Camera::render (camera.rs:100): flatten + gs_scene_create + gs_render + PPM (P3) text to `path`. */
int gsh_scene_render_ppm(gsh_scene* scene, const char* path, int device, uint64_t seed, gs_stats* stats);

/* Binned-SAH BVH over n boxes (6 doubles each: min xyz, max xyz).  nodes_out has room for 2n-1 nodes,
 * order_out for n indices; returns the node count through n_nodes_out. */
int gsh_bvh_build(const double* bounds, size_t n, int max_leaf_size, gs_bvh_node* nodes_out, size_t* n_nodes_out,
                  uint32_t* order_out);

/* PermutationTable of noise::Perlin::default() / Perlin::new(seed) (texture.rs:104). */
void gsh_perlin_table(uint32_t seed, uint8_t* out256);
/* The seeded scene-construction RNG (k-th draw after seeding), for cross-checks against the oracle. */
double gsh_scene_rng(uint64_t seed, int k);

#ifdef __cplusplus
}
#endif
#endif /* GRAYSHIFT_HOST_H */
