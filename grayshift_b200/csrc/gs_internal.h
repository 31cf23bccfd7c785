// grayshift_b200/csrc/gs_internal.h -- declarations shared by the C-ABI translation unit (abi.cu) and
// the kernel translation units (render_kernels.cu, trace_closest.cu).  Not part of the public ABI.
#pragma once
#include "gs_device.cuh"

namespace gs {

struct RenderParams {
	DeviceScene sc;
	DCamera cam;
	unsigned long long seed;
	unsigned sample_begin, sample_count;   // per-pixel sample index range of this launch
	unsigned long long total_items;        // n_pixels * sample_count
	unsigned chunk;                        // items fetched per warp-level atomic
	float4* accum;                         // W*H (sum r, sum g, sum b, count)
	unsigned long long* work_counter;      // dynamic work distribution
	unsigned long long* stats;             // [0] rays, [1] paths, [2] texel fetches, [3] node visits
	// adaptive rounds (Camera::sample's batch loop, camera.rs:135-165, run batch by batch over the pixels still sampling -- all rounds
	// inside ONE cooperative launch: render phase, grid-wide sync, stop test, grid-wide sync, next round)
	uint32_t* pixel_list;                  // two lists of n_pixels entries each, back to back: [0] = the pixels of round 0; nullptr = every pixel (fixed count)
	unsigned* n_list;                      // [0], [1] = lengths of the two lists, [2] = index of the list the rounds leave for the tail kernel
	unsigned long long* work_counter2;     // a second work counter: rounds alternate between the two
	double2* lum;                          // per pixel (sum of luminance, sum of its square), camera.rs:132-133,146-147
	double confidence_sq, tolerance_sq;    // the stop test of camera.rs:150-164
	unsigned max_samples, n_pixels, tail_pixels;  // n_pixels = capacity of each list; the rounds end once <= tail_pixels pixels are left
	bool rounds_in_kernel;                 // false: this launch renders ONE batch (round0) and the host launches the stop test (A/B alternative)
	unsigned round0;
	bool small_scene;                      // nodes + primitives <= kSmallSceneBytes: the launch takes the 6-CTAs-per-SM build
	bool wide;                             // traverse sc.nodes8 (eight-wide tree) instead of sc.nodes
	uint32_t philox_rk[20];                // the ten Philox round-key pairs of `seed` (filled by launch_render_fixed)
};

struct AdaptiveTailParams {  // the last few pixels: one warp per pixel runs the rest of its batch loop (adaptive_tail_kernel)
	DeviceScene sc;
	DCamera cam;
	unsigned long long seed;
	double confidence_sq, tolerance_sq;
	unsigned batch_size, max_samples;
	const uint32_t* lists;     // the two pixel lists of the rounds (n_pixels entries each)
	const unsigned* n_list;    // [0], [1] their lengths, [2] which of them holds the pixels still sampling
	unsigned n_pixels;
	float4* accum;
	const double2* lum;
	unsigned long long* work_counter;
	unsigned long long* stats;
};

// launch wrappers (defined next to the kernels); all asynchronous on `stream`
cudaError_t launch_render_fixed(const RenderParams& p, int sm_count, bool count_nodes, bool triangles, cudaStream_t stream);
cudaError_t launch_adaptive_init(uint32_t* list, unsigned* n_list, unsigned pixel_begin, unsigned pixel_stride, unsigned n_pixels, cudaStream_t stream);
cudaError_t launch_adaptive_tail(const AdaptiveTailParams& p, unsigned n_upper, unsigned n_known, int sm_count, bool wide, cudaStream_t stream);
cudaError_t launch_adaptive_decide(const RenderParams& p, unsigned cur, cudaStream_t stream);  // host-driven rounds only
cudaError_t launch_resolve(const float4* accum, int n_pixels, float* out_rgb, unsigned* out_counts, cudaStream_t stream);
cudaError_t launch_resolve_rgb8(const float4* accum, int n_pixels, unsigned char* out_rgb8, unsigned* out_counts, cudaStream_t stream);
cudaError_t launch_camera_rays(const DCamera& cam, unsigned long long seed, unsigned sample, int pixel_begin, int pixel_end, double* rays_out,
                               cudaStream_t stream);
cudaError_t launch_path_radiance(const DeviceScene& sc, const DCamera& cam, unsigned long long seed, const unsigned* pixel,
                                 const unsigned* sample, size_t n, float* out_rgb, bool wide, cudaStream_t stream);
// exact = the parity build (trace_closest.cu: -fmad=false, IEEE division / square root); otherwise the render kernels' own
// arithmetic (render_kernels.cu: FMA contraction, rcp/rsqrt.approx.f64 + Newton) over the same rays
cudaError_t launch_trace_closest(const DeviceScene& sc, const double* rays, size_t n, double tmin, double tmax, int* prim_id, double* t,
                                 bool wide, cudaStream_t stream);
cudaError_t launch_trace_closest_fast(const DeviceScene& sc, const double* rays, size_t n, double tmin, double tmax, int* prim_id, double* t,
                                      bool wide, cudaStream_t stream);
int render_fixed_blocks_per_sm(bool count_nodes, bool triangles);

// Multi-GPU resolve: the W*H float4 accumulators of up to kMaxPeers devices are summed while they are resolved -- the kernel runs
// on one device and reads the others' buffers through peer-mapped pointers (NVLink / NVSwitch), so the cross-GPU sum, the division
// by the sample count (camera.rs:167) and, for the byte output, write_color's transfer (color.rs:8-26) are one pass over the data.
constexpr int kMaxPeers = 16;
struct PeerAccums { const float4* acc[kMaxPeers]; int n; };
cudaError_t launch_reduce_resolve(const PeerAccums& peers, int n_pixels, float* out_rgb, unsigned char* out_rgb8, unsigned* out_counts,
                                  float4* sum_out, cudaStream_t stream);

// ---- wavefront pipeline (wavefront_kernels.cu) -------------------------------------------------------
// Path pool: structure of arrays in HBM, one slot per in-flight path, every field a 16-byte record so each
// access is one 128-bit load/store and consecutive slots coalesce.
struct PathPool {
	double4* org;   // ray origin x, y, z, time
	double4* dir;   // ray direction x, y, z, closest t found by the traversal (surface hit or medium event)
	float4* thr;    // throughput (prod of attenuations) r, g, b, -
	float4* rad;    // radiance so far r, g, b, -
	uint4* ids;     // pixel, sample, draw counter, bounce | kSlotActive
	uint4* hit;     // primitive ref (kNoHit = none), medium index (-1 = none), u bits, v bits
	int n;
};
constexpr unsigned kSlotActive = 0x80000000u;

struct WaveCounters {              // one small device struct per scene
	unsigned long long next_item;  // next (pixel, sample) item to start
	unsigned next_ray;             // traverse kernel: next slot to fetch
	unsigned n_free;               // slots released by the shade kernel of this wave
	int n_live;                    // slots holding a path
};

constexpr int kShadeClasses = 8;  // wavefront shading queues: miss, Lambertian (solid / textured), metal, glass, light, isotropic, spare

struct WavefrontParams {
	DeviceScene sc;
	DCamera cam;
	PathPool pool;
	int* shade_queue;      // kShadeClasses segments of pool.n slots each: the rays of a wave sorted by shading class; nullptr = unsorted
	unsigned* shade_count; // kShadeClasses counters
	unsigned long long seed;
	unsigned sample_begin, sample_count;
	unsigned long long total_items;
	float4* accum;
	WaveCounters* ctr;
	int* free_list;        // slots to (re)generate; nullptr = every slot (first wave)
	unsigned long long* stats;
};

cudaError_t launch_wf_generate(const WavefrontParams& p, bool first_wave, cudaStream_t stream);
cudaError_t launch_wf_traverse(const WavefrontParams& p, int sm_count, bool count_nodes, bool triangles, cudaStream_t stream);
cudaError_t launch_wf_classify(const WavefrontParams& p, cudaStream_t stream);
cudaError_t launch_wf_shade(const WavefrontParams& p, cudaStream_t stream);

}  // namespace gs
