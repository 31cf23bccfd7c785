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
};

struct AdaptiveParams {
	DeviceScene sc;
	DCamera cam;
	unsigned long long seed;
	double confidence_sq, tolerance_sq;
	unsigned batch_size, max_samples;
	unsigned pixel_begin, pixel_stride, n_pixels;  // this rank renders pixels pixel_begin + k*pixel_stride
	float4* accum;
	unsigned long long* work_counter;
	unsigned long long* stats;
};

// launch wrappers (defined next to the kernels); all asynchronous on `stream`
cudaError_t launch_render_fixed(const RenderParams& p, int sm_count, bool count_nodes, cudaStream_t stream);
cudaError_t launch_render_adaptive(const AdaptiveParams& p, int sm_count, cudaStream_t stream);
cudaError_t launch_resolve(const float4* accum, int n_pixels, float* out_rgb, unsigned* out_counts, cudaStream_t stream);
cudaError_t launch_camera_rays(const DCamera& cam, unsigned long long seed, unsigned sample, int pixel_begin, int pixel_end, double* rays_out,
                               cudaStream_t stream);
cudaError_t launch_path_radiance(const DeviceScene& sc, const DCamera& cam, unsigned long long seed, const unsigned* pixel,
                                 const unsigned* sample, size_t n, float* out_rgb, cudaStream_t stream);
cudaError_t launch_trace_closest(const DeviceScene& sc, const double* rays, size_t n, double tmin, double tmax, int* prim_id, double* t,
                                 cudaStream_t stream);
int render_fixed_blocks_per_sm(bool count_nodes);

}  // namespace gs
