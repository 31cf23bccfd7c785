// grayshift_b200/csrc/trace_closest.cu -- parity kernel K1: closest-hit primitive id and t for a batch of
// rays (gs_trace_closest).  Same traversal and intersection code as the render megakernel
// (gs_device.cuh), compiled in this translation unit with -fmad=false so the f64 primitive
// arithmetic is evaluated without FMA contraction, like the Rust reference.
#define GS_EXACT_F64 1  // IEEE division and square root, like the reference
#include "gs_internal.h"

namespace gs {

__global__ void __launch_bounds__(128) trace_closest_kernel(const __grid_constant__ DeviceScene sc, const double* __restrict__ rays, size_t n,
                                                            double tmin, double tmax, int* __restrict__ prim_id, double* __restrict__ t_out) {
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const double* rp = rays + k * 7;
	Ray r;
	r.ox = rp[0]; r.oy = rp[1]; r.oz = rp[2]; r.dx = rp[3]; r.dy = rp[4]; r.dz = rp[5]; r.time = rp[6];
	Hit h = trace_closest<2>(sc, r, tmin, tmax);
	if (h.ref == kNoHit) { prim_id[k] = -1; t_out[k] = 0.0; return; }
	const uint32_t index = h.ref & kRefIndexMask;
	const uint32_t kind = h.ref >> kRefKindShift;
	prim_id[k] = kind == 0u ? sc.sph_aux[index].y : (kind == 1u ? sc.quad_aux[index].y : sc.tri_aux[index].y);
	t_out[k] = h.t;
}

cudaError_t launch_trace_closest(const DeviceScene& sc, const double* rays, size_t n, double tmin, double tmax, int* prim_id, double* t,
                                 cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	trace_closest_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(sc, rays, n, tmin, tmax, prim_id, t);
	return cudaGetLastError();
}

}  // namespace gs
