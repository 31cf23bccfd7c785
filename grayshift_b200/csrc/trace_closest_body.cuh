// grayshift_b200/csrc/trace_closest_body.cuh -- the closest-hit batch kernel, compiled twice: in trace_closest.cu with
// -fmad=false and IEEE f64 division / square root (the parity build), and in render_kernels.cu with the render kernels' own
// arithmetic (FMA contraction, rcp/rsqrt.approx.f64 + Newton).  GS_TRACE_NAME names the kernel and its launcher.
#define GS_TC_CAT_(a, b) a##b
#define GS_TC_CAT(a, b) GS_TC_CAT_(a, b)
namespace gs {

template <int kF>
__global__ void __launch_bounds__(128) GS_TC_CAT(GS_TRACE_NAME, _kernel)(const __grid_constant__ DeviceScene sc, const double* __restrict__ rays, size_t n,
                                                                        double tmin, double tmax, int* __restrict__ prim_id, double* __restrict__ t_out) {
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	const double* rp = rays + k * 7;
	Ray r;
	r.ox = rp[0]; r.oy = rp[1]; r.oz = rp[2]; r.dx = rp[3]; r.dy = rp[4]; r.dz = rp[5]; r.time = rp[6];
	Hit h = trace_closest<kF>(sc, r, tmin, tmax);
	if (h.ref == kNoHit) { prim_id[k] = -1; t_out[k] = 0.0; return; }
	const uint32_t index = h.ref & kRefIndexMask;
	const uint32_t kind = h.ref >> kRefKindShift;
	prim_id[k] = kind == kKindQuad ? sc.quad_aux[index].y : (kind == kKindTriangle ? sc.tri_aux[index].y : sc.sph_aux[index].y);
	t_out[k] = h.t;
}

cudaError_t GS_TC_CAT(launch_, GS_TRACE_NAME)(const DeviceScene& sc, const double* rays, size_t n, double tmin, double tmax, int* prim_id, double* t,
                                              bool wide, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	const unsigned grid = (unsigned)((n + 127) / 128);
	if (wide) GS_TC_CAT(GS_TRACE_NAME, _kernel)<18><<<grid, 128, 0, stream>>>(sc, rays, n, tmin, tmax, prim_id, t);
	else GS_TC_CAT(GS_TRACE_NAME, _kernel)<2><<<grid, 128, 0, stream>>>(sc, rays, n, tmin, tmax, prim_id, t);
	return cudaGetLastError();
}

}  // namespace gs
#undef GS_TC_CAT
#undef GS_TC_CAT_
