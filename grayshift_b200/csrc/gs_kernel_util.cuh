// grayshift_b200/csrc/gs_kernel_util.cuh -- small helpers shared by the kernel translation units.
#pragma once
#include "gs_shade.cuh"

namespace gs {

#ifndef GS_BLOCK
#define GS_BLOCK 128  // threads per CTA of the render kernels (a multiple of 32, at least 64)
#endif
constexpr int kBlock = GS_BLOCK;
constexpr unsigned kFull = 0xffffffffu;

// One 128-bit vector reduction into the (sum r, sum g, sum b, count) accumulator (sm_90+).
__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
	asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// Perlin permutation table -> shared memory (256 B per CTA).
// The Perlin tables of a CTA: perm[256] followed by the sixteen gradient vectors of perlin_grad (gs_shade.cuh) as three
// arrays of sixteen floats (x, y, z components), kPermBytes in all, 16-byte aligned.
constexpr int kPermBytes = 256 + 3 * 16 * 4;
__host__ __device__ constexpr int perm_bytes(bool grad) { return grad ? kPermBytes : 256; }  // kernels that never read the gradient table carry perm[] only
template <bool kGrad = true>
__device__ __forceinline__ void load_perm(uint8_t* s_perm, const uint8_t* g_perm) {
	if (threadIdx.x < 64) {
		uint32_t v = g_perm ? reinterpret_cast<const uint32_t*>(g_perm)[threadIdx.x] : 0u;
		reinterpret_cast<uint32_t*>(s_perm)[threadIdx.x] = v;
	}
	if (kGrad && threadIdx.x < 16) {
		// hash & 15 -> edge gradient: cases 12..15 repeat 0, 1, 9, 11; then (+-u) + (+-v) with u = x (h < 8) or y, v = y (h < 4) or z,
		// signs in bits 0 and 1
		const uint32_t h = (uint32_t)(0xB910BA9876543210ull >> (threadIdx.x * 4u)) & 15u;
		const float su = (h & 1u) ? -1.f : 1.f, sv = (h & 2u) ? -1.f : 1.f;
		float gx = 0.f, gy = 0.f, gz = 0.f;
		if (h < 8u) gx = su; else gy = su;
		if (h < 4u) gy = sv; else gz = sv;
		float* g = reinterpret_cast<float*>(s_perm + 256);
		g[threadIdx.x] = gx; g[16 + threadIdx.x] = gy; g[32 + threadIdx.x] = gz;
	}
	__syncthreads();
}

// Per-thread counters -> device totals: warp shuffle reduction, one atomic per warp and counter.
__device__ __forceinline__ void flush_stats(unsigned long long* stats, const PathCounters& pc, unsigned paths) {
	// one REDUX per counter (a shuffle tree costs five shuffles each plus their divergence fall-backs: 400 instructions
	// of a kernel that is sensitive to its code size)
	const unsigned r = __reduce_add_sync(kFull, pc.rays), t = __reduce_add_sync(kFull, pc.tex_fetches);
	const unsigned n = __reduce_add_sync(kFull, pc.node_visits), p = __reduce_add_sync(kFull, paths);
	if ((threadIdx.x & 31) == 0) {
		if (r) atomicAdd(stats + 0, (unsigned long long)r);
		if (p) atomicAdd(stats + 1, (unsigned long long)p);
		if (t) atomicAdd(stats + 2, (unsigned long long)t);
		if (n) atomicAdd(stats + 3, (unsigned long long)n);
	}
}

}  // namespace gs
