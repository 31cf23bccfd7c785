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
__device__ __forceinline__ void load_perm(uint8_t* s_perm, const uint8_t* g_perm) {
	if (threadIdx.x < 64) {
		uint32_t v = g_perm ? reinterpret_cast<const uint32_t*>(g_perm)[threadIdx.x] : 0u;
		reinterpret_cast<uint32_t*>(s_perm)[threadIdx.x] = v;
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
