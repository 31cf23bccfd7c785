// grayshift_b200/csrc/wavefront_kernels.cu -- the wavefront form of the fixed-count renderer (sm_100a).
//
// Why it exists (ncu on the megakernel, profiles/r01_v0_*, r01_v1_*): 107 KB of SASS against a 32 KB instruction
// cache (42 % of warp stalls were "no instruction"), 124-128 registers (16 warps/SM), and 10 of 32 lanes active per
// instruction because one kernel interleaves traversal, five materials, four textures, media and ray generation.
// Splitting along the reference's own call structure gives three small kernels per wave over a pool of in-flight
// paths kept in HBM as a structure of arrays (PathPool, 112 B per path, all accesses 128-bit and coalesced):
//
//   wf_generate_kernel  Camera::get_ray (camera.rs:204-226) for every free slot: path regeneration
//   wf_traverse_kernel  ConstantMedium::hit + BVHNode/AABB/Sphere/Quad::hit (world.hit, camera.rs:177): persistent
//                       warps that fetch the next ray the moment a lane finishes its ray (dynamic ray fetch)
//   wf_classify_kernel  material-sorted queue compaction: every traversed ray is filed under its shading class (miss, Lambertian
//                       with a solid / another texture, metal, glass, light, isotropic) -- lanes of a warp that share a class
//                       find each other with match.any, one atomic per class and warp reserves their places in the class's queue
//   wf_shade_kernel     emitted + scatter + value_at + sample_background (camera.rs:178-201) over the queues, class by class
//                       (a warp shades one material), accumulation of finished paths (camera.rs:141), and the free list for
//                       the next generate
//
// A slot's path survives from wave to wave; the host launches waves until every (pixel, sample) item has been
// started and the pool is empty.  Random numbers are the same Philox streams as in the megakernel and the oracle
// (counter = pixel, sample, draw), so the two pipelines render the same image up to fp32 summation order.
#include "gs_internal.h"
#include "gs_kernel_util.cuh"

namespace gs {

constexpr int kRefill = 8;  // traverse: fetch new rays once this many lanes of a warp are idle

__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31u)) - 1u; }

// ---- generate ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) wf_generate_kernel(const __grid_constant__ WavefrontParams P, const int first_wave) {
	const unsigned i = blockIdx.x * kBlock + threadIdx.x;
	const unsigned lane = threadIdx.x & 31u;
	const unsigned n_todo = first_wave ? (unsigned)P.pool.n : P.ctr->n_free;
	const bool has = i < n_todo;
	const int slot = has ? (first_wave ? (int)i : P.free_list[i]) : -1;
	const unsigned m = __ballot_sync(kFull, has);
	if (m == 0u) return;
	unsigned long long base = 0;
	const int leader = __ffs(m) - 1;
	if ((int)lane == leader) base = atomicAdd(&P.ctr->next_item, (unsigned long long)__popc(m));
	base = __shfl_sync(kFull, base, leader);
	const unsigned long long item = base + __popc(m & lanemask_lt());
	const bool start = has && item < P.total_items;
	if (start) {
		PathState st;
		const uint32_t pixel = (uint32_t)(item / P.sample_count);
		const uint32_t sample = (uint32_t)(item % P.sample_count) + P.sample_begin;
		path_begin(st, P.cam, P.seed, pixel, sample);
		P.pool.org[slot] = make_double4(st.ray.ox, st.ray.oy, st.ray.oz, st.ray.time);
		P.pool.dir[slot] = make_double4(st.ray.dx, st.ray.dy, st.ray.dz, 0.0);
		P.pool.thr[slot] = make_float4(1.f, 1.f, 1.f, 0.f);
		P.pool.rad[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
		P.pool.ids[slot] = make_uint4(pixel, sample, st.rng.draw, 0u | kSlotActive);
	} else if (has && first_wave) {
		P.pool.ids[slot] = make_uint4(0u, 0u, 0u, 0u);
	}
	const unsigned ms = __ballot_sync(kFull, start);
	if (ms && (int)lane == leader) {
		atomicAdd(&P.ctr->n_live, __popc(ms));
		atomicAdd(P.stats + 1, (unsigned long long)__popc(ms));  // paths started
	}
}

// ---- traverse ------------------------------------------------------------------------------------------
template <int kF>
__global__ void __launch_bounds__(kBlock) wf_traverse_kernel(const __grid_constant__ WavefrontParams P) {
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt = lanemask_lt();
	const unsigned n = (unsigned)P.pool.n;
	int slot = -1;
	Ray ray;
	Trav2 tv;  // the wavefront pipeline walks the binary tree
	RayConsts rc;
	int medium = -1;
	int stack_mem[kMaxStack];
	const LocalStack stack{stack_mem, nullptr};
	PathCounters pc = {0u, 0u, 0u};
	bool drained = false;
	tv.park();

	for (;;) {
		const unsigned idle = __ballot_sync(kFull, slot < 0);
		if (!drained && (__popc(idle) >= kRefill)) {
			unsigned base = 0;
			if (lane == 0) base = atomicAdd(&P.ctr->next_ray, (unsigned)__popc(idle));
			base = __shfl_sync(kFull, base, 0);
			if (base + (unsigned)__popc(idle) >= n) drained = true;
			if (slot < 0) {
				const unsigned s = base + __popc(idle & lt);
				if (s < n) {
					const uint4 id = P.pool.ids[s];
					if (id.w & kSlotActive) {
						const double4 o = P.pool.org[s], d = P.pool.dir[s];
						ray.ox = o.x; ray.oy = o.y; ray.oz = o.z; ray.time = o.w;
						ray.dx = d.x; ray.dy = d.y; ray.dz = d.z;
						// media first (volume.rs:32-63): the closest medium event bounds the surface search
						double tmax = 1.7976931348623157e308;
						medium = -1;
						++pc.rays;
						if (P.sc.n_media > 0) {
							const double a = ray.dx * ray.dx + ray.dy * ray.dy + ray.dz * ray.dz;
							const double inv_a = f64_rcp(a);
							// the media draw from the path's own stream (same order as the megakernel's segment_begin); the
							// advanced draw counter goes back into the pool for the shade kernel
							PathRng rng;
							rng.pixel = id.x; rng.sample = id.y; rng.draw = id.z;
							rng.k0 = (uint32_t)P.seed; rng.k1 = (uint32_t)(P.seed >> 32);
							if (id.z & 3u) rng.buf = philox4x32_10(id.x, id.y, id.z >> 2, 0u, rng.k0, rng.k1);
							for (int mi = 0; mi < P.sc.n_media; ++mi) {
								const float u01 = rng.next_f32();
								double tm;
								if (medium_event(P.sc, P.sc.media[mi], ray, a, inv_a, 0.001, tmax, u01, tm)) { tmax = tm; medium = mi; }
							}
							reinterpret_cast<uint32_t*>(P.pool.ids + s)[2] = rng.draw;
						}
						tv.init(P.sc, ray, stack, tmax);
						rc = ray_consts(ray, tmax);
						slot = (int)s;
					}
				}
			}
		}
		if (__ballot_sync(kFull, slot >= 0) == 0u) {
			if (drained) break;
			continue;
		}
		if (slot >= 0) {
			trav_round<kF, 0>(P.sc, ray, rc, 0.001, tv, stack, &pc.node_visits);
			if (tv.done()) {
				// closest event: surface hit (ref), else medium event (t = the medium's candidate), else miss
				reinterpret_cast<double*>(P.pool.dir + slot)[3] = tv.best.t;
				P.pool.hit[slot] = make_uint4(tv.best.ref, (unsigned)medium, __float_as_uint(tv.best.u), __float_as_uint(tv.best.v));
				slot = -1;
			}
		}
	}
	flush_stats(P.stats, pc, 0u);
}

// ---- classify: material-sorted queue compaction -------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) wf_classify_kernel(const __grid_constant__ WavefrontParams P) {
	const unsigned s = blockIdx.x * kBlock + threadIdx.x;
	const unsigned lane = threadIdx.x & 31u;
	int cls = -1;  // -1: nothing to shade in this slot
	if (s < (unsigned)P.pool.n && (P.pool.ids[s].w & kSlotActive)) {
		const uint4 hw = P.pool.hit[s];
		if (hw.x == kNoHit && (int)hw.y < 0) cls = 0;  // miss: sample_background
		else {
			int mat;
			if (hw.x != kNoHit) {
				const uint32_t index = hw.x & kRefIndexMask, kind = hw.x >> kRefKindShift;
				mat = kind == kKindQuad ? P.sc.quad_aux[index].x : (kind == kKindTriangle ? P.sc.tri_aux[index].x : P.sc.sph_aux[index].x);
			} else mat = P.sc.media[hw.y].phase_mat;
			const DMaterial m = P.sc.materials[mat];
			if (m.kind == 0) cls = P.sc.textures[m.tex].kind == 0 ? 1 : 2;  // Lambertian: solid colour / checker, image, noise
			else cls = 2 + m.kind;                                             // metal 3, dielectric 4, light 5, isotropic 6
		}
	}
	const unsigned live = __ballot_sync(kFull, cls >= 0);
	if (cls >= 0) {
		const unsigned peers = __match_any_sync(live, cls);  // the lanes of this warp in the same class
		const int leader = __ffs(peers) - 1;
		unsigned base = 0;
		if ((int)lane == leader) base = atomicAdd(P.shade_count + cls, (unsigned)__popc(peers));
		base = __shfl_sync(peers, base, leader);
		P.shade_queue[(size_t)cls * P.pool.n + base + __popc(peers & lanemask_lt())] = (int)s;
	}
}

// ---- shade -----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) wf_shade_kernel(const __grid_constant__ WavefrontParams P) {
	__shared__ __align__(16) uint8_t s_perm[kPermBytes];
	load_perm(s_perm, P.sc.perlin_perm);
	unsigned s = blockIdx.x * kBlock + threadIdx.x;
	const unsigned lane = threadIdx.x & 31u;
	PathCounters pc = {0u, 0u, 0u};
	bool terminated = false;
	uint4 id = make_uint4(0u, 0u, 0u, 0u);
	if (P.shade_queue) {
		// thread i takes the i-th entry of the concatenated class queues; each class starts on a warp boundary, so a warp
		// never mixes two classes
		unsigned i = s, slot = 0xffffffffu;
#pragma unroll
		for (int c = 0; c < kShadeClasses; ++c) {
			const unsigned n = P.shade_count[c], padded = (n + 31u) & ~31u;
			if (slot == 0xffffffffu) {
				if (i < n) slot = (unsigned)P.shade_queue[(size_t)c * P.pool.n + i];
				else if (i < padded) slot = 0xfffffffeu;  // padding lane of this class's last warp
				else i -= padded;
			}
		}
		s = slot < 0xfffffffeu ? slot : (unsigned)P.pool.n;
	}
	if (s < (unsigned)P.pool.n) id = P.pool.ids[s];
	if (id.w & kSlotActive) {
		PathState st;
		const double4 o = P.pool.org[s], d = P.pool.dir[s];
		const float4 thr = P.pool.thr[s], rad = P.pool.rad[s];
		const uint4 hw = P.pool.hit[s];
		st.ray.ox = o.x; st.ray.oy = o.y; st.ray.oz = o.z; st.ray.time = o.w;
		st.ray.dx = d.x; st.ray.dy = d.y; st.ray.dz = d.z;
		st.T = mk3(thr.x, thr.y, thr.z);
		st.L = mk3(rad.x, rad.y, rad.z);
		st.bounce = id.w & 0xffffu;
		st.rng.pixel = id.x; st.rng.sample = id.y; st.rng.draw = id.z;
		st.rng.k0 = (uint32_t)P.seed; st.rng.k1 = (uint32_t)(P.seed >> 32);
		// mid-block: rebuild the Philox block the next draws come from (draw k = word k&3 of block k>>2)
		if (id.z & 3u) st.rng.buf = philox4x32_10(id.x, id.y, id.z >> 2, 0u, st.rng.k0, st.rng.k1);
		Hit h;
		h.t = d.w; h.ref = hw.x; h.u = __uint_as_float(hw.z); h.v = __uint_as_float(hw.w);
		bool finished = segment_shade<true>(P.sc, P.cam, s_perm, st, pc, h, d.w, (int)hw.y);
		if (!finished && st.bounce >= P.cam.max_depth) finished = true;  // the next level returns ZERO (camera.rs:175)
		if (finished) {
			red_add_v4(P.accum + st.pixel(), st.L.x, st.L.y, st.L.z, 1.0f);  // pixel_color += sample_color (camera.rs:141)
			P.pool.ids[s] = make_uint4(id.x, id.y, st.rng.draw, 0u);
			terminated = true;
		} else {
			P.pool.org[s] = make_double4(st.ray.ox, st.ray.oy, st.ray.oz, st.ray.time);
			P.pool.dir[s] = make_double4(st.ray.dx, st.ray.dy, st.ray.dz, 0.0);
			P.pool.thr[s] = make_float4(st.T.x, st.T.y, st.T.z, 0.f);
			P.pool.rad[s] = make_float4(st.L.x, st.L.y, st.L.z, 0.f);
			P.pool.ids[s] = make_uint4(id.x, id.y, st.rng.draw, st.bounce | kSlotActive);
		}
	}
	// released slots -> free list (warp-aggregated: one atomic per warp)
	const unsigned m = __ballot_sync(kFull, terminated);
	if (m) {
		unsigned base = 0;
		const int leader = __ffs(m) - 1;
		if ((int)lane == leader) {
			base = atomicAdd(&P.ctr->n_free, (unsigned)__popc(m));
			atomicSub(&P.ctr->n_live, __popc(m));
		}
		base = __shfl_sync(kFull, base, leader);
		if (terminated) P.free_list[base + __popc(m & lanemask_lt())] = (int)s;
	}
	flush_stats(P.stats, pc, 0u);
}

// ---- launch wrappers ---------------------------------------------------------------------------------------
cudaError_t launch_wf_generate(const WavefrontParams& p, bool first_wave, cudaStream_t stream) {
	wf_generate_kernel<<<(p.pool.n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(p, first_wave ? 1 : 0);
	return cudaGetLastError();
}

template <int kF>
static cudaError_t wf_traverse_launch(const WavefrontParams& p, int sm_count, cudaStream_t stream) {
	int per_sm = 0;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_traverse_kernel<kF>, kBlock, 0);
	if (per_sm < 1) per_sm = 1;
	long long grid = (long long)sm_count * per_sm;  // persistent: SMs x resident CTAs
	const long long needed = (p.pool.n + kBlock - 1) / kBlock;
	if (needed < grid) grid = needed;
	wf_traverse_kernel<kF><<<(unsigned)grid, kBlock, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_wf_traverse(const WavefrontParams& p, int sm_count, bool count_nodes, bool triangles, cudaStream_t stream) {
	switch ((count_nodes ? 1 : 0) | (triangles ? 2 : 0)) {
		case 0: return wf_traverse_launch<0>(p, sm_count, stream);
		case 1: return wf_traverse_launch<1>(p, sm_count, stream);
		case 2: return wf_traverse_launch<2>(p, sm_count, stream);
		default: return wf_traverse_launch<3>(p, sm_count, stream);
	}
}

cudaError_t launch_wf_classify(const WavefrontParams& p, cudaStream_t stream) {
	wf_classify_kernel<<<(p.pool.n + kBlock - 1) / kBlock, kBlock, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_wf_shade(const WavefrontParams& p, cudaStream_t stream) {
	// sorted: every class is padded to a whole warp, so up to kShadeClasses * 31 more threads than slots
	const int threads = p.pool.n + (p.shade_queue ? kShadeClasses * 32 : 0);
	wf_shade_kernel<<<(threads + kBlock - 1) / kBlock, kBlock, 0, stream>>>(p);
	return cudaGetLastError();
}

}  // namespace gs
