// grayshift_b200/csrc/render_kernels.cu -- the render kernels for sm_100a.
//
// render_fixed_kernel: persistent megakernel for fixed sample counts (tolerance == 0).  Replaces the
//   rayon per-pixel loop + Camera::sample + ray_color (camera.rs:100-202).  One warp = 32 independent
//   paths; a lane whose path ends takes the next (pixel, sample) item at the top of the next
//   iteration (path regeneration), so the warp stays full until the whole image is exhausted.
//   Work items are handed out in chunks through one global 64-bit atomic per warp per chunk.
//   Each lane keeps (sum r, sum g, sum b, n) for its current pixel in shared memory and flushes with one
//   128-bit vector reduction (red.global.add.v4.f32) when its pixel changes.
// adaptive sampling (camera.rs:125-171, exact batch / stop-test semantics, Q3-Q5): the same megakernel, launched ONCE
//   (cooperatively) for all batches: a round renders one batch of every pixel still sampling, a grid-wide sync, the stop
//   test rebuilds the list, a second sync, next round -- no host round trip; adaptive_tail_kernel (one warp per pixel)
//   finishes the last few pixels.  See below.
// Grids are sized to SM count x resident CTAs (persistent), never to the image.
#include <cooperative_groups.h>

#include <atomic>
#include <cstdlib>

#include "gs_internal.h"
#include "gs_kernel_util.cuh"

// "variable st is used before its value is set": every lane starts in state kDead and reads its PathState only after
// path_begin wrote it; zero-initialising it to please the front end changes ptxas's register allocation (+36 B of spills).
#pragma nv_diag_suppress 549

namespace gs {

// Lane states of the megakernel.  A lane is a path; the warp alternates between two phases:
//   phase A (shade / regenerate / set up): lanes whose ray has finished are shaded, finished paths are
//            accumulated and replaced by the next (pixel, sample) item, and every lane with a fresh ray
//            evaluates the media and initialises its traversal;
//   phase B (traverse): all lanes with an unfinished ray walk the BVH.  The phase ends as soon as kMinShade
//            lanes have finished (or nobody is traversing): the others KEEP their traversal state (node, sp,
//            closest hit, stack) and continue in the next phase B.
enum LaneState : int { kDead = 0, kTraverse = 1, kShade = 2, kSetup = 3 };
// Tuning knobs (overridable with -D for the A/B sweeps of scripts/build_variants.sh + scripts/gpu_ab.sh; defaults = the
// measured best, profiles/r01_v5_variant_sweeps.log).
#ifndef GS_MIN_SHADE
#define GS_MIN_SHADE 24   // leave phase B once this many lanes have finished their ray (32 = classic while-while, every ray to completion)
#endif
#ifndef GS_MIN_BLOCKS
#define GS_MIN_BLOCKS 5   // __launch_bounds__ min CTAs per SM (register cap = 65536 / (128 * GS_MIN_BLOCKS)).  The kernel is latency-
                          // bound, so occupancy pays until spills eat the L1: 5 = 96 registers, 20 warps/SM, no spills -> best on
                          // C4 (1518 M paths/s; its 0.5 MB scene is the one that does not fit the L1); 6 = 80 registers with 256 B
                          // of spills in the shading code -> 1446 on C4 but +5 % on C1 and +7 % on C3 (small scenes); 4 = 110
                          // registers -> 1403.  profiles/r01_v5_variant_sweeps.log.  Scenes of <= 128 KB get the build with one
                          // more CTA per SM (kernel variants with bit 3 set).
#endif
#ifndef GS_TOP_SMEM
#define GS_TOP_SMEM 0     // nodes of the top BVH levels (breadth-first prefix) staged in shared memory per CTA; 0 = off.  Measured -9 %
                          // with 64 or 256 nodes: the top of the tree is L1-resident anyway and shared memory takes L1 capacity away
#endif
constexpr int kTopNodes = GS_TOP_SMEM;
static_assert(kTopNodes <= kNodePad, "gs_scene_create pads the node array to kNodePad entries");
#ifndef GS_MIN_SHADE_WIDE
#define GS_MIN_SHADE_WIDE 20  // the same threshold in the kernels that walk the eight-wide tree: a wide node step costs three times a binary one,
                              // so waiting for the 24th ray is dearer there (16 / 20 / 24: 386.9 / 388.4 / 377.2 M paths/s on config 5, profiles/r02_v6_variants.log)
#endif
#ifndef GS_NOISE_COOP
#define GS_NOISE_COOP 1   // NoiseTextures are evaluated by the whole warp after the shading phase (noise_turb_coop); 0 = by the lane that met one
#endif
#ifndef GS_RNG_WINDOW
#define GS_RNG_WINDOW 0   // the megakernel's paths draw from a two-block window in shared memory, filled once per level at a converged point
                          // (PathRngW, gs_device.cuh); 0 = one block in registers, computed at the draw that needs it
#endif
#ifndef GS_PHILOX_RK
#define GS_PHILOX_RK 0    // Philox round keys from the parameter block, rounds unrolled (philox4x32_10_rk: 4 instructions per round instead of 7.5):
                          // 1 = every kernel variant, 2 = the small-scene variants (kF & 8) only, 0 = none (key schedule in registers, rolled loop).
                          // Measured (profiles/r02_v15_ab.log sections 6 and 12): +3.9 % config 1, +1.1 % config 3, +4 % scene 3, +2 % config 2 -- but
                          // -7...-15 % on config 4, whose kernel falls off the instruction cache with ten unrolled copies; and on the final code the
                          // small-scene variants split: +3.9 % C1, +1.2 % C3, but -6.7 % C2, -2.3 % scene 5 (204 B of spills).  Left off.
#endif
#ifndef GS_NODE_UNROLL2
#define GS_NODE_UNROLL2 0
#endif
#ifndef GS_BAIL
#define GS_BAIL 6         // resumable traversal: the node loop of a round ends once no more than this many lanes still descend
#endif
#ifndef GS_SMEM_STACK
#define GS_SMEM_STACK 0   // entries of the traversal stack kept in shared memory ([entry][thread]); the rest stays in local memory
#endif
constexpr int kSmemStack = GS_SMEM_STACK;
constexpr int kMinShade = GS_MIN_SHADE;
constexpr int kBail = GS_BAIL;

// kF: bit 0 = count BVH node visits, bit 1 = the scene holds triangles, bit 2 = one batch of the adaptive sampler (work items
// come from a device-resident list of the pixels still sampling, and the luminance sums of camera.rs:146-147 are kept),
// bit 4 = traverse the eight-wide compressed tree (DNode8 / Trav8) instead of the binary one -- chosen per scene by
// gs_scene_create: a large scene's binary nodes come from L2 and every fetch is a dependent round trip
// bit 3 = the scene's nodes and primitives fit the L1 several times over: one more CTA per SM (80 registers, some spills into the
// shading code) -- measured +3...4 % on C1 / C3 and -3 % on C4, whose 0.5 MB working set feels the spills in the L1
#ifdef GS_MAXNREG
#define GS_RENDER_BOUNDS(kF) __maxnreg__(GS_MAXNREG)  // A/B builds: an explicit register cap instead of the CTA count
#else
#define GS_RENDER_BOUNDS(kF) __launch_bounds__(kBlock, ((kF) & 8) ? GS_MIN_BLOCKS + 1 : GS_MIN_BLOCKS)
#endif
template <int kF>
__global__ void GS_RENDER_BOUNDS(kF) render_fixed_kernel(const __grid_constant__ RenderParams P) {
	constexpr bool kAdaptive = (kF & 4) != 0;
	using Trav = typename std::conditional<(kF & 16) != 0, Trav8, Trav2>::type;
	// adaptive: one iteration of the round loop below per batch.  What changes from round to round is CTA-uniform and lives in
	// shared memory -- as registers these seven values spilled the traversal loop (-27 % on the final scene)
	__shared__ struct { unsigned long long total_items; const uint32_t* pixel_list; unsigned long long* work_counter; unsigned round; } s_rnd;
	if (kAdaptive && threadIdx.x == 0) {
		// (host-driven rounds: this launch is round P.round0 on list P.round0 & 1; in-kernel rounds start at 0)
		s_rnd.total_items = (unsigned long long)__ldcg(P.n_list + (P.round0 & 1u)) * P.sample_count;
		s_rnd.pixel_list = P.pixel_list + (size_t)(P.round0 & 1u) * P.n_pixels; s_rnd.work_counter = P.work_counter; s_rnd.round = P.round0;
	}
	__shared__ __align__(16) uint8_t s_perm[perm_bytes((tune_of(kF) & 1) != 0)];
	__shared__ float4 s_top[kTopNodes > 0 ? kTopNodes * 4 : 1];
	if (kTopNodes > 0)
		for (int i = threadIdx.x; i < kTopNodes * 4; i += kBlock) s_top[i] = __ldg(reinterpret_cast<const float4*>(P.sc.nodes) + i);
	load_perm<(tune_of(kF) & 1) != 0>(s_perm, P.sc.perlin_perm);
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	constexpr int kShadeTune = tune_of(kF) | (GS_NOISE_COOP ? 4 : 0);
	PathCounters pc = {0u, 0u, 0u};
#if GS_RNG_WINDOW
	__shared__ uint32_t s_rng[9][kBlock];  // PathRngW: two Philox blocks per lane + the index of the newer one (4.5 KB per CTA)
	PathStateT<PathRngW<kBlock>> st;
	st.rng.win = &s_rng[0][threadIdx.x];
#else
	PathStateT<PathRngT<GS_PHILOX_RK == 1 || (GS_PHILOX_RK == 2 && (kF & 8) != 0)>> st;
	st.rng.rk = P.philox_rk;
#endif
	int state = kDead;
	Trav tv;
	tv.park();
	int seg_medium = -1;
	int stack_mem[Trav::kStackWords - kSmemStack > 0 ? Trav::kStackWords - kSmemStack : 1];
	__shared__ int s_stack[kSmemStack > 0 ? kSmemStack : 1][kBlock];
	const StackT<kSmemStack, kBlock> stack{stack_mem, &s_stack[0][threadIdx.x]};
	// The kernel is register-bound (96 registers = 20 warps per SM), so what is touched once per PATH lives in shared
	// memory instead of in registers held through every traversal: the lane's running sum for its current pixel
	// (float4 + pixel id, [field][thread]: conflict-free) and the warp's current chunk of work items [next, end) with
	// its "no more chunks" flag (warp-uniform).  2.6 KB per CTA; parking the whole path state the same way (12 KB) was
	// measured slower -- the L1 capacity it takes away costs more than the spills it saves.
	__shared__ float s_sum[4][kBlock];
	__shared__ double s_lum[kAdaptive ? 2 : 1][kAdaptive ? kBlock : 1];  // adaptive: the lane's share of sum, sq_sum (f64 like the reference)
	__shared__ uint32_t s_pix[kBlock];
	__shared__ unsigned long long s_chunk[kBlock / 32][4];  // next item, end of the chunk, "no more chunks", (pixel << 32 | sample) of the next item
	unsigned long long* const chunk = s_chunk[threadIdx.x >> 5];
	s_pix[threadIdx.x] = 0xffffffffu;
	if (lane == 0) { chunk[0] = 0ull; chunk[1] = 0ull; chunk[2] = 0ull; chunk[3] = 0ull; }
	__syncwarp();

	// off / sample_count for off <= sample_count + 31 (an item's sample index plus a lane rank): with 32 or more samples per pixel the
	// quotient is 0 or 1, so repeated subtraction -- one compare -- replaces the 32-bit division (~20 instructions, twice per
	// regeneration, in the kernel's hottest straight-line code); a launch of k < 32 samples per pixel loops up to 32 / k times
	// (the eight-wide kernels keep the division: with the loop ptxas spills 68 B more into their traversal, -1.6 % on config 5)
	auto quot = [&](uint32_t off) -> uint32_t {
		if (kF & 16) return off / P.sample_count;
		uint32_t q = 0u;
		while (off >= P.sample_count) { off -= P.sample_count; ++q; }
		return q;
	};
	auto flush_pixel = [&]() {
		const unsigned t = threadIdx.x;
		if (s_pix[t] != 0xffffffffu) {
			red_add_v4(P.accum + s_pix[t], s_sum[0][t], s_sum[1][t], s_sum[2][t], s_sum[3][t]);
			if (kAdaptive) { atomicAdd(&P.lum[s_pix[t]].x, s_lum[0][t]); atomicAdd(&P.lum[s_pix[t]].y, s_lum[1][t]); }
		}
	};
	auto finish_path = [&]() {  // pixel_color += sample_color, camera.rs:141
		const unsigned t = threadIdx.x;
		float ar = st.L.x, ag = st.L.y, ab = st.L.z, an = 1.f;
		double lum = 0.0, lum_sq = 0.0;
		if (kAdaptive) {  // color::luminance, color.rs:31-33 (0.144 sic)
			lum = 0.299 * (double)st.L.x + 0.587 * (double)st.L.y + 0.144 * (double)st.L.z;
			lum_sq = lum * lum;
		}
		if (st.pixel() != s_pix[t]) {
			flush_pixel();
			s_pix[t] = st.pixel();
		} else {
			ar += s_sum[0][t]; ag += s_sum[1][t]; ab += s_sum[2][t]; an += s_sum[3][t];
			if (kAdaptive) { lum += s_lum[0][t]; lum_sq += s_lum[1][t]; }
		}
		s_sum[0][t] = ar; s_sum[1][t] = ag; s_sum[2][t] = ab; s_sum[3][t] = an;
		if (kAdaptive) { s_lum[0][t] = lum; s_lum[1][t] = lum_sq; }
		state = kDead;
	};

	for (;;) {  // rounds of the adaptive sampler (the fixed-count kernels leave after the first)
	for (;;) {
		// ---- phase A1: shade the lanes whose traversal finished --------------------------------------
		bool finished = false;
		int noise_pend = -1;
		if (state == kShade) {
			// the RNG key is a kernel parameter: re-reading it here ends the live range of two registers before phase B
			st.rng.k0 = (uint32_t)P.seed; st.rng.k1 = (uint32_t)(P.seed >> 32);
			// (tv.best.t doubles as the medium's event distance: it starts as the medium candidate and only a closer
			// surface hit lowers it)
			finished = segment_shade<(kF & 2) != 0, kShadeTune>(P.sc, P.cam, s_perm, st, pc, tv.best, tv.best.t, seg_medium, &noise_pend);
			if (!(kF & 1)) pc.tex_fetches = 0u;  // texel fetches are a statistic of the counting variant only (one register less here)
		}
		// NoiseTextures met in this phase: evaluated here by the whole warp, one pending lane after the other (gs_shade.cuh)
		if (kShadeTune & 4) noise_apply_deferred<(kShadeTune & 1) != 0>(P.sc, s_perm, st, noise_pend, lane);
		if (state == kShade) {
			if (finished) finish_path();
			else state = kSetup;
		}
		// ---- phase A2: dead lanes take the next (pixel, sample) items ------------------------------------
		const unsigned dead = __ballot_sync(kFull, state == kDead);
		if (dead) {
			unsigned long long warp_next = chunk[0], warp_end = chunk[1];
			bool exhausted = chunk[2] != 0ull;
			if (!(exhausted && warp_next >= warp_end)) {
				// item = pixel * sample_count + sample.  The 64-bit division runs once per CHUNK (warp-uniform); within a chunk
				// the (pixel, sample) of the next item is carried along and a lane adds its rank with one 32-bit division
				// (sample_count <= 2^31, checked by the ABI) -- the per-lane 64-bit div/mod pair was ~150 instructions at
				// seven active lanes.
				uint32_t pixel0 = (uint32_t)(chunk[3] >> 32), sample0 = (uint32_t)chunk[3];
				if (warp_next >= warp_end) {
					unsigned long long base = 0;
					if (lane == 0) base = atomicAdd(kAdaptive ? s_rnd.work_counter : P.work_counter, (unsigned long long)P.chunk);
					base = __shfl_sync(kFull, base, 0);
					const unsigned long long total_items = kAdaptive ? s_rnd.total_items : P.total_items;
					if (base >= total_items) exhausted = true;
					else {
						warp_next = base; warp_end = min(base + (unsigned long long)P.chunk, total_items);
						pixel0 = (uint32_t)(base / P.sample_count);
						sample0 = (uint32_t)(base - (unsigned long long)pixel0 * P.sample_count);
					}
				}
				const unsigned long long avail = warp_end - warp_next;
				if (state == kDead) {
					const unsigned rank = __popc(dead & lt_mask);
					if (rank < avail) {
						const uint32_t off = sample0 + rank, q = quot(off);
						uint32_t pixel = pixel0 + q;
						const uint32_t sample = off - q * P.sample_count + (kAdaptive ? s_rnd.round * P.sample_count : P.sample_begin);
						// the lists are rewritten every round by other SMs; the gpu-scope fences around the grid syncs invalidate this SM's
						// L1 (CCTL.IVALL), so an ordinary cached load sees the new list -- a warp's lanes share a pixel, i.e. one line.
						// (Reading it from L2 with ld.cg cost every regeneration an L2 round trip: -15 % on the final scene.)
						if (kAdaptive) pixel = s_rnd.pixel_list[pixel];
						path_begin<tune_of(kF)>(st, P.cam, P.seed, pixel, sample);
						state = kSetup;
					}
				}
				const unsigned long long need = (unsigned long long)__popc(dead);
				const uint32_t adv = (uint32_t)(need < avail ? need : avail);
				warp_next += adv;
				{
					const uint32_t off = sample0 + adv, q = quot(off);
					pixel0 += q; sample0 = off - q * P.sample_count;
				}
				__syncwarp();
				if (lane == 0) { chunk[0] = warp_next; chunk[1] = warp_end; chunk[2] = exhausted ? 1ull : 0ull; chunk[3] = ((unsigned long long)pixel0 << 32) | sample0; }
				__syncwarp();
			}
		}
		// ---- phase A3: depth test, media, traversal set-up for every fresh ray ----------------------------
		if (state == kSetup) {
			double seg_tmax;
			if (segment_begin(P.sc, P.cam, st, pc, seg_tmax, seg_medium)) {
				tv.init(P.sc, st.ray, stack, seg_tmax);
				state = kTraverse;
			} else finish_path();  // depth exhausted; the lane picks up new work in the next phase A
		}
		const unsigned trav0 = __ballot_sync(kFull, state == kTraverse);
		if (trav0 == 0u) {
			if (chunk[2] != 0ull && chunk[0] >= chunk[1] && __ballot_sync(kFull, state != kDead) == 0u) break;
			continue;
		}
		// ---- phase B: traverse until enough lanes have finished their ray ---------------------------------
		RayConsts rc;
		if (state == kTraverse) rc = ray_consts(st.ray, tv.best.t);
#ifdef GS_DIVERGENCE_STATS
		// Diagnostic build (scripts/build_variants.sh ... -DGS_DIVERGENCE_STATS): the same while-while policy written
		// as a converged loop so that every warp-level iteration can record what the 32 lanes are doing.  Totals go to
		// stats[4..]: node iterations, lanes stepping / waiting on a leaf / finished or without a ray in them;
		// leaf rounds, lanes intersecting spheres / quads in them, lanes finished.
		{
			unsigned long long c_iter = 0, c_step = 0, c_wait = 0, c_done = 0, l_round = 0, l_sph = 0, l_quad = 0, l_done = 0;
			const bool trav = state == kTraverse;
			for (;;) {
				for (;;) {
					const bool step = trav && tv.wants_node();
					const unsigned ms = __ballot_sync(kFull, step);
					if (ms == 0u) break;
					const unsigned mw = __ballot_sync(kFull, trav && tv.wants_leaf());
					c_iter += 1; c_step += __popc(ms); c_wait += __popc(mw); c_done += 32 - __popc(ms) - __popc(mw);
					if (step) tv.template node_step<(kF & 1) != 0, kTopNodes>(P.sc, rc, stack, &pc.node_visits, s_top);
				}
				const bool leaf = trav && tv.wants_leaf();
				const unsigned ml = __ballot_sync(kFull, leaf);
				if (ml == 0u) break;
				const unsigned kind = leaf ? tv.pending_leaf_kind() : 9u;
				const unsigned m0 = __ballot_sync(kFull, kind == 0u);
				l_round += 1; l_sph += __popc(m0); l_quad += __popc(ml) - __popc(m0); l_done += 32 - __popc(ml);
				if (leaf) tv.template leaf_step<(kF & 2) != 0>(P.sc, st.ray, rc, 0.001, stack);
			}
			if (trav) state = kShade;
			if (lane == 0) {
				atomicAdd(P.stats + 4, c_iter); atomicAdd(P.stats + 5, c_step); atomicAdd(P.stats + 6, c_wait); atomicAdd(P.stats + 7, c_done);
				atomicAdd(P.stats + 8, l_round); atomicAdd(P.stats + 9, l_sph); atomicAdd(P.stats + 10, l_quad); atomicAdd(P.stats + 11, l_done);
				atomicAdd(P.stats + 12, 1ull); atomicAdd(P.stats + 13, (unsigned long long)__popc(trav0));
			}
		}
#elif GS_MIN_SHADE >= 32
		// Every lane runs its ray to completion (classic while-while); measured best on B200.
		if (state == kTraverse) {
			while (!tv.done()) trav_round<kF, kTopNodes>(P.sc, st.ray, rc, 0.001, tv, stack, &pc.node_visits, s_top);
			state = kShade;
		}
#else
		// Converged, resumable traversal.  Measured lane occupancy of the classic loop on C4 (GS_DIVERGENCE_STATS build,
		// profiles/r01_v5_divergence.log): per node-loop iteration 8.2 lanes step, 4.8 wait on a leaf for the others and
		// 19.0 have FINISHED their ray -- ray lengths are close to exponential (mean 11 node steps, the longest of 32 ~44),
		// so the warp mostly waits for its longest ray.  Here all 32 lanes stay in one loop, two ballots per iteration
		// tell how many lanes step / wait on a leaf / are finished, and the warp
		//   * leaves for phase A as soon as kMinShade lanes are finished (the others keep node, sp, best, stack and
		//     resume in the next phase B),
		//   * runs the leaf code once kLeafT lanes wait on a leaf (or nobody can step).
		// (Lanes without a ray at entry -- only at the very end of the image -- do not count as finished: with fewer than
		// kMinShade rays in the warp, all of them run to completion.)
		constexpr int kExit = (kF & 16) ? GS_MIN_SHADE_WIDE : kMinShade;
		const int n_trav0 = __popc(trav0), exit_at = n_trav0 < kExit ? n_trav0 : kExit;
		const bool trav = state == kTraverse;
		for (;;) {
			// node phase: the classic tight per-lane loop, except that the lanes still descending give up once no more than
			// kBail of them are left (__activemask() = the lanes still inside this loop; the others wait at its exit), so
			// a round lasts as long as the kBail-th slowest lane needs, not the slowest.  The stragglers sit out the leaf
			// phase and carry on in the next round.  (A fully converged loop with two ballots per node step was measured
			// 10-15 % slower than the classic loop: 12 more instructions per step, and ptxas spilled the ray constants.)
			if (trav && tv.wants_node()) {
#if GS_NODE_UNROLL2
				// two node steps per look at the lanes left (the vote + count + branch are 4 of a step's 58 instructions)
				do {
					tv.template node_step<(kF & 1) != 0, kTopNodes>(P.sc, rc, stack, &pc.node_visits, s_top);
					if (!tv.wants_node()) break;
					tv.template node_step<(kF & 1) != 0, kTopNodes>(P.sc, rc, stack, &pc.node_visits, s_top);
				} while (tv.wants_node() && __popc(__activemask()) > kBail);
#else
				do tv.template node_step<(kF & 1) != 0, kTopNodes>(P.sc, rc, stack, &pc.node_visits, s_top);
				while (tv.wants_node() && __popc(__activemask()) > kBail);
#endif
			}
			__syncwarp();
			const bool leaf = trav && tv.wants_leaf();
			const unsigned m_live = __ballot_sync(kFull, trav && !tv.done());
			if (n_trav0 - __popc(m_live) >= exit_at) break;
			if (leaf) tv.template leaf_step<(kF & 2) != 0>(P.sc, st.ray, rc, 0.001, stack);
		}
		if (state == kTraverse && tv.done()) state = kShade;
#endif
	}
	flush_pixel();
	if (!kAdaptive) {
		if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.stats + 1, P.total_items);  // every work item starts exactly one path
		break;
	}
	if (!P.rounds_in_kernel) {  // host-driven rounds: the stop test is the next launch
		if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.stats + 1, s_rnd.total_items);
		break;
	}
	// ---- between two batches: Camera::sample's stop test (camera.rs:150-164) for every pixel of this round's list ----
	// (the whole grid is co-resident: cooperative launch, persistent CTAs)
	const unsigned cur = s_rnd.round & 1u, nxt = cur ^ 1u;
	const uint32_t* const pixel_list = s_rnd.pixel_list;
	s_pix[threadIdx.x] = 0xffffffffu;
	if (lane == 0) { chunk[0] = 0ull; chunk[1] = 0ull; chunk[2] = 0ull; chunk[3] = 0ull; }
	__threadfence();
	cooperative_groups::this_grid().sync();
	{
		const unsigned n = __ldcg(P.n_list + cur);
		uint32_t* const next = P.pixel_list + (size_t)nxt * P.n_pixels;
		for (unsigned base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
			const unsigned k = base + lane;
			bool more = false;
			uint32_t pixel = 0;
			if (k < n) {
				pixel = __ldcg(pixel_list + k);
				const double sample_count = (double)__ldcg(&P.accum[pixel].w);  // exact: one 1.0f per path, far below 2^24
				const double2 sl = __ldcg(P.lum + pixel);
				const double mean = sl.x / sample_count;  // camera.rs:150-158
				const double variance_sq = 1.0 / (sample_count - 1.0) * (sl.y - sl.x * sl.x / sample_count);
				const double convergence_sq = P.confidence_sq * variance_sq / sample_count;
				const bool converged = convergence_sq < mean * mean * P.tolerance_sq;
				const unsigned sc_u32 = sample_count >= 4294967295.0 ? 4294967295u : (unsigned)sample_count;
				more = !converged && !(sc_u32 > P.max_samples);  // camera.rs:162-164
			}
			const unsigned m = __ballot_sync(kFull, more);  // one atomic per warp; the order of the list does not matter
			if (m) {
				unsigned slot = 0;
				if (lane == 0) slot = atomicAdd(P.n_list + nxt, (unsigned)__popc(m));
				slot = __shfl_sync(kFull, slot, 0);
				if (more) next[slot + __popc(m & ((1u << lane) - 1u))] = pixel;
			}
		}
		if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(P.stats + 1, (unsigned long long)n * P.sample_count);
	}
	__threadfence();
	cooperative_groups::this_grid().sync();
	const unsigned left = __ldcg(P.n_list + nxt);
	if (left <= P.tail_pixels) {  // the last few pixels (or none): adaptive_tail_kernel takes list `nxt`
		if (blockIdx.x == 0 && threadIdx.x == 0) P.n_list[2] = nxt;
		break;
	}
	// next round: the other list, the other work counter; this round's are reset for the round after
	__syncthreads();  // everybody has read this round's s_rnd
	if (threadIdx.x == 0) {
		const unsigned round = s_rnd.round + 1u;
		s_rnd.round = round;
		s_rnd.pixel_list = P.pixel_list + (size_t)nxt * P.n_pixels;
		s_rnd.total_items = (unsigned long long)left * P.sample_count;
		s_rnd.work_counter = (round & 1u) ? P.work_counter2 : P.work_counter;
		if (blockIdx.x == 0) {
			P.n_list[cur] = 0u;  // becomes the output list of the coming round's stop test; nobody reads it before the next grid sync
			*((round & 1u) ? P.work_counter : P.work_counter2) = 0ull;  // last used a round ago
		}
	}
	__syncthreads();
	state = kDead;
	tv.park();
	}
	flush_stats(P.stats, pc, 0u);
}

// ---- the adaptive sampler's control flow, camera.rs:125-171 ------------------------------------------------------
// Camera::sample loops per pixel: take batch_size samples, test the luminance statistics, stop or take another batch.
// On the device the loop is turned inside out: ONE launch of the fixed-count megakernel per batch over the list of
// pixels still sampling (same (pixel, global sample index) Philox streams, so every pixel sees exactly the samples the
// per-pixel loop would give it), then this kernel applies the reference's stop test to each of them and writes the
// list for the next batch.  The per-pixel state is the accumulator (sum rgb, n) plus (sum lum, sum lum^2) in f64.
__global__ void adaptive_init_kernel(uint32_t* __restrict__ list, unsigned* __restrict__ n_list, unsigned pixel_begin, unsigned pixel_stride,
                                     unsigned n_pixels) {
	const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k < n_pixels) list[k] = pixel_begin + k * pixel_stride;
	if (k == 0) { n_list[0] = n_pixels; n_list[1] = 0u; n_list[2] = 0u; }
}

// The stop test as a kernel of its own: the host-driven alternative to the in-kernel rounds (one render launch + this per batch;
// GS_ADAPTIVE_ROUNDS=host), kept for A/B measurements.
__global__ void adaptive_decide_kernel(const uint32_t* __restrict__ list, const unsigned* __restrict__ n_in, uint32_t* __restrict__ next,
                                       unsigned* __restrict__ n_out, const float4* __restrict__ accum, const double2* __restrict__ lum,
                                       double confidence_sq, double tolerance_sq, unsigned max_samples) {
	const unsigned n = *n_in;
	const unsigned lane = threadIdx.x & 31u;
	for (unsigned base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
		const unsigned k = base + lane;
		bool more = false;
		uint32_t pixel = 0;
		if (k < n) {
			pixel = list[k];
			const double sample_count = (double)accum[pixel].w;  // exact: one 1.0f per path, far below 2^24
			const double2 s = lum[pixel];
			const double mean = s.x / sample_count;  // camera.rs:150-158
			const double variance_sq = 1.0 / (sample_count - 1.0) * (s.y - s.x * s.x / sample_count);
			const double convergence_sq = confidence_sq * variance_sq / sample_count;
			const bool converged = convergence_sq < mean * mean * tolerance_sq;
			const unsigned sc_u32 = sample_count >= 4294967295.0 ? 4294967295u : (unsigned)sample_count;
			more = !converged && !(sc_u32 > max_samples);  // camera.rs:162-164
		}
		const unsigned m = __ballot_sync(kFull, more);  // one atomic per warp; the order of the list does not matter
		if (m) {
			unsigned slot = 0;
			if (lane == 0) slot = atomicAdd(n_out, (unsigned)__popc(m));
			slot = __shfl_sync(kFull, slot, 0);
			if (more) next[slot + __popc(m & ((1u << lane) - 1u))] = pixel;
		}
	}
}

// The tail of an adaptive render: when only a few pixels are still sampling, a launch per batch costs more than it
// renders (a batch of 32 samples is one warp's worth of work and a whole launch's worth of latency).  Then one warp
// takes each remaining pixel and runs the rest of its loop (camera.rs:135-165) on its own, starting from the sums the
// batch rounds left in accum / lum: the batch's samples spread over the lanes, regeneration inside the batch.
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
	return v;
}

template <int kF>  // 2 = binary tree, 18 = eight-wide tree (both with the triangle branch: this kernel is not the hot one)
__global__ void __launch_bounds__(kBlock) adaptive_tail_kernel(const __grid_constant__ AdaptiveTailParams P) {
	__shared__ __align__(16) uint8_t s_perm[kPermBytes];
	load_perm(s_perm, P.sc.perlin_perm);
	const unsigned lane = threadIdx.x & 31u;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned which = P.n_list[2] & 1u;
	const unsigned n_list = P.n_list[which];
	const uint32_t* const list = P.lists + (size_t)which * P.n_pixels;
	PathCounters pc = {0u, 0u, 0u};
	unsigned paths = 0;
	for (;;) {
		unsigned long long k = 0;
		if (lane == 0) k = atomicAdd(P.work_counter, 1ull);
		k = __shfl_sync(kFull, k, 0);
		if (k >= n_list) break;
		const uint32_t pixel = list[k];
		const float4 a0 = P.accum[pixel];
		const double2 l0 = P.lum[pixel];
		double sum = 0.0, sq_sum = 0.0, cr = 0.0, cg = 0.0, cb = 0.0;  // this lane's share of what is added now
		double sample_count = (double)a0.w;
		unsigned next = (unsigned)a0.w;  // global sample index of the next sample of this pixel
		for (;;) {
			sample_count += (double)P.batch_size;  // camera.rs:137
			const unsigned end = next + P.batch_size;
			PathState st;
			bool alive = false;
			for (;;) {
				const unsigned dead = __ballot_sync(kFull, !alive);
				if (dead) {
					const unsigned avail = end - next;
					if (!alive) {
						const unsigned rank = __popc(dead & lt_mask);
						if (rank < avail) { path_begin(st, P.cam, P.seed, pixel, next + rank); alive = true; ++paths; }
					}
					const unsigned need = __popc(dead);
					next += need < avail ? need : avail;
				}
				if (!__any_sync(kFull, alive)) break;
				if (alive && path_step<kF>(P.sc, P.cam, s_perm, st, pc)) {
					cr += (double)st.L.x; cg += (double)st.L.y; cb += (double)st.L.z;
					const double lum = 0.299 * (double)st.L.x + 0.587 * (double)st.L.y + 0.144 * (double)st.L.z;  // color.rs:31-33
					sum += lum; sq_sum += lum * lum;
					alive = false;
				}
			}
			const double tsum = l0.x + warp_sum(sum), tsq = l0.y + warp_sum(sq_sum);
			const double mean = tsum / sample_count;  // camera.rs:150-158
			const double variance_sq = 1.0 / (sample_count - 1.0) * (tsq - tsum * tsum / sample_count);
			const double convergence_sq = P.confidence_sq * variance_sq / sample_count;
			if (convergence_sq < mean * mean * P.tolerance_sq) break;
			const unsigned sc_u32 = sample_count >= 4294967295.0 ? 4294967295u : (unsigned)sample_count;
			if (sc_u32 > P.max_samples) break;  // camera.rs:162-164
		}
		const double tr = warp_sum(cr), tg = warp_sum(cg), tb = warp_sum(cb);
		if (lane == 0) P.accum[pixel] = make_float4(a0.x + (float)tr, a0.y + (float)tg, a0.z + (float)tb, (float)sample_count);
	}
	flush_stats(P.stats, pc, paths);
}

// The same loop when only a few hundred pixels are left and the GPU is mostly idle: a pixel's batches are sequential by
// definition and a batch lasts as long as its longest path (up to max_depth bounces), so a pixel that needs thirty more
// batches holds the whole render for thirty path latencies (measured: 5.0 of scene 3's 7.0 ms, 1.2 of scene 1's 2.4).
// Here a CTA of kWarps warps takes one pixel and renders its next kWarps batches AT ONCE -- batch b is the global sample
// indices [b * batch_size, (b + 1) * batch_size) whoever renders it -- and then replays Camera::sample's loop over them in
// order (camera.rs:137-164): the batches up to the one whose stop test fires are added, the rest are dropped, so the pixel
// ends with exactly the samples the sequential loop gives it.  The dropped batches are work the reference would not have
// done; they are not counted in the statistics (paths, rays, node visits are those of the batches kept).
template <int kF, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) adaptive_tail_ahead_kernel(const __grid_constant__ AdaptiveTailParams P) {
	__shared__ __align__(16) uint8_t s_perm[kPermBytes];
	__shared__ double s_part[kWarps][5];  // per batch: sum lum, sum lum^2, r, g, b
	__shared__ unsigned long long s_k;
	load_perm(s_perm, P.sc.perlin_perm);
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const unsigned lt_mask = (1u << lane) - 1u;
	const unsigned which = P.n_list[2] & 1u;
	const unsigned n_list = P.n_list[which];
	const uint32_t* const list = P.lists + (size_t)which * P.n_pixels;
	PathCounters pc = {0u, 0u, 0u};
	unsigned paths = 0;
	for (;;) {
		if (threadIdx.x == 0) s_k = atomicAdd(P.work_counter, 1ull);
		__syncthreads();
		const unsigned long long k = s_k;
		__syncthreads();  // everybody has read s_k before it is written again
		if (k >= n_list) break;
		const uint32_t pixel = list[k];
		const float4 a0 = P.accum[pixel];
		const double2 l0 = P.lum[pixel];
		double sample_count = (double)a0.w, tsum = l0.x, tsq = l0.y, tr = 0.0, tg = 0.0, tb = 0.0;  // the same on every thread
		unsigned base = (unsigned)a0.w;  // global sample index of the pixel's next sample
		for (;;) {
			// this warp's batch
			unsigned next = base + warp * P.batch_size;
			const unsigned end = next + P.batch_size;
			double sum = 0.0, sq_sum = 0.0, cr = 0.0, cg = 0.0, cb = 0.0;
			PathCounters pcb = {0u, 0u, 0u};
			unsigned started = 0;
			PathState st;
			bool alive = false;
			for (;;) {
				const unsigned dead = __ballot_sync(kFull, !alive);
				if (dead) {
					const unsigned avail = end - next;
					if (!alive) {
						const unsigned rank = __popc(dead & lt_mask);
						if (rank < avail) { path_begin(st, P.cam, P.seed, pixel, next + rank); alive = true; ++started; }
					}
					const unsigned need = __popc(dead);
					next += need < avail ? need : avail;
				}
				if (!__any_sync(kFull, alive)) break;
				if (alive && path_step<kF>(P.sc, P.cam, s_perm, st, pcb)) {
					cr += (double)st.L.x; cg += (double)st.L.y; cb += (double)st.L.z;
					const double lum = 0.299 * (double)st.L.x + 0.587 * (double)st.L.y + 0.144 * (double)st.L.z;  // color.rs:31-33
					sum += lum; sq_sum += lum * lum;
					alive = false;
				}
			}
			const double w0 = warp_sum(sum), w1 = warp_sum(sq_sum), w2 = warp_sum(cr), w3 = warp_sum(cg), w4 = warp_sum(cb);
			if (lane == 0) { s_part[warp][0] = w0; s_part[warp][1] = w1; s_part[warp][2] = w2; s_part[warp][3] = w3; s_part[warp][4] = w4; }
			__syncthreads();
			bool stop = false;
			unsigned kept = 0;
			for (int w = 0; w < kWarps && !stop; ++w) {
				sample_count += (double)P.batch_size;  // camera.rs:137
				tsum += s_part[w][0]; tsq += s_part[w][1]; tr += s_part[w][2]; tg += s_part[w][3]; tb += s_part[w][4];
				kept = (unsigned)w + 1u;
				const double mean = tsum / sample_count;  // camera.rs:150-158
				const double variance_sq = 1.0 / (sample_count - 1.0) * (tsq - tsum * tsum / sample_count);
				const double convergence_sq = P.confidence_sq * variance_sq / sample_count;
				if (convergence_sq < mean * mean * P.tolerance_sq) stop = true;
				const unsigned sc_u32 = sample_count >= 4294967295.0 ? 4294967295u : (unsigned)sample_count;
				if (sc_u32 > P.max_samples) stop = true;  // camera.rs:162-164
			}
			if (warp < kept) { pc.rays += pcb.rays; pc.tex_fetches += pcb.tex_fetches; pc.node_visits += pcb.node_visits; paths += started; }
			base += (unsigned)kWarps * P.batch_size;
			__syncthreads();  // s_part is rewritten by the next round
			if (stop) break;
		}
		if (threadIdx.x == 0) P.accum[pixel] = make_float4(a0.x + (float)tr, a0.y + (float)tg, a0.z + (float)tb, (float)sample_count);
	}
	flush_stats(P.stats, pc, paths);
}

// pixel_color /= sample_count (camera.rs:167)
__global__ void resolve_kernel(const float4* __restrict__ accum, int n_pixels, float* __restrict__ out_rgb, unsigned* __restrict__ out_counts) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	float4 a = accum[i];
	float inv = a.w > 0.f ? 1.0f / a.w : 0.f;
	out_rgb[i * 3 + 0] = a.x * inv; out_rgb[i * 3 + 1] = a.y * inv; out_rgb[i * 3 + 2] = a.z * inv;
	if (out_counts) out_counts[i] = (unsigned)a.w;
}

// the same fused with write_color's transfer (color.rs:8-26): the bytes of the reference's image.ppm
__global__ void resolve_rgb8_kernel(const float4* __restrict__ accum, int n_pixels, uchar3* __restrict__ out_rgb8, unsigned* __restrict__ out_counts) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	float4 a = accum[i];
	float inv = a.w > 0.f ? 1.0f / a.w : 0.f;
	const float c[3] = {a.x * inv, a.y * inv, a.z * inv};
	unsigned char q[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		double g = c[k] > 0.0f ? sqrt((double)c[k]) : 0.0;  // linear_to_gamma
		g = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);          // INTENSITY.clamp
		q[k] = (unsigned char)(int)(256.0 * g);
	}
	out_rgb8[i] = make_uchar3(q[0], q[1], q[2]);
	if (out_counts) out_counts[i] = (unsigned)a.w;
}

// Cross-GPU sum fused with the resolve: acc[0] is local, acc[1..] are the peers' accumulators read over NVLink (one coalesced
// 128-bit load per pixel and device); optional outputs: linear mean radiance, the PPM bytes of write_color, counts, the summed
// accumulator itself.
__global__ void reduce_resolve_kernel(const __grid_constant__ PeerAccums P, int n_pixels, float* __restrict__ out_rgb, uchar3* __restrict__ out_rgb8,
                                      unsigned* __restrict__ out_counts, float4* __restrict__ sum_out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pixels) return;
	float4 a = P.acc[0][i];
	for (int d = 1; d < P.n; ++d) {
		const float4 b = P.acc[d][i];
		a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
	}
	if (sum_out) sum_out[i] = a;
	const float inv = a.w > 0.f ? 1.0f / a.w : 0.f;
	const float c[3] = {a.x * inv, a.y * inv, a.z * inv};
	if (out_rgb) { out_rgb[i * 3 + 0] = c[0]; out_rgb[i * 3 + 1] = c[1]; out_rgb[i * 3 + 2] = c[2]; }
	if (out_rgb8) {
		unsigned char q[3];
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			double g = c[k] > 0.0f ? sqrt((double)c[k]) : 0.0;  // linear_to_gamma
			g = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);          // INTENSITY.clamp
			q[k] = (unsigned char)(int)(256.0 * g);
		}
		out_rgb8[i] = make_uchar3(q[0], q[1], q[2]);
	}
	if (out_counts) out_counts[i] = (unsigned)a.w;
}

__global__ void camera_rays_kernel(const __grid_constant__ DCamera cam, unsigned long long seed, unsigned sample, int pixel_begin, int pixel_end,
                                   double* __restrict__ rays_out) {
	int n = pixel_begin + blockIdx.x * blockDim.x + threadIdx.x;
	if (n >= pixel_end) return;
	PathRng rng;
	rng.start(seed, (uint32_t)n, sample);
	Ray r = get_ray<false>(cam, n % cam.width, n / cam.width, rng);
	double* o = rays_out + (size_t)(n - pixel_begin) * 7;
	o[0] = r.ox; o[1] = r.oy; o[2] = r.oz; o[3] = r.dx; o[4] = r.dy; o[5] = r.dz; o[6] = r.time;
}

template <int kF>
__global__ void __launch_bounds__(kBlock) path_radiance_kernel(const __grid_constant__ DeviceScene sc, const __grid_constant__ DCamera cam,
                                                               unsigned long long seed, const unsigned* __restrict__ pixel,
                                                               const unsigned* __restrict__ sample, size_t n, float* __restrict__ out_rgb) {
	__shared__ __align__(16) uint8_t s_perm[kPermBytes];
	load_perm(s_perm, sc.perlin_perm);
	size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n) return;
	PathCounters pc = {0u, 0u, 0u};
	PathState st;
	path_begin(st, cam, seed, pixel[k], sample[k]);
	while (!path_step<kF>(sc, cam, s_perm, st, pc)) {}
	out_rgb[k * 3 + 0] = st.L.x; out_rgb[k * 3 + 1] = st.L.y; out_rgb[k * 3 + 2] = st.L.z;
}

// ---- launch wrappers ------------------------------------------------------------------------------
// kernel variant: bit 0 = count node visits, bit 1 = triangles, bit 2 = adaptive batch (never combined with bit 0),
// bit 3 = small scene (never combined with bit 0 or bit 4), bit 4 = eight-wide tree
#define GS_FIXED_VARIANTS(X) X(0) X(1) X(2) X(3) X(4) X(6) X(8) X(10) X(12) X(14) X(16) X(17) X(18) X(19) X(20) X(22)

static int fixed_variant(const RenderParams& p, bool count_nodes, bool triangles) {
	const int wide = p.wide ? 16 : 0, small = (p.small_scene && !p.wide) ? 8 : 0;
	if (p.pixel_list) return 4 | (triangles ? 2 : 0) | small | wide;
	if (count_nodes) return 1 | (triangles ? 2 : 0) | wide;
	return (triangles ? 2 : 0) | small | wide;
}

// Shared-memory carve-out of a render kernel: exactly what its resident CTAs need, so that the rest of the SM's 256 KB is L1.
// Left alone the driver reserves more (the hint that reproduces its choice is 12-15 % of the maximum; the 5 CTAs x (2.9 KB + 1 KB
// reserved) of the config-4 kernel are 8.5 %): config 4, whose 0.5 MB of nodes and primitives do not fit the L1, gains 1.3 %
// (1701 vs 1680 M paths/s, profiles/r02_v15_carveout.log); scenes that fit are unchanged.  GS_CARVEOUT (percent) overrides.
template <class K>
static void set_carveout(K kernel, int blocks_per_sm) {
	static const int forced = [] { const char* e = std::getenv("GS_CARVEOUT"); return e ? std::atoi(e) : -1; }();
	int dev = 0, max_smem = 0, reserved = 0;
	cudaFuncAttributes fa;
	if (cudaGetDevice(&dev) != cudaSuccess || cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) { cudaGetLastError(); return; }
	cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
	cudaDeviceGetAttribute(&reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev);
	if (max_smem <= 0) return;
	const long long need = (long long)blocks_per_sm * ((long long)fa.sharedSizeBytes + reserved);
	int pct = (int)((need * 100 + max_smem - 1) / max_smem);
	if (pct > 100) pct = 100;
	cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, forced >= 0 ? forced : pct);
}

static int fixed_blocks_per_sm(int variant) {
	constexpr int kMaxDev = 64;
	static std::atomic<int> cached[kMaxDev][32];  // blocks per SM of (device, variant), 0 = not asked yet; the carve-out is set on the first ask
	int dev = 0;
	cudaGetDevice(&dev);
	const bool cache = dev >= 0 && dev < kMaxDev && variant >= 0 && variant < 32;
	if (cache) { const int c = cached[dev][variant].load(std::memory_order_acquire); if (c > 0) return c; }
	int n = 0;
	switch (variant) {
#define X(v) case v: \
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, render_fixed_kernel<v>, kBlock, 0); \
		if (n > 0) { set_carveout(render_fixed_kernel<v>, n); int m = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, render_fixed_kernel<v>, kBlock, 0); if (m > 0 && m < n) n = m; } \
		break;
		GS_FIXED_VARIANTS(X)
#undef X
		default: break;
	}
	if (n <= 0) n = 1;
	if (cache) cached[dev][variant].store(n, std::memory_order_release);
	return n;
}

int render_fixed_blocks_per_sm(bool count_nodes, bool triangles) { return fixed_blocks_per_sm((count_nodes ? 1 : 0) | (triangles ? 2 : 0)); }

cudaError_t launch_render_fixed(const RenderParams& p_in, int sm_count, bool count_nodes, bool triangles, cudaStream_t stream) {
	RenderParams p = p_in;
	philox_round_keys(p.seed, p.philox_rk);
	const int variant = fixed_variant(p, count_nodes, triangles);
	const int per_sm = fixed_blocks_per_sm(variant);
	unsigned long long warps_needed = (p.total_items + 31ull) / 32ull;  // adaptive: total_items = the upper bound (every listed pixel)
	unsigned long long blocks_needed = (warps_needed + (kBlock / 32) - 1) / (kBlock / 32);
	unsigned long long grid = (unsigned long long)sm_count * per_sm;  // persistent: SMs x resident CTAs
	if (blocks_needed < grid) grid = blocks_needed ? blocks_needed : 1;
	// the adaptive variants run every batch round in this one launch and synchronise the grid in between: cooperative launch
	// (all CTAs co-resident -- the grid is sized to the occupancy anyway)
	void* args[] = {const_cast<RenderParams*>(&p)};
	switch (variant) {
#define X(v) case v: \
		if (((v) & 4) && p.rounds_in_kernel) return cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&render_fixed_kernel<v>), dim3((unsigned)grid), dim3(kBlock), args, 0, stream); \
		render_fixed_kernel<v><<<(unsigned)grid, kBlock, 0, stream>>>(p); break;
		GS_FIXED_VARIANTS(X)
#undef X
		default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

cudaError_t launch_adaptive_init(uint32_t* list, unsigned* n_list, unsigned pixel_begin, unsigned pixel_stride, unsigned n_pixels, cudaStream_t stream) {
	adaptive_init_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(list, n_list, pixel_begin, pixel_stride, n_pixels);
	return cudaGetLastError();
}

// n_known: the length of the list if the host has read it (host-driven rounds), else 0.  Few pixels on an idle GPU:
// several batches per pixel at once (adaptive_tail_ahead_kernel), as many as fit beside each other.
cudaError_t launch_adaptive_tail(const AdaptiveTailParams& p, unsigned n_upper, unsigned n_known, int sm_count, bool wide, cudaStream_t stream) {
	// tuning knobs: up to how many pixels a CTA of 4 / of 8 warps per pixel is used (0 = never).  Measured on the reference's
	// scenes 1 and 3 (2450 pixels left after the first batch; profiles/r02_v8_adaptive_tail_ahead.log): one warp per pixel 2.39 /
	// 7.02 ms for the whole render, four 1.83 / 3.66, eight 2.29 / 3.73.
	static const unsigned ahead4 = [] { const char* e = std::getenv("GS_TAIL_AHEAD4"); return e ? (unsigned)std::atoi(e) : 8192u; }();
	static const unsigned ahead8 = [] { const char* e = std::getenv("GS_TAIL_AHEAD8"); return e ? (unsigned)std::atoi(e) : 512u; }();
	(void)sm_count;
	if (n_known > 0 && (n_known <= ahead4 || n_known <= ahead8)) {
		const bool eight = n_known <= ahead8;
		const unsigned grid = n_known;
		if (eight) {
			if (wide) adaptive_tail_ahead_kernel<18, 8><<<grid, 256, 0, stream>>>(p);
			else adaptive_tail_ahead_kernel<2, 8><<<grid, 256, 0, stream>>>(p);
		} else {
			if (wide) adaptive_tail_ahead_kernel<18, 4><<<grid, 128, 0, stream>>>(p);
			else adaptive_tail_ahead_kernel<2, 4><<<grid, 128, 0, stream>>>(p);
		}
		return cudaGetLastError();
	}
	int per_sm = 0;
	if (wide) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_tail_kernel<18>, kBlock, 0);
	else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adaptive_tail_kernel<2>, kBlock, 0);
	if (per_sm < 1) per_sm = 1;
	unsigned long long blocks_needed = ((unsigned long long)n_upper + (kBlock / 32) - 1) / (kBlock / 32);
	unsigned long long grid = (unsigned long long)sm_count * per_sm;
	if (blocks_needed < grid) grid = blocks_needed ? blocks_needed : 1;
	if (wide) adaptive_tail_kernel<18><<<(unsigned)grid, kBlock, 0, stream>>>(p);
	else adaptive_tail_kernel<2><<<(unsigned)grid, kBlock, 0, stream>>>(p);
	return cudaGetLastError();
}

cudaError_t launch_adaptive_decide(const RenderParams& p, unsigned cur, cudaStream_t stream) {
	const unsigned blocks = (p.n_pixels + 255) / 256;
	adaptive_decide_kernel<<<blocks < 1184u ? (blocks ? blocks : 1u) : 1184u, 256, 0, stream>>>(
		p.pixel_list + (size_t)cur * p.n_pixels, p.n_list + cur, p.pixel_list + (size_t)(cur ^ 1u) * p.n_pixels, p.n_list + (cur ^ 1u), p.accum, p.lum,
		p.confidence_sq, p.tolerance_sq, p.max_samples);
	return cudaGetLastError();
}

cudaError_t launch_resolve(const float4* accum, int n_pixels, float* out_rgb, unsigned* out_counts, cudaStream_t stream) {
	resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(accum, n_pixels, out_rgb, out_counts);
	return cudaGetLastError();
}

cudaError_t launch_resolve_rgb8(const float4* accum, int n_pixels, unsigned char* out_rgb8, unsigned* out_counts, cudaStream_t stream) {
	resolve_rgb8_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(accum, n_pixels, reinterpret_cast<uchar3*>(out_rgb8), out_counts);
	return cudaGetLastError();
}

cudaError_t launch_reduce_resolve(const PeerAccums& peers, int n_pixels, float* out_rgb, unsigned char* out_rgb8, unsigned* out_counts,
                                  float4* sum_out, cudaStream_t stream) {
	reduce_resolve_kernel<<<(n_pixels + 255) / 256, 256, 0, stream>>>(peers, n_pixels, out_rgb, reinterpret_cast<uchar3*>(out_rgb8), out_counts, sum_out);
	return cudaGetLastError();
}

cudaError_t launch_camera_rays(const DCamera& cam, unsigned long long seed, unsigned sample, int pixel_begin, int pixel_end, double* rays_out,
                               cudaStream_t stream) {
	int n = pixel_end - pixel_begin;
	if (n <= 0) return cudaSuccess;
	camera_rays_kernel<<<(n + 127) / 128, 128, 0, stream>>>(cam, seed, sample, pixel_begin, pixel_end, rays_out);
	return cudaGetLastError();
}

cudaError_t launch_path_radiance(const DeviceScene& sc, const DCamera& cam, unsigned long long seed, const unsigned* pixel,
                                 const unsigned* sample, size_t n, float* out_rgb, bool wide, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
	if (wide) path_radiance_kernel<18><<<grid, kBlock, 0, stream>>>(sc, cam, seed, pixel, sample, n, out_rgb);
	else path_radiance_kernel<2><<<grid, kBlock, 0, stream>>>(sc, cam, seed, pixel, sample, n, out_rgb);
	return cudaGetLastError();
}

}  // namespace gs

// The closest-hit batch kernel once more, with THIS translation unit's arithmetic (what the render kernels really run:
// FMA contraction, approximate-plus-Newton f64 division / square root): gs_trace_closest_fast, the parity hook that shows
// the timed code picks the same primitives as the contraction-free build.
#define GS_TRACE_NAME trace_closest_fast
#include "trace_closest_body.cuh"
