// grayshift_b200/csrc/gs_host.h -- host-side internals shared by abi.cu (entry points, launches), scene_layout.cu
// (validation, re-layout of the flat scene description, the two trees, upload) and multi_gpu.cu.  Not part of the ABI.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/grayshift.h"
#include "gs_internal.h"

namespace gs {

int fail(int code, const std::string& msg);
#define GS_CUDA(call)                                                                                          \
	do {                                                                                                       \
		cudaError_t e_ = (call);                                                                               \
		if (e_ != cudaSuccess) return ::gs::fail(GS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

// Independent iterations over the host threads (scene creation of a 10^6-primitive scene spends its time in such loops);
// small ranges stay on the calling thread.
unsigned host_threads();
// (a process-wide pool of worker threads that sleep between loops: a scene creation runs some thirty of these loops, and
// starting and joining sixteen threads for each was a tenth of its time).  pool_run hands out [0, n) in pieces of `grain`;
// the caller takes part; a loop started while the pool is busy -- from another host thread, or from inside a loop body --
// simply runs on its own thread.
void pool_run(size_t n, size_t grain, void (*fn)(void* ctx, size_t lo, size_t hi), void* ctx);
template <class F> void parallel_for(size_t n, F&& body, size_t serial_below = 32768) {
	const unsigned threads = host_threads();
	if (n < serial_below || threads < 2) { for (size_t i = 0; i < n; ++i) body(i); return; }
	using Body = typename std::remove_reference<F>::type;
	size_t grain = n / ((size_t)threads * 8);
	if (grain < serial_below / 16) grain = serial_below / 16;
	if (grain < 1) grain = 1;
	pool_run(n, grain, [](void* ctx, size_t lo, size_t hi) { Body& b = *static_cast<Body*>(ctx); for (size_t i = lo; i < hi; ++i) b(i); },
	         const_cast<void*>(static_cast<const void*>(&body)));
}

// Device memory comes from the device's stream-ordered pool with the release threshold lifted, so creating and
// destroying scenes reuses the same HBM instead of mapping / unmapping it each time (cudaMalloc + cudaFree of a
// 2.6 MB scene measured 25 + 55 ms per create/destroy pair next to a 226 ms render; from the pool it is < 1 ms).
// Every allocation names the stream it is ordered on: the scene's upload stream for the immutable arrays (consumers wait
// on gs_scene::ready), the caller's stream for per-render scratch (calls on different streams are chained through
// gs_scene::last).
cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t stream);
void dev_free(void* p, cudaStream_t stream);

template <class T> struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	// grow-only scratch, ordered on `stream`
	cudaError_t alloc(size_t count, cudaStream_t stream) {
		if (count <= n && p) return cudaSuccess;
		release(stream);
		n = count;
		return dev_alloc(reinterpret_cast<void**>(&p), n * sizeof(T), stream);
	}
	void release(cudaStream_t stream) { if (p) dev_free(p, stream); p = nullptr; n = 0; }
};

// Page-locked staging memory from a small process-wide pool (pinning 60 MB costs tens of milliseconds; a scene is created
// per frame by callers that treat gs_scene_create + gs_render as one call).
struct Staging {
	uint8_t* p = nullptr;
	size_t cap = 0;
	int slot = -1;
};
cudaError_t staging_acquire(size_t bytes, Staging& out);
void staging_release(Staging& s, cudaEvent_t busy_until);  // the block is reused once the event has completed
void staging_forget_event(cudaEvent_t ev);                  // an event that is about to be destroyed no longer guards a block
const char* last_error_cstr();

constexpr int kCounterSlots = 256;
constexpr int kMaxChains = 4;      // independent chains of adaptive rounds (gs_render_adaptive_accumulate)
constexpr size_t kSmallSceneBytes = 128u << 10;  // C1 (486 spheres) is 47 KB, C3 2 KB, C4 (the one that wants the spill-free build) 490 KB
constexpr unsigned kAdaptiveTailPixels = 8192;   // adaptive renders: below this many unfinished pixels, switch to one warp per pixel (measured 256 / 2048 / 8192: profiles/r02_v2_adaptive_times.log)
constexpr int kStatSlots = 16;                   // [0..3] = gs_stats totals; [4..] only written by the GS_DIVERGENCE_STATS diagnostic build
constexpr size_t kWideMinPrims = 65536;          // scenes with at least this many primitives traverse the eight-wide tree (override: GS_BVH_WIDTH);
                                                 // measured crossover on the sphere cloud: binary +9 % at 30 k, eight-wide +3 % at 100 k, +10 % at 300 k
                                                 // (profiles/r02_v5_tree_threshold.log)

}  // namespace gs

struct gs_scene {
	int device = 0;
	int sm_count = 0;
	cudaStream_t up = nullptr;    // uploads and allocations of the immutable arrays
	cudaEvent_t ready = nullptr;  // recorded on `up` after the last upload; every entry makes its stream wait on it
	cudaEvent_t last = nullptr;   // recorded after the work an entry issued: orders entries on different streams (scratch reuse)
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // gs_render's timing pair
	bool last_valid = false;

	// ---- host copies kept for (re)building node arrays: re-padding when a camera sits outside the assumed origin range,
	// and the tree that was not built at creation (gs_scene_set_option "bvh_width")
	std::unique_ptr<gs_bvh_node[]> h_nodes;  // copy of the caller's tree (filled by a helper thread while the layout runs)
	size_t n_nodes = 0;
	std::vector<int> h_node_to_dnode;   // ABI node -> DNode index (internal nodes), -1 for leaves
	std::unique_ptr<int[]> h_leaf_code;  // ABI node -> leaf code (leaves; entries of internal nodes are never read)
	std::vector<int> h_bfs;             // internal ABI nodes in BFS order (= binary device order)
	// eight-wide topology: slot k of wide node h holds ABI node w_kids[8h + k] (-1 = empty), w_child_base[h] = wide index of its
	// first internal kid, w_prim_base[3h + kind] = first device primitive of that kind among its single-primitive leaf kids
	std::vector<int> w_kids;
	std::vector<int> w_child_base;
	std::vector<uint32_t> w_prim_base;
	bool w_any_link = false;  // some leaf cannot be addressed as base + offset from its wide node (build_nodes8 then emits the side table)
	std::unique_ptr<int[]> h_leaf_direct;  // ABI leaf -> device primitive index if the leaf is one freshly placed primitive the wide node can address by offset, else -1
	int wide_depth = 0;
	double origin_limit = 0.0;
	double max_abs_coord = 0.0;

	// ---- device: one allocation for all immutable arrays except the node arrays
	void* blob = nullptr;
	size_t blob_bytes = 0;
	gs::DNode* nodes = nullptr;        size_t n_dnodes = 0;
	gs::DNode8* nodes8 = nullptr;      size_t n_dnodes8 = 0;
	int* leaf_links8 = nullptr;
	std::vector<gs_image> h_images;  // sizes/formats for validation of camera.hdri_image
	gs::DeviceScene ds{};
	bool wide = false;           // which tree the kernels traverse

	// adaptive sampler: the two lists of pixels still sampling, their lengths, per-pixel (sum lum, sum lum^2)
	cudaStream_t aux[gs::kMaxChains - 1] = {};        // streams of the adaptive sampler's extra chains, made on first use
	cudaEvent_t aux_done[gs::kMaxChains - 1] = {}, aux_begin = nullptr;
	bool aux_ready = false;
	gs::DevBuf<uint32_t> px_list[2];
	gs::DevBuf<unsigned> px_count;
	gs::DevBuf<double2> lum;

	// render scratch
	gs::DevBuf<float4> accum;
	gs::DevBuf<float> out_rgb;
	gs::DevBuf<unsigned> out_counts;
	unsigned long long* counters = nullptr;  // kCounterSlots work counters   (allocated with the blob's stream)
	unsigned long long* stats = nullptr;     // kStatSlots device totals
	int next_counter = 0;
	uint64_t launches = 0;
	bool count_nodes = false;
	bool has_triangles = false;  // selects the kernel variants that carry the triangle branch
	bool small_scene = false;    // nodes + primitives fit the L1 with room to spare: the kernels built for one more CTA per SM
	uint64_t h2d_bytes = 0;

	// gs_render_multi: this device's accumulator (plain cudaMalloc: peers map it), a landing buffer for peers that cannot be mapped
	float4* multi_accum = nullptr;
	size_t multi_accum_n = 0;
	gs::DevBuf<float4> multi_landing;

	// wavefront pipeline: path pool (structure of arrays) + counters.  Off by default: measured against the
	// megakernel on B200 it wins 6 % on C4 and loses ~2x on C1/C3; option "wavefront" = 1 enables it.
	bool use_wavefront = false;
	int pool_capacity = 1 << 20;  // in-flight paths (112 B each)
	gs::DevBuf<double4> pool_org, pool_dir;
	gs::DevBuf<float4> pool_thr, pool_rad;
	gs::DevBuf<uint4> pool_ids, pool_hit;
	gs::DevBuf<int> free_list;
	gs::DevBuf<int> shade_queue;        // material-sorted shading queues (kShadeClasses x pool slots)
	gs::DevBuf<unsigned> shade_count;
	bool wavefront_sort = false;        // option "wavefront_sort": measured 6...9 % slower than the unsorted queues (profiles/r02_v5_wavefront_sort.log)
	int adaptive_chains = 0;            // option "adaptive_chains": independent chains of adaptive rounds, 0 = chosen by the image size
	bool tail_ahead = true;             // option "tail_ahead": adaptive tail renders several batches of a pixel at once when few pixels are left
	gs::DevBuf<gs::WaveCounters> wave_ctr;
	gs::WaveCounters* h_wave_ctr = nullptr;  // pinned

	~gs_scene();
};

namespace gs {
int validate_desc(const gs_scene_desc* d);
// Re-layout + upload of everything but the node arrays; fills s->ds, the host topology and s->ready's predecessors.
int layout_scene(gs_scene* s, const gs_scene_desc* d);
// (Re)build the node array of the binary / eight-wide tree for ray origins within origin_limit of the world origin.
int build_nodes2(gs_scene* s, double origin_limit);
int build_nodes8(gs_scene* s, double origin_limit);
int build_current_nodes(gs_scene* s, double origin_limit);  // whichever trees exist
int select_tree(gs_scene* s, int width);                    // 2 or 8; builds the tree if it does not exist yet
}  // namespace gs
