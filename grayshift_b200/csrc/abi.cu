// grayshift_b200/csrc/abi.cu -- the entry points of the C ABI in include/grayshift.h: scene creation (the re-layout and
// the two trees live in scene_layout.cu), kernel launches, resolve.  Host code only; the kernels live in render_kernels.cu
// and trace_closest.cu.  There is no CPU compute path here: without a CUDA device every compute entry fails with
// GS_ERR_NO_DEVICE.
//
// Stream discipline.  The immutable scene arrays are uploaded on the scene's own stream and guarded by gs_scene::ready;
// every entry that issues device work makes the caller's stream wait on `ready` first (gs_scene_create returns while the
// copy may still be in flight) and on gs_scene::last, the event recorded behind the previous entry's work -- entries
// issued on DIFFERENT streams therefore run one after the other: the per-scene scratch (pixel lists, work counters,
// staging for the output) is not re-entrant, and that ordering is what makes it safe.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <mutex>
#include <thread>

#include "gs_host.h"

using namespace gs;

namespace {

int convert_camera(const gs_scene* s, const gs_camera* c, DCamera& d) {
	if (!c) return fail(GS_ERR_INVALID, "null camera");
	if (c->image_width <= 0 || c->image_height <= 0) return fail(GS_ERR_INVALID, "camera: bad image size");
	if ((uint64_t)c->image_width * (uint64_t)c->image_height >= 0xffffffffull) return fail(GS_ERR_UNSUPPORTED, "image too large");
	for (int a = 0; a < 3; ++a) {
		d.center[a] = c->center[a]; d.pixel00[a] = c->starting_pixel_pos[a]; d.du[a] = c->pixel_delta_u[a]; d.dv[a] = c->pixel_delta_v[a];
		d.disk_u[a] = c->defocus_disk_u[a]; d.disk_v[a] = c->defocus_disk_v[a]; d.bg[a] = (float)c->background_color[a];
	}
	d.defocus_angle = c->defocus_angle;
	d.width = c->image_width; d.height = c->image_height; d.max_depth = c->max_depth;
	d.bg_kind = c->background_kind;
	d.hdri_image = 0;
	if (c->background_kind == GS_BG_HDRI) {
		if (c->hdri_image < 0 || (size_t)c->hdri_image >= s->h_images.size() || s->h_images[c->hdri_image].format != GS_IMG_F32_RGB)
			return fail(GS_ERR_INVALID, "camera: hdri_image must index a GS_IMG_F32_RGB image");
		d.hdri_image = c->hdri_image;
	} else if (c->background_kind != GS_BG_SOLID) return fail(GS_ERR_INVALID, "camera: bad background kind");
	// rotate_vector's matrix, util.rs:67-86 (radians)
	double sx = std::sin(c->hdri_rotation[0]), cx = std::cos(c->hdri_rotation[0]);
	double sy = std::sin(c->hdri_rotation[1]), cy = std::cos(c->hdri_rotation[1]);
	double sz = std::sin(c->hdri_rotation[2]), cz = std::cos(c->hdri_rotation[2]);
	double m[9] = {cy * cz, cx * sz + sx * sy * cz, sx * sz - cx * sy * cz,
	               -cy * sz, cx * cz - sx * sy * sz, sx * cz + cx * sy * sz,
	               sy, -sx * cy, cx * cy};
	for (int i = 0; i < 9; ++i) d.rot[i] = (float)m[i];
	return GS_OK;
}

int ensure_origin_range(gs_scene* s, const gs_camera* c) {
	double m = 0.0;
	for (int a = 0; a < 3; ++a) {
		double reach = std::fabs(c->center[a]) + std::fabs(c->defocus_disk_u[a]) + std::fabs(c->defocus_disk_v[a]);
		m = std::fmax(m, reach);
	}
	if (m <= s->origin_limit) return GS_OK;
	GS_CUDA(cudaDeviceSynchronize());  // the node arrays are rewritten in place: nothing may be reading them
	return build_current_nodes(s, 2.0 * m);
}

int check_device(gs_scene* s) {
	if (!s) return fail(GS_ERR_INVALID, "null scene");
	GS_CUDA(cudaSetDevice(s->device));
	return GS_OK;
}

// order this entry's work on `st` behind the scene upload and behind the previous entry (see the header comment)
int begin_entry(gs_scene* s, cudaStream_t st) {
	GS_CUDA(cudaStreamWaitEvent(st, s->ready, 0));
	if (s->last_valid) GS_CUDA(cudaStreamWaitEvent(st, s->last, 0));
	return GS_OK;
}
int end_entry(gs_scene* s, cudaStream_t st) {
	GS_CUDA(cudaEventRecord(s->last, st));
	s->last_valid = true;
	return GS_OK;
}

// scratch of the parity hooks: allocated and freed on the scene's stream within the call
template <class T> struct Scratch {
	DevBuf<T> b;
	cudaStream_t st;
	explicit Scratch(cudaStream_t s) : st(s) {}
	~Scratch() { b.release(st); }
};

// Work items a warp takes from the queue at a time.  A launch ends when the last chunk taken is done, and a warp needs about
// 2 us per path while the SMs are full.  Fixed-count launches: 256 (64 costs the 500-spp step 3.5 %: a division, an atomic and
// a pixel flush more often; even at 16 spp per launch 256 is not worse).  Rounds of the adaptive sampler, a few million paths of
// 32 per pixel each: 128 (256 / 128 / 64 / 32 with two chains: final scene 400x400 125.7 / 116.3 / 115.3 / 117.3 ms, 800x800
// 363.9 / 364.0 / 369.4 / 380.4, scene 1 1.79 / 1.60 / 1.57 / 1.54; profiles/r02_v12_adaptive_chunk.log).
unsigned chunk_items(bool adaptive_round) {
	static const int forced = [] { const char* e = std::getenv("GS_CHUNK"); return e ? std::atoi(e) : 0; }();  // tuning knob
	if (forced >= 32) return (unsigned)forced;
	return adaptive_round ? 128u : 256u;
}

unsigned long long* next_counter(gs_scene* s) {
	unsigned long long* p = s->counters + s->next_counter;
	s->next_counter = (s->next_counter + 1) % kCounterSlots;
	return p;
}

}  // namespace

extern "C" {

const char* gs_last_error(void) { return last_error_cstr(); }
int gs_abi_version(void) { return GS_ABI_VERSION; }

int gs_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int gs_scene_create(const gs_scene_desc* d, int device, gs_scene** out) {
	if (!out) return fail(GS_ERR_INVALID, "null out pointer");
	*out = nullptr;
	int rc = validate_desc(d);
	if (rc) return rc;
	int ndev = gs_device_count();
	if (ndev == 0) return fail(GS_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
	if (device < 0 || device >= ndev) return fail(GS_ERR_INVALID, "device ordinal out of range");
	GS_CUDA(cudaSetDevice(device));
	// two attribute queries instead of cudaGetDeviceProperties, which fills in every property of the device and was
	// the slowest call of a scene creation
	int cc_major = 0, sm_count = 0;
	GS_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, device));
	GS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
	if (cc_major < 10) return fail(GS_ERR_UNSUPPORTED, "device " + std::to_string(device) + " is not sm_100 class; this library carries sm_100a code only");

	gs_scene* s = new gs_scene();
	s->device = device;
	s->sm_count = sm_count;
	struct Guard { gs_scene* s; bool ok = false; ~Guard() { if (!ok) delete s; } } guard{s};
	GS_CUDA(cudaStreamCreateWithFlags(&s->up, cudaStreamNonBlocking));
	GS_CUDA(cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
	GS_CUDA(cudaEventCreateWithFlags(&s->last, cudaEventDisableTiming));
	GS_CUDA(cudaEventCreate(&s->ev0));
	GS_CUDA(cudaEventCreate(&s->ev1));
	rc = layout_scene(s, d);
	if (rc) return rc;
	rc = build_current_nodes(s, 4.0 * std::fmax(s->max_abs_coord, 1.0));
	if (rc) return rc;
	guard.ok = true;
	*out = s;
	return GS_OK;
}

void gs_scene_destroy(gs_scene* scene) { delete scene; }

int gs_scene_set_option(gs_scene* scene, const char* key, int64_t value) {
	if (!scene || !key) return fail(GS_ERR_INVALID, "null argument");
	if (std::strcmp(key, "count_nodes") == 0) { scene->count_nodes = value != 0; return GS_OK; }
	if (std::strcmp(key, "wavefront") == 0) {
		if (value != 0 && !scene->nodes) {  // the wavefront pipeline walks the binary tree
			int rc = check_device(scene);
			if (rc) return rc;
			GS_CUDA(cudaDeviceSynchronize());
			rc = build_nodes2(scene, scene->origin_limit);
			if (rc) return rc;
		}
		scene->use_wavefront = value != 0;
		return GS_OK;
	}
	if (std::strcmp(key, "wavefront_sort") == 0) { scene->wavefront_sort = value != 0; return GS_OK; }
	if (std::strcmp(key, "tail_ahead") == 0) { scene->tail_ahead = value != 0; return GS_OK; }
	if (std::strcmp(key, "adaptive_chains") == 0) {
		if (value < 0 || value > kMaxChains) return fail(GS_ERR_INVALID, "adaptive_chains: 0 (by image size) to 4");
		scene->adaptive_chains = (int)value;
		return GS_OK;
	}
	if (std::strcmp(key, "pool_capacity") == 0) {
		if (value < 128 || value > (1ll << 26)) return fail(GS_ERR_INVALID, "pool_capacity out of range");
		scene->pool_capacity = (int)value;
		return GS_OK;
	}
	if (std::strcmp(key, "bvh_width") == 0) {
		int rc = check_device(scene);
		if (rc) return rc;
		return select_tree(scene, (int)value);
	}
	return fail(GS_ERR_INVALID, std::string("unknown option '") + key + "'");
}

int gs_scene_get_info(gs_scene* scene, const char* key, int64_t* value) {
	if (!scene || !key || !value) return fail(GS_ERR_INVALID, "null argument");
	if (std::strcmp(key, "bvh_width") == 0) { *value = scene->wide ? 8 : 2; return GS_OK; }
	if (std::strcmp(key, "nodes2") == 0) { *value = (int64_t)scene->h_bfs.size(); return GS_OK; }
	if (std::strcmp(key, "nodes8") == 0) { *value = (int64_t)scene->w_child_base.size(); return GS_OK; }
	if (std::strcmp(key, "wide_depth") == 0) { *value = scene->wide_depth; return GS_OK; }
	if (std::strcmp(key, "scene_bytes") == 0) { *value = (int64_t)(scene->blob_bytes + scene->n_dnodes * sizeof(DNode) + scene->n_dnodes8 * sizeof(DNode8)); return GS_OK; }
	return fail(GS_ERR_INVALID, std::string("unknown info key '") + key + "'");
}

// The wavefront pipeline: waves of generate -> traverse -> shade over the path pool until every item has been
// started and the pool is empty.  The host looks at the counters every few waves (one small pinned read).
static int render_wavefront(gs_scene* s, WavefrontParams& p, cudaStream_t st) {
	int n = s->pool_capacity;
	if ((unsigned long long)n > p.total_items) n = (int)((p.total_items + 127ull) / 128ull * 128ull);
	GS_CUDA(s->pool_org.alloc(n, st)); GS_CUDA(s->pool_dir.alloc(n, st)); GS_CUDA(s->pool_thr.alloc(n, st)); GS_CUDA(s->pool_rad.alloc(n, st));
	GS_CUDA(s->pool_ids.alloc(n, st)); GS_CUDA(s->pool_hit.alloc(n, st)); GS_CUDA(s->free_list.alloc(n, st)); GS_CUDA(s->wave_ctr.alloc(1, st));
	if (!s->h_wave_ctr) GS_CUDA(cudaMallocHost(&s->h_wave_ctr, sizeof(WaveCounters)));
	p.pool.org = s->pool_org.p; p.pool.dir = s->pool_dir.p; p.pool.thr = s->pool_thr.p; p.pool.rad = s->pool_rad.p;
	p.pool.ids = s->pool_ids.p; p.pool.hit = s->pool_hit.p; p.pool.n = n;
	p.ctr = s->wave_ctr.p; p.free_list = s->free_list.p;
	p.shade_queue = nullptr; p.shade_count = nullptr;
	if (s->wavefront_sort) {
		GS_CUDA(s->shade_queue.alloc((size_t)n * kShadeClasses, st)); GS_CUDA(s->shade_count.alloc(kShadeClasses, st));
		p.shade_queue = s->shade_queue.p; p.shade_count = s->shade_count.p;
	}
	GS_CUDA(cudaMemsetAsync(p.ctr, 0, sizeof(WaveCounters), st));
	const int kCheckEvery = 4;
	for (int wave = 0;; ++wave) {
		GS_CUDA(launch_wf_generate(p, wave == 0, st));
		GS_CUDA(cudaMemsetAsync(&p.ctr->next_ray, 0, 2 * sizeof(unsigned), st));  // next_ray and n_free
		GS_CUDA(launch_wf_traverse(p, s->sm_count, s->count_nodes, s->has_triangles, st));
		if (p.shade_queue) {
			GS_CUDA(cudaMemsetAsync(p.shade_count, 0, kShadeClasses * sizeof(unsigned), st));
			GS_CUDA(launch_wf_classify(p, st));
			s->launches += 1;
		}
		GS_CUDA(launch_wf_shade(p, st));
		s->launches += 3;
		if (wave % kCheckEvery == kCheckEvery - 1) {
			GS_CUDA(cudaMemcpyAsync(s->h_wave_ctr, p.ctr, sizeof(WaveCounters), cudaMemcpyDeviceToHost, st));
			GS_CUDA(cudaStreamSynchronize(st));
			if (s->h_wave_ctr->next_item >= p.total_items && s->h_wave_ctr->n_live <= 0) break;
		}
		if (wave > 1000000) return fail(GS_ERR_CUDA, "wavefront did not terminate");
	}
	return GS_OK;
}

// the fixed-count launch without the entry bookkeeping (shared with gs_render_multi)
int gs_issue_fixed(gs_scene* s, const DCamera& cam, uint64_t seed, uint32_t sample_begin, uint32_t sample_count, void* d_accum, cudaStream_t st) {
	RenderParams p;
	p.sc = s->ds;
	p.cam = cam;
	p.seed = seed; p.sample_begin = sample_begin; p.sample_count = sample_count;
	p.total_items = (unsigned long long)cam.width * cam.height * sample_count;
	if (s->use_wavefront) {
		WavefrontParams w;
		w.sc = s->ds; w.cam = p.cam; w.seed = seed; w.sample_begin = sample_begin; w.sample_count = sample_count;
		w.total_items = p.total_items; w.accum = static_cast<float4*>(d_accum); w.stats = s->stats;
		return render_wavefront(s, w, st);
	}
	p.chunk = chunk_items(false);
	p.accum = static_cast<float4*>(d_accum);
	p.work_counter = next_counter(s);
	p.stats = s->stats;
	p.pixel_list = nullptr; p.n_list = nullptr; p.lum = nullptr;  // every pixel, no luminance statistics
	p.rounds_in_kernel = false; p.round0 = 0; p.work_counter2 = nullptr; p.n_pixels = 0; p.tail_pixels = 0; p.max_samples = 0; p.confidence_sq = p.tolerance_sq = 0.0;
	p.small_scene = s->small_scene;
	p.wide = s->wide;
	GS_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
	GS_CUDA(launch_render_fixed(p, s->sm_count, s->count_nodes, s->has_triangles, st));
	s->launches += 1;
	return GS_OK;
}

int gs_render_accumulate(gs_scene* s, const gs_camera* camera, uint64_t seed, uint32_t sample_begin, uint32_t sample_count, void* d_accum,
                         void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum) return fail(GS_ERR_INVALID, "null accumulator");
	if (sample_count == 0) return GS_OK;
	if (sample_count > 0x7fffffffu || (uint64_t)sample_begin + sample_count > 0xffffffffull) return fail(GS_ERR_UNSUPPORTED, "sample indices beyond 32 bits");
	DCamera cam;
	rc = convert_camera(s, camera, cam);
	if (rc) return rc;
	rc = ensure_origin_range(s, camera);
	if (rc) return rc;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	rc = begin_entry(s, st);
	if (rc) return rc;
	rc = gs_issue_fixed(s, cam, seed, sample_begin, sample_count, d_accum, st);
	if (rc) return rc;
	return end_entry(s, st);
}

int gs_render_adaptive_accumulate(gs_scene* s, const gs_camera* camera, const gs_sample_settings* settings, uint64_t seed,
                                  uint32_t pixel_begin, uint32_t pixel_stride, void* d_accum, void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum || !settings) return fail(GS_ERR_INVALID, "null argument");
	if (settings->batch_size == 0) return fail(GS_ERR_INVALID, "batch_size must be > 0 (the reference would loop forever)");
	if (pixel_stride == 0) return fail(GS_ERR_INVALID, "pixel_stride must be > 0");
	RenderParams p;
	rc = convert_camera(s, camera, p.cam);
	if (rc) return rc;
	rc = ensure_origin_range(s, camera);
	if (rc) return rc;
	p.sc = s->ds;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	const unsigned total = (unsigned)camera->image_width * (unsigned)camera->image_height;
	const unsigned n_pixels = pixel_begin >= total ? 0u : (total - pixel_begin + pixel_stride - 1) / pixel_stride;
	if (n_pixels == 0) return GS_OK;
	const double confidence_sq = settings->confidence * settings->confidence;  // camera.rs:128-129
	const double tolerance_sq = settings->tolerance * settings->tolerance;
	// Camera::sample's loop (camera.rs:135-165) leaves after the first batch that takes `sample_count as u32` past
	// max_samples: at most max_samples / batch_size + 1 batches.
	const unsigned long long rounds = (unsigned long long)settings->max_samples / settings->batch_size + 1ull;
	// the per-pixel sample count lives in the accumulator's fourth float: exact up to 2^24
	if (rounds * settings->batch_size > (1ull << 24)) return fail(GS_ERR_UNSUPPORTED, "adaptive sampling beyond 2^24 samples per pixel");
	rc = begin_entry(s, st);
	if (rc) return rc;
	static const unsigned tail_pixels = [] { const char* e = std::getenv("GS_TAIL_PIXELS"); return e ? (unsigned)std::atoi(e) : kAdaptiveTailPixels; }();  // tuning knob
	// Who drives the batch rounds.  Default: the host (one render launch + the stop test per batch, a 4-byte "anyone left?" read after
	// batches 1, 2, 4, 8 ...).  GS_ADAPTIVE_ROUNDS=kernel: all rounds inside one cooperative launch with grid-wide syncs in between --
	// no host round trip, but measured 2...15 % slower on every scene of main.rs (profiles/r02_v4_adaptive_rounds_host_vs_kernel.log:
	// the final scene at 400x400 141 vs 165 ms): CTAs that finish a round early spin at the grid barrier next to the ones still
	// rendering and every arrival invalidates the SM's L1, where a kernel boundary simply frees the SM.
	static const bool host_rounds = [] { const char* e = std::getenv("GS_ADAPTIVE_ROUNDS"); return !(e && std::strcmp(e, "kernel") == 0); }();
	// Chains.  A round's launch ends with a drain: the queue is empty, the SMs go idle one by one, and the launch lasts until its
	// longest path is done (the final scene at 400x400: 5 M paths per round, the SMs busy 79 % of the launch,
	// profiles/r02_v10_scene8_adaptive_round_raw.csv).  The stop test is per pixel, so the pixels are dealt out to a few independent
	// chains of rounds, each on its own stream: while one chain's launch drains, the next chain's launch takes over the idle SMs.
	static const unsigned max_chains = [] { const char* e = std::getenv("GS_ADAPTIVE_CHAINS"); const int v = e ? std::atoi(e) : 2; return (unsigned)(v < 1 ? 1 : (v > kMaxChains ? kMaxChains : v)); }();
	// (measured, profiles/r02_v11_adaptive_chains.log, 1 / 2 / 3 / 4 chains: the final scene at 400x400 142 / 123 / 125 / 129 ms, at 800x800
	// 397 / 366 / 377 / 389, scene 5 94 / 76 / 77 / 77, Cornell 90 / 87 / 91 / 94; renders of 90 k pixels lose 0.2 ms to the extra launches,
	// hence the 65536 pixels a chain must have)
	unsigned K = host_rounds ? std::min(max_chains, std::max(1u, n_pixels / 65536u)) : 1u;
	if (host_rounds && s->adaptive_chains > 0) K = std::min((unsigned)s->adaptive_chains, std::max(1u, n_pixels / 64u));  // option "adaptive_chains"
	if (K > 1 && !s->aux_ready) {
		for (int c = 0; c < kMaxChains - 1; ++c) {
			GS_CUDA(cudaStreamCreateWithFlags(&s->aux[c], cudaStreamNonBlocking));
			GS_CUDA(cudaEventCreateWithFlags(&s->aux_done[c], cudaEventDisableTiming));
		}
		GS_CUDA(cudaEventCreateWithFlags(&s->aux_begin, cudaEventDisableTiming));
		s->aux_ready = true;
	}
	GS_CUDA(s->px_list[0].alloc((size_t)n_pixels * 2, st));  // per chain the two lists of its rounds, back to back
	GS_CUDA(s->px_count.alloc(4 * kMaxChains, st));
	GS_CUDA(s->lum.alloc(total, st));
	GS_CUDA(cudaMemsetAsync(s->lum.p, 0, (size_t)total * sizeof(double2), st));
	p.seed = seed;
	p.sample_begin = 0;
	p.sample_count = settings->batch_size;
	p.chunk = 256;  // (set per chain below)
	p.accum = static_cast<float4*>(d_accum);
	p.stats = s->stats;
	p.lum = s->lum.p;
	p.small_scene = s->small_scene;
	p.wide = s->wide;
	p.confidence_sq = confidence_sq; p.tolerance_sq = tolerance_sq;
	p.max_samples = settings->max_samples;
	p.tail_pixels = tail_pixels;
	p.rounds_in_kernel = !host_rounds;
	p.round0 = 0;
	struct Chain {
		cudaStream_t st = nullptr;
		RenderParams p;
		unsigned n = 0;           // pixels of this chain
		unsigned which = 0;       // the list holding the pixels still sampling
		unsigned left = 0;        // host copy of its length (valid after a check)
		unsigned known_left = 0;  // the number of pixels handed to the tail kernel, when the host has read it
		bool active = true, need_tail = false;
	} chain[kMaxChains];
	size_t list_at = 0;
	for (unsigned c = 0; c < K; ++c) {
		Chain& ch = chain[c];
		ch.st = c == 0 ? st : s->aux[c - 1];
		ch.n = (n_pixels - c + K - 1) / K;  // list entries c, c + K, c + 2K ... of the whole render
		ch.p = p;
		ch.p.n_pixels = ch.n;
		ch.p.total_items = (unsigned long long)ch.n * settings->batch_size;  // upper bound (sizes the grid); the kernel reads the list lengths
		ch.p.chunk = chunk_items(true);
		ch.p.pixel_list = s->px_list[0].p + list_at;
		ch.p.n_list = s->px_count.p + 4 * c;
		ch.p.work_counter = next_counter(s);
		ch.p.work_counter2 = next_counter(s);
		ch.need_tail = !host_rounds;
		list_at += (size_t)ch.n * 2;
		GS_CUDA(launch_adaptive_init(ch.p.pixel_list, ch.p.n_list, pixel_begin + c * pixel_stride, pixel_stride * K, ch.n, st));
		GS_CUDA(cudaMemsetAsync(ch.p.work_counter, 0, sizeof(unsigned long long), st));
		GS_CUDA(cudaMemsetAsync(ch.p.work_counter2, 0, sizeof(unsigned long long), st));
		s->launches += 1;
	}
	if (K > 1) {
		GS_CUDA(cudaEventRecord(s->aux_begin, st));
		for (unsigned c = 1; c < K; ++c) GS_CUDA(cudaStreamWaitEvent(chain[c].st, s->aux_begin, 0));
	}
	static const bool trace = std::getenv("GS_TRACE") != nullptr;
	if (host_rounds) {
		// one launch per batch and chain + the stop test; after batches 1, 2, 4, 8 ... a 4-byte read per chain tells whether to go on,
		// to hand the last few pixels to the tail kernel, or to stop (synchronises the streams)
		for (unsigned long long r = 0; r < rounds; ++r) {
			const unsigned cur = (unsigned)(r & 1ull), nxt = cur ^ 1u;
			bool any = false;
			for (unsigned c = 0; c < K; ++c) {
				Chain& ch = chain[c];
				if (!ch.active) continue;
				any = true;
				ch.p.round0 = (unsigned)r;
				ch.p.work_counter = next_counter(s);
				GS_CUDA(cudaMemsetAsync(ch.p.work_counter, 0, sizeof(unsigned long long), ch.st));
				GS_CUDA(cudaMemsetAsync(ch.p.n_list + nxt, 0, sizeof(unsigned), ch.st));
				GS_CUDA(launch_render_fixed(ch.p, s->sm_count, false, s->has_triangles, ch.st));
				GS_CUDA(launch_adaptive_decide(ch.p, cur, ch.st));
				s->launches += 2;
				ch.which = nxt;
			}
			if (!any) break;
			if (((r + 1ull) & r) == 0ull && r + 1 < rounds) {
				for (unsigned c = 0; c < K; ++c)
					if (chain[c].active) GS_CUDA(cudaMemcpyAsync(&chain[c].left, chain[c].p.n_list + nxt, sizeof(unsigned), cudaMemcpyDeviceToHost, chain[c].st));
				for (unsigned c = 0; c < K; ++c) {
					Chain& ch = chain[c];
					if (!ch.active) continue;
					GS_CUDA(cudaStreamSynchronize(ch.st));
					const unsigned share = tail_pixels / K;
					if (trace) std::fprintf(stderr, "[gs_render_adaptive] chain %u: %u of %u pixels left after %llu batches%s\n", c, ch.left, ch.n, r + 1ull, ch.left <= share ? ": tail kernel" : "");
					if (ch.left <= share) { ch.known_left = ch.left; ch.need_tail = ch.left > 0; ch.active = false; }
				}
			}
		}
		// (a chain that ran all its rounds has stopped every pixel: max_samples, camera.rs:162-164)
	} else
	// (alternative) every batch round in one cooperative launch (render, grid sync, stop test, grid sync ...) until at most
	// tail_pixels pixels are left; then one warp per remaining pixel finishes its loop.  No host round trip.
	{ GS_CUDA(launch_render_fixed(chain[0].p, s->sm_count, false, s->has_triangles, st)); s->launches += 1; }
	for (unsigned c = 0; c < K; ++c) {
		Chain& ch = chain[c];
		if (ch.need_tail) {
			// which list the tail kernel reads: it looks at bit 0 only, so a byte fill will do -- no host buffer to keep alive
			if (host_rounds) GS_CUDA(cudaMemsetAsync(ch.p.n_list + 2, (int)ch.which, sizeof(unsigned), ch.st));
			AdaptiveTailParams tp;
			tp.sc = s->ds; tp.cam = p.cam; tp.seed = seed;
			tp.confidence_sq = confidence_sq; tp.tolerance_sq = tolerance_sq;
			tp.batch_size = settings->batch_size; tp.max_samples = settings->max_samples;
			tp.lists = ch.p.pixel_list; tp.n_list = ch.p.n_list; tp.n_pixels = ch.n;
			tp.accum = static_cast<float4*>(d_accum); tp.lum = s->lum.p;
			tp.work_counter = next_counter(s); tp.stats = s->stats;
			GS_CUDA(cudaMemsetAsync(tp.work_counter, 0, sizeof(unsigned long long), ch.st));
			GS_CUDA(launch_adaptive_tail(tp, std::min(ch.n, tail_pixels), s->tail_ahead ? ch.known_left : 0u, s->sm_count, s->wide, ch.st));
			s->launches += 1;
		}
		if (c > 0) {  // the caller's stream continues after every chain
			GS_CUDA(cudaEventRecord(s->aux_done[c - 1], ch.st));
			GS_CUDA(cudaStreamWaitEvent(st, s->aux_done[c - 1], 0));
		}
	}
	return end_entry(s, st);
}

int gs_resolve(gs_scene* s, const void* d_accum, int32_t width, int32_t height, float* out_rgb, uint32_t* out_counts, void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum || !out_rgb || width <= 0 || height <= 0) return fail(GS_ERR_INVALID, "bad resolve arguments");
	const int n = width * height;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	rc = begin_entry(s, st);
	if (rc) return rc;
	GS_CUDA(s->out_rgb.alloc((size_t)n * 3, st));
	GS_CUDA(s->out_counts.alloc((size_t)n, st));
	GS_CUDA(launch_resolve(static_cast<const float4*>(d_accum), n, s->out_rgb.p, s->out_counts.p, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb, s->out_rgb.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
	if (out_counts) GS_CUDA(cudaMemcpyAsync(out_counts, s->out_counts.p, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

int gs_resolve_rgb8(gs_scene* s, const void* d_accum, int32_t width, int32_t height, uint8_t* out_rgb8, uint32_t* out_counts, void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum || !out_rgb8 || width <= 0 || height <= 0) return fail(GS_ERR_INVALID, "bad resolve arguments");
	const int n = width * height;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	rc = begin_entry(s, st);
	if (rc) return rc;
	GS_CUDA(s->out_rgb.alloc((size_t)n * 3, st));  // reused as the byte buffer (n*3 bytes fit in n*3 floats)
	GS_CUDA(s->out_counts.alloc((size_t)n, st));
	unsigned char* d_rgb8 = reinterpret_cast<unsigned char*>(s->out_rgb.p);
	GS_CUDA(launch_resolve_rgb8(static_cast<const float4*>(d_accum), n, d_rgb8, s->out_counts.p, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb8, d_rgb8, (size_t)n * 3, cudaMemcpyDeviceToHost, st));
	if (out_counts) GS_CUDA(cudaMemcpyAsync(out_counts, s->out_counts.p, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

#ifdef GS_DIVERGENCE_STATS
// diagnostic builds only: the raw device counters, read and reset
int gs_debug_counters(gs_scene* s, unsigned long long* out) {
	GS_CUDA(cudaDeviceSynchronize());
	GS_CUDA(cudaMemcpy(out, s->stats, kStatSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	GS_CUDA(cudaMemset(s->stats + 4, 0, (kStatSlots - 4) * sizeof(unsigned long long)));
	return GS_OK;
}
#endif

int gs_stats_read(gs_scene* s, gs_stats* stats) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!stats) return fail(GS_ERR_INVALID, "null stats");
	GS_CUDA(cudaDeviceSynchronize());
	unsigned long long h[4];
	GS_CUDA(cudaMemcpy(h, s->stats, sizeof h, cudaMemcpyDeviceToHost));
	GS_CUDA(cudaMemset(s->stats, 0, sizeof h));
	GS_CUDA(cudaDeviceSynchronize());
	std::memset(stats, 0, sizeof *stats);
	stats->rays = h[0]; stats->paths = h[1]; stats->tex_fetches = h[2]; stats->node_visits = h[3];
	stats->kernel_launches = s->launches;
	stats->h2d_bytes = s->h2d_bytes;
	s->launches = 0;
	return GS_OK;
}

int gs_render(gs_scene* s, const gs_camera* camera, const gs_sample_settings* settings, uint64_t seed, float* out_rgb, uint32_t* out_counts,
              gs_stats* stats) {
	auto t0 = std::chrono::steady_clock::now();
	int rc = check_device(s);
	if (rc) return rc;
	if (!camera || !settings || !out_rgb) return fail(GS_ERR_INVALID, "null argument");
	if (settings->batch_size == 0) return fail(GS_ERR_INVALID, "batch_size must be > 0 (the reference would loop forever)");
	if (camera->image_width <= 0 || camera->image_height <= 0) return fail(GS_ERR_INVALID, "camera: bad image size");
	const size_t n = (size_t)camera->image_width * camera->image_height;
	cudaStream_t st = s->up;  // the scene's own stream: a blocking call has no caller stream to respect
	rc = begin_entry(s, st);
	if (rc) return rc;
	GS_CUDA(s->accum.alloc(n, st));
	GS_CUDA(cudaMemsetAsync(s->accum.p, 0, n * sizeof(float4), st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaEventRecord(s->ev0, st));
	if (settings->max_samples < settings->batch_size) {
		// After the first batch `sample_count as u32 > max_samples` holds whatever the convergence test says
		// (camera.rs:153-164): exactly batch_size samples per pixel -> the fixed-count megakernel.  (tolerance
		// == 0 alone is not enough: on constant pixels the variance rounds below zero and `x < 0` stops early.)
		rc = gs_render_accumulate(s, camera, seed, 0, settings->batch_size, s->accum.p, st);
	} else {
		rc = gs_render_adaptive_accumulate(s, camera, settings, seed, 0, 1, s->accum.p, st);
	}
	if (rc) return rc;
	GS_CUDA(cudaEventRecord(s->ev1, st));
	rc = gs_resolve(s, s->accum.p, camera->image_width, camera->image_height, out_rgb, out_counts, st);
	if (rc) return rc;
	float ms = 0.f;
	GS_CUDA(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
	if (stats) {
		rc = gs_stats_read(s, stats);
		if (rc) return rc;
		stats->kernel_ms = ms;
		stats->d2h_bytes = n * 3 * sizeof(float) + (out_counts ? n * sizeof(unsigned) : 0);
		stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	}
	return GS_OK;
}


// ---- several GPUs, one call -------------------------------------------------------------------------------------
}  // extern "C"
namespace {

// NCCL on demand (GS_MULTI_REDUCE=nccl): the library does not link libnccl; the handful of entry points it needs are
// looked up in whatever libnccl.so.2 the process already holds or the loader finds.
struct NcclApi {
	typedef int (*InitAll)(void** comms, int ndev, const int* devlist);
	typedef int (*Group)();
	typedef int (*Reduce)(const void* send, void* recv, size_t count, int dtype, int op, int root, void* comm, cudaStream_t stream);
	typedef const char* (*ErrStr)(int);
	typedef int (*Destroy)(void* comm);
	InitAll init_all = nullptr; Group group_start = nullptr, group_end = nullptr; Reduce reduce = nullptr; ErrStr err = nullptr; Destroy destroy = nullptr;
	bool ok = false;
	std::string why;
};
NcclApi& nccl_api() {
	static NcclApi api = [] {
		NcclApi a;
		void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!h) { a.why = std::string("libnccl not found: ") + dlerror(); return a; }
		a.init_all = (NcclApi::InitAll)dlsym(h, "ncclCommInitAll");
		a.group_start = (NcclApi::Group)dlsym(h, "ncclGroupStart");
		a.group_end = (NcclApi::Group)dlsym(h, "ncclGroupEnd");
		a.reduce = (NcclApi::Reduce)dlsym(h, "ncclReduce");
		a.err = (NcclApi::ErrStr)dlsym(h, "ncclGetErrorString");
		a.destroy = (NcclApi::Destroy)dlsym(h, "ncclCommDestroy");
		a.ok = a.init_all && a.group_start && a.group_end && a.reduce && a.err;
		if (!a.ok) a.why = "libnccl lacks an expected entry point";
		return a;
	}();
	return api;
}
struct NcclComms { std::vector<int> devs; std::vector<void*> comms; };
std::mutex g_nccl_mu;
std::vector<NcclComms> g_nccl_comms;  // one communicator set per device list, kept for the life of the process

int nccl_comms_for(const std::vector<int>& devs, std::vector<void*>& out) {
	NcclApi& api = nccl_api();
	if (!api.ok) return fail(GS_ERR_NCCL, api.why);
	std::lock_guard<std::mutex> lock(g_nccl_mu);
	for (NcclComms& c : g_nccl_comms)
		if (c.devs == devs) { out = c.comms; return GS_OK; }
	NcclComms c;
	c.devs = devs;
	c.comms.assign(devs.size(), nullptr);
	const int r = api.init_all(c.comms.data(), (int)devs.size(), devs.data());
	if (r != 0) return fail(GS_ERR_NCCL, std::string("ncclCommInitAll: ") + api.err(r));
	g_nccl_comms.push_back(c);
	out = c.comms;
	return GS_OK;
}

}  // namespace
extern "C" {

int gs_render_multi(gs_scene* const* scenes, int n_scenes, const gs_camera* camera, const gs_sample_settings* settings, uint64_t seed,
                    float* out_rgb, uint32_t* out_counts, gs_stats* stats) {
	auto t0 = std::chrono::steady_clock::now();
	if (!scenes || n_scenes < 1 || n_scenes > kMaxPeers) return fail(GS_ERR_INVALID, "gs_render_multi: 1.." + std::to_string(kMaxPeers) + " scenes");
	if (!camera || !settings || !out_rgb) return fail(GS_ERR_INVALID, "null argument");
	if (settings->batch_size == 0) return fail(GS_ERR_INVALID, "batch_size must be > 0 (the reference would loop forever)");
	if (camera->image_width <= 0 || camera->image_height <= 0) return fail(GS_ERR_INVALID, "camera: bad image size");
	std::vector<int> devs;
	for (int g = 0; g < n_scenes; ++g) {
		if (!scenes[g]) return fail(GS_ERR_INVALID, "null scene");
		for (int d : devs) if (d == scenes[g]->device) return fail(GS_ERR_INVALID, "gs_render_multi: two scenes on one device");
		devs.push_back(scenes[g]->device);
	}
	const size_t n = (size_t)camera->image_width * camera->image_height;
	const bool fixed = settings->max_samples < settings->batch_size;
	const int G = n_scenes;
	static const bool use_nccl = [] { const char* e = std::getenv("GS_MULTI_REDUCE"); return e && std::strcmp(e, "nccl") == 0; }();

	// ---- every device renders its share into its own accumulator ----
	auto render_share = [&](int g) -> int {
		gs_scene* s = scenes[g];
		int rc = check_device(s);
		if (rc) return rc;
		DCamera cam;
		rc = convert_camera(s, camera, cam);
		if (rc) return rc;
		rc = ensure_origin_range(s, camera);
		if (rc) return rc;
		if (s->multi_accum_n < n) {
			if (s->multi_accum) { GS_CUDA(cudaDeviceSynchronize()); GS_CUDA(cudaFree(s->multi_accum)); s->multi_accum = nullptr; }
			GS_CUDA(cudaMalloc(reinterpret_cast<void**>(&s->multi_accum), n * sizeof(float4)));
			s->multi_accum_n = n;
		}
		cudaStream_t st = s->up;
		rc = begin_entry(s, st);
		if (rc) return rc;
		GS_CUDA(cudaMemsetAsync(s->multi_accum, 0, n * sizeof(float4), st));
		GS_CUDA(cudaEventRecord(s->ev0, st));
		rc = end_entry(s, st);
		if (rc) return rc;
		if (fixed) {
			const uint32_t b = (uint32_t)(((uint64_t)settings->batch_size * (uint64_t)g) / (uint64_t)G);
			const uint32_t e = (uint32_t)(((uint64_t)settings->batch_size * (uint64_t)(g + 1)) / (uint64_t)G);
			rc = gs_render_accumulate(s, camera, seed, b, e - b, s->multi_accum, st);
		} else rc = gs_render_adaptive_accumulate(s, camera, settings, seed, (uint32_t)g, (uint32_t)G, s->multi_accum, st);
		if (rc) return rc;
		GS_CUDA(cudaEventRecord(s->ev1, st));
		return end_entry(s, st);
	};
	if (fixed || G == 1) {  // launches are asynchronous: one thread keeps every device busy
		for (int g = 0; g < G; ++g) { int rc = render_share(g); if (rc) return rc; }
	} else {  // the adaptive sampler reads a 4-byte "anyone left?" back between rounds: one host thread per device
		std::vector<int> rcs(G, GS_OK);
		std::vector<std::string> msgs(G);
		std::vector<std::thread> pool;
		for (int g = 0; g < G; ++g) pool.emplace_back([&, g] { rcs[g] = render_share(g); if (rcs[g]) msgs[g] = gs_last_error(); });
		for (auto& t : pool) t.join();
		for (int g = 0; g < G; ++g) if (rcs[g]) return fail(rcs[g], msgs[g]);
	}

	// ---- the exchange: the first device sums while it resolves ----
	gs_scene* s0 = scenes[0];
	int rc = check_device(s0);
	if (rc) return rc;
	cudaStream_t st0 = s0->up;
	rc = begin_entry(s0, st0);
	if (rc) return rc;
	PeerAccums peers;
	peers.n = G;
	peers.acc[0] = s0->multi_accum;
	bool reduced_by_nccl = false;
	if (use_nccl && G > 1) {
		std::vector<void*> comms;
		rc = nccl_comms_for(devs, comms);
		if (rc) return rc;
		NcclApi& api = nccl_api();
		int r = api.group_start();
		for (int g = 0; g < G && r == 0; ++g) {
			GS_CUDA(cudaSetDevice(scenes[g]->device));
			r = api.reduce(scenes[g]->multi_accum, scenes[g]->multi_accum, n * 4, /*ncclFloat*/ 7, /*ncclSum*/ 0, 0, comms[g], scenes[g]->up);
		}
		const int r2 = api.group_end();
		GS_CUDA(cudaSetDevice(s0->device));
		if (r != 0 || r2 != 0) return fail(GS_ERR_NCCL, std::string("ncclReduce: ") + api.err(r != 0 ? r : r2));
		for (int g = 1; g < G; ++g) {  // the peers' streams carry their part of the collective: chain them behind it
			GS_CUDA(cudaSetDevice(scenes[g]->device));
			rc = end_entry(scenes[g], scenes[g]->up);
			if (rc) return rc;
		}
		GS_CUDA(cudaSetDevice(s0->device));
		peers.n = 1;
		reduced_by_nccl = true;
	}
	if (!reduced_by_nccl) {
		size_t landing = 0;
		for (int g = 1; g < G; ++g) {
			GS_CUDA(cudaStreamWaitEvent(st0, scenes[g]->last, 0));
			int can = 0;
			GS_CUDA(cudaDeviceCanAccessPeer(&can, s0->device, scenes[g]->device));
			if (can) {
				cudaError_t e = cudaDeviceEnablePeerAccess(scenes[g]->device, 0);
				if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
				if (e != cudaSuccess) return fail(GS_ERR_NCCL, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
				peers.acc[g] = scenes[g]->multi_accum;
			} else {
				peers.acc[g] = nullptr;
				++landing;
			}
		}
		if (landing) {  // devices that cannot be mapped: their buffers are copied next to ours first
			GS_CUDA(s0->multi_landing.alloc(landing * n, st0));
			size_t k = 0;
			for (int g = 1; g < G; ++g) {
				if (peers.acc[g]) continue;
				float4* dst = s0->multi_landing.p + k * n;
				cudaError_t e = cudaMemcpyPeerAsync(dst, s0->device, scenes[g]->multi_accum, scenes[g]->device, n * sizeof(float4), st0);
				if (e != cudaSuccess) return fail(GS_ERR_NCCL, std::string("cudaMemcpyPeerAsync: ") + cudaGetErrorString(e));
				peers.acc[g] = dst;
				++k;
			}
		}
	}
	GS_CUDA(s0->out_rgb.alloc(n * 3, st0));
	GS_CUDA(s0->out_counts.alloc(n, st0));
	cudaEvent_t x0, x1;
	GS_CUDA(cudaEventCreate(&x0)); GS_CUDA(cudaEventCreate(&x1));
	struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{x0, x1};
	GS_CUDA(cudaEventRecord(x0, st0));
	GS_CUDA(launch_reduce_resolve(peers, (int)n, s0->out_rgb.p, nullptr, s0->out_counts.p, nullptr, st0));
	GS_CUDA(cudaEventRecord(x1, st0));
	s0->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb, s0->out_rgb.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st0));
	if (out_counts) GS_CUDA(cudaMemcpyAsync(out_counts, s0->out_counts.p, n * sizeof(unsigned), cudaMemcpyDeviceToHost, st0));
	rc = end_entry(s0, st0);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st0));
	if (stats) {
		std::memset(stats, 0, sizeof *stats);
		float xms = 0.f;
		GS_CUDA(cudaEventElapsedTime(&xms, x0, x1));
		for (int g = 0; g < G; ++g) {
			gs_stats one;
			rc = gs_stats_read(scenes[g], &one);
			if (rc) return rc;
			stats->paths += one.paths; stats->rays += one.rays; stats->tex_fetches += one.tex_fetches; stats->node_visits += one.node_visits;
			stats->kernel_launches += one.kernel_launches; stats->h2d_bytes += one.h2d_bytes;
			float ms = 0.f;
			GS_CUDA(cudaSetDevice(scenes[g]->device));
			GS_CUDA(cudaEventElapsedTime(&ms, scenes[g]->ev0, scenes[g]->ev1));
			stats->kernel_ms = std::max(stats->kernel_ms, (double)ms);
		}
		stats->exchange_ms = xms;
		stats->d2h_bytes = n * 3 * sizeof(float) + (out_counts ? n * sizeof(unsigned) : 0);
		stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
	}
	return GS_OK;
}

int gs_tonemap(const float* rgb, size_t n_pixels, uint8_t* out_rgb8) {  // color.rs:8-26
	if (!rgb || !out_rgb8) return fail(GS_ERR_INVALID, "null argument");
	for (size_t i = 0; i < n_pixels * 3; ++i) {
		double c = rgb[i];
		double g = c > 0.0 ? std::sqrt(c) : 0.0;
		double cl = g < 0.0 ? 0.0 : (g > 0.999 ? 0.999 : g);
		out_rgb8[i] = (uint8_t)(int)(256.0 * cl);
	}
	return GS_OK;
}

static int trace_closest_impl(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t, bool fast) {
	int rc = check_device(s);
	if (rc) return rc;
	if (n == 0) return GS_OK;
	if (!rays || !prim_id || !t) return fail(GS_ERR_INVALID, "null argument");
	double m = 0.0;
	for (size_t k = 0; k < n; ++k) for (int a = 0; a < 3; ++a) m = std::fmax(m, std::fabs(rays[k * 7 + a]));
	if (!std::isfinite(m)) return fail(GS_ERR_INVALID, "non-finite ray origin");
	if (m > s->origin_limit) {
		GS_CUDA(cudaDeviceSynchronize());
		rc = build_current_nodes(s, 2.0 * m);
		if (rc) return rc;
	}
	cudaStream_t st = s->up;
	rc = begin_entry(s, st);
	if (rc) return rc;
	Scratch<double> d_rays(st), d_t(st);
	Scratch<int> d_id(st);
	GS_CUDA(d_rays.b.alloc(n * 7, st)); GS_CUDA(d_t.b.alloc(n, st)); GS_CUDA(d_id.b.alloc(n, st));
	GS_CUDA(cudaMemcpyAsync(d_rays.b.p, rays, n * 7 * sizeof(double), cudaMemcpyHostToDevice, st));
	if (fast) GS_CUDA(launch_trace_closest_fast(s->ds, d_rays.b.p, n, tmin, tmax, d_id.b.p, d_t.b.p, s->wide, st));
	else GS_CUDA(launch_trace_closest(s->ds, d_rays.b.p, n, tmin, tmax, d_id.b.p, d_t.b.p, s->wide, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(prim_id, d_id.b.p, n * sizeof(int), cudaMemcpyDeviceToHost, st));
	GS_CUDA(cudaMemcpyAsync(t, d_t.b.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

int gs_trace_closest(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t) {
	return trace_closest_impl(s, rays, n, tmin, tmax, prim_id, t, false);
}
int gs_trace_closest_fast(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t) {
	return trace_closest_impl(s, rays, n, tmin, tmax, prim_id, t, true);
}

int gs_camera_rays(gs_scene* s, const gs_camera* camera, uint64_t seed, uint32_t sample, int32_t pixel_begin, int32_t pixel_end, double* rays_out) {
	int rc = check_device(s);
	if (rc) return rc;
	DCamera cam;
	rc = convert_camera(s, camera, cam);
	if (rc) return rc;
	if (pixel_begin < 0 || pixel_end < pixel_begin || pixel_end > camera->image_width * camera->image_height || !rays_out) return fail(GS_ERR_INVALID, "bad pixel range");
	size_t n = (size_t)(pixel_end - pixel_begin);
	if (n == 0) return GS_OK;
	cudaStream_t st = s->up;
	rc = begin_entry(s, st);
	if (rc) return rc;
	Scratch<double> d_rays(st);
	GS_CUDA(d_rays.b.alloc(n * 7, st));
	GS_CUDA(launch_camera_rays(cam, seed, sample, pixel_begin, pixel_end, d_rays.b.p, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(rays_out, d_rays.b.p, n * 7 * sizeof(double), cudaMemcpyDeviceToHost, st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

int gs_path_radiance(gs_scene* s, const gs_camera* camera, uint64_t seed, const uint32_t* pixel, const uint32_t* sample, size_t n, float* out_rgb) {
	int rc = check_device(s);
	if (rc) return rc;
	if (n == 0) return GS_OK;
	if (!pixel || !sample || !out_rgb) return fail(GS_ERR_INVALID, "null argument");
	DCamera cam;
	rc = convert_camera(s, camera, cam);
	if (rc) return rc;
	rc = ensure_origin_range(s, camera);
	if (rc) return rc;
	const uint32_t total = (uint32_t)camera->image_width * (uint32_t)camera->image_height;
	for (size_t k = 0; k < n; ++k) if (pixel[k] >= total) return fail(GS_ERR_INVALID, "pixel index out of range");
	cudaStream_t st = s->up;
	rc = begin_entry(s, st);
	if (rc) return rc;
	Scratch<unsigned> d_px(st), d_sm(st);
	Scratch<float> d_out(st);
	GS_CUDA(d_px.b.alloc(n, st)); GS_CUDA(d_sm.b.alloc(n, st)); GS_CUDA(d_out.b.alloc(n * 3, st));
	GS_CUDA(cudaMemcpyAsync(d_px.b.p, pixel, n * sizeof(unsigned), cudaMemcpyHostToDevice, st));
	GS_CUDA(cudaMemcpyAsync(d_sm.b.p, sample, n * sizeof(unsigned), cudaMemcpyHostToDevice, st));
	GS_CUDA(launch_path_radiance(s->ds, cam, seed, d_px.b.p, d_sm.b.p, n, d_out.b.p, s->wide, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb, d_out.b.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
	rc = end_entry(s, st);
	if (rc) return rc;
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

}  // extern "C"
