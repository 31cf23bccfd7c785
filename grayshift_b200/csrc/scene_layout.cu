// grayshift_b200/csrc/scene_layout.cu -- gs_scene_create's work: validation, re-layout of the host's flat scene description
// for the GPU, the two acceleration structures (binary DNode tree, eight-wide compressed DNode8 tree), upload.  Host code.
//
// Data path: every device array is written ONCE, directly into page-locked staging memory (process-wide pool), by loops
// spread over the host threads; the immutable arrays share one device allocation and travel in one cudaMemcpyAsync on the
// scene's upload stream; gs_scene::ready is recorded behind it and every render entry makes its stream wait on that event,
// so gs_scene_create returns while the copy is still in flight and the first render's launch overlaps it.
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <pthread.h>

#include "gs_host.h"

namespace gs {

namespace {
thread_local std::string g_err;
}
int fail(int code, const std::string& msg) { g_err = msg; return code; }
const char* last_error_cstr() { return g_err.c_str(); }

unsigned host_threads() {
	static const unsigned n = [] {
		unsigned t = std::thread::hardware_concurrency();
		if (const char* e = std::getenv("GS_HOST_THREADS")) t = (unsigned)std::max(1, std::atoi(e));
		return t > 32u ? 32u : (t ? t : 1u);
	}();
	return n;
}

// ---- the worker pool behind parallel_for --------------------------------------------------------------------
namespace {
class HostPool {
public:
	explicit HostPool(unsigned helpers) {
		for (unsigned i = 0; i < helpers; ++i) workers_.emplace_back([this] { work(); });
	}
	~HostPool() {
		{ std::lock_guard<std::mutex> lk(m_); stop_ = true; ++generation_; }
		wake_.notify_all();
		for (auto& t : workers_) t.join();
	}
	// false: busy (another loop is running), the caller should run its loop itself
	bool run(size_t n, size_t grain, void (*fn)(void*, size_t, size_t), void* ctx) {
		std::unique_lock<std::mutex> busy(busy_, std::try_to_lock);
		if (!busy.owns_lock()) return false;
		{
			std::lock_guard<std::mutex> lk(m_);
			fn_ = fn; ctx_ = ctx; n_ = n; grain_ = grain;
			next_.store(0, std::memory_order_relaxed);
			pending_ = (unsigned)workers_.size();
			++generation_;
		}
		wake_.notify_all();
		drain();
		std::unique_lock<std::mutex> lk(m_);
		done_.wait(lk, [this] { return pending_ == 0; });
		return true;
	}

private:
	void drain() {
		for (;;) {
			const size_t lo = next_.fetch_add(grain_, std::memory_order_relaxed);
			if (lo >= n_) break;
			fn_(ctx_, lo, lo + grain_ < n_ ? lo + grain_ : n_);
		}
	}
	void work() {
		uint64_t seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> lk(m_);
				wake_.wait(lk, [&] { return generation_ != seen; });
				seen = generation_;
				if (stop_) return;
			}
			drain();
			{
				std::lock_guard<std::mutex> lk(m_);
				if (--pending_ == 0) done_.notify_one();
			}
		}
	}
	std::vector<std::thread> workers_;
	std::mutex m_, busy_;
	std::condition_variable wake_, done_;
	uint64_t generation_ = 0;
	bool stop_ = false;
	unsigned pending_ = 0;
	void (*fn_)(void*, size_t, size_t) = nullptr;
	void* ctx_ = nullptr;
	size_t n_ = 0, grain_ = 1;
	std::atomic<size_t> next_{0};
};
}  // namespace

namespace { std::atomic<bool> g_forked{false}; }
void pool_run(size_t n, size_t grain, void (*fn)(void*, size_t, size_t), void* ctx) {
	// Never destroyed: its threads sleep on its condition variable until the process ends.  A forked child has the object
	// but not the threads, and runs its loops on the calling thread.
	static HostPool* const pool = [] {
		pthread_atfork(nullptr, nullptr, [] { g_forked.store(true); });
		return new HostPool(host_threads() - 1);  // host_threads() >= 2 here (parallel_for runs small or single-threaded loops itself)
	}();
	if (g_forked.load(std::memory_order_relaxed) || !pool->run(n, grain, fn, ctx)) fn(ctx, 0, n);
}

cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t stream) {
	static std::mutex mu;
	static bool tuned[64] = {false};
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	{
		std::lock_guard<std::mutex> lock(mu);
		if (dev < 64 && !tuned[dev]) {
			cudaMemPool_t pool;
			e = cudaDeviceGetDefaultMemPool(&pool, dev);
			if (e != cudaSuccess) return e;
			unsigned long long keep = ~0ull;
			e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
			if (e != cudaSuccess) return e;
			tuned[dev] = true;
		}
	}
	return cudaMallocAsync(p, bytes ? bytes : 1, stream);
}
void dev_free(void* p, cudaStream_t stream) { if (p) cudaFreeAsync(p, stream); }

// ---- page-locked staging pool --------------------------------------------------------------------------
namespace {
struct StagingSlot { uint8_t* p = nullptr; size_t cap = 0; bool in_use = false; cudaEvent_t busy = nullptr; int busy_device = -1; };
std::mutex g_staging_mu;
std::vector<StagingSlot> g_staging;
constexpr size_t kStagingKeep = 6;  // blocks kept for reuse
}  // namespace

cudaError_t staging_acquire(size_t bytes, Staging& out) {
	bytes = bytes ? bytes : 1;
	std::lock_guard<std::mutex> lock(g_staging_mu);
	int best = -1;
	for (size_t i = 0; i < g_staging.size(); ++i) {
		StagingSlot& s = g_staging[i];
		if (s.in_use || s.cap < bytes) continue;
		if (s.busy && cudaEventQuery(s.busy) != cudaSuccess) { cudaGetLastError(); continue; }  // a copy out of it is still in flight
		if (best < 0 || s.cap < g_staging[best].cap) best = (int)i;
	}
	if (best < 0) {
		// drop an idle block that is too small rather than growing without bound
		for (size_t i = 0; i < g_staging.size() && g_staging.size() >= kStagingKeep; ++i) {
			StagingSlot& s = g_staging[i];
			if (s.in_use || (s.busy && cudaEventQuery(s.busy) != cudaSuccess)) { cudaGetLastError(); continue; }
			cudaFreeHost(s.p);
			g_staging.erase(g_staging.begin() + (long)i);
			break;
		}
		StagingSlot s;
		const size_t cap = (bytes + (4u << 20) - 1) & ~((size_t)(4u << 20) - 1);
		cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&s.p), cap, cudaHostAllocPortable);
		if (e != cudaSuccess) return e;
		s.cap = cap;
		g_staging.push_back(s);
		best = (int)g_staging.size() - 1;
	}
	g_staging[best].in_use = true;
	g_staging[best].busy = nullptr;
	out.p = g_staging[best].p; out.cap = g_staging[best].cap; out.slot = best;
	return cudaSuccess;
}

void staging_release(Staging& st, cudaEvent_t busy_until) {
	if (!st.p) return;
	std::lock_guard<std::mutex> lock(g_staging_mu);
	for (StagingSlot& s : g_staging)
		if (s.p == st.p) { s.in_use = false; s.busy = busy_until; break; }
	st = Staging{};
}
// a scene that goes away takes its event with it: forget it (the scene has synchronised the device by then)
void staging_forget_event(cudaEvent_t ev) {
	std::lock_guard<std::mutex> lock(g_staging_mu);
	for (StagingSlot& s : g_staging)
		if (s.busy == ev) s.busy = nullptr;
}

namespace {

// the neighbouring floats of a FINITE float, on the bit pattern (std::nextafterf is a library call, and the node builders
// round some ten million box planes outward per scene)
inline float next_below(float f) {
	uint32_t u; std::memcpy(&u, &f, 4);
	if ((u << 1) == 0u) u = 0x80000001u;      // +-0 -> the smallest negative subnormal
	else if (u >> 31) ++u;                      // negative: away from zero
	else --u;                                   // positive: towards zero
	std::memcpy(&f, &u, 4); return f;
}
inline float next_above(float f) { return -next_below(-f); }
float float_down(double x) { float f = (float)x; if ((double)f > x) f = std::isfinite(f) ? next_below(f) : std::nextafterf(f, -INFINITY); return f; }
float float_up(double x) { float f = (float)x; if ((double)f < x) f = std::isfinite(f) ? next_above(f) : std::nextafterf(f, INFINITY); return f; }
bool finite3(const double* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

DSphere make_dsphere(const gs_sphere& s) {
	DSphere d;
	d.cx = s.center[0]; d.cy = s.center[1]; d.cz = s.center[2];
	d.r = s.radius;  // signed: outward_normal = (p - c) / radius (sphere.rs:91)
	return d;
}

// Quad::new (quad.rs:25-38) + Plane::new (plane.rs:15-18), then the two barycentric functionals
// alpha = w.(planar x v) = planar.(v x w), beta = w.(u x planar) = planar.(w x u) (quad.rs:88-89).
DQuad make_dquad(const gs_quad& g) {
	const double* q = g.q; const double* u = g.u; const double* v = g.v;
	double n[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
	double nn = n[0] * n[0] + n[1] * n[1] + n[2] * n[2];
	double len = std::sqrt(nn);
	double normal[3] = {n[0] / len, n[1] / len, n[2] / len};
	double w[3] = {n[0] / nn, n[1] / nn, n[2] / nn};
	DQuad d;
	d.nx = normal[0]; d.ny = normal[1]; d.nz = normal[2];
	d.d = normal[0] * q[0] + normal[1] * q[1] + normal[2] * q[2];
	d.ax = v[1] * w[2] - v[2] * w[1]; d.ay = v[2] * w[0] - v[0] * w[2]; d.az = v[0] * w[1] - v[1] * w[0];
	d.bx = w[1] * u[2] - w[2] * u[1]; d.by = w[2] * u[0] - w[0] * u[2]; d.bz = w[0] * u[1] - w[1] * u[0];
	d.qa = q[0] * d.ax + q[1] * d.ay + q[2] * d.az;
	d.qb = q[0] * d.bx + q[1] * d.by + q[2] * d.bz;
	return d;
}

// Triangle::new (triangle.rs:20-28) keeps a, b, c; the two edges hit() recomputes per call are stored instead.
DTriangle make_dtriangle(const gs_triangle& t) {
	DTriangle d;
	d.ax = t.a[0]; d.ay = t.a[1]; d.az = t.a[2]; d.pad0 = 0.0;
	d.e1x = t.c[0] - t.a[0]; d.e1y = t.c[1] - t.a[1]; d.e1z = t.c[2] - t.a[2]; d.pad1 = 0.0;  // edge_1 = c - a
	d.e2x = t.b[0] - t.a[0]; d.e2y = t.b[1] - t.a[1]; d.e2z = t.b[2] - t.a[2]; d.pad2 = 0.0;  // edge_2 = b - a
	return d;
}

// first failing index of a parallel check (deterministic message: the smallest one)
template <class F> long first_bad(size_t n, F&& bad) {
	std::atomic<long> worst{-1};
	parallel_for(n, [&](size_t i) {
		if (bad(i)) {
			long cur = worst.load(std::memory_order_relaxed);
			while ((cur < 0 || (long)i < cur) && !worst.compare_exchange_weak(cur, (long)i)) {}
		}
	});
	return worst.load();
}

// Per-sphere flags (bit 0 moving, bit 1 baked with a RotateY) gathered by validate_desc's pass over the sphere array and
// used by layout_scene on the same thread right after: a compact table the layout can index at random without touching
// the 100-byte sphere records again.  Keyed by the array it was made from; layout_scene rebuilds it if the key differs.
thread_local std::vector<uint8_t> t_sphere_flags;
thread_local const gs_sphere* t_sphere_flags_of = nullptr;
inline uint8_t sphere_flags_of(const gs_sphere& s) {
	return (uint8_t)((s.moving ? 1 : 0) | ((s.uv_rotation[0] != 1.0 || s.uv_rotation[1] != 0.0) ? 2 : 0));
}

bool trace_enabled() { static const bool t = std::getenv("GS_TRACE") != nullptr; return t; }
struct Lap {
	const char* who;
	std::chrono::steady_clock::time_point prev = std::chrono::steady_clock::now();
	explicit Lap(const char* w) : who(w) {}
	void operator()(const char* what) {
		if (!trace_enabled()) return;
		auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[%s] %-34s %8.3f ms\n", who, what, std::chrono::duration<double, std::milli>(now - prev).count());
		prev = now;
	}
};

}  // namespace

// ---- validation: nothing aborts across the boundary, so everything the layout and the kernels index is checked here ----
int validate_desc(const gs_scene_desc* d) {
	Lap lap("validate_desc");
	if (!d) return fail(GS_ERR_INVALID, "null scene description");
	if (d->abi_version != GS_ABI_VERSION) return fail(GS_ERR_INVALID, "abi_version mismatch");
	if (d->n_spheres + d->n_quads + d->n_triangles == 0) return fail(GS_ERR_INVALID, "empty scene (the reference recurses forever on an empty list, BVH.rs:18-65)");
	if (d->n_spheres >= (1u << 25) || d->n_quads >= (1u << 25) || d->n_triangles >= (1u << 25)) return fail(GS_ERR_UNSUPPORTED, "too many primitives (2^25 per kind)");
	if (d->n_prim_refs >= (1u << 27) - 16u) return fail(GS_ERR_UNSUPPORTED, "too many primitive references");
	if (d->n_nodes >= (1u << 30)) return fail(GS_ERR_UNSUPPORTED, "too many BVH nodes");
	// every array whose count is non-zero must be there
	if (d->n_spheres && !d->spheres) return fail(GS_ERR_INVALID, "null sphere array");
	if (d->n_quads && !d->quads) return fail(GS_ERR_INVALID, "null quad array");
	if (d->n_triangles && !d->triangles) return fail(GS_ERR_INVALID, "null triangle array");
	if (d->n_nodes == 0 || !d->nodes) return fail(GS_ERR_INVALID, "no BVH nodes");
	if (d->n_prim_refs == 0 || !d->prim_refs) return fail(GS_ERR_INVALID, "no primitive references");
	if (d->n_boundary_spheres && !d->boundary_spheres) return fail(GS_ERR_INVALID, "null boundary sphere array");
	if (d->n_boundary_quads && !d->boundary_quads) return fail(GS_ERR_INVALID, "null boundary quad array");
	if (d->n_media && !d->media) return fail(GS_ERR_INVALID, "null media array");
	if (d->n_materials == 0 || !d->materials) return fail(GS_ERR_INVALID, "no materials");
	if (d->n_textures && !d->textures) return fail(GS_ERR_INVALID, "null texture array");
	if (d->n_images && !d->images) return fail(GS_ERR_INVALID, "null image array");
	auto bad_mat = [&](int m) { return m < 0 || (size_t)m >= d->n_materials; };
	long b;
	t_sphere_flags.resize(d->n_spheres);
	t_sphere_flags_of = d->spheres;
	uint8_t* sphere_flags = t_sphere_flags.data();
	b = first_bad(d->n_spheres, [&](size_t i) {
		const gs_sphere& s = d->spheres[i];
		sphere_flags[i] = sphere_flags_of(s);
		return !finite3(s.center) || !finite3(s.path) || !std::isfinite(s.radius) || s.radius == 0.0 || !std::isfinite(s.uv_rotation[0]) || !std::isfinite(s.uv_rotation[1]) || bad_mat(s.material);
	});
	if (b >= 0) return fail(GS_ERR_INVALID, "sphere " + std::to_string(b) + ": non-finite, zero radius or bad material index");
	b = first_bad(d->n_quads, [&](size_t i) { const gs_quad& q = d->quads[i]; return !finite3(q.q) || !finite3(q.u) || !finite3(q.v) || bad_mat(q.material); });
	if (b >= 0) return fail(GS_ERR_INVALID, "quad " + std::to_string(b) + ": non-finite or bad material index");
	b = first_bad(d->n_triangles, [&](size_t i) { const gs_triangle& t = d->triangles[i]; return !finite3(t.a) || !finite3(t.b) || !finite3(t.c) || bad_mat(t.material); });
	if (b >= 0) return fail(GS_ERR_INVALID, "triangle " + std::to_string(b) + ": non-finite or bad material index");
	for (size_t i = 0; i < d->n_boundary_spheres; ++i) {
		const gs_sphere& s = d->boundary_spheres[i];
		if (!finite3(s.center) || !finite3(s.path) || !std::isfinite(s.radius) || s.radius == 0.0) return fail(GS_ERR_INVALID, "boundary sphere " + std::to_string(i) + ": non-finite or zero radius");
	}
	for (size_t i = 0; i < d->n_boundary_quads; ++i) {
		const gs_quad& q = d->boundary_quads[i];
		if (!finite3(q.q) || !finite3(q.u) || !finite3(q.v)) return fail(GS_ERR_INVALID, "boundary quad " + std::to_string(i) + ": non-finite");
	}
	for (size_t i = 0; i < d->n_media; ++i) {
		const gs_medium& m = d->media[i];
		if (m.sphere_first < 0 || m.sphere_count < 0 || (uint64_t)m.sphere_first + (uint64_t)m.sphere_count > d->n_boundary_spheres ||
		    m.quad_first < 0 || m.quad_count < 0 || (uint64_t)m.quad_first + (uint64_t)m.quad_count > d->n_boundary_quads)
			return fail(GS_ERR_INVALID, "medium " + std::to_string(i) + ": boundary range out of bounds");
		if (!(m.density != 0.0) || !std::isfinite(m.density)) return fail(GS_ERR_INVALID, "medium: bad density");
		if (bad_mat(m.phase_material)) return fail(GS_ERR_INVALID, "medium: bad phase material");
	}
	// (one material and one texture per sphere is a common way to build a scene: these tables can have 10^6 rows)
	auto material_fault = [&](size_t i) -> int {
		const gs_material& m = d->materials[i];
		if (m.kind < 0 || m.kind > GS_MAT_ISOTROPIC) return 1;
		const bool needs_tex = m.kind == GS_MAT_LAMBERTIAN || m.kind == GS_MAT_DIFFUSE_LIGHT || m.kind == GS_MAT_ISOTROPIC;
		if (needs_tex && (m.texture < 0 || (size_t)m.texture >= d->n_textures)) return 2;
		return 0;
	};
	b = first_bad(d->n_materials, [&](size_t i) { return material_fault(i) != 0; });
	if (b >= 0) return fail(GS_ERR_INVALID, "material " + std::to_string(b) + (material_fault((size_t)b) == 1 ? ": bad kind" : ": bad texture index"));
	auto texture_fault = [&](size_t i) -> int {
		const gs_texture& t = d->textures[i];
		if (t.kind < 0 || t.kind > GS_TEX_NOISE) return 1;
		// children must come earlier in the table: rules out cycles (Arc graphs are acyclic)
		if (t.kind == GS_TEX_CHECKER && (t.a < 0 || t.b < 0 || (size_t)t.a >= i || (size_t)t.b >= i)) return 2;
		if (t.kind == GS_TEX_IMAGE && (t.a < 0 || (size_t)t.a >= d->n_images || d->images[t.a].format != GS_IMG_U8_RGB)) return 3;
		if (t.kind == GS_TEX_NOISE && !d->perlin_perm) return 4;
		return 0;
	};
	b = first_bad(d->n_textures, [&](size_t i) { return texture_fault(i) != 0; });
	if (b >= 0) {
		static const char* msg[] = {"", "texture: bad kind", "checker texture: children must precede it", "image texture: bad image index/format", "noise texture without perlin_perm"};
		return fail(GS_ERR_INVALID, std::string(msg[texture_fault((size_t)b)]) + " (texture " + std::to_string(b) + ")");
	}
	for (size_t i = 0; i < d->n_images; ++i) {
		const gs_image& im = d->images[i];
		if (im.width <= 0 || im.height <= 0 || !im.pixels) return fail(GS_ERR_INVALID, "image: bad size or null pixels");
		if (im.format != GS_IMG_U8_RGB && im.format != GS_IMG_F32_RGB) return fail(GS_ERR_INVALID, "image: bad format");
	}
	auto node_fault = [&](size_t i) -> int {
		const gs_bvh_node& n = d->nodes[i];
		if (!finite3(n.bounds_min) || !finite3(n.bounds_max)) return 1;
		if (n.ref_count < 0) return 2;  // would get neither a node index nor a leaf code
		if (n.ref_count > 0) {
			if (n.ref_count > 15) return 3;
			if (n.first_ref < 0 || (uint64_t)n.first_ref + (uint64_t)n.ref_count > d->n_prim_refs) return 4;
		} else if (n.left <= (int)i || n.right <= (int)i || (size_t)n.left >= d->n_nodes || (size_t)n.right >= d->n_nodes || n.left == n.right) return 5;
		return 0;
	};
	b = first_bad(d->n_nodes, [&](size_t i) { return node_fault(i) != 0; });
	if (b >= 0) {
		static const char* msg[] = {"", "non-finite bounds (the reference panics, BVH.rs:54)", "negative ref_count", "leaf with more than 15 primitives",
		                            "leaf reference range out of bounds", "children must be two distinct nodes with larger indices"};
		const int why = node_fault((size_t)b);
		return fail(why == 3 ? GS_ERR_UNSUPPORTED : GS_ERR_INVALID, "BVH node " + std::to_string(b) + ": " + msg[why]);
	}
	b = first_bad(d->n_prim_refs, [&](size_t i) {
		uint32_t r = d->prim_refs[i], kind = r >> 30, idx = r & kRefIndexMask;
		return kind > 2 || (kind == 0 && idx >= d->n_spheres) || (kind == 1 && idx >= d->n_quads) || (kind == 2 && idx >= d->n_triangles);
	});
	if (b >= 0) return fail(GS_ERR_INVALID, "bad primitive reference " + std::to_string(b));
	lap("all checks");
	return GS_OK;
}

// ---- the eight-wide topology ------------------------------------------------------------------------------
// Every binary node that stays internal becomes a wide node: it starts with its two children and keeps replacing the
// internal child of largest surface area by that child's own two children until it has eight (or only leaves).  The
// children then take the slots whose octant (bit a of the slot = upper side of axis a) matches the position of their
// centre relative to the node's centre best (greedy over the 8 x 8 scores), so that (slot XOR ray octant) orders them
// near-to-far.  Internal children are numbered in slot order, breadth-first: child_base + rank among the internal slots.
namespace {
struct WideBuild {
	std::vector<int> kids;        // 8 per wide node, ABI node index or -1
	std::vector<int> child_base;  // wide index of the first internal kid
	std::vector<int> of_bin;      // wide node -> the binary node it replaces
	std::vector<size_t> level_begin;  // wide nodes of level l are [level_begin[l], level_begin[l + 1]) (breadth-first numbering)
	int depth = 0;                // levels of the eight-wide tree
	int binary_depth = 0;         // deepest node of the binary tree (root = 0)
};

void build_wide_topology(const gs_bvh_node* N, size_t n_nodes, WideBuild& w) {
	// surface measure of a node's box, from the node record itself (a separate area array would be one more cache miss per
	// candidate: this pass is bound by the latency of its scattered reads of the 2 * n_prims node records)
	auto area_of = [&](int a) {
		const double dx = N[a].bounds_max[0] - N[a].bounds_min[0], dy = N[a].bounds_max[1] - N[a].bounds_min[1], dz = N[a].bounds_max[2] - N[a].bounds_min[2];
		return (float)(dx * dy + dy * dz + dz * dx);
	};
	auto prefetch_node = [&](int a) {
		const char* p = reinterpret_cast<const char*>(N + a);
		__builtin_prefetch(p); __builtin_prefetch(p + sizeof(gs_bvh_node) - 1);
	};
	auto internal = [&](int a) { return N[a].ref_count == 0; };
	// experiment knob: fewer children per node (the slots stay eight)
	static const int arity = [] { const char* e = std::getenv("GS_WIDE_ARITY"); const int v = e ? std::atoi(e) : 8; return v < 2 ? 2 : (v > 8 ? 8 : v); }();
	// one wide node: collapse + octant slots; returns the number of internal kids
	// (depth_in = binary depth of node a; kid_depth[] = binary depths of the kids, in slot order)
	auto make_node = [&](int a, int depth_in, int* slot_kid, int* slot_depth) -> int {
		int kids[8] = {N[a].left, N[a].right, -1, -1, -1, -1, -1, -1};
		int kd[8] = {depth_in + 1, depth_in + 1, 0, 0, 0, 0, 0, 0};
		float ka[8];  // area of an internal kid, -1 for a leaf (never opened)
		auto note = [&](int k) {
			const int c = kids[k];
			if (internal(c)) { ka[k] = area_of(c); prefetch_node(N[c].left); prefetch_node(N[c].right); }
			else ka[k] = -1.0f;
		};
		prefetch_node(kids[0]); prefetch_node(kids[1]);
		note(0); note(1);
		int n = 2;
		while (n < arity) {
			int best = -1;
			for (int k = 0; k < n; ++k)
				if (ka[k] >= 0.0f && (best < 0 || ka[k] > ka[best])) best = k;
			if (best < 0) break;
			const int c = kids[best];
			kids[best] = N[c].left;
			kids[n] = N[c].right;
			kd[best] += 1; kd[n] = kd[best];
			note(best); note(n);
			++n;
		}
		double centre[3];
		for (int x = 0; x < 3; ++x) centre[x] = 0.5 * (N[a].bounds_min[x] + N[a].bounds_max[x]);
		double score[8][8];
		for (int k = 0; k < n; ++k) {
			double rel[3];
			for (int x = 0; x < 3; ++x) rel[x] = 0.5 * (N[kids[k]].bounds_min[x] + N[kids[k]].bounds_max[x]) - centre[x];
			for (int s = 0; s < 8; ++s) score[k][s] = ((s & 1) ? rel[0] : -rel[0]) + ((s & 2) ? rel[1] : -rel[1]) + ((s & 4) ? rel[2] : -rel[2]);
		}
		for (int s = 0; s < 8; ++s) slot_kid[s] = -1;
		// greedy assignment: repeatedly the best (kid, free slot) pair, ties to the lowest kid and then the lowest slot.
		// Each unplaced kid remembers its best free slot; only the kids that lose theirs look again.
		int best_slot[8];
		auto find_slot = [&](int k) {
			int bs = -1;
			for (int s = 0; s < 8; ++s)
				if (slot_kid[s] < 0 && (bs < 0 || score[k][s] > score[k][bs])) bs = s;
			best_slot[k] = bs;
		};
		for (int k = 0; k < n; ++k) find_slot(k);
		uint32_t unplaced = (1u << n) - 1u;
		while (unplaced) {
			int bk = -1;
			for (int k = 0; k < n; ++k)
				if ((unplaced >> k & 1u) && (bk < 0 || score[k][best_slot[k]] > score[bk][best_slot[bk]])) bk = k;
			const int bs = best_slot[bk];
			slot_kid[bs] = kids[bk];
			slot_depth[bs] = kd[bk];
			unplaced &= ~(1u << bk);
			for (int k = 0; k < n; ++k)
				if ((unplaced >> k & 1u) && best_slot[k] == bs) find_slot(k);
		}
		int n_int = 0;
		for (int s = 0; s < 8; ++s) if (slot_kid[s] >= 0 && internal(slot_kid[s])) ++n_int;
		return n_int;
	};
	// level by level: the nodes of a level are independent of each other (host threads), their internal kids -- in node
	// order, then slot order -- are the next level: the same numbering a breadth-first queue gives
	std::vector<int> level{0}, next, n_int, level_depth{0}, next_depth, kid_depth;
	std::atomic<int> deepest{0};
	size_t base = 0;
	while (!level.empty()) {
		const size_t L = level.size();
		w.level_begin.push_back(base);
		w.kids.resize((base + L) * 8);
		w.child_base.resize(base + L);
		w.of_bin.insert(w.of_bin.end(), level.begin(), level.end());
		n_int.assign(L, 0);
		kid_depth.assign(L * 8, 0);
		parallel_for(L, [&](size_t i) {
			if (i + 8 < L) prefetch_node(level[i + 8]);
			if (i + 4 < L) { const gs_bvh_node& f = N[level[i + 4]]; prefetch_node(f.left); prefetch_node(f.right); }  // every level entry is internal
			n_int[i] = make_node(level[i], level_depth[i], w.kids.data() + (base + i) * 8, kid_depth.data() + i * 8);
			int m = 0;
			for (int s = 0; s < 8; ++s) m = std::max(m, kid_depth[i * 8 + s]);
			int cur = deepest.load(std::memory_order_relaxed);
			while (m > cur && !deepest.compare_exchange_weak(cur, m)) {}
		}, 2048);
		size_t total = 0;
		for (size_t i = 0; i < L; ++i) { w.child_base[base + i] = n_int[i] ? (int)(base + L + total) : 0; total += (size_t)n_int[i]; }
		next.assign(total, 0);
		next_depth.assign(total, 0);
		parallel_for(L, [&](size_t i) {
			if (!n_int[i]) return;
			const size_t at = (size_t)w.child_base[base + i] - (base + L);
			int* out = next.data() + at;
			int* outd = next_depth.data() + at;
			const int* kid = w.kids.data() + (base + i) * 8;
			for (int s = 0; s < 8; ++s) if (kid[s] >= 0 && internal(kid[s])) { *out++ = kid[s]; *outd++ = kid_depth[i * 8 + s]; }
		}, 2048);
		base += L;
		level.swap(next);
		level_depth.swap(next_depth);
		++w.depth;
	}
	w.level_begin.push_back(base);
	w.binary_depth = deepest.load();
}
}  // namespace

// ---- re-layout + upload ---------------------------------------------------------------------------------
int layout_scene(gs_scene* s, const gs_scene_desc* d) {
	Lap lap("gs_scene_create");
	const size_t nN = d->n_nodes;
	const gs_bvh_node* N = d->nodes;
	// host copy of the tree (kept for re-padding and for the tree not built now): uninitialised storage, filled by helper
	// threads while this thread goes on with the caller's array (128 MB for 10^6 primitives: first-touch page faults dominate)
	s->h_nodes.reset(new gs_bvh_node[nN]);
	s->n_nodes = nN;
	std::vector<std::thread> copiers;
	{
		const unsigned T = std::max(1u, std::min(4u, host_threads() / 4));
		const size_t chunk = (nN + T - 1) / T;
		gs_bvh_node* dst = s->h_nodes.get();
		for (unsigned t = 0; t < T; ++t) {
			const size_t lo = (size_t)t * chunk, hi = std::min(nN, lo + chunk);
			if (lo < hi) copiers.emplace_back([=] { std::memcpy(dst + lo, N + lo, (hi - lo) * sizeof(gs_bvh_node)); });
		}
	}
	struct Join { std::vector<std::thread>& t; ~Join() { for (auto& x : t) if (x.joinable()) x.join(); } } join_copiers{copiers};
	lap("copy nodes (started)");
	// tree check: children carry larger indices than their parent (validated), so "every node but the root has exactly one
	// parent" already means one tree hanging off node 0.  (Its depth comes out of the topology pass below.)
	{
		std::unique_ptr<std::atomic<uint8_t>[]> parents(new std::atomic<uint8_t>[nN]);
		parallel_for(nN, [&](size_t i) { parents[i].store(0, std::memory_order_relaxed); });
		std::atomic<bool> twice{false};
		parallel_for(nN, [&](size_t i) {
			if (N[i].ref_count > 0) return;
			if (parents[N[i].left].fetch_add(1, std::memory_order_relaxed) != 0 || parents[N[i].right].fetch_add(1, std::memory_order_relaxed) != 0) twice = true;
		});
		if (twice.load()) return fail(GS_ERR_INVALID, "BVH is not a tree (node reached twice)");
		const long orphan = first_bad(nN, [&](size_t i) { return (i == 0) != (parents[i].load(std::memory_order_relaxed) == 0); });
		if (orphan >= 0) return fail(GS_ERR_INVALID, orphan == 0 ? "BVH is not a tree (the root has a parent)" : "BVH node " + std::to_string(orphan) + " is not reachable from the root");
	}
	lap("tree check");

	// ---- eight-wide topology (also fixes the order of the primitives on the device) ----
	WideBuild wb;
	const bool root_leaf = N[0].ref_count > 0;
	if (!root_leaf) build_wide_topology(N, nN, wb);
	if (wb.binary_depth > kMaxStack - 4) return fail(GS_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
	const size_t nW = wb.of_bin.size();
	lap("wide topology");

	// ---- device order of the primitives: depth-first over the wide tree; at a node first the primitives of its
	// single-primitive leaf children (addressable as base + offset from the node), then those of its larger leaves ----
	std::vector<int> sph_new, quad_new, tri_new;  // "placed already?" tables of the serial walk only
	std::vector<int> sph_src, quad_src, tri_src;
	std::vector<uint32_t> refs;
	s->h_leaf_code.reset(new int[nN]);     // uninitialised: every leaf's entry is written below, nothing else is read
	s->h_leaf_direct.reset(new int[nN]);
	s->w_prim_base.assign(nW * 3, 0u);
	bool any_moving = false, any_rot = false;
	std::atomic<bool> link_seen{false};
	if (t_sphere_flags_of != d->spheres || t_sphere_flags.size() != d->n_spheres) {  // not reached through gs_scene_create, which validates first
		t_sphere_flags.resize(d->n_spheres);
		uint8_t* f = t_sphere_flags.data();
		parallel_for(d->n_spheres, [&](size_t i) { f[i] = sphere_flags_of(d->spheres[i]); });
		t_sphere_flags_of = d->spheres;
	}
	const uint8_t* sflag = t_sphere_flags.data();
	auto emit_leaf = [&](int leaf) {
		const gs_bvh_node& nd = N[leaf];
		const uint32_t first = (uint32_t)refs.size();
		bool fresh = false;
		for (int k = 0; k < nd.ref_count; ++k) {
			const uint32_t r = d->prim_refs[nd.first_ref + k], kind = r >> 30, idx = r & kRefIndexMask;
			if (kind == 0) {
				if (sph_new[idx] < 0) { sph_new[idx] = (int)sph_src.size(); sph_src.push_back((int)idx); fresh = true; }
				const bool moving = (sflag[idx] & 1) != 0;
				any_moving |= moving;
				refs.push_back(((moving ? kKindMovingSphere : kKindSphere) << kRefKindShift) | (uint32_t)sph_new[idx]);
			} else if (kind == 1) {
				if (quad_new[idx] < 0) { quad_new[idx] = (int)quad_src.size(); quad_src.push_back((int)idx); fresh = true; }
				refs.push_back((kKindQuad << kRefKindShift) | (uint32_t)quad_new[idx]);
			} else {
				if (tri_new[idx] < 0) { tri_new[idx] = (int)tri_src.size(); tri_src.push_back((int)idx); fresh = true; }
				refs.push_back((kKindTriangle << kRefKindShift) | (uint32_t)tri_new[idx]);
			}
		}
		s->h_leaf_code[leaf] = leaf_code(first, (uint32_t)nd.ref_count);
		s->h_leaf_direct[leaf] = -1;
		if (nd.ref_count == 1) {
			// single-primitive leaf: (kind << 25 | index) in the link itself, count field 0 -- no dependent load from `refs`
			const uint32_t r = refs.back(), kind = r >> kRefKindShift, idx = r & kRefIndexMask;
			s->h_leaf_code[leaf] = leaf_code((kind << 25) | idx, 0u);
			if (fresh && kind != kKindMovingSphere) s->h_leaf_direct[leaf] = (int)idx;
		}
		if (s->h_leaf_direct[leaf] < 0) link_seen.store(true, std::memory_order_relaxed);
	};
	// Does every primitive sit in exactly one leaf?  (True for any tree without spatial splits.)  Then the order below can be
	// computed from subtree sizes and filled by all host threads; otherwise the serial walk with its "seen before" test.
	bool unique_refs = !root_leaf && d->n_prim_refs == d->n_spheres + d->n_quads + d->n_triangles;
	if (unique_refs) {
		std::vector<std::atomic<uint8_t>> seen_s(d->n_spheres), seen_q(d->n_quads), seen_t(d->n_triangles);
		std::atomic<bool> dup{false};
		parallel_for(d->n_prim_refs, [&](size_t i) {
			const uint32_t r = d->prim_refs[i], kind = r >> 30, idx = r & kRefIndexMask;
			std::atomic<uint8_t>& f = kind == 0 ? seen_s[idx] : (kind == 1 ? seen_q[idx] : seen_t[idx]);
			if (f.exchange(1, std::memory_order_relaxed)) dup = true;
		});
		unique_refs = !dup.load();
	}
	lap("  order: unique-reference check");
	if (root_leaf || !unique_refs) {
		sph_new.assign(d->n_spheres, -1); quad_new.assign(d->n_quads, -1); tri_new.assign(d->n_triangles, -1);
		sph_src.reserve(d->n_spheres); quad_src.reserve(d->n_quads); tri_src.reserve(d->n_triangles);
		refs.reserve(d->n_prim_refs);
	}
	if (root_leaf) emit_leaf(0);
	else if (unique_refs) {
		// per wide node: primitives (per kind) and references of its own leaf kids, singles and larger leaves apart
		struct Own { uint32_t single[3], multi[3], refs_single, refs_multi; };
		std::vector<Own> own(nW);
		parallel_for(nW, [&](size_t h) {
			Own o{};
			const int* kids = wb.kids.data() + 8 * h;
			for (int k = 0; k < 8; ++k) {
				const int c = kids[k];
				if (c < 0 || N[c].ref_count == 0) continue;
				for (int j = 0; j < N[c].ref_count; ++j) {
					const uint32_t kind = d->prim_refs[N[c].first_ref + j] >> 30;
					if (N[c].ref_count == 1) ++o.single[kind]; else ++o.multi[kind];
				}
				if (N[c].ref_count == 1) ++o.refs_single; else o.refs_multi += (uint32_t)N[c].ref_count;
			}
			own[h] = o;
		}, 4096);
		lap("  order: own counts");
		// subtree totals bottom-up (kids carry larger indices), then start offsets top-down in the walk's order:
		// own singles, own larger leaves, then the internal kids' subtrees in slot order
		struct Tot { uint32_t prim[3], refs; };
		std::vector<Tot> sub(nW), start(nW);
		// (level by level: the nodes of a level only read the level below / write the level below, so each level is a parallel loop)
		auto subtree_total = [&](size_t h) {
			Tot t{{own[h].single[0] + own[h].multi[0], own[h].single[1] + own[h].multi[1], own[h].single[2] + own[h].multi[2]}, own[h].refs_single + own[h].refs_multi};
			const int* kids = wb.kids.data() + 8 * h;
			int rank = 0;
			for (int k = 0; k < 8; ++k) {
				if (kids[k] < 0 || N[kids[k]].ref_count != 0) continue;
				const Tot& c = sub[(size_t)wb.child_base[h] + rank++];
				for (int x = 0; x < 3; ++x) t.prim[x] += c.prim[x];
				t.refs += c.refs;
			}
			sub[h] = t;
		};
		auto place_children = [&](size_t h) {
			Tot cur = start[h];
			for (int x = 0; x < 3; ++x) cur.prim[x] += own[h].single[x] + own[h].multi[x];
			cur.refs += own[h].refs_single + own[h].refs_multi;
			const int* kids = wb.kids.data() + 8 * h;
			int rank = 0;
			for (int k = 0; k < 8; ++k) {
				if (kids[k] < 0 || N[kids[k]].ref_count != 0) continue;
				const size_t c = (size_t)wb.child_base[h] + rank++;
				start[c] = cur;
				for (int x = 0; x < 3; ++x) cur.prim[x] += sub[c].prim[x];
				cur.refs += sub[c].refs;
			}
		};
		const size_t n_levels = wb.level_begin.size() - 1;
		for (size_t l = n_levels; l-- > 0;) {
			const size_t lo = wb.level_begin[l], cnt = wb.level_begin[l + 1] - lo;
			parallel_for(cnt, [&](size_t i) { subtree_total(lo + i); }, 4096);
		}
		start[0] = Tot{{0, 0, 0}, 0};
		for (size_t l = 0; l < n_levels; ++l) {
			const size_t lo = wb.level_begin[l], cnt = wb.level_begin[l + 1] - lo;
			parallel_for(cnt, [&](size_t i) { place_children(lo + i); }, 4096);
		}
		lap("  order: subtree offsets");
		sph_src.assign(d->n_spheres, 0); quad_src.assign(d->n_quads, 0); tri_src.assign(d->n_triangles, 0);
		refs.assign(d->n_prim_refs, 0u);
		std::atomic<bool> moving_seen{false};
		parallel_for(nW, [&](size_t h) {
			uint32_t next_prim[3] = {start[h].prim[0], start[h].prim[1], start[h].prim[2]};
			uint32_t next_ref = start[h].refs;
			for (int x = 0; x < 3; ++x) s->w_prim_base[3 * h + x] = next_prim[x];
			const int* kids = wb.kids.data() + 8 * h;
			if (h + 3 < nW)
				for (int k = 0; k < 8; ++k) if (kids[24 + k] >= 0) __builtin_prefetch(N + kids[24 + k]);
			for (int pass = 0; pass < 2; ++pass)
				for (int k = 0; k < 8; ++k) {
					const int c = kids[k];
					if (c < 0 || N[c].ref_count == 0 || (N[c].ref_count == 1) != (pass == 0)) continue;
					const uint32_t first = next_ref;
					uint32_t last_ref = 0;
					for (int j = 0; j < N[c].ref_count; ++j) {
						const uint32_t r = d->prim_refs[N[c].first_ref + j], kind = r >> 30, idx = r & kRefIndexMask;
						const uint32_t dev = next_prim[kind]++;
						uint32_t dkind = kind;
						if (kind == 0) {
							sph_src[dev] = (int)idx;
							if (sflag[idx] & 1) { dkind = kKindMovingSphere; moving_seen = true; }
						} else if (kind == 1) quad_src[dev] = (int)idx;
						else tri_src[dev] = (int)idx;
						last_ref = (dkind << kRefKindShift) | dev;
						refs[next_ref++] = last_ref;
					}
					s->h_leaf_code[c] = leaf_code(first, (uint32_t)N[c].ref_count);
					s->h_leaf_direct[c] = -1;
					if (N[c].ref_count == 1) {
						const uint32_t kind = last_ref >> kRefKindShift, idx = last_ref & kRefIndexMask;
						s->h_leaf_code[c] = leaf_code((kind << 25) | idx, 0u);
						if (kind != kKindMovingSphere) s->h_leaf_direct[c] = (int)idx;
					}
					if (s->h_leaf_direct[c] < 0) link_seen.store(true, std::memory_order_relaxed);
				}
		}, 4096);
		any_moving = moving_seen.load();
	} else {
		std::vector<int> stack{0};
		while (!stack.empty()) {
			const int h = stack.back(); stack.pop_back();
			s->w_prim_base[3 * (size_t)h + 0] = (uint32_t)sph_src.size();
			s->w_prim_base[3 * (size_t)h + 1] = (uint32_t)quad_src.size();
			s->w_prim_base[3 * (size_t)h + 2] = (uint32_t)tri_src.size();
			const int* kids = wb.kids.data() + 8 * (size_t)h;
			for (int pass = 0; pass < 2; ++pass)
				for (int k = 0; k < 8; ++k) {
					const int c = kids[k];
					if (c < 0 || N[c].ref_count == 0) continue;
					if ((N[c].ref_count == 1) == (pass == 0)) emit_leaf(c);
				}
			int rank = 0;
			for (int k = 0; k < 8; ++k) if (kids[k] >= 0 && N[kids[k]].ref_count == 0) ++rank;
			for (int k = 7, r = rank; k >= 0; --k)  // first internal child on top of the stack
				if (kids[k] >= 0 && N[kids[k]].ref_count == 0) stack.push_back(wb.child_base[h] + --r);
		}
	}
	{
		uint8_t all = 0;  // over every sphere of the description, placed or not: the table is only an optimisation
		for (size_t i = 0; i < d->n_spheres; ++i) all |= sflag[i];
		any_rot = (all & 2) != 0;
	}
	s->w_any_link = link_seen.load();
	s->w_kids.swap(wb.kids);
	s->w_child_base.swap(wb.child_base);
	s->wide_depth = wb.depth;
	lap("primitive order");

	// ---- coordinate range for the fp32 box padding ----
	double mx = 0.0;
	{
		const unsigned T = host_threads();
		std::vector<double> part(T, 0.0);
		const size_t chunk = (nN + T - 1) / T;
		parallel_for(T, [&](size_t t) {
			double m = 0.0;
			for (size_t i = t * chunk; i < std::min(nN, (t + 1) * chunk); ++i)
				for (int a = 0; a < 3; ++a) m = std::fmax(m, std::fmax(std::fabs(N[i].bounds_min[a]), std::fabs(N[i].bounds_max[a])));
			part[t] = m;
		}, 2);
		for (double m : part) mx = std::fmax(mx, m);
	}
	for (size_t i = 0; i < d->n_boundary_spheres; ++i) {
		const gs_sphere& b = d->boundary_spheres[i];
		for (int a = 0; a < 3; ++a) mx = std::fmax(mx, std::fabs(b.center[a]) + std::fabs(b.path[a]) + std::fabs(b.radius));
	}
	for (size_t i = 0; i < d->n_boundary_quads; ++i) {
		const gs_quad& b = d->boundary_quads[i];
		for (int a = 0; a < 3; ++a) mx = std::fmax(mx, std::fabs(b.q[a]) + std::fabs(b.u[a]) + std::fabs(b.v[a]));
	}
	s->max_abs_coord = mx;

	// ---- one device allocation, one staging block, sections aligned to 256 bytes ----
	size_t total = 0;
	auto section = [&](size_t bytes) { const size_t off = total; total += (bytes + 255) & ~(size_t)255; return off; };
	const size_t nS = sph_src.size(), nQ = quad_src.size(), nT = tri_src.size();
	const size_t o_refs = section(refs.size() * 4);
	const size_t o_sph = section(nS * sizeof(DSphere)), o_sph_path = section(any_moving ? nS * sizeof(double4) : 0), o_sph_aux = section(nS * sizeof(int2));
	const size_t o_sph_rot = section(any_rot ? nS * sizeof(float2) : 0);
	const size_t o_quad = section(nQ * sizeof(DQuad)), o_quad_aux = section(nQ * sizeof(int2)), o_quad_order = section(nQ * sizeof(int));
	const size_t o_tri = section(nT * sizeof(DTriangle)), o_tri_aux = section(nT * sizeof(int2));
	const size_t o_bsph = section(d->n_boundary_spheres * sizeof(DSphere)), o_bpath = section(d->n_boundary_spheres * sizeof(double4));
	const size_t o_bquad = section(d->n_boundary_quads * sizeof(DQuad));
	const size_t o_media = section(d->n_media * sizeof(DMedium)), o_mat = section(d->n_materials * sizeof(DMaterial)), o_tex = section(d->n_textures * sizeof(DTexture));
	const size_t o_img = section(d->n_images * sizeof(DImage)), o_perm = section(d->perlin_perm ? 256 : 0);
	std::vector<size_t> o_texels(d->n_images);
	for (size_t i = 0; i < d->n_images; ++i) {
		const size_t px = (size_t)d->images[i].width * d->images[i].height;
		o_texels[i] = section(px * (d->images[i].format == GS_IMG_U8_RGB ? sizeof(uchar4) : sizeof(float4)));
	}
	const size_t o_counters = section(kCounterSlots * sizeof(unsigned long long)), o_stats = section(kStatSlots * sizeof(unsigned long long));

	if (trace_enabled())
		std::fprintf(stderr, "[gs_scene_create] blob %zu B: refs %zu, spheres %zu (+path %d, +uvrot %d), quads %zu, tris %zu, images %zu, nodes in %zu\n", total,
		             refs.size(), nS, (int)any_moving, (int)any_rot, nQ, nT, (size_t)d->n_images, nN);
	Staging stg;
	GS_CUDA(staging_acquire(total, stg));
	struct StageGuard { Staging& g; cudaEvent_t ev = nullptr; ~StageGuard() { staging_release(g, ev); } } stage_guard{stg};
	GS_CUDA(dev_alloc(&s->blob, total, s->up));
	s->blob_bytes = total;
	uint8_t* H = stg.p;
	uint8_t* D = static_cast<uint8_t*>(s->blob);
	lap("staging + device allocation");

	std::memcpy(H + o_refs, refs.data(), refs.size() * 4);
	{
		DSphere* sp = reinterpret_cast<DSphere*>(H + o_sph); double4* pa = reinterpret_cast<double4*>(H + o_sph_path);
		int2* au = reinterpret_cast<int2*>(H + o_sph_aux); float2* ro = reinterpret_cast<float2*>(H + o_sph_rot);
		parallel_for(nS, [&](size_t i) {
			if (i + 8 < nS) {  // the device order is a walk of the tree, not the caller's order: start a later record's misses now
				const char* nx = reinterpret_cast<const char*>(d->spheres + sph_src[i + 8]);
				__builtin_prefetch(nx); __builtin_prefetch(nx + sizeof(gs_sphere) - 1);
			}
			const gs_sphere& g = d->spheres[sph_src[i]];
			sp[i] = make_dsphere(g);
			if (any_moving) pa[i] = make_double4(g.path[0], g.path[1], g.path[2], g.moving ? 1.0 : 0.0);
			au[i] = make_int2(g.material, g.prim_id);
			if (any_rot) ro[i] = make_float2((float)g.uv_rotation[0], (float)g.uv_rotation[1]);
		});
		DQuad* qd = reinterpret_cast<DQuad*>(H + o_quad); int2* qa = reinterpret_cast<int2*>(H + o_quad_aux); int* qo = reinterpret_cast<int*>(H + o_quad_order);
		parallel_for(nQ, [&](size_t i) {
			const gs_quad& g = d->quads[quad_src[i]];
			qd[i] = make_dquad(g); qa[i] = make_int2(g.material, g.prim_id); qo[i] = g.order;
		});
		DTriangle* td = reinterpret_cast<DTriangle*>(H + o_tri); int2* ta = reinterpret_cast<int2*>(H + o_tri_aux);
		parallel_for(nT, [&](size_t i) {
			const gs_triangle& g = d->triangles[tri_src[i]];
			td[i] = make_dtriangle(g); ta[i] = make_int2(g.material, g.prim_id);
		});
	}
	{
		DSphere* bs = reinterpret_cast<DSphere*>(H + o_bsph); double4* bp = reinterpret_cast<double4*>(H + o_bpath);
		for (size_t i = 0; i < d->n_boundary_spheres; ++i) {
			const gs_sphere& g = d->boundary_spheres[i];
			bs[i] = make_dsphere(g);
			bp[i] = make_double4(g.path[0], g.path[1], g.path[2], g.moving ? 1.0 : 0.0);
		}
		DQuad* bq = reinterpret_cast<DQuad*>(H + o_bquad);
		for (size_t i = 0; i < d->n_boundary_quads; ++i) bq[i] = make_dquad(d->boundary_quads[i]);
		DMedium* md = reinterpret_cast<DMedium*>(H + o_media);
		for (size_t i = 0; i < d->n_media; ++i) {
			const gs_medium& m = d->media[i];
			DMedium dm;
			dm.sph_first = m.sphere_first; dm.sph_count = m.sphere_count; dm.quad_first = m.quad_first; dm.quad_count = m.quad_count;
			dm.neg_inv_density = -1.0 / m.density;  // volume.rs:18
			dm.phase_mat = m.phase_material;
			dm.flags = (m.sphere_count == 1 && m.quad_count == 0 && d->boundary_spheres[m.sphere_first].moving) ? 1 : 0;
			md[i] = dm;
		}
		DMaterial* mats = reinterpret_cast<DMaterial*>(H + o_mat);
		parallel_for(d->n_materials, [&](size_t i) {
			const gs_material& m = d->materials[i];
			DMaterial dm;
			dm.kind = m.kind; dm.tex = m.texture;
			dm.ar = (float)m.albedo[0]; dm.ag = (float)m.albedo[1]; dm.ab = (float)m.albedo[2];
			dm.param = (float)m.param; dm.pad0 = dm.pad1 = 0;
			mats[i] = dm;
		});
		DTexture* texs = reinterpret_cast<DTexture*>(H + o_tex);
		parallel_for(d->n_textures, [&](size_t i) {
			const gs_texture& t = d->textures[i];
			DTexture dt;
			dt.kind = t.kind; dt.a = t.a; dt.b = t.b; dt.scale = (float)t.scale;
			dt.cr = (float)t.color[0]; dt.cg = (float)t.color[1]; dt.cb = (float)t.color[2]; dt.pad = 0.f;
			dt.scale_inv_d = t.kind == GS_TEX_CHECKER ? 1.0 / t.scale : 0.0;  // texture.rs:42,50
			dt.scale_d = t.scale;
			texs[i] = dt;
		});
		DImage* imgs = reinterpret_cast<DImage*>(H + o_img);
		for (size_t i = 0; i < d->n_images; ++i) {
			const gs_image& im = d->images[i];
			const size_t n = (size_t)im.width * im.height;
			if (im.format == GS_IMG_U8_RGB) {
				uchar4* t = reinterpret_cast<uchar4*>(H + o_texels[i]);
				const uint8_t* p = static_cast<const uint8_t*>(im.pixels);
				parallel_for(n, [&](size_t k) { t[k] = make_uchar4(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], 255); });
			} else {
				float4* t = reinterpret_cast<float4*>(H + o_texels[i]);
				const float* p = static_cast<const float*>(im.pixels);
				parallel_for(n, [&](size_t k) { t[k] = make_float4(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], 0.f); });
			}
			DImage di; di.w = im.width; di.h = im.height; di.format = im.format; di.pad = 0; di.texels = D + o_texels[i];
			imgs[i] = di;
			gs_image keep = im; keep.pixels = nullptr;
			s->h_images.push_back(keep);
		}
		if (d->perlin_perm) std::memcpy(H + o_perm, d->perlin_perm, 256);
		std::memset(H + o_counters, 0, kCounterSlots * sizeof(unsigned long long));
		std::memset(H + o_stats, 0, kStatSlots * sizeof(unsigned long long));
	}
	lap("fill staging");
	GS_CUDA(cudaMemcpyAsync(D, H, total, cudaMemcpyHostToDevice, s->up));
	s->h2d_bytes += total;

	DeviceScene& ds = s->ds;
	ds = DeviceScene{};
	ds.refs = reinterpret_cast<const uint32_t*>(D + o_refs);
	ds.spheres = reinterpret_cast<const DSphere*>(D + o_sph);
	ds.sph_path = any_moving ? reinterpret_cast<const double4*>(D + o_sph_path) : nullptr;
	ds.sph_aux = reinterpret_cast<const int2*>(D + o_sph_aux);
	ds.sph_uvrot = any_rot ? reinterpret_cast<const float2*>(D + o_sph_rot) : nullptr;
	ds.quads = reinterpret_cast<const DQuad*>(D + o_quad);
	ds.quad_aux = reinterpret_cast<const int2*>(D + o_quad_aux);
	ds.quad_order = reinterpret_cast<const int*>(D + o_quad_order);
	ds.tris = reinterpret_cast<const DTriangle*>(D + o_tri);
	ds.tri_aux = reinterpret_cast<const int2*>(D + o_tri_aux);
	ds.bnd_spheres = reinterpret_cast<const DSphere*>(D + o_bsph);
	ds.bnd_sph_path = reinterpret_cast<const double4*>(D + o_bpath);
	ds.bnd_quads = reinterpret_cast<const DQuad*>(D + o_bquad);
	ds.media = reinterpret_cast<const DMedium*>(D + o_media);
	ds.materials = reinterpret_cast<const DMaterial*>(D + o_mat);
	ds.textures = reinterpret_cast<const DTexture*>(D + o_tex);
	ds.images = reinterpret_cast<const DImage*>(D + o_img);
	ds.perlin_perm = d->perlin_perm ? D + o_perm : nullptr;
	ds.n_media = (int)d->n_media;
	ds.root = root_leaf ? s->h_leaf_code[0] : 0;
	ds.prmt_hi = 0x47000000u;
	s->counters = reinterpret_cast<unsigned long long*>(D + o_counters);
	s->stats = reinterpret_cast<unsigned long long*>(D + o_stats);
	s->has_triangles = nT > 0;
	s->small_scene = nN * sizeof(DNode) / 2 + nS * sizeof(DSphere) + nQ * sizeof(DQuad) + nT * sizeof(DTriangle) <= kSmallSceneBytes;

	// which tree: by size, or as GS_BVH_WIDTH says; a single-leaf scene and a tree deeper than the group stack stay binary
	bool wide = nS + nQ + nT >= kWideMinPrims;
	if (const char* e = std::getenv("GS_BVH_WIDTH")) wide = std::atoi(e) == 8;
	if (root_leaf || s->wide_depth + 1 > kMaxStack8) wide = false;
	s->wide = wide;
	// the staging block of the arrays above is free again once the copy has run
	GS_CUDA(cudaEventRecord(s->ready, s->up));
	stage_guard.ev = s->ready;
	lap("upload issued");
	return GS_OK;
}

// ---- node arrays --------------------------------------------------------------------------------------------
// fp32 error budget of a slab distance fma(plane, 1/d, -o/d): origin and plane rounding (<= 1 ulp at origin_limit),
// subtraction, reciprocal and product (relative 2^-23 each over distances <= 2*origin_limit) -> 8 ulp(origin_limit) is
// safe.  (The render kernels take the reciprocal from rcp.approx.ftz.f32, whose error bound is that 2^-23; gs_device.cuh safe_rcp.)  The wide nodes add the rounding of the quantised decode (two more FMAs): 12 ulp, plus 1/64 of a grid step
// for the bias constant of Trav8::qf (its own error is 2^-9 of a step).
static double pad_for(double origin_limit, int ulps) {
	const float lim = (float)origin_limit;
	return (double)ulps * (double)(std::nextafterf(lim, INFINITY) - lim);
}

static int upload_nodes(gs_scene* s, void** dev, size_t bytes, const Staging& stg) {
	if (!*dev) GS_CUDA(dev_alloc(dev, bytes, s->up));
	GS_CUDA(cudaMemcpyAsync(*dev, stg.p, bytes, cudaMemcpyHostToDevice, s->up));
	s->h2d_bytes += bytes;
	return GS_OK;
}

int build_nodes2(gs_scene* s, double origin_limit) {
	Lap lap("build_nodes2");
	const gs_bvh_node* N = s->h_nodes.get();
	if (s->h_bfs.empty() && N[0].ref_count == 0) {  // breadth-first numbering of the internal nodes = device order (top levels contiguous)
		s->h_node_to_dnode.assign(s->n_nodes, -1);
		std::vector<int>& queue = s->h_bfs;
		queue.reserve(s->n_nodes / 2 + 1);
		queue.push_back(0);
		for (size_t h = 0; h < queue.size(); ++h) {
			const int n = queue[h];
			s->h_node_to_dnode[n] = (int)h;
			if (N[N[n].left].ref_count == 0) queue.push_back(N[n].left);
			if (N[N[n].right].ref_count == 0) queue.push_back(N[n].right);
		}
	}
	const size_t count = std::max(s->h_bfs.size(), (size_t)kNodePad);  // also covers the single-leaf tree: the pointer stays valid
	const double pad = pad_for(origin_limit, 8);
	Staging stg;
	GS_CUDA(staging_acquire(count * sizeof(DNode), stg));
	struct StageGuard { Staging& g; cudaEvent_t ev = nullptr; ~StageGuard() { staging_release(g, ev); } } guard{stg};
	DNode* out = reinterpret_cast<DNode*>(stg.p);
	auto child_box = [&](int abi, float* mnmx) {
		const gs_bvh_node& c = N[abi];
		for (int a = 0; a < 3; ++a) {
			const double lo = c.bounds_min[a] - pad, hi = c.bounds_max[a] + pad;
#if GS_NODE_CH
			// (centre, half extent): [centre - half, centre + half] covers the padded box
			const float centre = (float)(0.5 * (lo + hi));
			mnmx[a * 2] = centre; mnmx[a * 2 + 1] = float_up(std::fmax(hi - (double)centre, (double)centre - lo));
#else
			mnmx[a * 2] = float_down(lo); mnmx[a * 2 + 1] = float_up(hi);
#endif
		}
	};
	auto link = [&](int abi) { return s->h_node_to_dnode[abi] >= 0 ? s->h_node_to_dnode[abi] : s->h_leaf_code[abi]; };
	parallel_for(s->h_bfs.size(), [&](size_t k) {
		const gs_bvh_node& n = N[s->h_bfs[k]];
		float b0[6], b1[6];
		child_box(n.left, b0); child_box(n.right, b1);
		DNode d;
		d.xy0 = make_float4(b0[0], b0[1], b0[2], b0[3]);
		d.xy1 = make_float4(b1[0], b1[1], b1[2], b1[3]);
		d.z01 = make_float4(b0[4], b0[5], b1[4], b1[5]);
		d.link = make_int4(link(n.left), link(n.right), 0, 0);
		out[k] = d;
	});
	if (s->h_bfs.size() < count) std::memset(out + s->h_bfs.size(), 0, (count - s->h_bfs.size()) * sizeof(DNode));
	int rc = upload_nodes(s, reinterpret_cast<void**>(&s->nodes), count * sizeof(DNode), stg);
	if (rc) return rc;
	s->n_dnodes = count;
	s->ds.nodes = s->nodes;
	GS_CUDA(cudaEventRecord(s->ready, s->up));
	guard.ev = s->ready;
	lap("pad + upload issued");
	return GS_OK;
}

int build_nodes8(gs_scene* s, double origin_limit) {
	Lap lap("build_nodes8");
	const gs_bvh_node* N = s->h_nodes.get();
	const size_t nW = s->w_child_base.size();
	if (nW == 0) return fail(GS_ERR_UNSUPPORTED, "the eight-wide tree needs an internal root");
	if (s->wide_depth + 1 > kMaxStack8) return fail(GS_ERR_UNSUPPORTED, "eight-wide tree deeper than its traversal stack");
	const double pad = pad_for(origin_limit, 12);
	// does any leaf child need an explicit leaf code (several primitives, a moving sphere, a primitive placed elsewhere)?
	const bool any_link = s->w_any_link;
	const size_t node_bytes = (nW * sizeof(DNode8) + 255) & ~(size_t)255;
	const size_t link_bytes = any_link ? nW * 8 * sizeof(int) : 0;
	Staging stg;
	GS_CUDA(staging_acquire(node_bytes + link_bytes, stg));
	struct StageGuard { Staging& g; cudaEvent_t ev = nullptr; ~StageGuard() { staging_release(g, ev); } } guard{stg};
	DNode8* out = reinterpret_cast<DNode8*>(stg.p);
	int* links = any_link ? reinterpret_cast<int*>(stg.p + node_bytes) : nullptr;  // no side table without a leaf that needs one
	std::atomic<int> bad{0};
	parallel_for(nW, [&](size_t h) {
		const int* kids = s->w_kids.data() + 8 * h;
		if (h + 3 < nW)  // the kids' records are scattered over the node array: start the misses of a later node now
			for (int k = 0; k < 8; ++k) {
				const int c = kids[24 + k];
				if (c >= 0) { const char* p = reinterpret_cast<const char*>(N + c); __builtin_prefetch(p); __builtin_prefetch(p + sizeof(gs_bvh_node) - 1); }
			}
		float lo[8][3], hi[8][3];
		float corner[3] = {INFINITY, INFINITY, INFINITY};
		for (int k = 0; k < 8; ++k) {
			if (kids[k] < 0) continue;
			const gs_bvh_node& c = N[kids[k]];
			for (int a = 0; a < 3; ++a) {
				lo[k][a] = float_down(c.bounds_min[a] - pad); hi[k][a] = float_up(c.bounds_max[a] + pad);
				corner[a] = std::fmin(corner[a], lo[k][a]);
			}
		}
		DNode8 d;
		std::memset(&d, 0, sizeof d);
		uint32_t q[3][4];  // per axis: lo[0..3], lo[4..7], hi[0..3], hi[4..7]; empty slots lo = 255 > hi = 0
		for (int a = 0; a < 3; ++a) { q[a][0] = q[a][1] = 0xffffffffu; q[a][2] = q[a][3] = 0u; }
		uint32_t exps = 0;
		// Every stored plane lies at least `margin` of a grid step outside the box it bounds -- the lowest one too, which is
		// why the node's corner sits two margins below the lowest child plane: the decode's bias constant (Trav8::qf) is
		// off by up to 2^-9 of a step, and with an axis-parallel ray (1/d = 1e25) that error decides hit or miss outright.
		const double margin = 1.0 / 64.0;
		for (int a = 0; a < 3; ++a) {
			double top = -INFINITY;
			for (int k = 0; k < 8; ++k) if (kids[k] >= 0) top = std::fmax(top, (double)hi[k][a]);
			int e = -100;  // grid step 2^e: the smallest power of two whose 255 steps span the children (never below 2^-100)
			const double extent0 = top - (double)corner[a];
			if (extent0 > 0.0 && std::isfinite(extent0)) {  // ceil(log2(extent0 / 250)) from the exponent field; the loop below settles the rest
				const double x = extent0 / 250.0;
				int e0 = std::ilogb(x);
				if (std::ldexp(1.0, e0) < x) ++e0;
				e = std::max(e, e0);
			}
			float origin;
			for (;; ++e) {
				const double step = std::ldexp(1.0, e);
				origin = float_down((double)corner[a] - 2.0 * margin * step);
				if (std::ceil((top - (double)origin) / step + margin) <= 255.0) break;
			}
			if (e + 127 < 1 || e + 127 > 254) { bad = 1; return; }
			exps |= (uint32_t)(e + 127) << (8 * a);
			(a == 0 ? d.px : (a == 1 ? d.py : d.pz)) = origin;
			const double inv_step = std::ldexp(1.0, -e);
			for (int k = 0; k < 8; ++k) {
				if (kids[k] < 0) continue;
				const double ql = std::floor(((double)lo[k][a] - (double)origin) * inv_step - margin);
				const double qh = std::ceil(((double)hi[k][a] - (double)origin) * inv_step + margin);
				if (ql < 0.0 || qh > 255.0) { bad = 3; return; }
				const uint32_t bl = (uint32_t)ql, bh = (uint32_t)qh;
				const int w = k >> 2, sh = 8 * (k & 3);
				q[a][w] = (q[a][w] & ~(0xffu << sh)) | (bl << sh);
				q[a][2 + w] = (q[a][2 + w] & ~(0xffu << sh)) | (bh << sh);
			}
		}
		uint32_t imask = 0, meta[2] = {0, 0};
		for (int k = 0; k < 8; ++k) {
			const int c = kids[k];
			if (links) links[8 * h + k] = 0;
			if (c < 0) continue;
			if (N[c].ref_count == 0) { imask |= 1u << k; continue; }
			uint32_t m;
			if (s->h_leaf_direct[c] >= 0) {
				const uint32_t first = (uint32_t)(-1 - s->h_leaf_code[c]) >> 4, kind = first >> 25, idx = first & ((1u << 25) - 1u);
				const uint32_t off = idx - s->w_prim_base[3 * h + kind];
				if (off > 63u || kind > 2u) { bad = 2; return; }
				m = (kind << 6) | off;
			} else {
				m = 3u << 6;
				links[8 * h + k] = s->h_leaf_code[c];
			}
			meta[k >> 2] |= m << (8 * (k & 3));
		}
		d.exps_imask = exps | (imask << 24);
		for (int i = 0; i < 4; ++i) { d.qx[i] = q[0][i]; d.qy[i] = q[1][i]; d.qz[i] = q[2][i]; }
		d.meta[0] = meta[0]; d.meta[1] = meta[1];
		d.child_base = (uint32_t)s->w_child_base[h];
		d.sphere_base = s->w_prim_base[3 * h + 0];
		d.quad_base = s->w_prim_base[3 * h + 1];
		d.tri_base = s->w_prim_base[3 * h + 2];
		out[h] = d;
	});
	if (bad.load() == 1) return fail(GS_ERR_UNSUPPORTED, "eight-wide tree: box extent out of the fp32 exponent range");
	if (bad.load() == 2) return fail(GS_ERR_CUDA, "eight-wide tree: internal error (leaf offset)");
	if (bad.load() == 3) return fail(GS_ERR_CUDA, "eight-wide tree: internal error (plane outside the grid)");
	const size_t bytes = node_bytes + link_bytes;
	int rc = upload_nodes(s, reinterpret_cast<void**>(&s->nodes8), bytes, stg);
	if (rc) return rc;
	s->n_dnodes8 = nW;
	s->leaf_links8 = any_link ? reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(s->nodes8) + node_bytes) : nullptr;
	s->ds.nodes8 = s->nodes8;
	s->ds.leaf_links8 = s->leaf_links8;
	GS_CUDA(cudaEventRecord(s->ready, s->up));
	guard.ev = s->ready;
	lap("quantise + upload issued");
	return GS_OK;
}

int build_current_nodes(gs_scene* s, double origin_limit) {
	s->origin_limit = origin_limit;
	int rc = GS_OK;
	if (s->nodes || !s->wide) rc = build_nodes2(s, origin_limit);
	if (rc) return rc;
	if (s->nodes8 || s->wide) rc = build_nodes8(s, origin_limit);
	return rc;
}

int select_tree(gs_scene* s, int width) {
	if (width != 2 && width != 8) return fail(GS_ERR_INVALID, "bvh_width must be 2 or 8");
	if (width == 8 && !s->nodes8) {
		GS_CUDA(cudaDeviceSynchronize());
		int rc = build_nodes8(s, s->origin_limit);
		if (rc) return rc;
	}
	if (width == 2 && !s->nodes) {
		GS_CUDA(cudaDeviceSynchronize());
		int rc = build_nodes2(s, s->origin_limit);
		if (rc) return rc;
	}
	s->wide = width == 8;
	return GS_OK;
}

}  // namespace gs

gs_scene::~gs_scene() {
	cudaSetDevice(device);
	cudaDeviceSynchronize();  // nothing of this scene is in flight when its memory goes back to the pool
	cudaStream_t st = up;
	pool_org.release(st); pool_dir.release(st); pool_thr.release(st); pool_rad.release(st); pool_ids.release(st); pool_hit.release(st);
	free_list.release(st); wave_ctr.release(st); shade_queue.release(st); shade_count.release(st);
	if (h_wave_ctr) cudaFreeHost(h_wave_ctr);
	px_list[0].release(st); px_list[1].release(st); px_count.release(st); lum.release(st);
	accum.release(st); out_rgb.release(st); out_counts.release(st); multi_landing.release(st);
	if (multi_accum) cudaFree(multi_accum);
	gs::dev_free(blob, st); gs::dev_free(nodes, st); gs::dev_free(nodes8, st);
	if (ready) { gs::staging_forget_event(ready); cudaEventDestroy(ready); }
	if (last) cudaEventDestroy(last);
	if (ev0) cudaEventDestroy(ev0);
	if (ev1) cudaEventDestroy(ev1);
	for (int c = 0; c < gs::kMaxChains - 1; ++c) { if (aux_done[c]) cudaEventDestroy(aux_done[c]); if (aux[c]) cudaStreamDestroy(aux[c]); }
	if (aux_begin) cudaEventDestroy(aux_begin);
	if (up) { cudaStreamSynchronize(up); cudaStreamDestroy(up); }
}
