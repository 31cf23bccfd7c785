// grayshift_b200/csrc/abi.cu -- implementation of the C ABI in include/grayshift.h: validation, re-layout of
// the host's flat scene description for the GPU, upload, kernel launches, resolve.  Host code only;
// the kernels live in render_kernels.cu and trace_closest.cu.  There is no CPU compute path here:
// without a CUDA device every compute entry fails with GS_ERR_NO_DEVICE.
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/grayshift.h"
#include "gs_internal.h"

using namespace gs;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define GS_CUDA(call)                                                                                     \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) return fail(GS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
	} while (0)

// Independent iterations over the host threads (scene creation of a 10^6-primitive scene spends most of its time in
// such loops); small ranges stay on the calling thread.
template <class F> void parallel_for(size_t n, F&& body) {
	unsigned threads = std::thread::hardware_concurrency();
	threads = threads > 16u ? 16u : threads;
	if (n < 65536 || threads < 2) { for (size_t i = 0; i < n; ++i) body(i); return; }
	std::vector<std::thread> pool;
	const size_t chunk = (n + threads - 1) / threads;
	for (unsigned t = 0; t < threads; ++t) {
		const size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
		if (lo >= hi) break;
		pool.emplace_back([lo, hi, &body]() { for (size_t i = lo; i < hi; ++i) body(i); });
	}
	for (auto& th : pool) th.join();
}

float float_down(double x) { float f = (float)x; if ((double)f > x) f = std::nextafterf(f, -INFINITY); return f; }
float float_up(double x) { float f = (float)x; if ((double)f < x) f = std::nextafterf(f, INFINITY); return f; }

// Device memory comes from the device's stream-ordered pool with the release threshold lifted, so creating and
// destroying scenes reuses the same HBM instead of mapping / unmapping it each time (cudaMalloc + cudaFree of a
// 2.6 MB scene measured 25 + 55 ms per create/destroy pair next to a 226 ms render; from the pool it is < 1 ms).
// All pool traffic is ordered on the legacy default stream; gs_scene_create / gs_scene_destroy synchronize the device.
cudaError_t dev_alloc(void** p, size_t bytes) {
	static thread_local int tuned_device = -1;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (tuned_device != dev) {
		cudaMemPool_t pool;
		e = cudaDeviceGetDefaultMemPool(&pool, dev);
		if (e != cudaSuccess) return e;
		unsigned long long keep = ~0ull;
		e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
		if (e != cudaSuccess) return e;
		tuned_device = dev;
	}
	return cudaMallocAsync(p, bytes, 0);
}
void dev_free(void* p) { if (p) cudaFreeAsync(p, 0); }

template <class T> struct DevBuf {
	T* p = nullptr;
	size_t n = 0;
	cudaError_t upload(const std::vector<T>& h) {
		release();
		n = h.size();
		if (n == 0) return cudaSuccess;
		cudaError_t e = dev_alloc(reinterpret_cast<void**>(&p), n * sizeof(T));
		if (e != cudaSuccess) return e;
		return cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
	}
	cudaError_t alloc(size_t count) {
		if (count <= n && p) return cudaSuccess;
		// growing a buffer a previous (asynchronous) render may still be using on a non-blocking stream: the pool's free
		// is ordered on the legacy default stream only, so wait for the device first (rare path: the image got larger)
		if (p) cudaDeviceSynchronize();
		release();
		n = count;
		return dev_alloc(reinterpret_cast<void**>(&p), n * sizeof(T));
	}
	void release() { dev_free(p); p = nullptr; n = 0; }
};

constexpr int kCounterSlots = 64;
constexpr size_t kSmallSceneBytes = 128u << 10;     // C1 (486 spheres) is 47 KB, C3 2 KB, C4 (the one that wants the spill-free build) 490 KB
constexpr unsigned kAdaptiveTailPixels = 2048;  // adaptive renders: below this many unfinished pixels, switch to one warp per pixel
constexpr int kStatSlots = 16;  // [0..3] = gs_stats totals; [4..] only written by the GS_DIVERGENCE_STATS diagnostic build

}  // namespace

struct gs_scene {
	int device = 0;
	int sm_count = 0;
	// host copies kept for re-padding the node boxes when a camera sits outside the assumed origin range
	std::vector<gs_bvh_node> h_nodes;
	std::vector<int> h_node_to_dnode;   // ABI node -> DNode index (internal nodes), -1 for leaves
	std::vector<int> h_leaf_code;       // ABI node -> leaf code (leaves)
	std::vector<int> h_bfs;             // internal ABI nodes in BFS order
	double origin_limit = 0.0;
	double max_abs_coord = 0.0;

	DevBuf<DNode> nodes;
	DevBuf<DNode4> nodes4;  // GS_WIDE builds
	DevBuf<uint32_t> refs;
	DevBuf<DSphere> spheres;
	DevBuf<double4> sph_path;
	DevBuf<int2> sph_aux;
	DevBuf<DQuad> quads;
	DevBuf<int2> quad_aux;
	DevBuf<int> quad_order;
	DevBuf<DTriangle> tris;
	DevBuf<int2> tri_aux;
	DevBuf<DSphere> bnd_spheres;
	DevBuf<double4> bnd_sph_path;
	DevBuf<DQuad> bnd_quads;
	DevBuf<DMedium> media;
	DevBuf<DMaterial> materials;
	DevBuf<DTexture> textures;
	DevBuf<DImage> images;
	std::vector<void*> image_texels;
	std::vector<gs_image> h_images;  // sizes/formats for validation of camera.hdri_image
	DevBuf<uint8_t> perm;
	DeviceScene ds{};

	// adaptive sampler: the two lists of pixels still sampling, their lengths, per-pixel (sum lum, sum lum^2)
	DevBuf<uint32_t> px_list[2];
	DevBuf<unsigned> px_count;
	DevBuf<double2> lum;

	// render scratch
	DevBuf<float4> accum;
	DevBuf<float> out_rgb;
	DevBuf<unsigned> out_counts;
	DevBuf<unsigned long long> counters;  // kCounterSlots work counters
	DevBuf<unsigned long long> stats;     // 4 device totals
	int next_counter = 0;
	uint64_t launches = 0;
	bool count_nodes = false;
	bool has_triangles = false;  // selects the kernel variants that carry the triangle branch
	bool small_scene = false;    // nodes + primitives fit the L1 with room to spare: the kernels built for one more CTA per SM
	uint64_t h2d_bytes = 0;

	// wavefront pipeline: path pool (structure of arrays) + counters.  Off by default: measured against the
	// megakernel on B200 it wins 6 % on C4 and loses ~2x on C1/C3; option "wavefront" = 1 enables it.
	bool use_wavefront = false;
	int pool_capacity = 1 << 20;  // in-flight paths (112 B each)
	DevBuf<double4> pool_org, pool_dir;
	DevBuf<float4> pool_thr, pool_rad;
	DevBuf<uint4> pool_ids, pool_hit;
	DevBuf<int> free_list;
	DevBuf<WaveCounters> wave_ctr;
	WaveCounters* h_wave_ctr = nullptr;  // pinned

	~gs_scene() {
		cudaSetDevice(device);
		cudaDeviceSynchronize();  // nothing of this scene is in flight when its memory goes back to the pool
		for (void* t : image_texels) dev_free(t);
		pool_org.release(); pool_dir.release(); pool_thr.release(); pool_rad.release(); pool_ids.release(); pool_hit.release();
		free_list.release(); wave_ctr.release();
		if (h_wave_ctr) cudaFreeHost(h_wave_ctr);
		px_list[0].release(); px_list[1].release(); px_count.release(); lum.release();
		nodes.release(); nodes4.release(); refs.release(); spheres.release(); sph_path.release(); sph_aux.release(); quads.release(); quad_aux.release();
		quad_order.release(); tris.release(); tri_aux.release(); bnd_spheres.release(); bnd_sph_path.release(); bnd_quads.release(); media.release(); materials.release();
		textures.release(); images.release(); perm.release(); accum.release(); out_rgb.release(); out_counts.release(); counters.release();
		stats.release();
	}
};

namespace {

bool finite3(const double* v) { return std::isfinite(v[0]) && std::isfinite(v[1]) && std::isfinite(v[2]); }

DSphere make_dsphere(const gs_sphere& s) {
	DSphere d;
	d.cx = s.center[0]; d.cy = s.center[1]; d.cz = s.center[2];
	d.r = s.moving ? -std::fabs(s.radius) : s.radius;  // sign flags a moving sphere; radius enters only as r*r and 1/r
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

int validate(const gs_scene_desc* d) {
	if (!d) return fail(GS_ERR_INVALID, "null scene description");
	if (d->abi_version != GS_ABI_VERSION) return fail(GS_ERR_INVALID, "abi_version mismatch");
	if (d->n_spheres + d->n_quads + d->n_triangles == 0) return fail(GS_ERR_INVALID, "empty scene (the reference recurses forever on an empty list, BVH.rs:18-65)");
	if (d->n_spheres >= (1u << 30) || d->n_quads >= (1u << 30) || d->n_triangles >= (1u << 30)) return fail(GS_ERR_UNSUPPORTED, "too many primitives");
	if (d->n_triangles && !d->triangles) return fail(GS_ERR_INVALID, "null triangle array");
	if (d->n_nodes == 0 || !d->nodes) return fail(GS_ERR_INVALID, "no BVH nodes");
	if (d->n_materials == 0 || !d->materials) return fail(GS_ERR_INVALID, "no materials");
	auto bad_mat = [&](int m) { return m < 0 || (size_t)m >= d->n_materials; };
	for (size_t i = 0; i < d->n_spheres; ++i) {
		const gs_sphere& s = d->spheres[i];
		if (!finite3(s.center) || !finite3(s.path) || !std::isfinite(s.radius) || s.radius == 0.0) return fail(GS_ERR_INVALID, "sphere " + std::to_string(i) + ": non-finite or zero radius");
		if (bad_mat(s.material)) return fail(GS_ERR_INVALID, "sphere " + std::to_string(i) + ": bad material index");
	}
	for (size_t i = 0; i < d->n_quads; ++i) {
		const gs_quad& q = d->quads[i];
		if (!finite3(q.q) || !finite3(q.u) || !finite3(q.v)) return fail(GS_ERR_INVALID, "quad " + std::to_string(i) + ": non-finite");
		if (bad_mat(q.material)) return fail(GS_ERR_INVALID, "quad " + std::to_string(i) + ": bad material index");
	}
	for (size_t i = 0; i < d->n_triangles; ++i) {
		const gs_triangle& t = d->triangles[i];
		if (!finite3(t.a) || !finite3(t.b) || !finite3(t.c)) return fail(GS_ERR_INVALID, "triangle " + std::to_string(i) + ": non-finite");
		if (bad_mat(t.material)) return fail(GS_ERR_INVALID, "triangle " + std::to_string(i) + ": bad material index");
	}
	for (size_t i = 0; i < d->n_boundary_spheres; ++i)
		if (!finite3(d->boundary_spheres[i].center) || !std::isfinite(d->boundary_spheres[i].radius)) return fail(GS_ERR_INVALID, "boundary sphere: non-finite");
	for (size_t i = 0; i < d->n_media; ++i) {
		const gs_medium& m = d->media[i];
		if (m.sphere_first < 0 || m.sphere_count < 0 || (size_t)(m.sphere_first + m.sphere_count) > d->n_boundary_spheres ||
		    m.quad_first < 0 || m.quad_count < 0 || (size_t)(m.quad_first + m.quad_count) > d->n_boundary_quads)
			return fail(GS_ERR_INVALID, "medium " + std::to_string(i) + ": boundary range out of bounds");
		if (!(m.density != 0.0) || !std::isfinite(m.density)) return fail(GS_ERR_INVALID, "medium: bad density");
		if (bad_mat(m.phase_material)) return fail(GS_ERR_INVALID, "medium: bad phase material");
	}
	for (size_t i = 0; i < d->n_materials; ++i) {
		const gs_material& m = d->materials[i];
		if (m.kind < 0 || m.kind > GS_MAT_ISOTROPIC) return fail(GS_ERR_INVALID, "material " + std::to_string(i) + ": bad kind");
		bool needs_tex = m.kind == GS_MAT_LAMBERTIAN || m.kind == GS_MAT_DIFFUSE_LIGHT || m.kind == GS_MAT_ISOTROPIC;
		if (needs_tex && (m.texture < 0 || (size_t)m.texture >= d->n_textures)) return fail(GS_ERR_INVALID, "material " + std::to_string(i) + ": bad texture index");
	}
	for (size_t i = 0; i < d->n_textures; ++i) {
		const gs_texture& t = d->textures[i];
		if (t.kind < 0 || t.kind > GS_TEX_NOISE) return fail(GS_ERR_INVALID, "texture: bad kind");
		if (t.kind == GS_TEX_CHECKER) {
			// children must come earlier in the table: rules out cycles (Arc graphs are acyclic)
			if (t.a < 0 || t.b < 0 || (size_t)t.a >= i || (size_t)t.b >= i) return fail(GS_ERR_INVALID, "checker texture " + std::to_string(i) + ": children must precede it");
		}
		if (t.kind == GS_TEX_IMAGE) {
			if (t.a < 0 || (size_t)t.a >= d->n_images || d->images[t.a].format != GS_IMG_U8_RGB) return fail(GS_ERR_INVALID, "image texture: bad image index/format");
		}
		if (t.kind == GS_TEX_NOISE && !d->perlin_perm) return fail(GS_ERR_INVALID, "noise texture without perlin_perm");
	}
	for (size_t i = 0; i < d->n_images; ++i) {
		const gs_image& im = d->images[i];
		if (im.width <= 0 || im.height <= 0 || !im.pixels) return fail(GS_ERR_INVALID, "image: bad size or null pixels");
	}
	for (size_t i = 0; i < d->n_nodes; ++i) {
		const gs_bvh_node& n = d->nodes[i];
		if (!finite3(n.bounds_min) || !finite3(n.bounds_max)) return fail(GS_ERR_INVALID, "BVH node " + std::to_string(i) + ": non-finite bounds (the reference panics, BVH.rs:54)");
		if (n.ref_count > 0) {
			if (n.ref_count > 15) return fail(GS_ERR_UNSUPPORTED, "BVH leaf with more than 15 primitives");
			if (n.first_ref < 0 || (size_t)(n.first_ref + n.ref_count) > d->n_prim_refs) return fail(GS_ERR_INVALID, "BVH leaf: reference range out of bounds");
		} else {
			if (n.left <= (int)i || n.right <= (int)i || (size_t)n.left >= d->n_nodes || (size_t)n.right >= d->n_nodes)
				return fail(GS_ERR_INVALID, "BVH node " + std::to_string(i) + ": children must have larger indices");
		}
	}
	for (size_t i = 0; i < d->n_prim_refs; ++i) {
		uint32_t r = d->prim_refs[i], kind = r >> 30, idx = r & kRefIndexMask;
		if (kind > 2 || (kind == 0 && idx >= d->n_spheres) || (kind == 1 && idx >= d->n_quads) || (kind == 2 && idx >= d->n_triangles))
			return fail(GS_ERR_INVALID, "bad primitive reference");
	}
	return GS_OK;
}

// (Re)build the device node array from the ABI nodes with boxes padded for fp32 traversal of rays
// whose origins lie within `origin_limit` of the world origin.
#if GS_WIDE
// The wide tree: every binary node that stays internal takes its two children and, twice, replaces the internal child
// with the largest surface area by that child's own two children (up to four).  Child boxes are padded for the fp32
// slab test exactly like the binary nodes (plus the rounding of the quantised decode), then stored as 8-bit offsets
// on a power-of-two grid from the node's lower corner: lower planes rounded down, upper planes up -- conservative.
int build_nodes_wide(gs_scene* s, double origin_limit) {
	const float lim = (float)origin_limit;
	// 8 ulp as for the binary nodes + 4 for the two extra roundings of t = q * (step / d) + (corner - o) / d
	const double pad = 12.0 * (double)(std::nextafterf(lim, INFINITY) - lim);
	const std::vector<gs_bvh_node>& N = s->h_nodes;
	auto internal = [&](int a) { return N[a].ref_count == 0; };
	auto area = [&](int a) {
		const double dx = N[a].bounds_max[0] - N[a].bounds_min[0], dy = N[a].bounds_max[1] - N[a].bounds_min[1], dz = N[a].bounds_max[2] - N[a].bounds_min[2];
		return dx * dy + dy * dz + dz * dx;
	};
	std::vector<DNode4> out;
	if (!N.empty() && internal(0)) {
		std::vector<int> queue{0};          // binary nodes that become wide nodes, breadth-first = device order
		std::vector<std::array<int, 4>> kids_of;
		std::vector<int> n_kids_of;
		std::vector<int> wide_index(N.size(), -1);
		wide_index[0] = 0;
		for (size_t h = 0; h < queue.size(); ++h) {
			const int a = queue[h];
			std::array<int, 4> kids{N[a].left, N[a].right, -1, -1};
			int n = 2;
			while (n < 4) {
				int best = -1;
				for (int k = 0; k < n; ++k)
					if (internal(kids[k]) && (best < 0 || area(kids[k]) > area(kids[best]))) best = k;
				if (best < 0) break;
				const int c = kids[best];
				kids[best] = N[c].left;
				kids[n++] = N[c].right;
			}
			for (int k = 0; k < n; ++k)
				if (internal(kids[k])) { wide_index[kids[k]] = (int)queue.size(); queue.push_back(kids[k]); }
			kids_of.push_back(kids);
			n_kids_of.push_back(n);
		}
		// worst-case stack use: (children - 1) siblings pushed at every level of the deepest chain
		std::vector<int> need(queue.size(), 0);
		for (size_t h = queue.size(); h-- > 0;) {
			int deepest = 0;
			for (int k = 0; k < n_kids_of[h]; ++k)
				if (internal(kids_of[h][k])) deepest = std::max(deepest, need[wide_index[kids_of[h][k]]]);
			need[h] = n_kids_of[h] - 1 + deepest;
		}
		if (need[0] + 2 > kMaxStack) return fail(GS_ERR_UNSUPPORTED, "wide BVH needs a deeper traversal stack than the kernels have");
		out.resize(queue.size());
		for (size_t h = 0; h < queue.size(); ++h) {
			float lo[4][3], hi[4][3];
			float corner[3] = {INFINITY, INFINITY, INFINITY}, top[3] = {-INFINITY, -INFINITY, -INFINITY};
			const int n = n_kids_of[h];
			for (int k = 0; k < n; ++k) {
				const gs_bvh_node& c = N[kids_of[h][k]];
				for (int a = 0; a < 3; ++a) {
					lo[k][a] = float_down(c.bounds_min[a] - pad); hi[k][a] = float_up(c.bounds_max[a] + pad);
					corner[a] = std::fmin(corner[a], lo[k][a]); top[a] = std::fmax(top[a], hi[k][a]);
				}
			}
			DNode4 d;
			std::memset(&d, 0, sizeof d);
			d.ox = corner[0]; d.oy = corner[1]; d.oz = corner[2];
			uint32_t q[6] = {0, 0, 0, 0, 0, 0};  // qlox, qhix, qloy, qhiy, qloz, qhiz
			for (int a = 0; a < 3; ++a) {
				const double extent = (double)top[a] - (double)corner[a];
				int e = -100;  // grid step 2^e: the smallest power of two with 255 steps >= extent (never below 2^-100)
				if (extent > 0.0) e = std::max(e, (int)std::ceil(std::log2(extent / 255.0)));
				for (;;) {  // make sure every upper plane fits into 8 bits after rounding up
					const double step = std::ldexp(1.0, e);
					bool ok = true;
					for (int k = 0; k < n && ok; ++k) ok = std::ceil(((double)hi[k][a] - (double)corner[a]) / step) <= 255.0;
					if (ok) break;
					++e;
				}
				if (e + 127 < 1 || e + 127 > 254) return fail(GS_ERR_UNSUPPORTED, "wide BVH: box extent out of the fp32 exponent range");
				d.exps |= (uint32_t)(e + 127) << (8 * a);
				const double step = std::ldexp(1.0, e);
				for (int k = 0; k < n; ++k) {
					const double ql = std::floor(((double)lo[k][a] - (double)corner[a]) / step), qh = std::ceil(((double)hi[k][a] - (double)corner[a]) / step);
					q[2 * a] |= (uint32_t)std::min(255.0, std::max(0.0, ql)) << (8 * k);
					q[2 * a + 1] |= (uint32_t)std::min(255.0, std::max(0.0, qh)) << (8 * k);
				}
			}
			d.qlox = q[0]; d.qhix = q[1]; d.qloy = q[2]; d.qhiy = q[3]; d.qloz = q[4]; d.qhiz = q[5];
			int lk[4] = {0, 0, 0, 0};
			for (int k = 0; k < n; ++k) lk[k] = internal(kids_of[h][k]) ? wide_index[kids_of[h][k]] : s->h_leaf_code[kids_of[h][k]];
			d.link = make_int4(lk[0], lk[1], lk[2], lk[3]);
			out[h] = d;
		}
	}
	if (out.empty()) { DNode4 d; std::memset(&d, 0, sizeof d); out.push_back(d); }  // single-leaf tree: keep the pointer valid
	GS_CUDA(s->nodes4.upload(out));
	s->h2d_bytes += out.size() * sizeof(DNode4);
	s->ds.nodes4 = s->nodes4.p;
	s->ds.nodes = nullptr;
	return GS_OK;
}
#endif

int build_nodes(gs_scene* s, double origin_limit) {
	s->origin_limit = origin_limit;
#if GS_WIDE
	return build_nodes_wide(s, origin_limit);
#endif
	// fp32 error budget: origin and plane rounding (<= 1 ulp at origin_limit), subtraction, reciprocal
	// and product (relative 2^-23 each over distances <= 2*origin_limit) -> 8 ulp(origin_limit) is safe.
	const float lim = (float)origin_limit;
	const double pad = 8.0 * (double)(std::nextafterf(lim, INFINITY) - lim);
	std::vector<DNode> out(s->h_bfs.size());
	auto child_box = [&](int abi, float* mnmx) {
		const gs_bvh_node& c = s->h_nodes[abi];
		for (int a = 0; a < 3; ++a) { mnmx[a * 2] = float_down(c.bounds_min[a] - pad); mnmx[a * 2 + 1] = float_up(c.bounds_max[a] + pad); }
	};
	auto link = [&](int abi) { return s->h_node_to_dnode[abi] >= 0 ? s->h_node_to_dnode[abi] : s->h_leaf_code[abi]; };
	parallel_for(s->h_bfs.size(), [&](size_t k) {
		const gs_bvh_node& n = s->h_nodes[s->h_bfs[k]];
		float b0[6], b1[6];
		child_box(n.left, b0); child_box(n.right, b1);
		DNode d;
		d.xy0 = make_float4(b0[0], b0[1], b0[2], b0[3]);
		d.xy1 = make_float4(b1[0], b1[1], b1[2], b1[3]);
		d.z01 = make_float4(b0[4], b0[5], b1[4], b1[5]);
		d.link = make_int4(link(n.left), link(n.right), 0, 0);
		out[k] = d;
	});
	if (out.size() < (size_t)kNodePad) {  // also covers the single-leaf tree: the pointer stays valid
		DNode d; std::memset(&d, 0, sizeof d);
		out.resize(kNodePad, d);
	}
	GS_CUDA(s->nodes.upload(out));
	s->h2d_bytes += out.size() * sizeof(DNode);
	s->ds.nodes = s->nodes.p;
	return GS_OK;
}

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
	GS_CUDA(cudaDeviceSynchronize());
	return build_nodes(s, 2.0 * m);
}

int check_device(gs_scene* s) {
	if (!s) return fail(GS_ERR_INVALID, "null scene");
	GS_CUDA(cudaSetDevice(s->device));
	return GS_OK;
}

unsigned long long* next_counter(gs_scene* s) {
	unsigned long long* p = s->counters.p + s->next_counter;
	s->next_counter = (s->next_counter + 1) % kCounterSlots;
	return p;
}

}  // namespace

extern "C" {

const char* gs_last_error(void) { return g_err.c_str(); }
int gs_abi_version(void) { return GS_ABI_VERSION; }

int gs_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
	return n;
}

int gs_scene_create(const gs_scene_desc* d, int device, gs_scene** out) {
	if (!out) return fail(GS_ERR_INVALID, "null out pointer");
	*out = nullptr;
	// GS_TRACE=1: stage times of this call on stderr (host clock)
	static const bool trace = std::getenv("GS_TRACE") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!trace) return;
		auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[gs_scene_create] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	int rc = validate(d);
	lap("validate");
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

	lap("device attributes");
	// ---- tree walk: BFS numbering of internal nodes, DFS leaf order for the references ----
	s->h_nodes.assign(d->nodes, d->nodes + d->n_nodes);
	s->h_node_to_dnode.assign(d->n_nodes, -1);
	s->h_leaf_code.assign(d->n_nodes, 0);
	std::vector<int> depth(d->n_nodes, -1);
	{
		std::vector<int> queue;
		if (d->nodes[0].ref_count == 0) queue.push_back(0);
		depth[0] = 0;
		for (size_t h = 0; h < queue.size(); ++h) {
			int n = queue[h];
			s->h_node_to_dnode[n] = (int)h;
			for (int c : {d->nodes[n].left, d->nodes[n].right}) {
				if (depth[c] >= 0) return fail(GS_ERR_INVALID, "BVH is not a tree (node reached twice)");
				depth[c] = depth[n] + 1;
				if (depth[c] > kMaxStack - 4) return fail(GS_ERR_UNSUPPORTED, "BVH deeper than the traversal stack");
				if (d->nodes[c].ref_count == 0) queue.push_back(c);
			}
		}
		s->h_bfs = queue;
	}
	std::vector<uint32_t> refs;
	std::vector<int> sph_new(d->n_spheres, -1), quad_new(d->n_quads, -1), tri_new(d->n_triangles, -1);
	std::vector<DTriangle> tris; std::vector<int2> tri_aux;
	std::vector<DSphere> spheres; std::vector<double4> sph_path; std::vector<int2> sph_aux;
	std::vector<DQuad> quads; std::vector<int2> quad_aux; std::vector<int> quad_order;
	refs.reserve(d->n_prim_refs);
	spheres.reserve(d->n_spheres); sph_path.reserve(d->n_spheres); sph_aux.reserve(d->n_spheres);
	quads.reserve(d->n_quads); quad_aux.reserve(d->n_quads); quad_order.reserve(d->n_quads);
	tris.reserve(d->n_triangles); tri_aux.reserve(d->n_triangles);
	{
		std::vector<int> stack{0};
		while (!stack.empty()) {
			int n = stack.back(); stack.pop_back();
			const gs_bvh_node& nd = d->nodes[n];
			if (depth[n] < 0) continue;
			if (nd.ref_count > 0) {
				uint32_t first = (uint32_t)refs.size();
				for (int k = 0; k < nd.ref_count; ++k) {
					uint32_t r = d->prim_refs[nd.first_ref + k], kind = r >> 30, idx = r & kRefIndexMask;
					if (kind == 0) {
						if (sph_new[idx] < 0) {
							sph_new[idx] = (int)spheres.size();
							const gs_sphere& sp = d->spheres[idx];
							spheres.push_back(make_dsphere(sp));
							sph_path.push_back(make_double4(sp.path[0], sp.path[1], sp.path[2], 0.0));
							sph_aux.push_back(make_int2(sp.material, sp.prim_id));
						}
						refs.push_back(GS_PRIM_REF(0u, sph_new[idx]));
					} else if (kind == 2) {
						if (tri_new[idx] < 0) {
							tri_new[idx] = (int)tris.size();
							tris.push_back(make_dtriangle(d->triangles[idx]));
							tri_aux.push_back(make_int2(d->triangles[idx].material, d->triangles[idx].prim_id));
						}
						refs.push_back(GS_PRIM_REF(2u, tri_new[idx]));
					} else {
						if (quad_new[idx] < 0) {
							quad_new[idx] = (int)quads.size();
							const gs_quad& q = d->quads[idx];
							quads.push_back(make_dquad(q));
							quad_aux.push_back(make_int2(q.material, q.prim_id));
							quad_order.push_back(q.order);
						}
						refs.push_back(GS_PRIM_REF(1u, quad_new[idx]));
					}
				}
				if (first >= (1u << 27) - 1u) return fail(GS_ERR_UNSUPPORTED, "too many primitive references");
				s->h_leaf_code[n] = leaf_code(first, (uint32_t)nd.ref_count);
				if (nd.ref_count == 1) {
					// single-primitive leaf: put (kind << 25 | index) into the link itself, count field 0
					const uint32_t r = refs.back(), kind = r >> 30, idx = r & kRefIndexMask;
					if (idx < (1u << 25)) s->h_leaf_code[n] = leaf_code((kind << 25) | idx, 0u);
				}
			} else { stack.push_back(nd.right); stack.push_back(nd.left); }
		}
	}
	// ---- coordinate range for the fp32 box padding ----
	double mx = 0.0;
	for (size_t i = 0; i < d->n_nodes; ++i)
		for (int a = 0; a < 3; ++a) mx = std::fmax(mx, std::fmax(std::fabs(d->nodes[i].bounds_min[a]), std::fabs(d->nodes[i].bounds_max[a])));
	for (size_t i = 0; i < d->n_boundary_spheres; ++i) {
		const gs_sphere& b = d->boundary_spheres[i];
		for (int a = 0; a < 3; ++a) mx = std::fmax(mx, std::fabs(b.center[a]) + std::fabs(b.path[a]) + std::fabs(b.radius));
	}
	for (size_t i = 0; i < d->n_boundary_quads; ++i) {
		const gs_quad& b = d->boundary_quads[i];
		for (int a = 0; a < 3; ++a) mx = std::fmax(mx, std::fabs(b.q[a]) + std::fabs(b.u[a]) + std::fabs(b.v[a]));
	}
	s->max_abs_coord = mx;

	lap("re-layout (host)");
	// ---- upload ----
	GS_CUDA(s->refs.upload(refs));
	GS_CUDA(s->spheres.upload(spheres)); GS_CUDA(s->sph_path.upload(sph_path)); GS_CUDA(s->sph_aux.upload(sph_aux));
	GS_CUDA(s->quads.upload(quads)); GS_CUDA(s->quad_aux.upload(quad_aux)); GS_CUDA(s->quad_order.upload(quad_order));
	GS_CUDA(s->tris.upload(tris)); GS_CUDA(s->tri_aux.upload(tri_aux));
	s->h2d_bytes += refs.size() * 4 + spheres.size() * (sizeof(DSphere) + sizeof(double4) + sizeof(int2)) + quads.size() * (sizeof(DQuad) + sizeof(int2) + 4);
	{
		std::vector<DSphere> bs; std::vector<double4> bp; std::vector<DQuad> bq;
		for (size_t i = 0; i < d->n_boundary_spheres; ++i) {
			bs.push_back(make_dsphere(d->boundary_spheres[i]));
			bp.push_back(make_double4(d->boundary_spheres[i].path[0], d->boundary_spheres[i].path[1], d->boundary_spheres[i].path[2], 0.0));
		}
		for (size_t i = 0; i < d->n_boundary_quads; ++i) bq.push_back(make_dquad(d->boundary_quads[i]));
		GS_CUDA(s->bnd_spheres.upload(bs)); GS_CUDA(s->bnd_sph_path.upload(bp)); GS_CUDA(s->bnd_quads.upload(bq));
		std::vector<DMedium> md;
		for (size_t i = 0; i < d->n_media; ++i) {
			const gs_medium& m = d->media[i];
			DMedium dm;
			dm.sph_first = m.sphere_first; dm.sph_count = m.sphere_count; dm.quad_first = m.quad_first; dm.quad_count = m.quad_count;
			dm.neg_inv_density = -1.0 / m.density;  // volume.rs:18
			dm.phase_mat = m.phase_material; dm.index = m.medium_index;
			md.push_back(dm);
		}
		GS_CUDA(s->media.upload(md));
	}
	{
		std::vector<DMaterial> mats;
		for (size_t i = 0; i < d->n_materials; ++i) {
			const gs_material& m = d->materials[i];
			DMaterial dm;
			dm.kind = m.kind; dm.tex = m.texture;
			dm.ar = (float)m.albedo[0]; dm.ag = (float)m.albedo[1]; dm.ab = (float)m.albedo[2];
			dm.param = (float)m.param; dm.pad0 = dm.pad1 = 0;
			mats.push_back(dm);
		}
		GS_CUDA(s->materials.upload(mats));
		std::vector<DTexture> texs;
		for (size_t i = 0; i < d->n_textures; ++i) {
			const gs_texture& t = d->textures[i];
			DTexture dt;
			dt.kind = t.kind; dt.a = t.a; dt.b = t.b; dt.scale = (float)t.scale;
			dt.cr = (float)t.color[0]; dt.cg = (float)t.color[1]; dt.cb = (float)t.color[2]; dt.pad = 0.f;
			dt.scale_inv_d = t.kind == GS_TEX_CHECKER ? 1.0 / t.scale : 0.0;  // texture.rs:42,50
			dt.scale_d = t.scale;
			texs.push_back(dt);
		}
		GS_CUDA(s->textures.upload(texs));
		std::vector<DImage> imgs;
		for (size_t i = 0; i < d->n_images; ++i) {
			const gs_image& im = d->images[i];
			size_t n = (size_t)im.width * im.height;
			void* dev = nullptr;
			if (im.format == GS_IMG_U8_RGB) {
				std::vector<uchar4> t(n);
				const uint8_t* p = static_cast<const uint8_t*>(im.pixels);
				for (size_t k = 0; k < n; ++k) t[k] = make_uchar4(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], 255);
				GS_CUDA(dev_alloc(&dev, n * sizeof(uchar4)));
				s->image_texels.push_back(dev);
				GS_CUDA(cudaMemcpy(dev, t.data(), n * sizeof(uchar4), cudaMemcpyHostToDevice));
				s->h2d_bytes += n * sizeof(uchar4);
			} else if (im.format == GS_IMG_F32_RGB) {
				std::vector<float4> t(n);
				const float* p = static_cast<const float*>(im.pixels);
				for (size_t k = 0; k < n; ++k) t[k] = make_float4(p[k * 3], p[k * 3 + 1], p[k * 3 + 2], 0.f);
				GS_CUDA(dev_alloc(&dev, n * sizeof(float4)));
				s->image_texels.push_back(dev);
				GS_CUDA(cudaMemcpy(dev, t.data(), n * sizeof(float4), cudaMemcpyHostToDevice));
				s->h2d_bytes += n * sizeof(float4);
			} else return fail(GS_ERR_INVALID, "image: bad format");
			DImage di; di.w = im.width; di.h = im.height; di.format = im.format; di.pad = 0; di.texels = dev;
			imgs.push_back(di);
			gs_image keep = im; keep.pixels = nullptr;
			s->h_images.push_back(keep);
		}
		GS_CUDA(s->images.upload(imgs));
		if (d->perlin_perm) {
			std::vector<uint8_t> pm(d->perlin_perm, d->perlin_perm + 256);
			GS_CUDA(s->perm.upload(pm));
		}
	}
	lap("upload primitives + tables");
	GS_CUDA(s->counters.alloc(kCounterSlots));
	GS_CUDA(s->stats.alloc(kStatSlots));
	GS_CUDA(cudaMemset(s->stats.p, 0, kStatSlots * sizeof(unsigned long long)));

	DeviceScene& ds = s->ds;
	ds.refs = s->refs.p; ds.spheres = s->spheres.p; ds.sph_path = s->sph_path.p; ds.sph_aux = s->sph_aux.p;
	ds.quads = s->quads.p; ds.quad_aux = s->quad_aux.p; ds.quad_order = s->quad_order.p;
	ds.tris = s->tris.p; ds.tri_aux = s->tri_aux.p;
	s->has_triangles = !tris.empty();
	s->small_scene = d->n_nodes * sizeof(DNode) / 2 + spheres.size() * sizeof(DSphere) + quads.size() * sizeof(DQuad) + tris.size() * sizeof(DTriangle) <= kSmallSceneBytes;
	ds.bnd_spheres = s->bnd_spheres.p; ds.bnd_sph_path = s->bnd_sph_path.p; ds.bnd_quads = s->bnd_quads.p;
	ds.media = s->media.p; ds.materials = s->materials.p; ds.textures = s->textures.p; ds.images = s->images.p;
	ds.perlin_perm = s->perm.p; ds.n_media = (int)d->n_media;
	ds.root = d->nodes[0].ref_count > 0 ? s->h_leaf_code[0] : 0;
	rc = build_nodes(s, 4.0 * std::fmax(mx, 1.0));
	if (rc) return rc;
	lap("nodes: pad + upload");
	guard.ok = true;
	*out = s;
	return GS_OK;
}

void gs_scene_destroy(gs_scene* scene) { delete scene; }

int gs_scene_set_option(gs_scene* scene, const char* key, int64_t value) {
	if (!scene || !key) return fail(GS_ERR_INVALID, "null argument");
	if (std::strcmp(key, "count_nodes") == 0) { scene->count_nodes = value != 0; return GS_OK; }
	if (std::strcmp(key, "wavefront") == 0) { scene->use_wavefront = value != 0; return GS_OK; }
	if (std::strcmp(key, "pool_capacity") == 0) {
		if (value < 128 || value > (1ll << 26)) return fail(GS_ERR_INVALID, "pool_capacity out of range");
		scene->pool_capacity = (int)value;
		return GS_OK;
	}
	return fail(GS_ERR_INVALID, std::string("unknown option '") + key + "'");
}

// The wavefront pipeline: waves of generate -> traverse -> shade over the path pool until every item has been
// started and the pool is empty.  The host looks at the counters every few waves (one small pinned read).
static int render_wavefront(gs_scene* s, WavefrontParams& p, cudaStream_t st) {
	int n = s->pool_capacity;
	if ((unsigned long long)n > p.total_items) n = (int)((p.total_items + 127ull) / 128ull * 128ull);
	GS_CUDA(s->pool_org.alloc(n)); GS_CUDA(s->pool_dir.alloc(n)); GS_CUDA(s->pool_thr.alloc(n)); GS_CUDA(s->pool_rad.alloc(n));
	GS_CUDA(s->pool_ids.alloc(n)); GS_CUDA(s->pool_hit.alloc(n)); GS_CUDA(s->free_list.alloc(n)); GS_CUDA(s->wave_ctr.alloc(1));
	if (!s->h_wave_ctr) GS_CUDA(cudaMallocHost(&s->h_wave_ctr, sizeof(WaveCounters)));
	p.pool.org = s->pool_org.p; p.pool.dir = s->pool_dir.p; p.pool.thr = s->pool_thr.p; p.pool.rad = s->pool_rad.p;
	p.pool.ids = s->pool_ids.p; p.pool.hit = s->pool_hit.p; p.pool.n = n;
	p.ctr = s->wave_ctr.p; p.free_list = s->free_list.p;
	GS_CUDA(cudaMemsetAsync(p.ctr, 0, sizeof(WaveCounters), st));
	const int kCheckEvery = 4;
	for (int wave = 0;; ++wave) {
		GS_CUDA(launch_wf_generate(p, wave == 0, st));
		GS_CUDA(cudaMemsetAsync(&p.ctr->next_ray, 0, 2 * sizeof(unsigned), st));  // next_ray and n_free
		GS_CUDA(launch_wf_traverse(p, s->sm_count, s->count_nodes, s->has_triangles, st));
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

int gs_render_accumulate(gs_scene* s, const gs_camera* camera, uint64_t seed, uint32_t sample_begin, uint32_t sample_count, void* d_accum,
                         void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum) return fail(GS_ERR_INVALID, "null accumulator");
	if (sample_count == 0) return GS_OK;
	RenderParams p;
	p.sc = s->ds;
	rc = convert_camera(s, camera, p.cam);
	if (rc) return rc;
	rc = ensure_origin_range(s, camera);
	if (rc) return rc;
	p.sc = s->ds;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	p.seed = seed; p.sample_begin = sample_begin; p.sample_count = sample_count;
	p.total_items = (unsigned long long)camera->image_width * camera->image_height * sample_count;
	if (s->use_wavefront) {
		WavefrontParams w;
		w.sc = s->ds; w.cam = p.cam; w.seed = seed; w.sample_begin = sample_begin; w.sample_count = sample_count;
		w.total_items = p.total_items; w.accum = static_cast<float4*>(d_accum); w.stats = s->stats.p;
		return render_wavefront(s, w, st);
	}
	p.chunk = 256;
	p.accum = static_cast<float4*>(d_accum);
	p.work_counter = next_counter(s);
	p.stats = s->stats.p;
	p.pixel_list = nullptr; p.n_list = nullptr; p.lum = nullptr;  // every pixel, no luminance statistics
	p.small_scene = s->small_scene;
	GS_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
	GS_CUDA(launch_render_fixed(p, s->sm_count, s->count_nodes, s->has_triangles, st));
	s->launches += 1;
	return GS_OK;
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
	AdaptiveParams ap;
	ap.confidence_sq = settings->confidence * settings->confidence;  // camera.rs:128-129
	ap.tolerance_sq = settings->tolerance * settings->tolerance;
	ap.max_samples = settings->max_samples;
	ap.n_pixels = n_pixels;
	ap.accum = static_cast<const float4*>(d_accum);
	// Camera::sample's loop (camera.rs:135-165) leaves after the first batch that takes `sample_count as u32` past
	// max_samples: at most max_samples / batch_size + 1 batches.
	const unsigned long long rounds = (unsigned long long)settings->max_samples / settings->batch_size + 1ull;
	// the per-pixel sample count lives in the accumulator's fourth float: exact up to 2^24
	if (rounds * settings->batch_size > (1ull << 24)) return fail(GS_ERR_UNSUPPORTED, "adaptive sampling beyond 2^24 samples per pixel");
	GS_CUDA(s->px_list[0].alloc(n_pixels)); GS_CUDA(s->px_list[1].alloc(n_pixels));
	GS_CUDA(s->px_count.alloc(2));
	GS_CUDA(s->lum.alloc(total));
	GS_CUDA(cudaMemsetAsync(s->lum.p, 0, (size_t)total * sizeof(double2), st));
	GS_CUDA(launch_adaptive_init(s->px_list[0].p, s->px_count.p, pixel_begin, pixel_stride, n_pixels, st));
	p.seed = seed;
	p.sample_count = settings->batch_size;
	p.total_items = (unsigned long long)n_pixels * settings->batch_size;  // upper bound; the kernel reads the list length
	p.chunk = 256;
	p.accum = static_cast<float4*>(d_accum);
	p.stats = s->stats.p;
	p.lum = s->lum.p;
	p.small_scene = s->small_scene;
	for (unsigned long long r = 0; r < rounds; ++r) {
		const int cur = (int)(r & 1ull), nxt = cur ^ 1;
		if ((unsigned long long)settings->batch_size * (r + 1ull) > 0xffffffffull) return fail(GS_ERR_UNSUPPORTED, "sample index beyond 32 bits");
		p.sample_begin = (unsigned)(r * settings->batch_size);
		p.pixel_list = s->px_list[cur].p;
		p.n_list = s->px_count.p + cur;
		p.work_counter = next_counter(s);
		GS_CUDA(cudaMemsetAsync(p.work_counter, 0, sizeof(unsigned long long), st));
		GS_CUDA(cudaMemsetAsync(s->px_count.p + nxt, 0, sizeof(unsigned), st));
		GS_CUDA(launch_render_fixed(p, s->sm_count, false, s->has_triangles, st));
		GS_CUDA(launch_adaptive_decide(ap, s->px_list[cur].p, s->px_count.p + cur, s->px_list[nxt].p, s->px_count.p + nxt, s->lum.p, st));
		s->launches += 2;
		// after batches 1, 2, 4, 8, ... look whether any pixel is still sampling (one 4-byte read): most renders end long
		// before the cap, and a launch pair that finds an empty list still costs ~10 us
		if (((r + 1ull) & r) == 0ull && r + 1 < rounds) {
			unsigned left = 0;
			GS_CUDA(cudaMemcpyAsync(&left, s->px_count.p + nxt, sizeof left, cudaMemcpyDeviceToHost, st));
			GS_CUDA(cudaStreamSynchronize(st));
			if (left == 0) break;
			if (left <= kAdaptiveTailPixels) {  // a handful of hard pixels: one warp each finishes them in a single launch
				AdaptiveTailParams tp;
				tp.sc = s->ds; tp.cam = p.cam; tp.seed = seed;
				tp.confidence_sq = ap.confidence_sq; tp.tolerance_sq = ap.tolerance_sq;
				tp.batch_size = settings->batch_size; tp.max_samples = settings->max_samples;
				tp.list = s->px_list[nxt].p; tp.n_list = s->px_count.p + nxt;
				tp.accum = static_cast<float4*>(d_accum); tp.lum = s->lum.p;
				tp.work_counter = next_counter(s); tp.stats = s->stats.p;
				GS_CUDA(cudaMemsetAsync(tp.work_counter, 0, sizeof(unsigned long long), st));
				GS_CUDA(launch_adaptive_tail(tp, left, s->sm_count, st));
				s->launches += 1;
				break;
			}
		}
	}
	return GS_OK;
}

int gs_resolve(gs_scene* s, const void* d_accum, int32_t width, int32_t height, float* out_rgb, uint32_t* out_counts, void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum || !out_rgb || width <= 0 || height <= 0) return fail(GS_ERR_INVALID, "bad resolve arguments");
	const int n = width * height;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	GS_CUDA(s->out_rgb.alloc((size_t)n * 3));
	GS_CUDA(s->out_counts.alloc((size_t)n));
	GS_CUDA(launch_resolve(static_cast<const float4*>(d_accum), n, s->out_rgb.p, s->out_counts.p, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb, s->out_rgb.p, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
	if (out_counts) GS_CUDA(cudaMemcpyAsync(out_counts, s->out_counts.p, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

int gs_resolve_rgb8(gs_scene* s, const void* d_accum, int32_t width, int32_t height, uint8_t* out_rgb8, uint32_t* out_counts, void* cuda_stream) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!d_accum || !out_rgb8 || width <= 0 || height <= 0) return fail(GS_ERR_INVALID, "bad resolve arguments");
	const int n = width * height;
	cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
	GS_CUDA(s->out_rgb.alloc((size_t)n * 3));  // reused as the byte buffer (n*3 bytes fit in n*3 floats)
	GS_CUDA(s->out_counts.alloc((size_t)n));
	unsigned char* d_rgb8 = reinterpret_cast<unsigned char*>(s->out_rgb.p);
	GS_CUDA(launch_resolve_rgb8(static_cast<const float4*>(d_accum), n, d_rgb8, s->out_counts.p, st));
	s->launches += 1;
	GS_CUDA(cudaMemcpyAsync(out_rgb8, d_rgb8, (size_t)n * 3, cudaMemcpyDeviceToHost, st));
	if (out_counts) GS_CUDA(cudaMemcpyAsync(out_counts, s->out_counts.p, (size_t)n * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
	GS_CUDA(cudaStreamSynchronize(st));
	return GS_OK;
}

#ifdef GS_DIVERGENCE_STATS
// diagnostic builds only: the raw device counters, read and reset
int gs_debug_counters(gs_scene* s, unsigned long long* out) {
	GS_CUDA(cudaDeviceSynchronize());
	GS_CUDA(cudaMemcpy(out, s->stats.p, kStatSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	GS_CUDA(cudaMemset(s->stats.p + 4, 0, (kStatSlots - 4) * sizeof(unsigned long long)));
	return GS_OK;
}
#endif

int gs_stats_read(gs_scene* s, gs_stats* stats) {
	int rc = check_device(s);
	if (rc) return rc;
	if (!stats) return fail(GS_ERR_INVALID, "null stats");
	GS_CUDA(cudaDeviceSynchronize());
	unsigned long long h[4];
	GS_CUDA(cudaMemcpy(h, s->stats.p, sizeof h, cudaMemcpyDeviceToHost));
	GS_CUDA(cudaMemset(s->stats.p, 0, sizeof h));
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
	GS_CUDA(s->accum.alloc(n));
	GS_CUDA(cudaMemsetAsync(s->accum.p, 0, n * sizeof(float4), 0));
	cudaEvent_t e0, e1;
	GS_CUDA(cudaEventCreate(&e0)); GS_CUDA(cudaEventCreate(&e1));
	GS_CUDA(cudaEventRecord(e0, 0));
	if (settings->max_samples < settings->batch_size) {
		// After the first batch `sample_count as u32 > max_samples` holds whatever the convergence test says
		// (camera.rs:153-164): exactly batch_size samples per pixel -> the fixed-count megakernel.  (tolerance
		// == 0 alone is not enough: on constant pixels the variance rounds below zero and `x < 0` stops early.)
		rc = gs_render_accumulate(s, camera, seed, 0, settings->batch_size, s->accum.p, nullptr);
	} else {
		rc = gs_render_adaptive_accumulate(s, camera, settings, seed, 0, 1, s->accum.p, nullptr);
	}
	if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
	GS_CUDA(cudaEventRecord(e1, 0));
	rc = gs_resolve(s, s->accum.p, camera->image_width, camera->image_height, out_rgb, out_counts, nullptr);
	if (rc) { cudaEventDestroy(e0); cudaEventDestroy(e1); return rc; }
	float ms = 0.f;
	GS_CUDA(cudaEventElapsedTime(&ms, e0, e1));
	cudaEventDestroy(e0); cudaEventDestroy(e1);
	if (stats) {
		rc = gs_stats_read(s, stats);
		if (rc) return rc;
		stats->kernel_ms = ms;
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

int gs_trace_closest(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t) {
	int rc = check_device(s);
	if (rc) return rc;
	if (n == 0) return GS_OK;
	if (!rays || !prim_id || !t) return fail(GS_ERR_INVALID, "null argument");
	double m = 0.0;
	for (size_t k = 0; k < n; ++k) for (int a = 0; a < 3; ++a) m = std::fmax(m, std::fabs(rays[k * 7 + a]));
	if (!std::isfinite(m)) return fail(GS_ERR_INVALID, "non-finite ray origin");
	if (m > s->origin_limit) { rc = build_nodes(s, 2.0 * m); if (rc) return rc; }
	DevBuf<double> d_rays, d_t; DevBuf<int> d_id;
	struct Rel { DevBuf<double>&a, &b; DevBuf<int>& c; ~Rel() { a.release(); b.release(); c.release(); } } rel{d_rays, d_t, d_id};
	GS_CUDA(d_rays.alloc(n * 7)); GS_CUDA(d_t.alloc(n)); GS_CUDA(d_id.alloc(n));
	GS_CUDA(cudaMemcpy(d_rays.p, rays, n * 7 * sizeof(double), cudaMemcpyHostToDevice));
	GS_CUDA(launch_trace_closest(s->ds, d_rays.p, n, tmin, tmax, d_id.p, d_t.p, 0));
	s->launches += 1;
	GS_CUDA(cudaMemcpy(prim_id, d_id.p, n * sizeof(int), cudaMemcpyDeviceToHost));
	GS_CUDA(cudaMemcpy(t, d_t.p, n * sizeof(double), cudaMemcpyDeviceToHost));
	return GS_OK;
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
	DevBuf<double> d_rays;
	struct Rel { DevBuf<double>& a; ~Rel() { a.release(); } } rel{d_rays};
	GS_CUDA(d_rays.alloc(n * 7));
	GS_CUDA(launch_camera_rays(cam, seed, sample, pixel_begin, pixel_end, d_rays.p, 0));
	s->launches += 1;
	GS_CUDA(cudaMemcpy(rays_out, d_rays.p, n * 7 * sizeof(double), cudaMemcpyDeviceToHost));
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
	DevBuf<unsigned> d_px, d_sm; DevBuf<float> d_out;
	struct Rel { DevBuf<unsigned>&a, &b; DevBuf<float>& c; ~Rel() { a.release(); b.release(); c.release(); } } rel{d_px, d_sm, d_out};
	GS_CUDA(d_px.alloc(n)); GS_CUDA(d_sm.alloc(n)); GS_CUDA(d_out.alloc(n * 3));
	GS_CUDA(cudaMemcpy(d_px.p, pixel, n * sizeof(unsigned), cudaMemcpyHostToDevice));
	GS_CUDA(cudaMemcpy(d_sm.p, sample, n * sizeof(unsigned), cudaMemcpyHostToDevice));
	GS_CUDA(launch_path_radiance(s->ds, cam, seed, d_px.p, d_sm.p, n, d_out.p, 0));
	s->launches += 1;
	GS_CUDA(cudaMemcpy(out_rgb, d_out.p, n * 3 * sizeof(float), cudaMemcpyDeviceToHost));
	return GS_OK;
}

}  // extern "C"
