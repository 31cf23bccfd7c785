// grayshift_b200/csrc/gs_device.cuh -- device-side data layout, RNG, intersection and BVH traversal
// shared by the render megakernel (render_kernels.cu) and the parity kernel (trace_closest.cu,
// compiled with -fmad=false).  sm_100a only.
//
// Layout in HBM (all immutable after gs_scene_create; C1-C4 working sets are L1/L2 resident,
// C5 ~100 MB fits the 126 MB L2):
//   nodes     DNode[ ]     64 B, two child boxes per node (fp32, padded outward), 4 x 128-bit loads   (binary tree: scenes
//                          whose nodes live in the L1)
//   nodes8    DNode8[ ]    96 B, eight children per node, boxes as 8-bit offsets on a power-of-two grid from the node's
//                          corner, children in octant order; 5 x 128-bit loads per visit (6th chunk only for quad /
//                          triangle leaves)                                                        (large scenes)
//   refs      uint32[ ]    leaf -> primitive references, kind<<30 | index, leaf order
//   spheres   DSphere[ ]   32 B f64 (cx,cy,cz,r; r keeps the sign the caller gave it, sphere.rs:91), leaf order, 2 x 128-bit
//                          loads; a MOVING sphere is flagged by reference kind 3 (same array)
//   sph_path  double4[ ]   only read for moving spheres
//   quads     DQuad[ ]     96 B f64 (plane n,d; barycentric functionals A,qa,B,qb), 6 x 128-bit loads
//   *_aux     int2[ ]      (material, prim_id) per primitive, read once per accepted hit
//   materials/textures     32 B records; images as uchar4 / float4 texels (one 32/128-bit load per lookup)
//
// Precision: boxes are tested in fp32 against conservatively padded bounds; primitives are
// intersected in f64 with the reference's formulas (B200 keeps a half-rate FP64 pipe and the
// reference is f64 throughout: ground sphere r=1000, t_min=0.001); shading is fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gs {

// ------------------------------------------------------------------------------------------
// Records
// ------------------------------------------------------------------------------------------
#ifndef GS_NODE_CH
#define GS_NODE_CH 1  // binary node boxes as (centre, half extent) instead of (min, max); see Trav2::node_step
#endif
struct __align__(16) DNode {
	float4 xy0;  // child 0: min.x, max.x, min.y, max.y      (GS_NODE_CH: centre.x, half.x, centre.y, half.y)
	float4 xy1;  // child 1
	float4 z01;  // c0.min.z, c0.max.z, c1.min.z, c1.max.z   (GS_NODE_CH: c0 centre.z, half.z, c1 centre.z, half.z)
	int4 link;   // child0, child1 (>= 0 internal node index; < 0 leaf code), unused, unused
};
// Eight-wide compressed node (large scenes: config 5's 10^6 spheres, whose binary nodes come from L2).  96 bytes:
//   chunk 0   lower corner of the node's box (fp32, padded outward), biased fp32 exponents of the grid step per axis, imask
//   chunk 1-3 per axis: the children's lower planes (slots 0-3, 4-7) and upper planes (slots 0-3, 4-7) in grid steps from
//             the corner, one byte each (lower planes rounded down, upper planes up: conservative); an empty slot holds
//             lower = 255 > upper = 0 and is never hit
//   chunk 4   meta byte per slot (leaf children: kind << 6 | offset into that kind's run of primitives; kind 3 = the leaf
//             code sits in leaf_links8[node * 8 + slot]: leaves of several primitives, moving spheres), child_base (internal
//             child in slot k is node child_base + popc(imask & ((1 << k) - 1))), sphere_base
//   chunk 5   quad_base, tri_base (read only when such a leaf is intersected)
// Children are assigned to slots by the octant of their centre relative to the node's, so that visiting the hit slots
// in the order of (slot XOR ray octant) is near-to-far without a sort (Ylitie, Karras, Laine 2017); a visit costs one
// dependent fetch where the binary tree needs three, and the traversal stack holds (node group, hit mask) pairs: at most
// one entry per level of a tree a third as deep.
struct __align__(16) DNode8 {
	float px, py, pz;
	uint32_t exps_imask;   // byte a = biased exponent of axis a's grid step (step = 2^(e - 127)), byte 3 = imask
	uint32_t qx[4];        // lo[0..3], lo[4..7], hi[0..3], hi[4..7]
	uint32_t qy[4];
	uint32_t qz[4];
	uint32_t meta[2];
	uint32_t child_base;
	uint32_t sphere_base;
	uint32_t quad_base, tri_base, pad0, pad1;
};
constexpr int kMaxStack8 = 40;  // (node group, hits) entries; gs_scene_create checks the tree against it
// leaf code: -1 - ((first_ref << 4) | count), 1 <= count <= 15; count == 0 marks a single-primitive leaf whose
// reference (kind << 25 | index) sits in the first_ref field itself
__host__ __device__ inline int leaf_code(uint32_t first, uint32_t count) { return -1 - (int)((first << 4) | count); }
// Negative (so the inner "internal node" loop stops on it) and not a valid leaf code: it would need
// first_ref = 2^27 - 1 with count 15, which gs_scene_create rejects.
constexpr int kStackSentinel = -2147483647 - 1;
constexpr int kMaxStack = 64;
constexpr int kNodePad = 512;  // the device node array holds at least this many entries (zero boxes beyond the tree), so a
                               // kernel may stage a fixed-size breadth-first prefix in shared memory without a bounds check

struct __align__(16) DSphere { double cx, cy, cz, r; };  // r is signed as given (a negative radius flips the normal, sphere.rs:91)
struct __align__(16) DQuad {
	double nx, ny, nz, d;    // unit plane normal and offset (quad.rs:32-36, plane.rs:15-18)
	double ax, ay, az, qa;   // alpha = P.A - qa with A = v x w    (== w . (planar x v), quad.rs:88)
	double bx, by, bz, qb;   // beta  = P.B - qb with B = w x u    (== w . (u x planar), quad.rs:89)
};
struct __align__(16) DTriangle {
	double ax, ay, az, pad0;     // vertex a
	double e1x, e1y, e1z, pad1;  // edge_1 = c - a (triangle.rs:35)
	double e2x, e2y, e2z, pad2;  // edge_2 = b - a (triangle.rs:36)
};
struct __align__(16) DMaterial { int kind; int tex; float ar, ag, ab; float param; int pad0, pad1; };
struct __align__(16) DTexture { int kind; int a; int b; float scale; float cr, cg, cb; float pad; double scale_inv_d; double scale_d; };
struct DImage { int w, h; int format; int pad; const void* texels; };  // uchar4 (u8) or float4 (f32) texels, row 0 = top
struct DMedium { int sph_first, sph_count, quad_first, quad_count; double neg_inv_density; int phase_mat; int flags; };  // flags bit 0: the single boundary sphere moves

constexpr uint32_t kRefKindShift = 30;
constexpr uint32_t kRefIndexMask = (1u << 30) - 1u;
// reference kinds on the device: 0 sphere, 1 quad, 2 triangle (= GS_PRIM_*), 3 = a MOVING sphere (index into `spheres`)
constexpr uint32_t kKindSphere = 0u, kKindQuad = 1u, kKindTriangle = 2u, kKindMovingSphere = 3u;

struct DeviceScene {
	const DNode* nodes;
	const DNode8* nodes8;       // the eight-wide tree (kernels with the wide traversal); either array may be null
	const int* leaf_links8;     // nodes8: leaf code of slot k of node n at [n * 8 + k] where the meta byte says kind 3
	const uint32_t* refs;
	const DSphere* spheres;
	const double4* sph_path;
	const int2* sph_aux;
	const float2* sph_uvrot;    // (cos, sin) of the RotateY an instance was baked with: sphere_uv works on the OBJECT-space normal
	const DQuad* quads;
	const int2* quad_aux;
	const int* quad_order;  // reference visiting rank, read only on exact-t ties
	const DTriangle* tris;
	const int2* tri_aux;
	const DSphere* bnd_spheres;
	const double4* bnd_sph_path;
	const DQuad* bnd_quads;
	const DMedium* media;
	const DMaterial* materials;
	const DTexture* textures;
	const DImage* images;
	const uint8_t* perlin_perm;
	int n_media;
	int root;  // node index or leaf code
	uint32_t prmt_hi;  // 0x47000000, passed as DATA so that the compiler keeps it in a register (see Trav8::qf)
};

struct DCamera {
	double center[3], pixel00[3], du[3], dv[3], disk_u[3], disk_v[3];
	double defocus_angle;
	int width, height;
	unsigned max_depth;
	int bg_kind;
	float bg[3];
	int hdri_image;
	float rot[9];  // rotate_vector's matrix (util.rs:67-86), rows x,y,z
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10, counter = (pixel, sample, block, stream), key = seed.  Draw k of a path is
// word k&3 of block k>>2 (the oracle uses the identical scheme).
// ------------------------------------------------------------------------------------------
// (An out-of-line copy was tried to shrink the megakernel's code: it cost 60 registers around the calls, so the
// rounds stay inline and rolled up: 10 iterations of 8 instructions instead of 25 inlined copies of 80.)
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#ifndef GS_PHILOX_UNROLL
#define GS_PHILOX_UNROLL 2  // tuning knob: 1 = rolled loop (11 instructions per round), 2 = two rounds per iteration (7.5 per round; measured +1 % C4,
                            // +5 % C1), 5 / 10 = more unrolling: faster on C1/C3 but -10...-15 % on C4, whose many code paths overflow the I-cache
#endif
#define GS_STR_(x) #x
#define GS_PRAGMA_UNROLL_(n) _Pragma(GS_STR_(unroll n))
	GS_PRAGMA_UNROLL_(GS_PHILOX_UNROLL)
	for (int r = 0; r < 10; ++r) {
		uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
	}
	return make_uint4(c0, c1, c2, c3);
}

// The same ten rounds with the round keys (k0 + r * 0x9E3779B9, k1 + r * 0xBB67AE85) read from `rk`: when that is the kernel's
// parameter block (RenderParams::philox_rk, filled by launch_render_fixed) each key is a constant-bank operand of the round's
// three-input XOR, and a round is four instructions -- two wide multiplies, two XORs -- with no key arithmetic and no loop.
__device__ __forceinline__ uint4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* __restrict__ rk) {
#pragma unroll
	for (int r = 0; r < 10; ++r) {
		uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		uint32_t n0 = hi1 ^ c1 ^ rk[2 * r], n2 = hi0 ^ c3 ^ rk[2 * r + 1];
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
	}
	return make_uint4(c0, c1, c2, c3);
}
__host__ __device__ inline void philox_round_keys(unsigned long long seed, uint32_t (&rk)[20]) {
	uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
	for (int r = 0; r < 10; ++r) { rk[2 * r] = k0; rk[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

// kRK: the round keys come from `rk` (see philox4x32_10_rk) instead of (k0, k1); the members a build does not use cost nothing.
template <bool kRK>
struct PathRngT {
	uint32_t pixel, sample, draw, k0, k1;
	const uint32_t* rk;
	uint4 buf;
	__device__ __forceinline__ uint4 block(uint32_t b) const {
		if (kRK) return philox4x32_10_rk(pixel, sample, b, 0u, rk);
		return philox4x32_10(pixel, sample, b, 0u, k0, k1);
	}
	__device__ __forceinline__ void start(uint64_t seed, uint32_t pixel_, uint32_t sample_) {
		pixel = pixel_; sample = sample_; draw = 0; k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
	}
	__device__ __forceinline__ uint32_t next_u32() {
		uint32_t w = draw & 3u;
		if (w == 0) buf = block(draw >> 2);
		++draw;
		return w == 0 ? buf.x : (w == 1 ? buf.y : (w == 2 ? buf.z : buf.w));
	}
	// The next K (<= 4) draws at once -- the same words next_u32() would return one by one.  With single draws every call
	// site carries its own copy of the ten Philox rounds and runs it for the quarter of the lanes whose draw counter sits
	// on a block boundary there (measured 2.8 active lanes on config 5, 7 % of the kernel's instructions); taking the draws
	// of a bounce together needs at most one block, computed once for all lanes that need one.
	// The words come from old[w..3] followed by the new block; with w == 0 nothing of the old block is left.
#ifndef GS_RNG_TAKE
#define GS_RNG_TAKE 1  // tuning knob: 0 = K single draws
#endif
	template <int K>
	__device__ __forceinline__ void take(uint32_t (&out)[K]) {
		static_assert(K >= 1 && K <= 3, "take: one to three draws");
#if GS_RNG_TAKE
		const uint32_t w = draw & 3u;
		const uint4 o = buf;
		if (w == 0u || w + (uint32_t)K > 4u) buf = block((draw + 3u) >> 2);  // the block after `o`, or of draw itself when w == 0
		const uint4 n = buf;
		const bool w0 = w == 0u, w1 = w == 1u, w2 = w == 2u;
		out[0] = w0 ? n.x : (w1 ? o.y : (w2 ? o.z : o.w));
		if (K > 1) out[1] = w0 ? n.y : (w1 ? o.z : (w2 ? o.w : n.x));
		if (K > 2) out[2] = w0 ? n.z : (w1 ? o.w : (w2 ? n.x : n.y));
		draw += (uint32_t)K;
#else
#pragma unroll
		for (int i = 0; i < K; ++i) out[i] = next_u32();
#endif
	}
	__device__ __forceinline__ static float to_f32(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }
	__device__ __forceinline__ static double to_f64(uint32_t u) { return (double)u * (1.0 / 4294967296.0); }
	__device__ __forceinline__ float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }  // [0,1), 24 bits
	__device__ __forceinline__ double next_f64() { return (double)next_u32() * (1.0 / 4294967296.0); }       // the oracle's value exactly
	__device__ __forceinline__ void ensure(int) {}  // blocks are computed where a draw needs one (see PathRngW)
};
using PathRng = PathRngT<false>;

// The same stream -- draw k of a path is word k & 3 of block k >> 2 -- with the blocks computed AHEAD, at a point where the warp's
// lanes are together, instead of at whichever of a dozen divergent draw sites a lane's counter happens to cross a block boundary
// (config 4, profiles/r02_v13: the Philox rounds and the word selection were 11 % of the megakernel's warp instructions at 12 active
// lanes).  A lane keeps a window of two blocks in shared memory, win[(k & 7) * kStride] = word of draw k for the blocks `top - 1` and
// `top` (win[8 * kStride] = top); a draw is one shared-memory load, and ensure(n) computes blocks until n more draws are covered.
// Two blocks hold at least five draws past any counter, so ensure(n <= 5) can always be met.  The megakernel calls ensure(5) once
// per level, after the media draws of segment_begin (all lanes with a fresh ray are there): enough for the first round of any
// material (<= 3 draws) plus two media draws of the next level; rejection loops and a third medium call it again before they draw more.
template <int kStride>
struct PathRngW {
	uint32_t pixel, sample, draw, k0, k1;
	uint32_t* win;  // this thread's column of the CTA's window array [9][kStride]
	__device__ __forceinline__ void ensure(int n) {
		int top = (int)win[8 * kStride];
		if ((top + 1) * 4 - (int)draw >= n) return;
		do {
			++top;
			const uint4 b = philox4x32_10(pixel, sample, (uint32_t)top, 0u, k0, k1);
			uint32_t* const w = win + (top & 1) * (4 * kStride);
			w[0] = b.x; w[kStride] = b.y; w[2 * kStride] = b.z; w[3 * kStride] = b.w;
		} while ((top + 1) * 4 - (int)draw < n);
		win[8 * kStride] = (uint32_t)top;
	}
	__device__ __forceinline__ void start(uint64_t seed, uint32_t pixel_, uint32_t sample_) {
		pixel = pixel_; sample = sample_; draw = 0; k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
		win[8 * kStride] = 0xffffffffu;  // top = -1: nothing computed
		ensure(5);                       // blocks 0 and 1: the camera ray (3 draws; 2 + a first lens pair + 1 with a lens) and two media
	}
	__device__ __forceinline__ uint32_t next_u32() {
		const uint32_t v = win[(draw & 7u) * kStride];
		++draw;
		return v;
	}
	template <int K>
	__device__ __forceinline__ void take(uint32_t (&out)[K]) {
#pragma unroll
		for (int i = 0; i < K; ++i) out[i] = next_u32();
	}
	__device__ __forceinline__ float next_f32() { return PathRng::to_f32(next_u32()); }
	__device__ __forceinline__ double next_f64() { return PathRng::to_f64(next_u32()); }
};

// ------------------------------------------------------------------------------------------
// Ray
// ------------------------------------------------------------------------------------------
struct Ray {
	double ox, oy, oz;
	double dx, dy, dz;
	double time;
};

struct Hit {
	double t;
	uint32_t ref;  // kind<<30 | index into spheres/quads; 0xffffffff = miss
	float u, v;    // triangle (u, v) for the textures (only kernels built for triangles touch them); spheres and quads
	               // compute their texture coordinates on demand
};
constexpr uint32_t kNoHit = 0xffffffffu;

// f64 division and square root.  The parity translation unit (trace_closest.cu, GS_EXACT_F64) uses the IEEE
// operators like the reference; the render kernels use the approximation instruction plus two Newton steps
// (<= 2 ulp, no special-case slow path: 8 instructions instead of ~15 plus a call; every operand here is a finite
// normal number).
#ifndef GS_FAST_F64
#define GS_FAST_F64 1  // tuning knob: 0 = IEEE division / square root in the render kernels too (measured 3-8 % slower)
#endif
#if defined(GS_EXACT_F64) || !GS_FAST_F64
__device__ __forceinline__ double f64_div(double n, double d) { return n / d; }
__device__ __forceinline__ double f64_rcp(double d) { return 1.0 / d; }
__device__ __forceinline__ double f64_sqrt(double x) { return sqrt(x); }
#else
__device__ __forceinline__ double f64_rcp(double d) {
	double x;
	asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
	double e = fma(-d, x, 1.0);
	x = fma(x, e, x);
	e = fma(-d, x, 1.0);
	return fma(x, e, x);
}
__device__ __forceinline__ double f64_div(double n, double d) {
	const double x = f64_rcp(d);
	const double q = n * x;
	return fma(fma(-d, q, n), x, q);
}
__device__ __forceinline__ double f64_sqrt(double v) {
	const double x = fmax(v, 1e-280);  // callers pass v >= 0; sqrt(0) becomes 1e-140
	double y;
	asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
	double g = x * y, h = 0.5 * y;
	double r = fma(-h, g, 0.5);
	g = fma(g, r, g); h = fma(h, r, h);
	r = fma(-h, g, 0.5);
	g = fma(g, r, g); h = fma(h, r, h);
	return fma(fma(-g, g, x), h, g);
}
#endif

// Sphere::hit, sphere.rs:64-92.  `a` = |d|^2 and its reciprocal are hoisted out by the caller (ray constants):
// the two roots (h -+ sqrt(disc)) / a become multiplications by 1/a (one f64 division per ray instead of up to
// two per candidate; differs from the reference's quotient by at most 1 ulp).  Strict interval (surrounds,
// interval.rs:33-35).
__device__ __forceinline__ bool hit_sphere(const DSphere& s, const double4* __restrict__ path, uint32_t index, bool moving, const Ray& r,
                                           double a, double inv_a, double tmin, double tmax, double& t_out) {
	double cx = s.cx, cy = s.cy, cz = s.cz;
	const double rad = s.r;  // enters only squared here; its sign matters for the normal (sphere.rs:91)
	if (moving) {  // center_start + time * center_path (sphere.rs:51-53)
		double4 p = path[index];
		cx = cx + r.time * p.x; cy = cy + r.time * p.y; cz = cz + r.time * p.z;
	}
	double ocx = cx - r.ox, ocy = cy - r.oy, ocz = cz - r.oz;
	double h = r.dx * ocx + r.dy * ocy + r.dz * ocz;
	double c = (ocx * ocx + ocy * ocy + ocz * ocz) - rad * rad;
	double disc = h * h - a * c;
	if (disc < 0.0) return false;
	double sq = f64_sqrt(disc);
	double t = (h - sq) * inv_a;
	if (!(tmin < t && t < tmax)) {
		t = (h + sq) * inv_a;
		if (!(tmin < t && t < tmax)) return false;
	}
	t_out = t;
	return true;
}

// Quad::hit + Plane::hit, quad.rs:84-93 and plane.rs:20-32.  inclusive interval (contains).
__device__ __forceinline__ bool hit_quad(const DQuad& q, const Ray& r, double tmin, double tmax, double& t_out, double& alpha, double& beta) {
	double den = q.nx * r.dx + q.ny * r.dy + q.nz * r.dz;
	if (fabs(den) < 1e-8) return false;
	double t = f64_div(q.d - (q.nx * r.ox + q.ny * r.oy + q.nz * r.oz), den);
	if (!(tmin <= t && t <= tmax)) return false;
	double px = r.ox + r.dx * t, py = r.oy + r.dy * t, pz = r.oz + r.dz * t;
	double al = (px * q.ax + py * q.ay + pz * q.az) - q.qa;
	double be = (px * q.bx + py * q.by + pz * q.bz) - q.qb;
	if (!(0.0 <= al && al <= 1.0) || !(0.0 <= be && be <= 1.0)) return false;
	t_out = t; alpha = al; beta = be;
	return true;
}

// Triangle::hit, triangle.rs:34-68: Moeller-Trumbore as the reference writes it (back faces culled, det < 1e-8),
// plus the interval test the reference forgets (strict like the spheres).  u, v are the reference's.
__device__ __forceinline__ bool hit_triangle(const DTriangle& tr, const Ray& r, double tmin, double tmax, double& t_out, double& u_out,
                                             double& v_out) {
	const double px = r.dy * tr.e2z - r.dz * tr.e2y, py = r.dz * tr.e2x - r.dx * tr.e2z, pz = r.dx * tr.e2y - r.dy * tr.e2x;  // d x edge_2
	const double det = tr.e1x * px + tr.e1y * py + tr.e1z * pz;
	if (det < 1e-8) return false;
	const double tx = r.ox - tr.ax, ty = r.oy - tr.ay, tz = r.oz - tr.az;
	double u = tx * px + ty * py + tz * pz;
	if (u < 0.0 || u > det) return false;
	const double qx = ty * tr.e1z - tz * tr.e1y, qy = tz * tr.e1x - tx * tr.e1z, qz = tx * tr.e1y - ty * tr.e1x;  // t_vec x edge_1
	double v = r.dx * qx + r.dy * qy + r.dz * qz;
	if (v < 0.0 || u + v > det) return false;
	double t = tr.e2x * qx + tr.e2y * qy + tr.e2z * qz;
	const double inv_det = 1.0 / det;
	t *= inv_det; u *= inv_det; v *= inv_det;
	if (!(tmin < t && t < tmax)) return false;
	t_out = t; u_out = u; v_out = v;
	return true;
}

// 256-bit read-only loads (sm_100: LDG.E.256).  Every lane walks its own ray, so a load instruction costs the L1 one tag
// lookup per active lane whatever its width: a 64-byte node in two 32-byte loads instead of four 16-byte ones halves the
// request rate that binds the traversal once the tree lives in L2 (config 5: L1TEX at 82 % of its peak with 16-byte loads).
// The pointer must be 32-byte aligned (all records here are 32, 64 or 96 bytes in 256-byte aligned arrays).
// Measured (profiles/r02_variants.log): +3...4 % on the 10^6-sphere scene, -3 % on the L1-resident final scene (the 8-register
// alignment of the destination costs the register-bound kernel a few spills), hence kWide: the kernels of the eight-wide tree
// -- the large scenes -- use them, the binary-tree kernels keep 128-bit loads.
#ifndef GS_LD256
#define GS_LD256 1  // tuning knob: 0 = 128-bit loads only, 1 = 256-bit loads in the eight-wide kernels, 2 = everywhere
#endif
template <bool kWide> struct Ld256 { static constexpr bool on = GS_LD256 == 2 || (GS_LD256 == 1 && kWide); };
template <bool kOn>
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
#if defined(__CUDA_ARCH__)
	if (kOn) {
		asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		    : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
		return;
	}
#endif
	a = __ldg(reinterpret_cast<const uint4*>(p)); b = __ldg(reinterpret_cast<const uint4*>(p) + 1);
}
template <bool kOn>
__device__ __forceinline__ void ldg256(const void* p, float4& a, float4& b) {
#if defined(__CUDA_ARCH__)
	if (kOn) {
		asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		    : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
		return;
	}
#endif
	a = __ldg(reinterpret_cast<const float4*>(p)); b = __ldg(reinterpret_cast<const float4*>(p) + 1);
}
template <bool kOn>
__device__ __forceinline__ void ldg256(const void* p, double2& a, double2& b) {
#if defined(__CUDA_ARCH__)
	if (kOn) {
		asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
		return;
	}
#endif
	a = __ldg(reinterpret_cast<const double2*>(p)); b = __ldg(reinterpret_cast<const double2*>(p) + 1);
}

template <bool kWide = false>
__device__ __forceinline__ DSphere load_sphere(const DSphere* __restrict__ p, uint32_t i) {
	double2 a, b;
	ldg256<Ld256<kWide>::on>(p + i, a, b);
	return DSphere{a.x, a.y, b.x, b.y};
}
template <bool kWide = false>
__device__ __forceinline__ DQuad load_quad(const DQuad* __restrict__ p, uint32_t i) {
	const double2* q = reinterpret_cast<const double2*>(p + i);
	double2 a, b, c, d, e, f;
	ldg256<Ld256<kWide>::on>(q, a, b); ldg256<Ld256<kWide>::on>(q + 2, c, d); ldg256<Ld256<kWide>::on>(q + 4, e, f);
	return DQuad{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, e.x, e.y, f.x, f.y};
}

template <bool kWide = false>
__device__ __forceinline__ DTriangle load_triangle(const DTriangle* __restrict__ p, uint32_t i) {
	const double2* q = reinterpret_cast<const double2*>(p + i);
	double2 a, b, c, d, e, f;
	ldg256<Ld256<kWide>::on>(q, a, b); ldg256<Ld256<kWide>::on>(q + 2, c, d); ldg256<Ld256<kWide>::on>(q + 4, e, f);
	return DTriangle{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, e.x, e.y, f.x, f.y};
}

#ifndef GS_APPROX_RCP
#define GS_APPROX_RCP 1  // 1 / d of the slab test with rcp.approx.ftz.f32 (one MUFU, relative error <= 2^-23) instead of the IEEE division
                         // (eight instructions plus a slow path), three times per ray and traversal phase: +2.7 % config 4, +2.0 % C3, +1.4 % C1,
                         // +0.9 % C5 (profiles/r02_v15_ab.log section 9).  The boxes' padding budgets a full 2^-23 for the reciprocal
                         // (scene_layout.cu, "fp32 error budget"), so the test stays conservative; the parity kernel (GS_EXACT_F64) keeps the division.
                         // (ftz: 1 / d underflows to 0 only for |d| > 8.5e37)
#endif
__device__ __forceinline__ float safe_rcp(float d) {
	// 1/0 = inf in the reference (AABB.rs:64); a huge finite inverse orders the slabs identically without NaNs
	// (1e25 keeps origin * inverse finite for any origin the padded boxes allow)
	float ad = fabsf(d);
#if GS_APPROX_RCP && defined(__CUDA_ARCH__) && !defined(GS_EXACT_F64)
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
	return ad < 1e-25f ? copysignf(1e25f, d) : r;
#else
	return ad < 1e-25f ? copysignf(1e25f, d) : 1.0f / d;
#endif
}

// ------------------------------------------------------------------------------------------
// Closest hit over the surface BVH: BVHNode::hit + AABB::hit + HittableList::hit collapsed into an
// ordered while-while traversal with a per-thread stack.  Result equals the reference's
// unordered recursion (BVH.rs:68-90); exact ties follow the reference's visiting order (Q9).
//
// The traversal is RESUMABLE: its whole state is (node, sp, best, stack), so the render megakernel
// can leave the loop when enough lanes of the warp have finished their ray, shade / regenerate
// those lanes, and re-enter with the unfinished rays exactly where they stopped (ncu on the first
// version showed 8.4 of 32 lanes active in the node loop: lanes idled until the warp's longest ray ended).
// ------------------------------------------------------------------------------------------
#ifndef GS_NODE_LOAD
#define GS_NODE_LOAD 1  // tuning knob: 1 = branch-free child selection (the link load is issued together with the box loads)
#endif

// The traversal stack of one thread.  Its first kSmem entries may live in shared memory, laid out [entry][thread] (no bank
// conflicts: a warp's lanes hit 32 consecutive words whatever their stack depths are... per entry), the rest in the per-thread
// local array (the round-1 profile: the all-local stack made up 30 % of the kernel's L1 traffic at 1.8 useful bytes per
// 32-byte sector).  kSmem = 0: local memory only (the parity and per-path kernels).
template <int kSmem, int kStride>
struct StackT {
	int* local;
	int* smem;  // &s_stack[0][thread]; entry i at smem[i * kStride]
	__device__ __forceinline__ int get(int i) const {
		if (kSmem > 0 && i < kSmem) return smem[i * kStride];
		return local[kSmem > 0 ? i - kSmem : i];
	}
	__device__ __forceinline__ void set(int i, int v) const {
		if (kSmem > 0 && i < kSmem) smem[i * kStride] = v;
		else local[kSmem > 0 ? i - kSmem : i] = v;
	}
};
using LocalStack = StackT<0, 1>;

struct RayConsts {  // per-ray values derived from the Ray; recomputed whenever a traversal is (re)entered
	float idx, idy, idz;     // 1 / direction
	float oix, oiy, oiz;     // origin * (1 / direction): slab distance = fma(plane, id, -oi)
	float tmax32;            // fp32 upper bound of the current closest t, for box culling
};

__device__ __forceinline__ float tmax_to_f32(double t) { return t >= 3.0e38 ? 3.4028235e38f : __double2float_ru(t); }

__device__ __forceinline__ RayConsts ray_consts(const Ray& r, double t_closest) {
	RayConsts c;
	c.idx = safe_rcp((float)r.dx); c.idy = safe_rcp((float)r.dy); c.idz = safe_rcp((float)r.dz);
	c.oix = (float)r.ox * c.idx; c.oiy = (float)r.oy * c.idy; c.oiz = (float)r.oz * c.idz;
	c.tmax32 = tmax_to_f32(t_closest);
	return c;
}

// One leaf: HittableList::hit over its primitives (hittable.rs:70-86) with the reference's interval rules.
// `leaf` is a leaf code.  kTris = the scene holds triangles (kernels for scenes without any are
// instantiated without the triangle branch: it costs registers and instruction-cache space in the hot loop).
template <bool kTris, bool kWide = false>
__device__ __forceinline__ void leaf_intersect(const DeviceScene& sc, const Ray& r, RayConsts& rc, double tmin, Hit& best, int leaf) {
	const uint32_t code = (uint32_t)(-1 - leaf);
	uint32_t first = code >> 4, count = code & 15u;
	// count == 0: a single-primitive leaf whose reference is stored in the link itself (kind << 25 | index):
	// no dependent load from `refs` between the node and the primitive
	const bool direct = count == 0u;
	if (direct) count = 1u;
	for (uint32_t k = 0; k < count; ++k) {
		const uint32_t ref = direct ? (((first >> 25) << kRefKindShift) | (first & ((1u << 25) - 1u))) : __ldg(sc.refs + first + k);
		const uint32_t index = ref & kRefIndexMask;
		const uint32_t kind = ref >> kRefKindShift;
		if (kTris && kind == kKindTriangle) {
			const DTriangle tr = load_triangle<kWide>(sc.tris, index);
			double t, u, v;
			if (hit_triangle(tr, r, tmin, best.t, t, u, v)) { best.t = t; best.ref = ref; best.u = (float)u; best.v = (float)v; }
		} else if (kind == kKindQuad) {
			const DQuad q = load_quad<kWide>(sc.quads, index);
			double t, al, be;
			if (hit_quad(q, r, tmin, best.t, t, al, be)) {
				// exact tie between two quads: the one the reference visits later wins (t <= closest, plane.rs:27)
				bool accept = true;
				if (t == best.t && best.ref != kNoHit && (best.ref >> kRefKindShift) == kKindQuad)
					accept = __ldg(sc.quad_order + index) > __ldg(sc.quad_order + (best.ref & kRefIndexMask));
				// (alpha, beta) are the quad's texture coordinates (quad.rs:101-102); the render kernels recompute them from
				// the hit point when an image texture asks for them instead of carrying two registers through the traversal
				if (accept) { best.t = t; best.ref = ref; }
			}
		} else {  // sphere, stationary (kind 0) or moving (kind 3)
			const DSphere s = load_sphere<kWide>(sc.spheres, index);
			// |d|^2 (Vec3::length_squared, sphere.rs:71) and its reciprocal are recomputed per test: eleven f64
			// instructions, against four registers held through the whole traversal of a register-bound kernel
			const double a = r.dx * r.dx + r.dy * r.dy + r.dz * r.dz, inv_a = f64_rcp(a);
			double t;
			if (hit_sphere(s, sc.sph_path, index, kind == kKindMovingSphere, r, a, inv_a, tmin, best.t, t)) { best.t = t; best.ref = ref; }
		}
	}
	if (best.ref != kNoHit) rc.tmax32 = __double2float_ru(best.t);
}

// ==========================================================================================
// Traversal policy 1: the binary tree (DNode).  State = (node, sp) + the per-thread stack of links.
// ==========================================================================================
#ifndef GS_PRMT_IMM
#define GS_PRMT_IMM 1  // eight-wide node test: PRMT selectors as immediates (0 = leave the operand choice to the compiler)
#endif
#ifndef GS_POP_CULL
#define GS_POP_CULL 0  // binary tree: push the far child's entry distance with its link and drop it at pop time if the closest hit
                       // has come nearer since (host simulation: 32...48 % of the pops are stale on C1 / C4 / C5, each one a wasted node
                       // or leaf visit) -- costs a second stack word per push
#endif
struct Trav2 {
	static constexpr int kStackWords = GS_POP_CULL ? 2 * kMaxStack : kMaxStack;
	int node;   // current node index / leaf code, kStackSentinel when the ray is finished
	int sp;     // stack entries in use
	Hit best;

	template <class Stack>
	__device__ __forceinline__ void init(const DeviceScene& sc, const Ray&, const Stack& stack, double tmax) {
		stack.set(0, kStackSentinel);
#if GS_POP_CULL
		stack.set(1, 0);  // entry distance 0: the sentinel is never culled
		sp = 2;
#else
		sp = 1;
#endif
		node = sc.root;
		best.t = tmax; best.ref = kNoHit; best.u = 0.f; best.v = 0.f;
	}
	__device__ __forceinline__ void park() { node = kStackSentinel; sp = 0; }
	__device__ __forceinline__ bool done() const { return node == kStackSentinel; }
	__device__ __forceinline__ bool wants_node() const { return node >= 0; }
	template <class Stack>
	__device__ __forceinline__ void pop(const Stack& stack, const RayConsts& rc) {
#if GS_POP_CULL
		float en;
		do { sp -= 2; node = stack.get(sp); en = __int_as_float(stack.get(sp + 1)); } while (en > rc.tmax32);
#else
		node = stack.get(--sp);
#endif
	}
	__device__ __forceinline__ bool wants_leaf() const { return node < 0 && node != kStackSentinel; }
	__device__ __forceinline__ uint32_t pending_leaf_kind() const { return ((uint32_t)(-1 - node) >> 4) >> 25; }  // diagnostics only

	// One internal node: AABB::hit (AABB.rs:58-113) for both children at once, near child first, far child pushed.
	// Slab distances in fp32 as fma(plane, 1/d, -o/d) against boxes padded by 8 ulp of the origin range
	// (scene_layout.cu build_nodes): conservative, never culls a true hit.  Precondition: wants_node().
	// kTop > 0: the first kTop nodes (breadth-first order = the top levels of the tree) are read from the CTA's
	// shared-memory copy `top`; one generic-address load serves both cases, so lanes do not diverge on it.
	template <bool kCount, int kTop, class Stack>
	__device__ __forceinline__ void node_step(const DeviceScene& sc, const RayConsts& rc, const Stack& stack, uint32_t* node_visits, const float4* top) {
		const float4* np = reinterpret_cast<const float4*>(sc.nodes + node);
		float4 xy0, xy1, z01;
		int link0, link1;
		if (kTop > 0) {
			if (node < kTop) np = top + node * 4;
			xy0 = np[0]; xy1 = np[1]; z01 = np[2];
			const int2 link = *reinterpret_cast<const int2*>(np + 3);
			link0 = link.x; link1 = link.y;
		} else {
			float4 lk;
			ldg256<Ld256<false>::on>(np, xy0, xy1); ldg256<Ld256<false>::on>(np + 2, z01, lk);
			link0 = __float_as_int(lk.x); link1 = __float_as_int(lk.y);
		}
		if (kCount) ++(*node_visits);
#if GS_NODE_CH
		// Boxes as (centre, half extent): the slab interval of an axis is tm -+ half * |1/d| around tm = (centre - o) / d --
		// three FMAs (the |.| and the sign are operand modifiers), where (min, max) planes cost two FMAs plus a min and a max.
		// The ALU pipe (min / max / compare / select) is the busier one in this loop: six fewer instructions per node and
		// twelve moved to the FMA pipe.  Same conservative padding as the (min, max) form.
		float tm, th;
		tm = fmaf(xy0.x, rc.idx, -rc.oix); th = xy0.y * fabsf(rc.idx);
		const float nx0 = tm - th, fx0 = tm + th;
		tm = fmaf(xy0.z, rc.idy, -rc.oiy); th = xy0.w * fabsf(rc.idy);
		const float ny0 = tm - th, fy0 = tm + th;
		tm = fmaf(z01.x, rc.idz, -rc.oiz); th = z01.y * fabsf(rc.idz);
		const float nz0 = tm - th, fz0 = tm + th;
		const float en0 = fmaxf(fmaxf(nx0, ny0), fmaxf(nz0, 0.0f));
		const float ex0 = fminf(fminf(fx0, fy0), fminf(fz0, rc.tmax32));
		tm = fmaf(xy1.x, rc.idx, -rc.oix); th = xy1.y * fabsf(rc.idx);
		const float nx1 = tm - th, fx1 = tm + th;
		tm = fmaf(xy1.z, rc.idy, -rc.oiy); th = xy1.w * fabsf(rc.idy);
		const float ny1 = tm - th, fy1 = tm + th;
		tm = fmaf(z01.z, rc.idz, -rc.oiz); th = z01.w * fabsf(rc.idz);
		const float nz1 = tm - th, fz1 = tm + th;
		const float en1 = fmaxf(fmaxf(nx1, ny1), fmaxf(nz1, 0.0f));
		const float ex1 = fminf(fminf(fx1, fy1), fminf(fz1, rc.tmax32));
#else
		float t0 = fmaf(xy0.x, rc.idx, -rc.oix), t1 = fmaf(xy0.y, rc.idx, -rc.oix);
		float t2 = fmaf(xy0.z, rc.idy, -rc.oiy), t3 = fmaf(xy0.w, rc.idy, -rc.oiy);
		float t4 = fmaf(z01.x, rc.idz, -rc.oiz), t5 = fmaf(z01.y, rc.idz, -rc.oiz);
		const float en0 = fmaxf(fmaxf(fminf(t0, t1), fminf(t2, t3)), fmaxf(fminf(t4, t5), 0.0f));
		const float ex0 = fminf(fminf(fmaxf(t0, t1), fmaxf(t2, t3)), fminf(fmaxf(t4, t5), rc.tmax32));
		t0 = fmaf(xy1.x, rc.idx, -rc.oix); t1 = fmaf(xy1.y, rc.idx, -rc.oix);
		t2 = fmaf(xy1.z, rc.idy, -rc.oiy); t3 = fmaf(xy1.w, rc.idy, -rc.oiy);
		t4 = fmaf(z01.z, rc.idz, -rc.oiz); t5 = fmaf(z01.w, rc.idz, -rc.oiz);
		const float en1 = fmaxf(fmaxf(fminf(t0, t1), fminf(t2, t3)), fmaxf(fminf(t4, t5), 0.0f));
		const float ex1 = fminf(fminf(fmaxf(t0, t1), fmaxf(t2, t3)), fminf(fmaxf(t4, t5), rc.tmax32));
#endif
		const bool h0 = en0 <= ex0, h1 = en1 <= ex1;
#if GS_NODE_LOAD == 1
		// Select form: both links are consumed unconditionally, so their load is issued with the box loads.  (In the
		// branchy form below the compiler sinks the link load under the "some child hit" branch and every descent pays two
		// dependent L1 round trips.)
		const bool first0 = h0 && (!h1 || !(en1 < en0));
		const int first = first0 ? link0 : link1, second = first0 ? link1 : link0;
#if GS_POP_CULL
		if (h0 && h1) { stack.set(sp, second); stack.set(sp + 1, __float_as_int(first0 ? en1 : en0)); sp += 2; }
#else
		if (h0 && h1) stack.set(sp++, second);
#endif
		if (h0 || h1) node = first;
		else pop(stack, rc);
#else
		if (!h0 && !h1) { pop(stack, rc); }
		else {
			int near = link0, far = link1;
			if (h0 && h1) {
				float far_en = en1;
				if (en1 < en0) { int tmp = near; near = far; far = tmp; far_en = en0; }
#if GS_POP_CULL
				stack.set(sp, far); stack.set(sp + 1, __float_as_int(far_en)); sp += 2;
#else
				(void)far_en;
				stack.set(sp++, far);
#endif
				node = near;
			} else node = h0 ? near : far;
		}
#endif
	}

	// The leaf under `node`, then pop.  Precondition: wants_leaf().
	template <bool kTris, class Stack>
	__device__ __forceinline__ void leaf_step(const DeviceScene& sc, const Ray& r, RayConsts& rc, double tmin, const Stack& stack) {
		leaf_intersect<kTris>(sc, r, rc, tmin, best, node);
		pop(stack, rc);
	}
};

// ==========================================================================================
// Traversal policy 2: the eight-wide compressed tree (DNode8).
// State: the current NODE GROUP (`base` = child_base of the node it came from; `grp` bits 0-7 = its internal children that were
// hit and are still to visit, stored at bit (slot ^ octinv) so that the highest set bit is the nearest in octant order; bits
// 8-15 = that node's imask in slot order, needed to turn a slot into a child index), the pending LEAF hits of the node visited
// last (`grp` bits 16-23, same bit order; `leaf_node` = that node), and a stack of (base, grp & 0xffff) pairs -- one entry per
// node whose group still has unvisited members, i.e. at most one per level.
// ==========================================================================================
struct Trav8 {
	static constexpr int kStackWords = 2 * kMaxStack8;
	uint32_t base, grp, leaf_node;
	int sp;      // stack WORDS in use (two per entry)
	Hit best;

	// bit a set where direction component a is >= 0, from the SIGN BIT (1 / d keeps it, and safe_rcp copies it for d = -0)
	__device__ __forceinline__ static uint32_t octinv_bits(float x, float y, float z) {
		return ((__float_as_uint(x) >> 31) | ((__float_as_uint(y) >> 31) << 1) | ((__float_as_uint(z) >> 31) << 2)) ^ 7u;
	}
	__device__ __forceinline__ static uint32_t octinv_of(const RayConsts& rc) { return octinv_bits(rc.idx, rc.idy, rc.idz); }
	template <class Stack>
	__device__ __forceinline__ void init(const DeviceScene&, const Ray& r, const Stack&, double tmax) {
		// a group holding only the root: base 0, imask bit 0, hit bit of slot 0
		base = 0u; grp = (1u << 8) | (1u << octinv_bits((float)r.dx, (float)r.dy, (float)r.dz)); leaf_node = 0u; sp = 0;
		best.t = tmax; best.ref = kNoHit; best.u = 0.f; best.v = 0.f;
	}
	__device__ __forceinline__ void park() { grp = 0u; sp = 0; }
	__device__ __forceinline__ bool done() const { return (grp & 0x00ff00ffu) == 0u; }  // steps refill from the stack, so empty = finished
	__device__ __forceinline__ bool wants_leaf() const { return (grp & 0x00ff0000u) != 0u; }
	__device__ __forceinline__ bool wants_node() const { return (grp & 0x00ff00ffu) != 0u && (grp & 0x00ff0000u) == 0u; }
	__device__ __forceinline__ uint32_t pending_leaf_kind() const { return 0u; }

	template <class Stack>
	__device__ __forceinline__ void refill(const Stack& stack) {  // nothing left in the current group: continue with the enclosing one
		if ((grp & 0x00ff00ffu) == 0u) {
			if (sp > 0) { sp -= 2; base = (uint32_t)stack.get(sp); grp = (uint32_t)stack.get(sp + 1); }
			else grp = 0u;
		}
	}

	// 32768 + q as a float: byte q goes to mantissa bits 8..15 of 0x47000000 (one PRMT, no integer-to-float conversion)
	// The selector is written as an immediate and the constant kept in a register (`k47`, read from DeviceScene::prmt_hi so that
	// the compiler cannot fold it): PRMT has one immediate slot, and left alone ptxas gives it to the constant and spends a MOV per PRMT on
	// the selector -- 48 extra instructions per node.
	__device__ __forceinline__ static float qf(uint32_t word, int j, uint32_t k47) {
#if GS_PRMT_IMM && defined(__CUDA_ARCH__)
		uint32_t d;
		if (j == 0) asm("prmt.b32 %0, %1, %2, 0x7404;" : "=r"(d) : "r"(word), "r"(k47));
		else if (j == 1) asm("prmt.b32 %0, %1, %2, 0x7414;" : "=r"(d) : "r"(word), "r"(k47));
		else if (j == 2) asm("prmt.b32 %0, %1, %2, 0x7424;" : "=r"(d) : "r"(word), "r"(k47));
		else asm("prmt.b32 %0, %1, %2, 0x7434;" : "=r"(d) : "r"(word), "r"(k47));
		return __uint_as_float(d);
#else
		(void)k47;
		return __uint_as_float(__byte_perm(word, 0x47000000u, j == 0 ? 0x7404 : (j == 1 ? 0x7414 : (j == 2 ? 0x7424 : 0x7434))));
#endif
	}

	// Visit the nearest unvisited member of the current group: fetch it, test its eight children's boxes, and make their
	// hits the new current group (the rest of the old one goes on the stack).  Precondition: wants_node().
	// Slab distances: plane = corner + q * step, so t = q * (step / d) + (corner - o) / d; with q taken as (32768 + q) the
	// constant absorbs -32768 * step / d (its rounding error, 2^-9 of a grid step, is covered by the builder's margin).
	template <bool kCount, int kTop, class Stack>
	__device__ __forceinline__ void node_step(const DeviceScene& sc, const RayConsts& rc, const Stack& stack, uint32_t* node_visits, const float4*) {
		const uint32_t octinv = octinv_of(rc);
		const uint32_t pos = 31u - __clz(grp & 0xffu);
		grp &= ~(1u << pos);
		const uint32_t slot = pos ^ octinv;
		const uint32_t node = base + __popc((grp >> 8) & ((1u << slot) - 1u));
		if (grp & 0xffu) { stack.set(sp, (int)base); stack.set(sp + 1, (int)(grp & 0xffffu)); sp += 2; }
		const uint4* np = reinterpret_cast<const uint4*>(sc.nodes8 + node);
		uint4 hd, qx, qy, qz;
		ldg256<Ld256<true>::on>(np, hd, qx); ldg256<Ld256<true>::on>(np + 2, qy, qz);
		const uint2 tail = __ldg(reinterpret_cast<const uint2*>(np + 4) + 1);  // child_base, sphere_base: only child_base is used here
		if (kCount) ++(*node_visits);
		const float ax = __uint_as_float((hd.w & 0xffu) << 23) * rc.idx;
		const float ay = __uint_as_float((hd.w & 0xff00u) << 15) * rc.idy;
		const float az = __uint_as_float((hd.w & 0xff0000u) << 7) * rc.idz;
		const float bx = fmaf(-32768.0f, ax, fmaf(__uint_as_float(hd.x), rc.idx, -rc.oix));
		const float by = fmaf(-32768.0f, ay, fmaf(__uint_as_float(hd.y), rc.idy, -rc.oiy));
		const float bz = fmaf(-32768.0f, az, fmaf(__uint_as_float(hd.z), rc.idz, -rc.oiz));
		// near / far planes per axis by the sign of the direction: whole words (four children each)
		const bool nx = !(octinv & 1u), ny = !(octinv & 2u), nz = !(octinv & 4u);
		const uint32_t nearx[2] = {nx ? qx.z : qx.x, nx ? qx.w : qx.y}, farx[2] = {nx ? qx.x : qx.z, nx ? qx.y : qx.w};
		const uint32_t neary[2] = {ny ? qy.z : qy.x, ny ? qy.w : qy.y}, fary[2] = {ny ? qy.x : qy.z, ny ? qy.y : qy.w};
		const uint32_t nearz[2] = {nz ? qz.z : qz.x, nz ? qz.w : qz.y}, farz[2] = {nz ? qz.x : qz.z, nz ? qz.y : qz.w};
		uint32_t hits = 0u;  // slot order
		const uint32_t k47 = GS_PRMT_IMM ? sc.prmt_hi : 0x47000000u;
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const int w = k >> 2, j = k & 3;
			const float en = fmaxf(fmaxf(fmaf(qf(nearx[w], j, k47), ax, bx), fmaf(qf(neary[w], j, k47), ay, by)), fmaxf(fmaf(qf(nearz[w], j, k47), az, bz), 0.0f));
			const float ex = fminf(fminf(fmaf(qf(farx[w], j, k47), ax, bx), fmaf(qf(fary[w], j, k47), ay, by)), fminf(fmaf(qf(farz[w], j, k47), az, bz), rc.tmax32));
			if (en <= ex) hits |= 1u << k;
		}
		// internal hits to the low byte, leaf hits to the third byte, then slot k -> bit (k ^ octinv) in both at once
		const uint32_t imask = hd.w >> 24;
		uint32_t m = (hits & imask) | ((hits & ~imask) << 16);
		if (octinv & 1u) m = ((m & 0x00550055u) << 1) | ((m >> 1) & 0x00550055u);
		if (octinv & 2u) m = ((m & 0x00330033u) << 2) | ((m >> 2) & 0x00330033u);
		if (octinv & 4u) m = ((m & 0x000f000fu) << 4) | ((m >> 4) & 0x000f000fu);
		base = tail.x;
		grp = m | (imask << 8);
		leaf_node = node;
		refill(stack);
	}

	// The pending leaf children of `leaf_node`, nearest first.  Precondition: wants_leaf().
	template <bool kTris, class Stack>
	__device__ __forceinline__ void leaf_step(const DeviceScene& sc, const Ray& r, RayConsts& rc, double tmin, const Stack& stack) {
		const uint32_t octinv = octinv_of(rc);
		const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(sc.nodes8 + leaf_node) + 4);  // meta[0], meta[1], child_base, sphere_base
		uint32_t pending = (grp >> 16) & 0xffu;
		grp &= 0xffffu;
		while (pending) {
			const uint32_t pos = 31u - __clz(pending);
			pending &= ~(1u << pos);
			const uint32_t slot = pos ^ octinv;
			const uint32_t meta = ((slot & 4u) ? c4.y : c4.x) >> ((slot & 3u) * 8u) & 0xffu;
			const uint32_t kind = meta >> 6, off = meta & 63u;
			int code;
			if (kind == kKindSphere) code = leaf_code((kKindSphere << 25) | (c4.w + off), 0u);
			else if (kind == 3u) code = __ldg(sc.leaf_links8 + leaf_node * 8u + slot);
			else {
				const uint2 c5 = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const uint4*>(sc.nodes8 + leaf_node) + 5));  // quad_base, tri_base
				code = leaf_code((kind << 25) | ((kind == kKindQuad ? c5.x : c5.y) + off), 0u);
			}
			leaf_intersect<kTris, true>(sc, r, rc, tmin, best, code);
		}
		refill(stack);
	}
};

// Closest hit over the surface BVH: BVHNode::hit + AABB::hit + HittableList::hit collapsed into an
// ordered while-while traversal with a per-thread stack.  Result equals the reference's
// unordered recursion (BVH.rs:68-90); exact ties follow the reference's visiting order (Q9).
//
// The traversal is RESUMABLE: its whole state is a Trav2 / Trav8 plus the stack, so the render megakernel
// can leave the loop when enough lanes of the warp have finished their ray, shade / regenerate
// those lanes, and re-enter with the unfinished rays exactly where they stopped (ncu on the first
// version showed 8.4 of 32 lanes active in the node loop: lanes idled until the warp's longest ray ended).
//
// One round of the classic while-while traversal: descend to the next leaf, intersect it.
// Precondition: !tv.done().  kF: bit 0 = count node visits, bit 1 = scene holds triangles.
template <int kF, int kTop, class Trav, class Stack>
__device__ __forceinline__ void trav_round(const DeviceScene& sc, const Ray& r, RayConsts& rc, double tmin, Trav& tv, const Stack& stack,
                                           uint32_t* node_visits, const float4* top = nullptr) {
	while (tv.wants_node()) tv.template node_step<(kF & 1) != 0, kTop>(sc, rc, stack, node_visits, top);
	if (tv.wants_leaf()) tv.template leaf_step<(kF & 2) != 0>(sc, r, rc, tmin, stack);
}

// Run-to-completion form (parity kernel, adaptive tail kernel, per-path hooks).  kF bit 4 = the eight-wide tree.
template <int kF>
__device__ __forceinline__ Hit trace_closest(const DeviceScene& sc, const Ray& r, double tmin, double tmax, uint32_t* node_visits = nullptr) {
	RayConsts rc = ray_consts(r, tmax);
	if (kF & 16) {
		int stack_mem[Trav8::kStackWords];
		const LocalStack stack{stack_mem, nullptr};
		Trav8 tv;
		tv.init(sc, r, stack, tmax);
		while (!tv.done()) trav_round<kF, 0>(sc, r, rc, tmin, tv, stack, node_visits);
		return tv.best;
	} else {
		int stack_mem[Trav2::kStackWords];
		const LocalStack stack{stack_mem, nullptr};
		Trav2 tv;
		tv.init(sc, r, stack, tmax);
		while (!tv.done()) trav_round<kF, 0>(sc, r, rc, tmin, tv, stack, node_visits);
		return tv.best;
	}
}

}  // namespace gs
