// grayshift_b200/csrc/gs_shade.cuh -- ray generation, media, materials, textures, environment map and
// the iterative form of Camera::ray_color for the render kernels.  fp32 shading on f64 hit points.
// Every function cites the reference file:line (relative to /root/reference) whose behaviour it keeps,
// quirks included.
#pragma once
#include "gs_device.cuh"

namespace gs {

struct f3 { float x, y, z; };
__device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ f3 unit(f3 a) { float inv = rsqrtf(dot(a, a)); return a * inv; }

constexpr float kPi = 3.14159265358979323846f;

// fp32 square root and reciprocal of the shading code.  The IEEE forms (sqrtf, 1.0f / x) are a MUFU plus a Newton step plus a
// test for a slow path that lives in the kernel as a called function; the twelve call sites cost the config-4 kernel its last
// 72 B of spills.  The approximate instructions (<= 1 ulp for rcp, <= 2 ulp for sqrt) are written out as PTX so that every
// kernel variant computes the same bits (nvcc's -prec-sqrt=false / -prec-div=false left the binary-tree and the eight-wide kernels
// with slightly different paths).  Shading is fp32 on f64 hit points and compared with the oracle to 2e-3 per path.
#ifndef GS_FAST_F32
#define GS_FAST_F32 1
#endif
__device__ __forceinline__ float sqrt_f32(float x) {
#if GS_FAST_F32 && defined(__CUDA_ARCH__)
	float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
	return sqrtf(x);
#endif
}
__device__ __forceinline__ float rcp_f32(float x) {
#if GS_FAST_F32 && defined(__CUDA_ARCH__)
	float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
#else
	return 1.0f / x;
#endif
}

// ---- random helpers, util.rs:18-60 ---------------------------------------------------------------
// Two forms of the three functions that draw several numbers in a row: *_single draws them one by one, *_take takes them
// together (PathRng::take: at most one Philox block per call site).  Same values.  Measured on the A/B harness
// (profiles/r02_v14_rng_take.log): *_take +3.9 % on config 1 (small-scene kernels) and +3.0 % on config 5 (eight-wide kernels), but
// -2 % on config 4, whose 96-register kernel answers any change of shape with spills -- so kTake follows the kernel variant.
template <class Rng>
__device__ __forceinline__ f3 random_unit_vector_single(Rng& rng) {  // util.rs:18-29 (3 draws per rejection round)
	for (;;) {
		float x = rng.next_f32() * 2.0f - 1.0f;
		float y = rng.next_f32() * 2.0f - 1.0f;
		float z = rng.next_f32() * 2.0f - 1.0f;
		float l2 = x * x + y * y + z * z;
		if (l2 < 1.0f) { float inv = rsqrtf(l2); return mk3(x * inv, y * inv, z * inv); }
		rng.ensure(5);  // another round, and two media draws after it
	}
}
template <class Rng>
__device__ __forceinline__ f3 random_cosine_direction_single(Rng& rng) {  // util.rs:48-60: NOT cosine weighted (Q1): r2^(1/4)
	float r1 = rng.next_f32();
	float r2 = rng.next_f32();
	float s, c;
	sincospif(2.0f * r1, &s, &c);
	float q = sqrt_f32(sqrt_f32(r2));
	return mk3(c * q, s * q, sqrt_f32(1.0f - r2));
}

// ---- Camera::get_ray, camera.rs:204-226 (f64, same draws as the oracle) --------------------------
template <class Rng>
__device__ __forceinline__ Ray get_ray_single(const DCamera& cam, int i, int j, Rng& rng) {
	double offset_x = rng.next_f64() - 0.5;
	double offset_y = rng.next_f64() - 0.5;
	double fi = (double)i + offset_x, fj = (double)j + offset_y;
	double psx = cam.pixel00[0] + fi * cam.du[0] + fj * cam.dv[0];
	double psy = cam.pixel00[1] + fi * cam.du[1] + fj * cam.dv[1];
	double psz = cam.pixel00[2] + fi * cam.du[2] + fj * cam.dv[2];
	Ray r;
	r.ox = cam.center[0]; r.oy = cam.center[1]; r.oz = cam.center[2];
	if (cam.defocus_angle > 0.0) {  // defocus_disk_sample: rejection in the unit disk, util.rs:36-46
		double vx, vy;
		for (;;) {
			vx = rng.next_f64() * 2.0 + -1.0;
			vy = rng.next_f64() * 2.0 + -1.0;
			if (vx * vx + vy * vy < 1.0) break;
			rng.ensure(5);  // another pair, the time, and two media draws
		}
		r.ox = cam.center[0] + vx * cam.disk_u[0] + vy * cam.disk_v[0];
		r.oy = cam.center[1] + vx * cam.disk_u[1] + vy * cam.disk_v[1];
		r.oz = cam.center[2] + vx * cam.disk_u[2] + vy * cam.disk_v[2];
	}
	r.dx = psx - r.ox; r.dy = psy - r.oy; r.dz = psz - r.oz;  // not normalised (Q6)
	r.time = rng.next_f64();                                  // always drawn (Q7)
	return r;
}

template <class Rng>
__device__ __forceinline__ f3 random_unit_vector_take(Rng& rng) {  // util.rs:18-29 (3 draws per rejection round)
	for (;;) {
		uint32_t u[3];
		rng.take(u);
		float x = PathRng::to_f32(u[0]) * 2.0f - 1.0f;
		float y = PathRng::to_f32(u[1]) * 2.0f - 1.0f;
		float z = PathRng::to_f32(u[2]) * 2.0f - 1.0f;
		float l2 = x * x + y * y + z * z;
		if (l2 < 1.0f) { float inv = rsqrtf(l2); return mk3(x * inv, y * inv, z * inv); }
		rng.ensure(5);  // another round, and two media draws after it
	}
}
template <class Rng>
__device__ __forceinline__ f3 random_cosine_direction_take(Rng& rng) {  // util.rs:48-60: NOT cosine weighted (Q1): r2^(1/4)
	uint32_t u[2];
	rng.take(u);
	float r1 = PathRng::to_f32(u[0]);
	float r2 = PathRng::to_f32(u[1]);
	float s, c;
	sincospif(2.0f * r1, &s, &c);
	float q = sqrt_f32(sqrt_f32(r2));
	return mk3(c * q, s * q, sqrt_f32(1.0f - r2));
}

// ---- Camera::get_ray, camera.rs:204-226 (f64, same draws as the oracle) --------------------------
template <class Rng>
__device__ __forceinline__ Ray get_ray_take(const DCamera& cam, int i, int j, Rng& rng) {
	// draws: ox, oy, [disk pairs,] time.  Without a lens the three come from one block in one go.
	const bool lens = cam.defocus_angle > 0.0;
	uint32_t u[3];
	if (lens) { uint32_t u2[2]; rng.take(u2); u[0] = u2[0]; u[1] = u2[1]; u[2] = 0u; }
	else rng.take(u);
	double offset_x = PathRng::to_f64(u[0]) - 0.5;
	double offset_y = PathRng::to_f64(u[1]) - 0.5;
	double fi = (double)i + offset_x, fj = (double)j + offset_y;
	double psx = cam.pixel00[0] + fi * cam.du[0] + fj * cam.dv[0];
	double psy = cam.pixel00[1] + fi * cam.du[1] + fj * cam.dv[1];
	double psz = cam.pixel00[2] + fi * cam.du[2] + fj * cam.dv[2];
	Ray r;
	r.ox = cam.center[0]; r.oy = cam.center[1]; r.oz = cam.center[2];
	if (lens) {  // defocus_disk_sample: rejection in the unit disk, util.rs:36-46
		double vx, vy;
		for (;;) {
			uint32_t d2[2];
			rng.take(d2);
			vx = PathRng::to_f64(d2[0]) * 2.0 + -1.0;
			vy = PathRng::to_f64(d2[1]) * 2.0 + -1.0;
			if (vx * vx + vy * vy < 1.0) break;
			rng.ensure(5);  // another pair, the time, and two media draws
		}
		r.ox = cam.center[0] + vx * cam.disk_u[0] + vy * cam.disk_v[0];
		r.oy = cam.center[1] + vx * cam.disk_u[1] + vy * cam.disk_v[1];
		r.oz = cam.center[2] + vx * cam.disk_u[2] + vy * cam.disk_v[2];
	}
	r.dx = psx - r.ox; r.dy = psy - r.oy; r.dz = psz - r.oz;  // not normalised (Q6)
	r.time = lens ? rng.next_f64() : PathRng::to_f64(u[2]);  // always drawn (Q7)
	return r;
}

template <bool kTake, class Rng> __device__ __forceinline__ f3 random_unit_vector(Rng& rng) { return kTake ? random_unit_vector_take(rng) : random_unit_vector_single(rng); }
template <bool kTake, class Rng> __device__ __forceinline__ f3 random_cosine_direction(Rng& rng) { return kTake ? random_cosine_direction_take(rng) : random_cosine_direction_single(rng); }
template <bool kTake, class Rng> __device__ __forceinline__ Ray get_ray(const DCamera& cam, int i, int j, Rng& rng) { return kTake ? get_ray_take(cam, i, j, rng) : get_ray_single(cam, i, j, rng); }

// ---- HDRI::sample, camera.rs:257-270 ----------------------------------------------------------------
__device__ __forceinline__ f3 background(const DeviceScene& sc, const DCamera& cam, f3 d, uint32_t& tex_fetches) {
	if (cam.bg_kind == 0) return mk3(cam.bg[0], cam.bg[1], cam.bg[2]);
	f3 r = mk3(d.x * cam.rot[0] + d.y * cam.rot[1] + d.z * cam.rot[2],
	           d.x * cam.rot[3] + d.y * cam.rot[4] + d.z * cam.rot[5],
	           d.x * cam.rot[6] + d.y * cam.rot[7] + d.z * cam.rot[8]);
	r = unit(r);
	float theta = atan2f(r.y, r.x);
	float phi = asinf(fminf(fmaxf(r.z, -1.0f), 1.0f));
	float u = 0.5f + theta * (0.5f / kPi);
	float v = 0.5f - phi * (1.0f / kPi);
	const DImage im = sc.images[cam.hdri_image];
	int x = (int)fmaxf(u * (float)im.w, 0.0f) % im.w;  // `as usize` saturates negatives to 0, then % W (camera.rs:265-266)
	int y = (int)fmaxf(v * (float)im.h, 0.0f) % im.h;
	float4 c = __ldg(reinterpret_cast<const float4*>(im.texels) + (size_t)y * im.w + x);
	++tex_fetches;
	return mk3(c.x, c.y, c.z);
}

// ---- Perlin (noise 0.9.0, restated; see oracle.cpp) on a shared-memory permutation table ------------
// The 12 edge gradients selected by hash & 15, cases 12..15 repeating cases 0, 1, 9, 11 (gradient_dot_v of noise 0.9's
// core::perlin::perlin_3d: 0|12, 1|13, 9|14, 11|15 -- the order of the "Improving Noise" paper, not the 13 <-> 14 swap
// of Perlin's Java bit trick; the oracle restates the same table).  Branch-free: the 16-way switch was inlined 8 times
// per octave and made up a fifth of the kernel's code.  The nibble table folds 12..15 onto 0, 1, 9, 11; for h < 12
// the gradient is (+-u) + (+-v) with u = x or y, v = y or z and the signs in bits 0 and 1.
// kLut: the gradient as a vector from the CTA's table (three LDS + three FMAs; load_perm fills it) instead of the selects.
// Same value bit for bit (one coefficient is 0, the others +-1, so the sum is (+-u) + (+-v) rounded once).  Measured
// (profiles/r02_v14_perlin_lut.log): +20 % on the reference's Perlin scene 3 and +25 % on scene 5, but -1.7 % on config 4, whose
// 96-register kernel pays for the loads' registers with spills; so only the small-scene kernel variants (kF & 8) use it.
template <bool kLut>
__device__ __forceinline__ float perlin_grad(const uint8_t* __restrict__ perm, uint32_t h, float x, float y, float z) {
	if (kLut) {
		const float* g = reinterpret_cast<const float*>(perm + 256) + (h & 15u);
		return fmaf(g[0], x, fmaf(g[16], y, g[32] * z));
	}
	h = (uint32_t)(0xB910BA9876543210ull >> ((h & 15u) * 4u)) & 15u;
	const float u = h < 8u ? x : y;
	const float v = h < 4u ? y : z;
	return ((h & 1u) ? -u : u) + ((h & 2u) ? -v : v);
}
template <bool kLut>
__device__ __forceinline__ float perlin_get(const uint8_t* __restrict__ perm, double px, double py, double pz) {
	double fx = floor(px), fy = floor(py), fz = floor(pz);
	int cx = (int)(long long)fx, cy = (int)(long long)fy, cz = (int)(long long)fz;
	float dx = (float)(px - fx), dy = (float)(py - fy), dz = (float)(pz - fz);
	uint32_t hx0 = perm[cx & 255], hx1 = perm[(cx + 1) & 255];
	uint32_t y0 = (uint32_t)(cy & 255), y1 = (uint32_t)((cy + 1) & 255), z0 = (uint32_t)(cz & 255), z1 = (uint32_t)((cz + 1) & 255);
	uint32_t h00 = perm[hx0 ^ y0], h10 = perm[hx1 ^ y0], h01 = perm[hx0 ^ y1], h11 = perm[hx1 ^ y1];
	float g000 = perlin_grad<kLut>(perm, perm[h00 ^ z0], dx, dy, dz), g100 = perlin_grad<kLut>(perm, perm[h10 ^ z0], dx - 1.f, dy, dz);
	float g010 = perlin_grad<kLut>(perm, perm[h01 ^ z0], dx, dy - 1.f, dz), g110 = perlin_grad<kLut>(perm, perm[h11 ^ z0], dx - 1.f, dy - 1.f, dz);
	float g001 = perlin_grad<kLut>(perm, perm[h00 ^ z1], dx, dy, dz - 1.f), g101 = perlin_grad<kLut>(perm, perm[h10 ^ z1], dx - 1.f, dy, dz - 1.f);
	float g011 = perlin_grad<kLut>(perm, perm[h01 ^ z1], dx, dy - 1.f, dz - 1.f), g111 = perlin_grad<kLut>(perm, perm[h11 ^ z1], dx - 1.f, dy - 1.f, dz - 1.f);
	float a = dx * dx * dx * (dx * (dx * 6.f - 15.f) + 10.f);
	float b = dy * dy * dy * (dy * (dy * 6.f - 15.f) + 10.f);
	float c = dz * dz * dz * (dz * (dz * 6.f - 15.f) + 10.f);
	float k1 = g100 - g000, k2 = g010 - g000, k3 = g001 - g000;
	float k4 = g000 + g110 - g100 - g010, k5 = g000 + g101 - g100 - g001, k6 = g000 + g011 - g010 - g001;
	float k7 = g100 + g010 + g001 + g111 - g000 - g110 - g101 - g011;
	float v = g000 + k1 * a + k2 * b + k3 * c + k4 * a * b + k5 * a * c + k6 * b * c + k7 * a * b * c;
	v *= 1.1547005383792515f;
	return fminf(fmaxf(v, -1.0f), 1.0f);
}

// turb(p, 7) without the final abs (texture.rs:116-130 keeps it for the caller), by one lane
template <bool kLut>
__device__ __forceinline__ float noise_turb(const uint8_t* __restrict__ perm, double sx, double sy, double sz) {
	float accum = 0.0f, weight = 1.0f;
#pragma unroll 1
	for (int i = 0; i < 7; ++i) {
		accum += weight * perlin_get<kLut>(perm, sx, sy, sz);
		weight *= 0.5f;
		sx *= 2.0; sy *= 2.0; sz *= 2.0;
	}
	return accum;
}

// ---- turb(p, 7) of ONE point by the whole warp ------------------------------------------------------------------
// The megakernel meets a NoiseTexture on ~2.5 of a warp's 32 lanes at a time (config 4: one marble sphere among 3400
// primitives), and the serial turbulence -- 7 octaves x 8 lattice corners, ~1400 instructions -- was 8 % of its warp
// instructions at those 2.5 lanes (profiles/r02_v13_c4_render_fixed_hotspots.txt).  Here the lane that holds the point
// (`src`) broadcasts it and lane (octave << 2 | corner pair) evaluates the two corners (jx, jy, 0) and (jx, jy, 1) of its
// octave, weighted with the trilinear factors of the fade curves: the octave's value is the sum over its four lanes (the
// same interpolant as perlin_get's polynomial form, summed in another order), then scale + clamp per octave exactly
// like perlin_get, then the sum over the octaves with the weights 2^-octave.  Lanes 28..31 idle.  All 32 lanes must call it.
// Returns turb's `accum` on every lane.
template <bool kLut>
__device__ __forceinline__ float noise_turb_coop(const uint8_t* __restrict__ perm, unsigned src, double px, double py, double pz, unsigned lane) {
	const double bx = __shfl_sync(0xffffffffu, px, src), by = __shfl_sync(0xffffffffu, py, src), bz = __shfl_sync(0xffffffffu, pz, src);
	const unsigned o = lane >> 2, jx = lane & 1u, jy = (lane >> 1) & 1u;
	const double mul = (double)(1u << o);  // the serial loop doubles the point per octave: exact either way
	const double qx = bx * mul, qy = by * mul, qz = bz * mul;
	const double fx = floor(qx), fy = floor(qy), fz = floor(qz);
	const int cx = (int)(long long)fx, cy = (int)(long long)fy, cz = (int)(long long)fz;
	const float dx = (float)(qx - fx), dy = (float)(qy - fy), dz = (float)(qz - fz);
	const uint32_t hx = perm[(cx + (int)jx) & 255];
	const uint32_t hxy = perm[hx ^ (uint32_t)((cy + (int)jy) & 255)];
	const uint32_t h0 = perm[hxy ^ (uint32_t)(cz & 255)], h1 = perm[hxy ^ (uint32_t)((cz + 1) & 255)];
	const float ex = dx - (float)jx, ey = dy - (float)jy;
	const float g0 = perlin_grad<kLut>(perm, h0, ex, ey, dz), g1 = perlin_grad<kLut>(perm, h1, ex, ey, dz - 1.f);
	const float a = dx * dx * dx * (dx * (dx * 6.f - 15.f) + 10.f);
	const float b = dy * dy * dy * (dy * (dy * 6.f - 15.f) + 10.f);
	const float c = dz * dz * dz * (dz * (dz * 6.f - 15.f) + 10.f);
	const float wx = jx ? a : 1.f - a, wy = jy ? b : 1.f - b;
	float v = wx * wy * fmaf(c, g1 - g0, g0);
	v += __shfl_xor_sync(0xffffffffu, v, 1);
	v += __shfl_xor_sync(0xffffffffu, v, 2);
	v *= 1.1547005383792515f;
	v = fminf(fmaxf(v, -1.0f), 1.0f);
	// (after the two exchanges the four lanes of an octave hold the same v, so each residue class lane & 3 sums the octaves on its own
	// and every lane ends with the total)
	float t = o < 7u ? v * __uint_as_float((127u - o) << 23) : 0.0f;  // weight 2^-octave
	t += __shfl_xor_sync(0xffffffffu, t, 4);
	t += __shfl_xor_sync(0xffffffffu, t, 8);
	t += __shfl_xor_sync(0xffffffffu, t, 16);
	return t;
}
// NoiseTexture's value from turb's sum: 0.5 * (1 + sin(scale * z + 10 * |turb|)) (texture.rs:107-130), shared by the serial and
// the cooperative form.
__device__ __forceinline__ float noise_value(double scale, double pz, float accum) {
	// sin of a phase that can reach a few hundred radians: reduce to [-pi, pi] in f64 (cheap, exact enough),
	// then the SFU sine (abs. error ~1e-6 there); libm's sinf drags in its 280-instruction slow path
	const double phase = scale * pz + (double)(10.0f * fabsf(accum));
	const float red = (float)(phase - 6.283185307179586 * rint(phase * 0.15915494309189535));
	return 0.5f * (1.0f + __sinf(red));
}

// ---- Texture::value_at, texture.rs:27-30, 57-70, 84-94, 107-130 --------------------------------------
struct Surface {
	double px, py, pz;  // hit position (f64: it is the next ray's origin)
	f3 n;               // shading normal, flipped against the ray (hittable.rs:35-36)
	f3 n_out;           // outward (unflipped) normal; spheres derive (u,v) from it
	float u, v;         // triangles only; see quad_index
	bool front, is_sphere;
	int material;
	int quad_index;     // >= 0: a quad, whose (u, v) = (alpha, beta) are computed from the hit point when a texture needs them
	int sphere_index;   // is_sphere: index into sph_uvrot (an instance's RotateY, needed by sphere_uv only)
};

// kDefer: a NoiseTexture is not evaluated here; its index is handed back in `pend` with the value (1, 1, 1), and the caller
// evaluates it with the whole warp at its next converged point (noise_turb_coop) and multiplies the result in.
template <bool kLut, bool kDefer = false>
__device__ __forceinline__ f3 texture_value(const DeviceScene& sc, const uint8_t* __restrict__ perm, int tex, const Surface& s, uint32_t& tex_fetches,
                                            int* pend = nullptr) {
#pragma unroll 1
	for (int depth = 0; depth < 16; ++depth) {
		const DTexture* tp = sc.textures + tex;
		const int4 head = __ldg(reinterpret_cast<const int4*>(tp));  // kind, a, b, scale(float bits)
		if (head.x == 0) {  // SolidColorTexture
			const float4 col = __ldg(reinterpret_cast<const float4*>(tp) + 1);
			return mk3(col.x, col.y, col.z);
		}
		if (head.x == 1) {  // CheckeredTexture: floor(scale_inv * p) as i32, (x+y+z) % 2 == 0 (texture.rs:58-63)
			const double2 sd = __ldg(reinterpret_cast<const double2*>(tp) + 2);  // scale_inv, scale
			int xi = __double2int_rd(sd.x * s.px), yi = __double2int_rd(sd.x * s.py), zi = __double2int_rd(sd.x * s.pz);
			int sum = xi + yi + zi;
			tex = (sum % 2 == 0) ? head.y : head.z;
			continue;
		}
		if (head.x == 2) {  // ImageTexture: nearest texel, v flipped, /255 (texture.rs:84-94); Q13 clamps the index
			float u = s.u, v = s.v;
			if (s.quad_index >= 0) {  // alpha = w.(planar x v), beta = w.(u x planar), quad.rs:88-89, as P.A - qa and P.B - qb
				const double2* qp = reinterpret_cast<const double2*>(sc.quads + s.quad_index);
				const double2 a01 = __ldg(qp + 2), a23 = __ldg(qp + 3), b01 = __ldg(qp + 4), b23 = __ldg(qp + 5);
				u = (float)((s.px * a01.x + s.py * a01.y + s.pz * a23.x) - a23.y);
				v = (float)((s.px * b01.x + s.py * b01.y + s.pz * b23.x) - b23.y);
			}
			if (s.is_sphere) {  // sphere_uv, sphere.rs:55-60, on the OBJECT-space outward normal: Sphere::hit computes (u, v) before
				// RotateY::hit turns the normal into world space (hittable.rs:179-211), so the baked rotation is undone here
				float nx = s.n_out.x, nz = s.n_out.z;
				if (sc.sph_uvrot) {  // null when no sphere of the scene sits under a RotateY
					const float2 cs = __ldg(sc.sph_uvrot + s.sphere_index);
					nx = cs.x * s.n_out.x - cs.y * s.n_out.z; nz = cs.y * s.n_out.x + cs.x * s.n_out.z;
				}
				float theta = acosf(fminf(fmaxf(-s.n_out.y, -1.0f), 1.0f));
				float phi = atan2f(-nz, nx) + kPi;
				u = phi * (0.5f / kPi); v = theta * (1.0f / kPi);
			}
			const DImage im = sc.images[head.y];
			float uc = fminf(fmaxf(u, 0.0f), 1.0f), vc = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
			int i = min((int)(uc * (float)im.w), im.w - 1), j = min((int)(vc * (float)im.h), im.h - 1);
			uchar4 px = __ldg(reinterpret_cast<const uchar4*>(im.texels) + (size_t)j * im.w + i);
			++tex_fetches;
			return mk3((float)px.x * (1.0f / 255.0f), (float)px.y * (1.0f / 255.0f), (float)px.z * (1.0f / 255.0f));
		}
		if (kDefer) { *pend = tex; return mk3(1.f, 1.f, 1.f); }
		{  // NoiseTexture: 0.5 * (1 + sin(scale*z + 10*turb(p, 7))) (texture.rs:107-130)
			const double2 sd = __ldg(reinterpret_cast<const double2*>(tp) + 2);
			const float g = noise_value(sd.y, s.pz, noise_turb<kLut>(perm, s.px, s.py, s.pz));
			return mk3(g, g, g);
		}
	}
	return mk3(0.f, 0.f, 0.f);
}

// ---- Material::emitted / scatter, material.rs:45-68, 87-102, 122-148, 165-169, 186-196 ------------
// returns false when the path ends here (no scatter); emission is added to `emit`.
// kTune bit 2: NoiseTextures are deferred (texture_value's kDefer).  `pend` = -1, or (texture index << 1 | 1 when the value is
// an emission, 0 when it is an attenuation); the deferred factor enters as 1 (attenuation) or 0 (emission) here.
template <int kTune, class Rng>
__device__ __forceinline__ bool shade(const DeviceScene& sc, const uint8_t* __restrict__ perm, const Surface& s, f3 d_in, Rng& rng,
                                      f3& emit, f3& atten, f3& d_out, uint32_t& tex_fetches, int& pend) {
	const DMaterial* mp = sc.materials + s.material;
	const int4 m0 = __ldg(reinterpret_cast<const int4*>(mp));        // kind, tex, albedo.r, albedo.g
	const float4 m1 = __ldg(reinterpret_cast<const float4*>(mp) + 1);  // albedo.b, param
	emit = mk3(0.f, 0.f, 0.f);
	// One copy of the texture evaluation (the Perlin turbulence is ~500 instructions) and of the unit-vector rejection loop
	// for all materials that use them: the megakernel is instruction-cache sensitive.  Textures draw no random numbers,
	// so the draw order of every material is unchanged.
	const int kind = m0.x;
	f3 tex = mk3(0.f, 0.f, 0.f), ruv = mk3(0.f, 0.f, 0.f);
	constexpr bool kDefer = (kTune & 4) != 0;
	int noise = -1;
	if (kind == 0 || kind >= 3) tex = texture_value<(kTune & 1) != 0, kDefer>(sc, perm, m0.y, s, tex_fetches, &noise);
	pend = -1;
	if (kDefer && noise >= 0) {
		pend = (noise << 1) | (kind == 3 ? 1 : 0);
		if (kind == 3) tex = mk3(0.f, 0.f, 0.f);
	}
	if (kind == 1 || kind >= 4) ruv = random_unit_vector<(kTune & 2) != 0>(rng);
	switch (kind) {
		case 0: {  // Lambertian
			atten = tex;
			// OrthonormalBasis::new, ONB.rs:10-23
			f3 w = unit(s.n);
			f3 a = fabsf(w.x) > 0.9f ? mk3(0.f, 1.f, 0.f) : mk3(1.f, 0.f, 0.f);
			f3 v = unit(cross(w, a));
			f3 u = cross(w, v);
			f3 l = random_cosine_direction<(kTune & 2) != 0>(rng);
			d_out = unit(u * l.x + v * l.y + w * l.z);
			return true;
		}
		case 1: {  // Metal: unit(reflect) + fuzz * random_unit_vector, drawn even when fuzz == 0 (Q7)
			f3 refl = d_in - s.n * (2.0f * dot(d_in, s.n));
			d_out = unit(refl) + ruv * m1.y;
			atten = mk3(__int_as_float(m0.z), __int_as_float(m0.w), m1.x);
			return dot(d_out, s.n) > 0.0f;
		}
		case 2: {  // Dielectric
			float ir = m1.y;
			float ri = s.front ? rcp_f32(ir) : ir;
			f3 ud = unit(d_in);
			float cos_theta = fminf(-dot(ud, s.n), 1.0f);
			float sin_theta = sqrt_f32(fmaxf(1.0f - cos_theta * cos_theta, 0.0f));
			bool cannot_refract = ri * sin_theta > 1.0f;
			float r0 = (1.0f - ri) * rcp_f32(1.0f + ri);
			r0 = r0 * r0;
			float m = 1.0f - cos_theta;
			float reflectance = r0 + (1.0f - r0) * (m * m * m * m * m);
			bool fresnel = reflectance > rng.next_f32();  // always drawn (Q7)
			if (cannot_refract || fresnel) d_out = ud - s.n * (2.0f * dot(ud, s.n));
			else {  // Vec3::refract, vec3.rs:57-62
				float ct = fminf(dot(s.n, -ud), 1.0f);
				f3 perp = (ud + s.n * ct) * ri;
				f3 par = s.n * (-sqrt_f32(fabsf(1.0f - dot(perp, perp))));
				d_out = perp + par;
			}
			atten = mk3(1.f, 1.f, 1.f);
			return true;
		}
		case 3: {  // DiffuseLight: two-sided, never scatters (Q8)
			emit = tex;
			return false;
		}
		default: {  // Isotropic
			atten = tex;
			d_out = ruv;
			return true;
		}
	}
}

// ---- ConstantMedium::hit, volume.rs:32-63, evaluated per ray before the surface traversal ---------
// Statistically identical to the reference's in-traversal evaluation: the free-flight draw is
// independent of everything else, so testing the medium first and clipping the surface search at
// its candidate t gives the same closest-event distribution.
#ifndef GS_BOUNDARY_INLINE
#define GS_BOUNDARY_INLINE 1  // tuning knob: 0 = out-of-line (smaller code, but the call spills ~600 B of registers)
#endif
#if GS_BOUNDARY_INLINE
__device__ __forceinline__
#else
static __device__ __noinline__
#endif
bool boundary_closest(const DeviceScene& sc, const DMedium& m, const Ray& r, double a, double inv_a, double lo, double hi, double& t_out) {
	bool found = false;
	double best = hi;
	for (int k = 0; k < m.sph_count; ++k) {
		DSphere s = load_sphere(sc.bnd_spheres, m.sph_first + k);
		double t;
		const bool moving = sc.bnd_sph_path[m.sph_first + k].w != 0.0;  // .w flags a moving boundary sphere
		if (hit_sphere(s, sc.bnd_sph_path, m.sph_first + k, moving, r, a, inv_a, lo, best, t)) { best = t; found = true; }
	}
	for (int k = 0; k < m.quad_count; ++k) {
		DQuad q = load_quad(sc.bnd_quads, m.quad_first + k);
		double t, al, be;
		if (hit_quad(q, r, lo, best, t, al, be)) { best = t; found = true; }
	}
	t_out = best;
	return found;
}

__device__ __forceinline__ bool medium_event(const DeviceScene& sc, const DMedium& m, const Ray& r, double a, double inv_a, double tmin,
                                             double tmax, float u01, double& t_out) {
	const double kMax = 1.7976931348623157e308;
	double t1, t2;
	if (m.sph_count == 1 && m.quad_count == 0) {
		// One stationary-or-moving sphere as the boundary (main.rs:714-738): both calls of volume.rs:33-37 are the two
		// roots of one quadratic.  hit(UNIVERSE) takes the near root (any finite value is inside (-MAX, MAX));
		// hit([t1 + 1e-4, MAX)) re-tests the near root (fails) and takes the far one if it lies beyond t1 + 1e-4.
		DSphere s = load_sphere(sc.bnd_spheres, m.sph_first);
		double cx = s.cx, cy = s.cy, cz = s.cz;
		const double rad = s.r;
		if (m.flags & 1) { double4 p = sc.bnd_sph_path[m.sph_first]; cx += r.time * p.x; cy += r.time * p.y; cz += r.time * p.z; }
		double ocx = cx - r.ox, ocy = cy - r.oy, ocz = cz - r.oz;
		double h = r.dx * ocx + r.dy * ocy + r.dz * ocz;
		double c = (ocx * ocx + ocy * ocy + ocz * ocz) - rad * rad;
		double disc = h * h - a * c;
		if (disc < 0.0) return false;
		double sq = f64_sqrt(disc);
		t1 = (h - sq) * inv_a;
		t2 = (h + sq) * inv_a;
		if (!(t1 > -kMax && t1 < kMax)) return false;
		if (!(t1 + 0.0001 < t2 && t2 < kMax)) return false;
	} else {
		// boundary.hit(UNIVERSE), then boundary.hit([t1 + 1e-4, MAX)), volume.rs:33-37 (one copy of the code, two passes)
		t1 = -kMax;
#pragma unroll 1
		for (int pass = 0; pass < 2; ++pass) {
			double t;
			if (!boundary_closest(sc, m, r, a, inv_a, pass == 0 ? -kMax : t1 + 0.0001, kMax, t)) return false;
			if (pass == 0) t1 = t; else t2 = t;
		}
	}
	if (t1 < tmin) t1 = tmin;
	if (t2 > tmax) t2 = tmax;
	if (t1 >= t2) return false;
	if (t1 < 0.0) t1 = 0.0;
	// volume.rs:46-52 in units of t: hit_dist / ray_len compared with (t2 - t1).  The free-flight length is a random
	// draw, so it is formed in f32 (24-bit uniform, logf) and only then promoted.
	const float inv_len = rsqrtf((float)a);
	const double dt = (double)((float)m.neg_inv_density * logf(u01) * inv_len);
	if (dt > t2 - t1) return false;
	t_out = t1 + dt;
	return true;
}

// ---- Camera::ray_color, camera.rs:174-202, iterative: L = sum_k (prod_{j<k} att_j) E_k (+ background) ---
struct PathCounters { uint32_t rays, tex_fetches, node_visits; };

template <class Rng>
struct PathStateT {
	Ray ray;
	f3 L, T;
	Rng rng;
	unsigned bounce;
	__device__ __forceinline__ uint32_t pixel() const { return rng.pixel; }
};
using PathState = PathStateT<PathRng>;

template <int kTune = 0, class PS>
__device__ __forceinline__ void path_begin(PS& st, const DCamera& cam, uint64_t seed, uint32_t pixel, uint32_t sample) {
	st.rng.start(seed, pixel, sample);
	st.ray = get_ray<(kTune & 2) != 0>(cam, (int)(pixel % (uint32_t)cam.width), (int)(pixel / (uint32_t)cam.width), st.rng);
	st.L = mk3(0.f, 0.f, 0.f);
	st.T = mk3(1.f, 1.f, 1.f);
	st.bounce = 0;
}

// Start of one level of the reference's recursion: the depth test (camera.rs:175) and the media.  Returns false
// when the path is finished (depth exhausted).  tmax = the closest medium event (or f64::MAX), medium_hit its index.
template <class PS>
__device__ __forceinline__ bool segment_begin(const DeviceScene& sc, const DCamera& cam, PS& st, PathCounters& pc, double& tmax,
                                              int& medium_hit) {
	if (st.bounce >= cam.max_depth) return false;  // `if depth <= 0 { return ZERO }`
	++pc.rays;
	tmax = 1.7976931348623157e308;  // Interval::new(0.001, f64::MAX), camera.rs:177
	medium_hit = -1;
	if (sc.n_media > 0) {
		const Ray& ray = st.ray;
		const double a = ray.dx * ray.dx + ray.dy * ray.dy + ray.dz * ray.dz;
		const double inv_a = f64_rcp(a);
		// One draw per medium from the path's own Philox stream, in medium order, between the scatter draws of the previous
		// level and those of this one (the reference draws inside ConstantMedium::hit, volume.rs:48, wherever its traversal
		// happens to be; a free-flight draw is independent of everything else).  A separate stream for the media cost one
		// more Philox block per level: a third of all Philox work on C4.
		for (int mi = 0; mi < sc.n_media; ++mi) {
			if (mi >= 2) st.rng.ensure(1);  // (PathRngW: two media draws are covered by the level before)
			const float u01 = st.rng.next_f32();
			double tm;
			if (medium_event(sc, sc.media[mi], ray, a, inv_a, 0.001, tmax, u01, tm)) { tmax = tm; medium_hit = mi; }
		}
	}
	st.rng.ensure(5);  // PathRngW: the blocks for this level's scatter and the next level's media, computed here for all lanes at once
	return true;
}

// The rest of the level: miss -> background; hit -> emitted + scatter.  `h` is the closest surface hit found with
// tmax = the medium candidate.  Returns true when the path is finished; st.L then holds the sample colour.
// kTune bit 2 (the megakernel): a NoiseTexture met here is left to the caller -- `*pend` as shade() describes it, the point
// in st.ray's origin; see noise_apply_deferred.
template <bool kTris, int kTune = 0, class PS>
__device__ __forceinline__ bool segment_shade(const DeviceScene& sc, const DCamera& cam, const uint8_t* __restrict__ perm, PS& st,
                                              PathCounters& pc, const Hit& h, double tmax, int medium_hit, int* pend = nullptr) {
	Ray& ray = st.ray;
	f3 d_in = mk3((float)ray.dx, (float)ray.dy, (float)ray.dz);
	if (h.ref == kNoHit && medium_hit < 0) {  // sample_background, camera.rs:228-233
		f3 bg = background(sc, cam, d_in, pc.tex_fetches);
		st.L = st.L + st.T * bg;
		return true;
	}
	Surface s;
	f3 n_out;
	if (h.ref != kNoHit) {
		const uint32_t index = h.ref & kRefIndexMask;
		s.px = ray.ox + ray.dx * h.t; s.py = ray.oy + ray.dy * h.t; s.pz = ray.oz + ray.dz * h.t;  // ray.at(t)
		s.sphere_index = 0;
		if (kTris && (h.ref >> kRefKindShift) == kKindTriangle) {
			// Triangle: normal = (b - a) x (c - a), handed on UNNORMALISED exactly like triangle.rs:21,63 (Lambertian
			// normalises it in its ONB; Metal and Dielectric reflect about it as it is)
			const DTriangle tr = load_triangle(sc.tris, index);
			n_out = mk3((float)(tr.e2y * tr.e1z - tr.e2z * tr.e1y), (float)(tr.e2z * tr.e1x - tr.e2x * tr.e1z), (float)(tr.e2x * tr.e1y - tr.e2y * tr.e1x));
			s.is_sphere = false; s.u = h.u; s.v = h.v; s.quad_index = -1;
			s.material = __ldg(sc.tri_aux + index).x;
		} else if ((h.ref >> kRefKindShift) != kKindQuad) {  // sphere: stationary (kind 0) or moving (kind 3)
			DSphere sp = load_sphere(sc.spheres, index);
			double cx = sp.cx, cy = sp.cy, cz = sp.cz;
			if ((h.ref >> kRefKindShift) == kKindMovingSphere) { double4 p = sc.sph_path[index]; cx += ray.time * p.x; cy += ray.time * p.y; cz += ray.time * p.z; }
			// outward_normal = (p - c) / r with the SIGNED radius (sphere.rs:91: a negative radius turns the normal inward, the
			// hollow-glass trick); difference in f64, quotient in f32
			const float inv_r = rcp_f32((float)sp.r);
			n_out = mk3((float)(s.px - cx) * inv_r, (float)(s.py - cy) * inv_r, (float)(s.pz - cz) * inv_r);
			s.is_sphere = true; s.u = 0.f; s.v = 0.f; s.quad_index = -1; s.sphere_index = (int)index;
			s.material = __ldg(sc.sph_aux + index).x;
		} else {
			const double2* qp = reinterpret_cast<const double2*>(sc.quads + index);
			double2 n01 = __ldg(qp), n23 = __ldg(qp + 1);
			n_out = mk3((float)n01.x, (float)n01.y, (float)n23.x);
			s.is_sphere = false; s.u = 0.f; s.v = 0.f; s.quad_index = (int)index;
			s.material = __ldg(sc.quad_aux + index).x;
		}
	} else {  // medium event: normal (1,0,0), u = v = 0, phase function as the material (volume.rs:54-62, Q10)
		s.px = ray.ox + ray.dx * tmax; s.py = ray.oy + ray.dy * tmax; s.pz = ray.oz + ray.dz * tmax;
		n_out = mk3(1.f, 0.f, 0.f);
		s.is_sphere = false; s.u = 0.f; s.v = 0.f; s.quad_index = -1; s.sphere_index = 0;
		s.material = sc.media[medium_hit].phase_mat;
	}
	s.front = dot(d_in, n_out) < 0.0f;  // HitRecord::new, hittable.rs:35-36
	s.n = s.front ? n_out : -n_out;
	s.n_out = n_out;
	f3 emit, atten, d_out;
	int noise;
	bool scattered = shade<kTune>(sc, perm, s, d_in, st.rng, emit, atten, d_out, pc.tex_fetches, noise);
	if (kTune & 4) *pend = noise;
	st.L = st.L + st.T * emit;
	if (!scattered) {
		if ((kTune & 4) && noise >= 0) { ray.ox = s.px; ray.oy = s.py; ray.oz = s.pz; }  // the deferred evaluation reads the point here
		return true;
	}
	st.T = st.T * atten;
	ray.ox = s.px; ray.oy = s.py; ray.oz = s.pz;
	ray.dx = (double)d_out.x; ray.dy = (double)d_out.y; ray.dz = (double)d_out.z;
	++st.bounce;
	return false;
}

#ifndef GS_NOISE_COOP_MAX
#define GS_NOISE_COOP_MAX 8  // pending lanes up to which the warp serves them together, one after the other (~100 instructions each,
                             // against ~1400 -- ~700 with the gradient table -- for any number of lanes evaluating on their own)
#endif
// The deferred NoiseTextures of a warp (segment_shade with kTune bit 2), at a converged point: every lane passes its `pend`
// (-1 = nothing) and its path state; the pending lanes are served one after the other by all 32 lanes, and the value is
// multiplied in where the serial code would have used it: T *= g (Lambertian / Isotropic albedo) or L += T * g (DiffuseLight).
template <bool kLut, class PS>
__device__ __forceinline__ void noise_apply_deferred(const DeviceScene& sc, const uint8_t* __restrict__ perm, PS& st, int pend, unsigned lane) {
	unsigned todo = __ballot_sync(0xffffffffu, pend >= 0);
	if (todo == 0u) return;
	float accum = 0.0f;
	if (GS_NOISE_COOP_MAX < 32 && __popc(todo) > GS_NOISE_COOP_MAX) {  // a warp full of marble (the reference's Perlin scenes): each lane for itself is cheaper
		if (pend >= 0) accum = noise_turb<kLut>(perm, st.ray.ox, st.ray.oy, st.ray.oz);
	} else {
		do {
			const unsigned src = __ffs(todo) - 1u;
			todo &= todo - 1u;
			const float a = noise_turb_coop<kLut>(perm, src, st.ray.ox, st.ray.oy, st.ray.oz, lane);
			if (lane == src) accum = a;
		} while (todo);
	}
	if (pend >= 0) {
		const double2 sd = __ldg(reinterpret_cast<const double2*>(sc.textures + (pend >> 1)) + 2);
		const float g = noise_value(sd.y, st.ray.oz, accum);
		if (pend & 1) st.L = st.L + st.T * mk3(g, g, g);
		else st.T = st.T * mk3(g, g, g);
	}
}

// One whole level, run to completion (adaptive kernel and the per-path hooks).
// kernel variant -> tuning of the shading code: bit 0 Perlin gradients from the CTA's table, bit 1 draws taken together.
// Small-scene kernels (kF & 8) both, eight-wide kernels (kF & 16) the draws, the others neither (see the measurements above).
#ifndef GS_TAKE_ALL
#define GS_TAKE_ALL 0  // A/B knob: draws taken together in every kernel variant
#endif
__host__ __device__ constexpr int tune_of(int kF) { return ((kF & 8) ? 3 : ((kF & 16) ? 2 : 0)) | (GS_TAKE_ALL ? 2 : 0); }

template <int kF>
__device__ __forceinline__ bool path_step(const DeviceScene& sc, const DCamera& cam, const uint8_t* __restrict__ perm, PathState& st,
                                          PathCounters& pc) {
	double tmax;
	int medium_hit;
	if (!segment_begin(sc, cam, st, pc, tmax, medium_hit)) return true;
	Hit h = trace_closest<kF>(sc, st.ray, 0.001, tmax, &pc.node_visits);
	return segment_shade<(kF & 2) != 0, tune_of(kF)>(sc, cam, perm, st, pc, h, tmax, medium_hit);
}

}  // namespace gs
