// tests/hostsim/cuda_shim.h -- TEST INFRASTRUCTURE.  Lets the device-side traversal (grayshift_b200/csrc/gs_device.cuh) and the
// scene layout (scene_layout.cu) compile as plain host C++, so that the eight-wide tree builder and the traversal code that
// walks it can be checked against the oracle on a machine without a GPU (pytest -m "not gpu").  Nothing under grayshift_b200/
// includes this; the product runs these functions on the device only.
#pragma once
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
	const uint64_t both = ((uint64_t)y << 32) | x;
	unsigned r = 0;
	for (int i = 0; i < 4; ++i) r |= (unsigned)((both >> (8 * ((s >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
	return r;
}
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline float __double2float_ru(double x) { float f = (float)x; if ((double)f < x) f = std::nextafterf(f, INFINITY); return f; }
