// grayshift_b200/csrc/trace_closest.cu -- parity kernel K1: closest-hit primitive id and t for a batch of
// rays (gs_trace_closest).  Same traversal and intersection code as the render megakernel
// (gs_device.cuh), compiled in this translation unit with -fmad=false so the f64 primitive
// arithmetic is evaluated without FMA contraction, like the Rust reference.
#define GS_EXACT_F64 1  // IEEE division and square root, like the reference
#include "gs_internal.h"
#define GS_TRACE_NAME trace_closest
#include "trace_closest_body.cuh"
