// tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE.  gs_scene_create's layout (validation, primitive order, binary and eight-wide
// node arrays) and the device traversal code, compiled for the host (cuda_shim.h, fake_cudart.cpp): hs_trace_closest runs the very
// functions the kernels run -- Trav2 / Trav8 node steps, leaf_intersect, the f64 primitive tests -- one ray at a time on the CPU.
#include "cuda_shim.h"
#define GS_EXACT_F64 1  // IEEE division / square root (the approximate versions are inline PTX)
#include "../../grayshift_b200/csrc/scene_layout.cu"

extern "C" {

const char* hs_last_error() { return gs::last_error_cstr(); }

// width = 2 or 8: the tree to build and traverse
int hs_scene_create(const gs_scene_desc* d, int width, gs_scene** out) {
	*out = nullptr;
	int rc = gs::validate_desc(d);
	if (rc) return rc;
	gs_scene* s = new gs_scene();
	rc = gs::layout_scene(s, d);
	if (!rc) {
		s->wide = width == 8;
		if (s->wide && s->w_child_base.empty()) { delete s; return gs::fail(GS_ERR_UNSUPPORTED, "single-leaf scene has no wide tree"); }
		rc = gs::build_current_nodes(s, 4.0 * std::fmax(s->max_abs_coord, 1.0));
	}
	if (rc) { delete s; return rc; }
	*out = s;
	return 0;
}
void hs_scene_destroy(gs_scene* s) { delete s; }
int hs_scene_info(gs_scene* s, int64_t* out) {  // nodes2, nodes8, wide_depth, any leaf links
	out[0] = (int64_t)s->h_bfs.size(); out[1] = (int64_t)s->w_child_base.size(); out[2] = s->wide_depth; out[3] = s->leaf_links8 != nullptr;
	return 0;
}

int hs_trace_closest(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, int32_t* prim_id, double* t, uint32_t* node_visits) {
	using namespace gs;
	for (size_t k = 0; k < n; ++k) {
		const double* rp = rays + k * 7;
		Ray r;
		r.ox = rp[0]; r.oy = rp[1]; r.oz = rp[2]; r.dx = rp[3]; r.dy = rp[4]; r.dz = rp[5]; r.time = rp[6];
		uint32_t visits = 0;
		const Hit h = s->wide ? trace_closest<19>(s->ds, r, tmin, tmax, &visits) : trace_closest<3>(s->ds, r, tmin, tmax, &visits);
		if (node_visits) node_visits[k] = visits;
		if (h.ref == kNoHit) { prim_id[k] = -1; t[k] = 0.0; continue; }
		const uint32_t index = h.ref & kRefIndexMask, kind = h.ref >> kRefKindShift;
		prim_id[k] = kind == kKindQuad ? s->ds.quad_aux[index].y : (kind == kKindTriangle ? s->ds.tri_aux[index].y : s->ds.sph_aux[index].y);
		t[k] = h.t;
	}
	return 0;
}
}
