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
	out[4] = (int64_t)s->blob_bytes; out[5] = (int64_t)s->h2d_bytes; out[6] = (int64_t)s->n_dnodes; out[7] = (int64_t)s->n_dnodes8;
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

// ---- experiment: how many popped nodes are already behind the closest hit?  (binary tree; decides whether storing the entry
// distance with a pushed link, to cull at pop time, is worth its stack traffic) ----
extern "C" int hs_pop_stats(gs_scene* s, const double* rays, size_t n, double tmin, double tmax, uint64_t* out) {
	using namespace gs;
	uint64_t visits = 0, pops = 0, stale = 0, leaf_tests = 0, stale_leaf = 0;
	if (!s->nodes) return 1;
	for (size_t k = 0; k < n; ++k) {
		const double* rp = rays + k * 7;
		Ray r;
		r.ox = rp[0]; r.oy = rp[1]; r.oz = rp[2]; r.dx = rp[3]; r.dy = rp[4]; r.dz = rp[5]; r.time = rp[6];
		RayConsts rc = ray_consts(r, tmax);
		int stack[kMaxStack]; float sten[kMaxStack];
		int sp = 0;
		int node = s->ds.root;
		Hit best; best.t = tmax; best.ref = kNoHit; best.u = best.v = 0.f;
		float cur_en = 0.f;
		for (;;) {
			if (node >= 0) {
				++visits;
				const DNode& nd = s->nodes[node];
				auto slab = [&](float c, float h, float id, float oi, float& tn, float& tf) { const float tm = fmaf(c, id, -oi), th = h * fabsf(id); tn = tm - th; tf = tm + th; };
				float n0x, f0x, n0y, f0y, n0z, f0z, n1x, f1x, n1y, f1y, n1z, f1z;
				slab(nd.xy0.x, nd.xy0.y, rc.idx, rc.oix, n0x, f0x); slab(nd.xy0.z, nd.xy0.w, rc.idy, rc.oiy, n0y, f0y); slab(nd.z01.x, nd.z01.y, rc.idz, rc.oiz, n0z, f0z);
				slab(nd.xy1.x, nd.xy1.y, rc.idx, rc.oix, n1x, f1x); slab(nd.xy1.z, nd.xy1.w, rc.idy, rc.oiy, n1y, f1y); slab(nd.z01.z, nd.z01.w, rc.idz, rc.oiz, n1z, f1z);
				const float en0 = fmaxf(fmaxf(n0x, n0y), fmaxf(n0z, 0.f)), ex0 = fminf(fminf(f0x, f0y), fminf(f0z, rc.tmax32));
				const float en1 = fmaxf(fmaxf(n1x, n1y), fmaxf(n1z, 0.f)), ex1 = fminf(fminf(f1x, f1y), fminf(f1z, rc.tmax32));
				const bool h0 = en0 <= ex0, h1 = en1 <= ex1;
				const bool first0 = h0 && (!h1 || !(en1 < en0));
				const int first = first0 ? nd.link.x : nd.link.y, second = first0 ? nd.link.y : nd.link.x;
				if (h0 && h1) { stack[sp] = second; sten[sp] = first0 ? en1 : en0; ++sp; }
				if (h0 || h1) { node = first; cur_en = first0 ? en0 : en1; continue; }
			} else {
				++leaf_tests;
				if (cur_en > rc.tmax32) ++stale_leaf;
				leaf_intersect<false>(s->ds, r, rc, tmin, best, node);
			}
			if (sp == 0) break;
			--sp; node = stack[sp]; cur_en = sten[sp]; ++pops;
			if (cur_en > rc.tmax32) ++stale;
		}
	}
	out[0] = visits; out[1] = pops; out[2] = stale; out[3] = leaf_tests; out[4] = stale_leaf;
	return 0;
}
