// grayshift_b200/host/flatten.cpp -- scene graph -> flat arrays of include/grayshift.h.
//
//  1. walk the graph in the reference's visiting order (HittableList in order, hittable.rs:75;
//     BVHNode left then right, BVH.rs:71-79), composing Translate / RotateY (hittable.rs:93-216) into one
//     rigid transform per primitive and baking it into world space (exact in R, ulps in f64; `t` is
//     unchanged because the reference never renormalises transformed directions);
//  2. hand every surface primitive its rank in that order (gs_quad.order / gs_sphere.order);
//  3. ConstantMedium boundaries go to the boundary arrays, not the BVH;
//  4. build a binned-SAH binary BVH over the world-space boxes (the reference's median-split tree is
//     available with FlattenOptions::reference_tree for A/B measurements);
//  5. materials / textures / images deduplicated by pointer identity (one row per shared Arc).
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <unordered_map>

#include "grayshift.hpp"

namespace grayshift {
namespace {

// rigid transform p -> R p + t with R a rotation about y: R = [[c,0,s],[0,1,0],[-s,0,c]] (hittable.rs:195-199)
struct Xform {
	double c = 1.0, s = 0.0;
	Vec3 t;
	Vec3 rot(const Vec3& p) const { return Vec3(c * p.x + s * p.z, p.y, -s * p.x + c * p.z); }
	Vec3 point(const Vec3& p) const { return rot(p) + t; }
	bool identity_rot() const { return c == 1.0 && s == 0.0; }
};

struct Flattener {
	FlatScene& out;
	std::unordered_map<const Material*, int> mat_ids;  // Arc sharing -> one table row (a scene may hold a material per primitive)
	std::unordered_map<const Texture*, int> tex_ids;
	int next_order = 0;
	std::vector<double> bounds;  // per surface primitive ref: min xyz, max xyz
	std::vector<uint32_t> refs;  // parallel to bounds
	// reference-tree emission (reference_tree option): filled during the ordered walk
	explicit Flattener(FlatScene& o) : out(o) {}

	int texture(const std::shared_ptr<Texture>& t) {
		if (!t) throw std::runtime_error("flatten: material without texture");
		auto it = tex_ids.find(t.get());
		if (it != tex_ids.end()) return it->second;
		gs_texture g{};
		g.kind = t->kind; g.a = g.b = -1;
		g.color[0] = t->albedo.x; g.color[1] = t->albedo.y; g.color[2] = t->albedo.z;
		g.scale = t->scale;
		if (t->kind == GS_TEX_CHECKER) { g.a = texture(t->even_texture); g.b = texture(t->odd_texture); }
		else if (t->kind == GS_TEX_IMAGE) g.a = out.image_index(t->image);
		else if (t->kind == GS_TEX_NOISE) out.has_noise = true;
		int id = (int)out.textures.size();  // children were pushed first: they precede their parent
		out.textures.push_back(g);
		tex_ids[t.get()] = id;
		return id;
	}
	int material(const std::shared_ptr<Material>& m) {
		if (!m) throw std::runtime_error("flatten: primitive without material");
		auto it = mat_ids.find(m.get());
		if (it != mat_ids.end()) return it->second;
		gs_material g{};
		g.kind = m->kind;
		g.texture = m->texture ? texture(m->texture) : -1;
		g.albedo[0] = m->albedo.x; g.albedo[1] = m->albedo.y; g.albedo[2] = m->albedo.z;
		g.param = m->param;
		int id = (int)out.materials.size();
		out.materials.push_back(g);
		mat_ids[m.get()] = id;
		return id;
	}

	gs_sphere bake_sphere(const Sphere& s, const Xform& x) {
		gs_sphere g{};
		Vec3 c = x.point(s.center_start), p = x.rot(s.center_path);
		g.center[0] = c.x; g.center[1] = c.y; g.center[2] = c.z;
		g.radius = s.radius;
		g.path[0] = p.x; g.path[1] = p.y; g.path[2] = p.z;
		g.material = material(s.material);
		g.prim_id = s.prim_id;
		g.moving = s.is_moving ? 1 : 0;
		// Sphere::hit takes (u, v) from the object-space normal, before RotateY::hit rotates the normal (hittable.rs:179-211)
		g.uv_rotation[0] = x.c; g.uv_rotation[1] = x.s;
		return g;
	}
	gs_quad bake_quad(const Quad& q, const Xform& x) {
		gs_quad g{};
		Vec3 qq = x.point(q.q), u = x.rot(q.u), v = x.rot(q.v);
		g.q[0] = qq.x; g.q[1] = qq.y; g.q[2] = qq.z;
		g.u[0] = u.x; g.u[1] = u.y; g.u[2] = u.z;
		g.v[0] = v.x; g.v[1] = v.y; g.v[2] = v.z;
		g.material = material(q.material);
		g.prim_id = q.prim_id;
		return g;
	}
	static void sphere_bounds(const gs_sphere& g, double* b) {
		double r = std::fabs(g.radius);
		for (int a = 0; a < 3; ++a) {
			double c0 = g.center[a], c1 = g.center[a] + (g.moving ? g.path[a] : 0.0);
			b[a] = std::fmin(c0, c1) - r; b[a + 3] = std::fmax(c0, c1) + r;
		}
	}
	static void quad_bounds(const gs_quad& g, double* b) {
		for (int a = 0; a < 3; ++a) {
			double p0 = g.q[a], p1 = g.q[a] + g.u[a], p2 = g.q[a] + g.v[a], p3 = g.q[a] + g.u[a] + g.v[a];
			b[a] = std::fmin(std::fmin(p0, p1), std::fmin(p2, p3));
			b[a + 3] = std::fmax(std::fmax(p0, p1), std::fmax(p2, p3));
		}
	}

	// surfaces: in the reference's visiting order
	void walk(const HittablePtr& h, const Xform& x) {
		switch (h->node_kind) {
			case NodeKind::Sphere: {
				gs_sphere g = bake_sphere(static_cast<const Sphere&>(*h), x);
				g.order = next_order++;
				double b[6]; sphere_bounds(g, b);
				bounds.insert(bounds.end(), b, b + 6);
				refs.push_back(GS_PRIM_REF(GS_PRIM_SPHERE, out.spheres.size()));
				out.spheres.push_back(g);
				break;
			}
			case NodeKind::Quad: {
				gs_quad g = bake_quad(static_cast<const Quad&>(*h), x);
				g.order = next_order++;
				double b[6]; quad_bounds(g, b);
				bounds.insert(bounds.end(), b, b + 6);
				refs.push_back(GS_PRIM_REF(GS_PRIM_QUAD, out.quads.size()));
				out.quads.push_back(g);
				break;
			}
			case NodeKind::Triangle: {
				const auto& t = static_cast<const Triangle&>(*h);
				gs_triangle g{};
				Vec3 a = x.point(t.a), b = x.point(t.b), c = x.point(t.c);
				g.a[0] = a.x; g.a[1] = a.y; g.a[2] = a.z; g.b[0] = b.x; g.b[1] = b.y; g.b[2] = b.z; g.c[0] = c.x; g.c[1] = c.y; g.c[2] = c.z;
				g.material = material(t.material);
				g.prim_id = t.prim_id;
				g.order = next_order++;
				double bd[6];
				for (int k = 0; k < 3; ++k) {
					bd[k] = std::fmin(g.a[k], std::fmin(g.b[k], g.c[k]));
					bd[k + 3] = std::fmax(g.a[k], std::fmax(g.b[k], g.c[k]));
				}
				bounds.insert(bounds.end(), bd, bd + 6);
				refs.push_back(GS_PRIM_REF(GS_PRIM_TRIANGLE, out.triangles.size()));
				out.triangles.push_back(g);
				break;
			}
			case NodeKind::List:
				for (const auto& o : static_cast<const HittableList&>(*h).objects) walk(o, x);
				break;
			case NodeKind::Bvh: {
				const auto& b = static_cast<const BVHNode&>(*h);
				walk(b.left, x);
				if (b.right) walk(b.right, x);
				break;
			}
			case NodeKind::Translate: {
				const auto& t = static_cast<const Translate&>(*h);
				Xform y = x;
				y.t = x.t + x.rot(t.offset);
				walk(t.object, y);
				break;
			}
			case NodeKind::RotateY: {
				const auto& r = static_cast<const RotateY&>(*h);
				Xform y = x;  // R_x * R_r: angles add
				y.c = x.c * r.cos_theta - x.s * r.sin_theta;
				y.s = x.s * r.cos_theta + x.c * r.sin_theta;
				walk(r.object, y);
				break;
			}
			case NodeKind::Medium: {
				const auto& m = static_cast<const ConstantMedium&>(*h);
				// the medium's fixed normal (1,0,0) (volume.rs:57) would be rotated by an enclosing RotateY
				if (!x.identity_rot()) throw std::runtime_error("flatten: ConstantMedium under RotateY is not supported");
				gs_medium g{};
				g.sphere_first = (int)out.boundary_spheres.size();
				g.quad_first = (int)out.boundary_quads.size();
				walk_boundary(m.boundary, x);
				g.sphere_count = (int)out.boundary_spheres.size() - g.sphere_first;
				g.quad_count = (int)out.boundary_quads.size() - g.quad_first;
				g.density = m.density;
				g.phase_material = material(m.phase_function);
				g.medium_index = m.medium_index;
				out.media.push_back(g);
				break;
			}
		}
	}
	void walk_boundary(const HittablePtr& h, const Xform& x) {
		switch (h->node_kind) {
			case NodeKind::Sphere: {
				gs_sphere g = bake_sphere(static_cast<const Sphere&>(*h), x);
				g.material = 0;  // a boundary's own material is never used (Q11)
				out.boundary_spheres.push_back(g);
				break;
			}
			case NodeKind::Quad: {
				gs_quad g = bake_quad(static_cast<const Quad&>(*h), x);
				g.material = 0;
				out.boundary_quads.push_back(g);
				break;
			}
			case NodeKind::List:
				for (const auto& o : static_cast<const HittableList&>(*h).objects) walk_boundary(o, x);
				break;
			case NodeKind::Bvh: {
				const auto& b = static_cast<const BVHNode&>(*h);
				walk_boundary(b.left, x);
				if (b.right) walk_boundary(b.right, x);
				break;
			}
			case NodeKind::Translate: {
				const auto& t = static_cast<const Translate&>(*h);
				Xform y = x;
				y.t = x.t + x.rot(t.offset);
				walk_boundary(t.object, y);
				break;
			}
			case NodeKind::RotateY: {
				const auto& r = static_cast<const RotateY&>(*h);
				Xform y = x;
				y.c = x.c * r.cos_theta - x.s * r.sin_theta;
				y.s = x.s * r.cos_theta + x.c * r.sin_theta;
				walk_boundary(r.object, y);
				break;
			}
			case NodeKind::Triangle:
				throw std::runtime_error("flatten: a triangle cannot bound a ConstantMedium (boundaries are spheres or quads)");
			case NodeKind::Medium:
				throw std::runtime_error("flatten: a ConstantMedium cannot bound another medium");
		}
	}
};

// ---- reference-tree emission: median split over the *baked* primitives in visiting order ---------------
// (used only for A/B measurements of tree quality; the SAH tree is the default)
void build_bvh_median(const std::vector<double>& bounds, std::vector<gs_bvh_node>& nodes, std::vector<uint32_t>& order) {
	const size_t n = bounds.size() / 6;
	order.resize(n);
	for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
	struct Rec {
		const std::vector<double>& b; std::vector<gs_bvh_node>& nodes; std::vector<uint32_t>& order;
		int build(size_t lo, size_t hi) {
			int id = (int)nodes.size();
			nodes.emplace_back();
			double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
			for (size_t k = lo; k < hi; ++k) for (int a = 0; a < 3; ++a) {
				mn[a] = std::fmin(mn[a], b[order[k] * 6 + a]); mx[a] = std::fmax(mx[a], b[order[k] * 6 + 3 + a]);
			}
			gs_bvh_node nd{};
			for (int a = 0; a < 3; ++a) { nd.bounds_min[a] = mn[a]; nd.bounds_max[a] = mx[a]; }
			if (hi - lo <= 2) { nd.first_ref = (int)lo; nd.ref_count = (int)(hi - lo); nodes[id] = nd; return id; }
			double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
			int axis = ex > ey ? (ex > ez ? 0 : 2) : (ey > ez ? 1 : 2);
			std::stable_sort(order.begin() + lo, order.begin() + hi, [&](uint32_t p, uint32_t q) { return b[p * 6 + axis] < b[q * 6 + axis]; });
			size_t mid = lo + (hi - lo) / 2;
			nd.left = build(lo, mid);
			nd.right = build(mid, hi);
			nodes[id] = nd;
			return id;
		}
	} rec{bounds, nodes, order};
	nodes.clear();
	rec.build(0, n);
}

}  // namespace

// ---- binned SAH builder -----------------------------------------------------------------------------------
void build_bvh_sah(const std::vector<double>& bounds, int max_leaf_size, std::vector<gs_bvh_node>& nodes, std::vector<uint32_t>& order) {
	const size_t n = bounds.size() / 6;
	if (n == 0) throw std::runtime_error("build_bvh_sah: no primitives");
	if (max_leaf_size < 1) max_leaf_size = 1;
	if (max_leaf_size > 15) max_leaf_size = 15;
	order.resize(n);
	for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
	std::vector<double> cent(n * 3);
	for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) cent[i * 3 + a] = 0.5 * (bounds[i * 6 + a] + bounds[i * 6 + 3 + a]);
	constexpr int kBins = 16;
	constexpr int kMaxDepth = 56;  // device traversal stack is 64 entries

	struct Box {
		double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
		void grow(const double* b) { for (int a = 0; a < 3; ++a) { mn[a] = std::fmin(mn[a], b[a]); mx[a] = std::fmax(mx[a], b[a + 3]); } }
		void grow(const Box& o) { for (int a = 0; a < 3; ++a) { mn[a] = std::fmin(mn[a], o.mn[a]); mx[a] = std::fmax(mx[a], o.mx[a]); } }
		double area() const {
			double dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
			if (dx < 0) return 0.0;
			return 2.0 * (dx * dy + dy * dz + dz * dx);
		}
	};

	// Nodes are allocated from a pre-sized array with an atomic counter, so subtrees can be built by several threads
	// (a tree over n primitives with single-primitive leaves has at most 2n - 1 nodes).  One task = one node: bound it,
	// pick the SAH split, partition its range of `order` in place, allocate the two children.
	struct Task { size_t lo, hi; int node, depth; };
	nodes.assign(2 * n, gs_bvh_node{});
	std::atomic<int> next_node{1};
	auto process = [&](const Task& tk, std::vector<Task>& out) {
		const size_t lo = tk.lo, hi = tk.hi, cnt = hi - lo;
		Box box, cbox;
		for (size_t k = lo; k < hi; ++k) {
			box.grow(&bounds[order[k] * 6]);
			const double* c = &cent[order[k] * 3];
			for (int a = 0; a < 3; ++a) { cbox.mn[a] = std::fmin(cbox.mn[a], c[a]); cbox.mx[a] = std::fmax(cbox.mx[a], c[a]); }
		}
		gs_bvh_node nd{};
		for (int a = 0; a < 3; ++a) { nd.bounds_min[a] = box.mn[a]; nd.bounds_max[a] = box.mx[a]; }
		auto make_leaf = [&]() { nd.first_ref = (int)lo; nd.ref_count = (int)cnt; nodes[tk.node] = nd; };
		if (cnt == 1) { make_leaf(); return; }

		size_t mid = lo;
		bool split_found = false;
		if (tk.depth < kMaxDepth - 24) {
			double best_cost = DBL_MAX; int best_axis = -1, best_bin = -1;
			for (int axis = 0; axis < 3; ++axis) {
				double cmin = cbox.mn[axis], cext = cbox.mx[axis] - cbox.mn[axis];
				if (!(cext > 0.0)) continue;
				Box bb[kBins]; size_t bc[kBins] = {0};
				double scale = (double)kBins / cext;
				for (size_t k = lo; k < hi; ++k) {
					int bi = (int)((cent[order[k] * 3 + axis] - cmin) * scale);
					if (bi >= kBins) bi = kBins - 1;
					if (bi < 0) bi = 0;
					bb[bi].grow(&bounds[order[k] * 6]); bc[bi]++;
				}
				double right_area[kBins]; size_t right_cnt[kBins];
				Box acc; size_t c = 0;
				for (int i = kBins - 1; i > 0; --i) { acc.grow(bb[i]); c += bc[i]; right_area[i] = acc.area(); right_cnt[i] = c; }
				Box accl; size_t cl = 0;
				for (int i = 1; i < kBins; ++i) {
					accl.grow(bb[i - 1]); cl += bc[i - 1];
					if (cl == 0 || right_cnt[i] == 0) continue;
					double cost = accl.area() * (double)cl + right_area[i] * (double)right_cnt[i];
					if (cost < best_cost) { best_cost = cost; best_axis = axis; best_bin = i; }
				}
			}
			if (best_axis >= 0) {
				// SAH termination: leaf cost = cnt * area(box); split cost = traversal (1.0 box pair) + children
				double leaf_cost = (double)cnt * box.area();
				double split_cost = 1.0 * box.area() + best_cost;
				if (cnt > (size_t)max_leaf_size || split_cost < leaf_cost) {
					double cmin = cbox.mn[best_axis], scale = (double)kBins / (cbox.mx[best_axis] - cbox.mn[best_axis]);
					auto it = std::partition(order.begin() + lo, order.begin() + hi, [&](uint32_t p) {
						int bi = (int)((cent[p * 3 + best_axis] - cmin) * scale);
						if (bi >= kBins) bi = kBins - 1;
						if (bi < 0) bi = 0;
						return bi < best_bin;
					});
					mid = (size_t)(it - order.begin());
					split_found = mid > lo && mid < hi;
				}
			}
		}
		if (!split_found) {
			if (cnt <= (size_t)max_leaf_size) { make_leaf(); return; }
			// coincident centroids or depth guard: median split by index along the widest centroid axis
			int axis = 0;
			double e0 = cbox.mx[0] - cbox.mn[0], e1 = cbox.mx[1] - cbox.mn[1], e2 = cbox.mx[2] - cbox.mn[2];
			if (e1 > e0 && e1 >= e2) axis = 1; else if (e2 > e0 && e2 > e1) axis = 2;
			mid = lo + cnt / 2;
			std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
			                 [&](uint32_t p, uint32_t q) { return cent[p * 3 + axis] < cent[q * 3 + axis]; });
		}
		nd.left = next_node.fetch_add(2);
		nd.right = nd.left + 1;
		nodes[tk.node] = nd;
		out.push_back({mid, hi, nd.right, tk.depth + 1});
		out.push_back({lo, mid, nd.left, tk.depth + 1});
	};

	unsigned threads = std::thread::hardware_concurrency();
	if (const char* e = std::getenv("GSH_BUILD_THREADS")) threads = (unsigned)std::max(1, std::atoi(e));
	threads = std::min(threads, 16u);
	const bool parallel = threads > 1 && n >= (std::getenv("GSH_BUILD_THREADS") ? 64u : 50000u);
	std::vector<Task> open{{0, n, 0, 0}};
	if (parallel) {
		// the top of the tree on this thread, widest task first, until there are enough subtrees to hand out
		while (open.size() < 8u * threads) {
			size_t big = 0;
			for (size_t i = 1; i < open.size(); ++i) if (open[i].hi - open[i].lo > open[big].hi - open[big].lo) big = i;
			if (open[big].hi - open[big].lo <= 4096) break;
			const Task tk = open[big];
			open.erase(open.begin() + (long)big);
			process(tk, open);
		}
		std::atomic<size_t> next_task{0};
		std::vector<std::thread> pool;
		std::exception_ptr failure;
		std::mutex failure_lock;
		for (unsigned t = 0; t < threads; ++t)
			pool.emplace_back([&]() {
				try {
					std::vector<Task> stack;
					for (size_t i; (i = next_task.fetch_add(1)) < open.size();) {
						stack.assign(1, open[i]);
						while (!stack.empty()) { const Task tk = stack.back(); stack.pop_back(); process(tk, stack); }
					}
				} catch (...) { std::lock_guard<std::mutex> g(failure_lock); failure = std::current_exception(); }
			});
		for (auto& th : pool) th.join();
		if (failure) std::rethrow_exception(failure);
	} else {
		while (!open.empty()) { const Task tk = open.back(); open.pop_back(); process(tk, open); }
	}
	nodes.resize((size_t)next_node.load());
	if (parallel) {
		// the allocation order depends on thread timing: renumber in depth-first pre-order (children after their parent,
		// as gs_scene_create requires) so that the same input always gives the same array
		std::vector<gs_bvh_node> sorted;
		sorted.reserve(nodes.size());
		std::vector<std::pair<int, int>> dfs{{0, -1}};  // (old index, parent's new index; the right child is pushed first)
		std::vector<int> renamed(nodes.size(), -1);
		while (!dfs.empty()) {
			const int old = dfs.back().first;
			dfs.pop_back();
			renamed[old] = (int)sorted.size();
			sorted.push_back(nodes[old]);
			if (nodes[old].ref_count == 0) { dfs.push_back({nodes[old].right, 0}); dfs.push_back({nodes[old].left, 0}); }
		}
		for (auto& nd : sorted) if (nd.ref_count == 0) { nd.left = renamed[nd.left]; nd.right = renamed[nd.right]; }
		nodes.swap(sorted);
	}
}

int FlatScene::image_index(const std::shared_ptr<ImageData>& img) {
	if (!img) throw std::runtime_error("flatten: null image");
	for (size_t i = 0; i < image_data.size(); ++i) if (image_data[i].get() == img.get()) return (int)i;
	image_data.push_back(img);
	gs_image g{};
	g.width = img->width; g.height = img->height; g.format = img->format;
	g.pixels = img->format == GS_IMG_U8_RGB ? (const void*)img->u8.data() : (const void*)img->f32.data();
	images.push_back(g);
	return (int)images.size() - 1;
}

void FlatScene::finish() {
	std::memset(&desc, 0, sizeof desc);
	desc.abi_version = GS_ABI_VERSION;
	desc.spheres = spheres.data(); desc.n_spheres = spheres.size();
	desc.quads = quads.data(); desc.n_quads = quads.size();
	desc.nodes = nodes.data(); desc.n_nodes = nodes.size();
	desc.prim_refs = prim_refs.data(); desc.n_prim_refs = prim_refs.size();
	desc.boundary_spheres = boundary_spheres.data(); desc.n_boundary_spheres = boundary_spheres.size();
	desc.boundary_quads = boundary_quads.data(); desc.n_boundary_quads = boundary_quads.size();
	desc.media = media.data(); desc.n_media = media.size();
	desc.materials = materials.data(); desc.n_materials = materials.size();
	desc.textures = textures.data(); desc.n_textures = textures.size();
	for (size_t i = 0; i < images.size(); ++i)
		images[i].pixels = image_data[i]->format == GS_IMG_U8_RGB ? (const void*)image_data[i]->u8.data() : (const void*)image_data[i]->f32.data();
	desc.images = images.data(); desc.n_images = images.size();
	desc.perlin_perm = has_noise ? perlin_perm : nullptr;
	desc.triangles = triangles.data(); desc.n_triangles = triangles.size();
}

std::unique_ptr<FlatScene> flatten(const HittablePtr& world, const FlattenOptions& opt) {
	if (!world) throw std::runtime_error("flatten: null world");
	auto flat = std::make_unique<FlatScene>();
	// GSH_TRACE=1: stage times on stderr
	const bool trace = std::getenv("GSH_TRACE") != nullptr;
	auto t_prev = std::chrono::steady_clock::now();
	auto lap = [&](const char* what) {
		if (!trace) return;
		const auto now = std::chrono::steady_clock::now();
		std::fprintf(stderr, "[flatten] %-32s %9.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
		t_prev = now;
	};
	Flattener f(*flat);
	f.walk(world, Xform());
	lap("walk + bake + dedupe");
	if (f.refs.empty()) throw std::runtime_error("flatten: the world holds no surface primitive");
	std::vector<uint32_t> order;
	if (opt.reference_tree) build_bvh_median(f.bounds, flat->nodes, order);
	else build_bvh_sah(f.bounds, opt.max_leaf_size, flat->nodes, order);
	lap("BVH build");
	flat->prim_refs.resize(order.size());
	for (size_t k = 0; k < order.size(); ++k) flat->prim_refs[k] = f.refs[order[k]];
	perlin_permutation_table(0, flat->perlin_perm);  // Perlin::default(), texture.rs:104
	flat->finish();
	lap("finish");
	return flat;
}

}  // namespace grayshift
