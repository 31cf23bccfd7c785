// oracle/graph_io.cpp -- neutral scene-graph text format ("gsgraph 1"), shared with the host
// mirror (grayshift_b200/host/graph_io.cpp implements the same grammar independently).
// TEST INFRASTRUCTURE ONLY (see oracle.hpp).
//
// Records are emitted in post-order DFS from the root, ids renumbered in emission order, so two
// independent builders of the same scene produce byte-identical dumps.  Doubles are %.17g.
//
//   gsgraph 1
//   image <id> u8|f32 <w> <h> <ppm:|hdr:|raw:><path>
//   texture <id> solid r g b | checker scale even odd | image <image> | noise scale
//   material <id> lambertian <tex> | metal r g b fuzz | dielectric ir | light <tex> | isotropic <tex>
//   node <id> sphere cx cy cz r <mat> <prim> | msphere c1 c2 r <mat> <prim> | quad q u v <mat> <prim>
//           | tri a b c <mat> <prim> | list n ids.. | bvh n ids.. | translate <child> ox oy oz
//           | rotate_y <child> deg | medium <child> density <phase mat> <medium index>
//   root <id>
//   camera aspect width confidence tolerance batch max depth vfov from(3) at(3) vup(3) defocus focus
//   background solid r g b | background hdri <image> rx ry rz
#include "oracle.hpp"

#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

namespace orc {
namespace {

std::string d17(double v) { char b[40]; std::snprintf(b, sizeof b, "%.17g", v); return b; }
std::string v17(const Vec3& v) { return d17(v.x) + " " + d17(v.y) + " " + d17(v.z); }

struct Dumper {
	std::ostringstream out;
	std::map<const void*, int> img_ids, tex_ids, mat_ids, node_ids;

	int image_u8(const std::shared_ptr<ImageU8>& im) {
		auto it = img_ids.find(im.get());
		if (it != img_ids.end()) return it->second;
		int id = (int)img_ids.size();
		img_ids[im.get()] = id;
		out << "image " << id << " u8 " << im->w << " " << im->h << " " << im->source << "\n";
		return id;
	}
	int image_f32(const std::shared_ptr<ImageF32>& im) {
		auto it = img_ids.find(im.get());
		if (it != img_ids.end()) return it->second;
		int id = (int)img_ids.size();
		img_ids[im.get()] = id;
		out << "image " << id << " f32 " << im->w << " " << im->h << " " << im->source << "\n";
		return id;
	}
	int texture(const std::shared_ptr<Texture>& t) {
		auto it = tex_ids.find(t.get());
		if (it != tex_ids.end()) return it->second;
		std::ostringstream rec;
		if (auto s = dynamic_cast<SolidColorTexture*>(t.get())) rec << "solid " << v17(s->albedo);
		else if (auto c = dynamic_cast<CheckeredTexture*>(t.get())) {
			int e = texture(c->even_texture), o = texture(c->odd_texture);
			rec << "checker " << d17(c->scale) << " " << e << " " << o;
		} else if (auto i = dynamic_cast<ImageTexture*>(t.get())) rec << "image " << image_u8(i->image);
		else if (auto n = dynamic_cast<NoiseTexture*>(t.get())) rec << "noise " << d17(n->scale);
		else throw std::runtime_error("dump: unknown texture");
		int id = (int)tex_ids.size();
		tex_ids[t.get()] = id;
		out << "texture " << id << " " << rec.str() << "\n";
		return id;
	}
	int material(const std::shared_ptr<Material>& m) {
		auto it = mat_ids.find(m.get());
		if (it != mat_ids.end()) return it->second;
		std::ostringstream rec;
		if (auto l = dynamic_cast<Lambertian*>(m.get())) rec << "lambertian " << texture(l->texture);
		else if (auto me = dynamic_cast<Metal*>(m.get())) rec << "metal " << v17(me->albedo) << " " << d17(me->fuzz);
		else if (auto d = dynamic_cast<Dielectric*>(m.get())) rec << "dielectric " << d17(d->refraction_index);
		else if (auto li = dynamic_cast<DiffuseLight*>(m.get())) rec << "light " << texture(li->texture);
		else if (auto is = dynamic_cast<Isotropic*>(m.get())) rec << "isotropic " << texture(is->texture);
		else throw std::runtime_error("dump: unknown material");
		int id = (int)mat_ids.size();
		mat_ids[m.get()] = id;
		out << "material " << id << " " << rec.str() << "\n";
		return id;
	}
	int node(const std::shared_ptr<Hittable>& h) {
		auto it = node_ids.find(h.get());
		if (it != node_ids.end()) return it->second;
		std::ostringstream rec;
		if (auto s = dynamic_cast<Sphere*>(h.get())) {
			int m = material(s->material);
			if (s->is_moving) rec << "msphere " << v17(s->center_start) << " " << v17(s->center_end) << " " << d17(s->radius) << " " << m << " " << s->prim_id;
			else rec << "sphere " << v17(s->center_start) << " " << d17(s->radius) << " " << m << " " << s->prim_id;
		} else if (auto q = dynamic_cast<Quad*>(h.get())) {
			int m = material(q->material);
			rec << "quad " << v17(q->q) << " " << v17(q->u) << " " << v17(q->v) << " " << m << " " << q->prim_id;
		} else if (auto t = dynamic_cast<Triangle*>(h.get())) {
			int m = material(t->material);
			rec << "tri " << v17(t->a) << " " << v17(t->b) << " " << v17(t->c) << " " << m << " " << t->prim_id;
		} else throw std::runtime_error("dump: unknown hittable");
		int id = (int)node_ids.size();
		node_ids[h.get()] = id;
		out << "node " << id << " " << rec.str() << "\n";
		return id;
	}
};

}  // namespace

void dump_graph(const Scene& scene, const std::string& path, const std::string&) {
	// BVH roots need the scene's node table to resolve the source list -> handle them here.
	struct D2 : Dumper {
		const Scene& sc;
		explicit D2(const Scene& s) : sc(s) {}
		int any(const std::shared_ptr<Hittable>& h) {
			auto it = node_ids.find(h.get());
			if (it != node_ids.end()) return it->second;
			std::ostringstream rec;
			if (auto b = dynamic_cast<BVHNode*>(h.get())) {
				std::vector<int> ids;
				for (int cid : b->source_children) ids.push_back(any(sc.nodes.at(cid)));
				rec << "bvh " << ids.size();
				for (int i : ids) rec << " " << i;
			} else if (auto l = dynamic_cast<HittableList*>(h.get())) {
				std::vector<int> ids;
				for (auto& o : l->objects) ids.push_back(any(o));
				rec << "list " << ids.size();
				for (int i : ids) rec << " " << i;
			} else if (auto tr = dynamic_cast<Translate*>(h.get())) {
				int c = any(tr->object);
				rec << "translate " << c << " " << v17(tr->offset);
			} else if (auto r = dynamic_cast<RotateY*>(h.get())) {
				int c = any(r->object);
				rec << "rotate_y " << c << " " << d17(r->angle_deg);
			} else if (auto cm = dynamic_cast<ConstantMedium*>(h.get())) {
				int c = any(cm->boundary);
				int m = material(cm->phase_function);
				rec << "medium " << c << " " << d17(cm->density) << " " << m << " " << cm->medium_index;
			} else return node(h);  // leaf primitives
			int id = (int)node_ids.size();
			node_ids[h.get()] = id;
			out << "node " << id << " " << rec.str() << "\n";
			return id;
		}
	} d(scene);
	d.out << "gsgraph 1\n";
	int root = d.any(scene.root);
	d.out << "root " << root << "\n";
	const Camera& c = scene.camera;
	d.out << "camera " << d17(c.arg_aspect) << " " << c.image_width << " " << d17(c.sample_settings.confidence) << " "
	      << d17(c.sample_settings.tolerance) << " " << c.sample_settings.batch_size << " " << c.sample_settings.max_samples << " "
	      << c.max_depth << " " << d17(c.arg_vfov) << " " << v17(c.arg_from) << " " << v17(c.arg_at) << " " << v17(c.arg_vup) << " "
	      << d17(c.defocus_angle) << " " << d17(c.arg_focus) << "\n";
	if (c.background.is_hdri) {
		int im = d.image_f32(c.background.hdri.image);
		d.out << "background hdri " << im << " " << v17(c.background.hdri.rotation) << "\n";
	} else d.out << "background solid " << v17(c.background.solid) << "\n";
	std::ofstream f(path);
	if (!f) throw std::runtime_error("cannot write " + path);
	f << d.out.str();
}

std::unique_ptr<Scene> load_graph(const std::string& path, bool include_media) {
	std::ifstream f(path);
	if (!f) throw std::runtime_error("cannot open " + path);
	auto s = std::make_unique<Scene>();
	s->name = path;
	std::map<int, std::shared_ptr<ImageU8>> img8;
	std::map<int, std::shared_ptr<ImageF32>> img32;
	std::map<int, std::shared_ptr<Texture>> tex;
	std::map<int, std::shared_ptr<Material>> mat;
	std::map<int, std::shared_ptr<Hittable>> nodes;  // nullptr = dropped medium
	std::string line;
	auto V = [](std::istringstream& is) { Vec3 v; is >> v.x >> v.y >> v.z; return v; };
	bool have_header = false, have_cam = false;
	struct CamRec { double aspect, conf, tol, vfov, defocus, focus; int width; uint32_t batch, max, depth; Vec3 from, at, vup; } cr{};
	Background bg;
	int root = -1;
	while (std::getline(f, line)) {
		if (line.empty() || line[0] == '#') continue;
		std::istringstream is(line);
		std::string kw; is >> kw;
		if (kw == "gsgraph") { int ver; is >> ver; if (ver != 1) throw std::runtime_error("gsgraph version"); have_header = true; }
		else if (kw == "image") {
			int id, w, h; std::string ty, src; is >> id >> ty >> w >> h; std::getline(is, src);
			src.erase(0, src.find_first_not_of(' '));
			std::string scheme = src.substr(0, 4), p = src.substr(4);
			if (ty == "u8") {
				auto im = scheme == "ppm:" ? load_ppm_u8(p) : load_raw_u8(p, w, h);
				im->source = src;
				if (im->w != w || im->h != h) throw std::runtime_error("image size mismatch: " + src);
				img8[id] = im; s->images_u8.push_back(im);
			} else {
				auto im = scheme == "hdr:" ? load_hdr_rgbe(p) : load_raw_f32(p, w, h);
				im->source = src;
				if (im->w != w || im->h != h) throw std::runtime_error("image size mismatch: " + src);
				img32[id] = im; s->images_f32.push_back(im);
			}
		} else if (kw == "texture") {
			int id; std::string ty; is >> id >> ty;
			if (ty == "solid") tex[id] = s->solid(V(is));
			else if (ty == "checker") { double sc; int e, o; is >> sc >> e >> o; tex[id] = s->add(std::make_shared<CheckeredTexture>(sc, tex.at(e), tex.at(o))); }
			else if (ty == "image") { int im; is >> im; tex[id] = s->add(std::make_shared<ImageTexture>(img8.at(im))); }
			else if (ty == "noise") { double sc; is >> sc; tex[id] = s->add(std::make_shared<NoiseTexture>(sc)); }
			else throw std::runtime_error("texture type " + ty);
		} else if (kw == "material") {
			int id; std::string ty; is >> id >> ty;
			if (ty == "lambertian") { int t; is >> t; mat[id] = s->lambertian(tex.at(t)); }
			else if (ty == "metal") { Vec3 a = V(is); double fz; is >> fz; mat[id] = s->metal(a, fz); }
			else if (ty == "dielectric") { double ir; is >> ir; mat[id] = s->dielectric(ir); }
			else if (ty == "light") { int t; is >> t; mat[id] = s->add(std::make_shared<DiffuseLight>(tex.at(t))); }
			else if (ty == "isotropic") { int t; is >> t; mat[id] = s->add(std::make_shared<Isotropic>(tex.at(t))); }
			else throw std::runtime_error("material type " + ty);
		} else if (kw == "node") {
			int id; std::string ty; is >> id >> ty;
			auto kids = [&](std::istringstream& in) {
				size_t n; in >> n; std::vector<std::shared_ptr<Hittable>> v;
				for (size_t k = 0; k < n; ++k) { int c; in >> c; auto o = nodes.at(c); if (o) v.push_back(o); }
				return v;
			};
			if (ty == "sphere") { Vec3 c = V(is); double r; int m, pid; is >> r >> m >> pid; auto sp = s->sphere(c, r, mat.at(m)); sp->prim_id = pid; nodes[id] = sp; }
			else if (ty == "msphere") { Vec3 c1 = V(is), c2 = V(is); double r; int m, pid; is >> r >> m >> pid; auto sp = s->moving_sphere(c1, c2, r, mat.at(m)); sp->prim_id = pid; nodes[id] = sp; }
			else if (ty == "quad") { Vec3 q = V(is), u = V(is), v = V(is); int m, pid; is >> m >> pid; auto qd = s->quad(q, u, v, mat.at(m)); qd->prim_id = pid; nodes[id] = qd; }
			else if (ty == "tri") { Vec3 a = V(is), b = V(is), c = V(is); int m, pid; is >> m >> pid; auto t = s->triangle(a, b, c, mat.at(m)); t->prim_id = pid; nodes[id] = t; }
			else if (ty == "list") { auto l = s->list(); for (auto& o : kids(is)) l->add(o); nodes[id] = l; }
			else if (ty == "bvh") { HittableList tmp; for (auto& o : kids(is)) tmp.add(o); nodes[id] = s->bvh_from_list(tmp); }
			else if (ty == "translate") { int c; is >> c; Vec3 off = V(is); auto o = nodes.at(c); nodes[id] = o ? s->translate(o, off) : nullptr; }
			else if (ty == "rotate_y") { int c; double deg; is >> c >> deg; auto o = nodes.at(c); nodes[id] = o ? s->rotate_y(o, deg) : nullptr; }
			else if (ty == "medium") {
				int c, m, mi; double dens; is >> c >> dens >> m >> mi;
				if (include_media) { auto md = s->medium(nodes.at(c), dens, mat.at(m)); md->medium_index = mi; nodes[id] = md; }
				else nodes[id] = nullptr;
			} else throw std::runtime_error("node type " + ty);
			if (!is) throw std::runtime_error("malformed record: " + line);
		} else if (kw == "root") is >> root;
		else if (kw == "camera") {
			is >> cr.aspect >> cr.width >> cr.conf >> cr.tol >> cr.batch >> cr.max >> cr.depth >> cr.vfov;
			cr.from = V(is); cr.at = V(is); cr.vup = V(is); is >> cr.defocus >> cr.focus;
			have_cam = true;
		} else if (kw == "background") {
			std::string ty; is >> ty;
			if (ty == "solid") { bg.is_hdri = false; bg.solid = V(is); }
			else { int im; is >> im; bg.is_hdri = true; bg.hdri.image = img32.at(im); bg.hdri.rotation = V(is); }
		} else throw std::runtime_error("unknown record: " + line);
	}
	if (!have_header || root < 0 || !have_cam) throw std::runtime_error(path + ": incomplete graph");
	s->root = nodes.at(root);
	if (!s->root) throw std::runtime_error(path + ": root dropped");
	s->camera = Camera(cr.aspect, cr.width, SampleSettings{cr.conf, cr.tol, cr.batch, cr.max}, cr.depth, cr.vfov, cr.from, cr.at, cr.vup,
	                   cr.defocus, cr.focus, bg);
	return s;
}

}  // namespace orc
