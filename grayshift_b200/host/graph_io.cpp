// grayshift_b200/host/graph_io.cpp -- text form of a host scene graph ("gsgraph 1").  One record per line,
// emitted in post-order from the root with ids assigned in emission order, doubles as %.17g:
//
//   image <id> u8|f32 <w> <h> <ppm:|hdr:|raw:><path>
//   texture <id> solid r g b | checker scale <even> <odd> | image <image> | noise scale
//   material <id> lambertian <tex> | metal r g b fuzz | dielectric ir | light <tex> | isotropic <tex>
//   node <id> sphere c(3) r <mat> <prim> | msphere c1(3) c2(3) r <mat> <prim> | quad q(3) u(3) v(3) <mat> <prim>
//           | tri a(3) b(3) c(3) <mat> <prim> | list n ids.. | bvh n ids.. | translate <child> off(3) | rotate_y <child> deg
//           | medium <boundary> density <phase mat> <medium index>
//   root <id>
//   camera aspect width confidence tolerance batch max depth vfov from(3) at(3) vup(3) defocus focus
//   background solid r g b | background hdri <image> rot(3)
//
// The parity tests load the same file into the CPU oracle, and compare this dump with the dump of the
// oracle's independently written scene builders byte for byte.
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>

#include "grayshift.hpp"

namespace grayshift {
namespace {

std::string num(double v) { char b[40]; std::snprintf(b, sizeof b, "%.17g", v); return b; }
std::string vec(const Vec3& v) { return num(v.x) + " " + num(v.y) + " " + num(v.z); }

struct Writer {
	std::ostringstream out;
	std::map<const void*, int> images, textures, materials, nodes;

	int image(const std::shared_ptr<ImageData>& im) {
		auto it = images.find(im.get());
		if (it != images.end()) return it->second;
		int id = (int)images.size();
		images[im.get()] = id;
		out << "image " << id << (im->format == GS_IMG_U8_RGB ? " u8 " : " f32 ") << im->width << " " << im->height << " " << im->source << "\n";
		return id;
	}
	int texture(const std::shared_ptr<Texture>& t) {
		auto it = textures.find(t.get());
		if (it != textures.end()) return it->second;
		std::string rec;
		switch (t->kind) {
			case GS_TEX_SOLID: rec = "solid " + vec(t->albedo); break;
			case GS_TEX_CHECKER: {
				int e = texture(t->even_texture);
				int o = texture(t->odd_texture);
				rec = "checker " + num(t->scale) + " " + std::to_string(e) + " " + std::to_string(o);
				break;
			}
			case GS_TEX_IMAGE: rec = "image " + std::to_string(image(t->image)); break;
			default: rec = "noise " + num(t->scale); break;
		}
		int id = (int)textures.size();
		textures[t.get()] = id;
		out << "texture " << id << " " << rec << "\n";
		return id;
	}
	int material(const std::shared_ptr<Material>& m) {
		auto it = materials.find(m.get());
		if (it != materials.end()) return it->second;
		std::string rec;
		switch (m->kind) {
			case GS_MAT_LAMBERTIAN: rec = "lambertian " + std::to_string(texture(m->texture)); break;
			case GS_MAT_METAL: rec = "metal " + vec(m->albedo) + " " + num(m->param); break;
			case GS_MAT_DIELECTRIC: rec = "dielectric " + num(m->param); break;
			case GS_MAT_DIFFUSE_LIGHT: rec = "light " + std::to_string(texture(m->texture)); break;
			default: rec = "isotropic " + std::to_string(texture(m->texture)); break;
		}
		int id = (int)materials.size();
		materials[m.get()] = id;
		out << "material " << id << " " << rec << "\n";
		return id;
	}
	int node(const HittablePtr& h) {
		auto it = nodes.find(h.get());
		if (it != nodes.end()) return it->second;
		std::string rec;
		auto children = [&](const char* tag, const std::vector<HittablePtr>& v) {
			std::vector<int> ids;
			for (const auto& o : v) ids.push_back(node(o));
			rec = std::string(tag) + " " + std::to_string(ids.size());
			for (int i : ids) rec += " " + std::to_string(i);
		};
		switch (h->node_kind) {
			case NodeKind::Sphere: {
				const auto& s = static_cast<const Sphere&>(*h);
				int m = material(s.material);
				if (s.is_moving) rec = "msphere " + vec(s.center_start) + " " + vec(s.center_end) + " " + num(s.radius);
				else rec = "sphere " + vec(s.center_start) + " " + num(s.radius);
				rec += " " + std::to_string(m) + " " + std::to_string(s.prim_id);
				break;
			}
			case NodeKind::Quad: {
				const auto& q = static_cast<const Quad&>(*h);
				int m = material(q.material);
				rec = "quad " + vec(q.q) + " " + vec(q.u) + " " + vec(q.v) + " " + std::to_string(m) + " " + std::to_string(q.prim_id);
				break;
			}
			case NodeKind::Triangle: {
				const auto& t = static_cast<const Triangle&>(*h);
				int m = material(t.material);
				rec = "tri " + vec(t.a) + " " + vec(t.b) + " " + vec(t.c) + " " + std::to_string(m) + " " + std::to_string(t.prim_id);
				break;
			}
			case NodeKind::List: children("list", static_cast<const HittableList&>(*h).objects); break;
			case NodeKind::Bvh: {
				const auto& b = static_cast<const BVHNode&>(*h);
				if (b.source.empty()) throw std::runtime_error("dump_graph: inner BVH node reached (only from_list roots are graph nodes)");
				children("bvh", b.source);
				break;
			}
			case NodeKind::Translate: {
				const auto& t = static_cast<const Translate&>(*h);
				int c = node(t.object);
				rec = "translate " + std::to_string(c) + " " + vec(t.offset);
				break;
			}
			case NodeKind::RotateY: {
				const auto& r = static_cast<const RotateY&>(*h);
				int c = node(r.object);
				rec = "rotate_y " + std::to_string(c) + " " + num(r.angle);
				break;
			}
			case NodeKind::Medium: {
				const auto& m = static_cast<const ConstantMedium&>(*h);
				int c = node(m.boundary);
				int pm = material(m.phase_function);
				rec = "medium " + std::to_string(c) + " " + num(m.density) + " " + std::to_string(pm) + " " + std::to_string(m.medium_index);
				break;
			}
		}
		int id = (int)nodes.size();
		nodes[h.get()] = id;
		out << "node " << id << " " << rec << "\n";
		return id;
	}
};

}  // namespace

void dump_graph(const HostScene& scene, const std::string& path) {
	Writer w;
	w.out << "gsgraph 1\n";
	int root = w.node(scene.world);
	w.out << "root " << root << "\n";
	const Camera& c = scene.camera;
	w.out << "camera " << num(c.arg_aspect) << " " << c.image_width << " " << num(c.sample_settings.confidence) << " "
	      << num(c.sample_settings.tolerance) << " " << c.sample_settings.batch_size << " " << c.sample_settings.max_samples << " "
	      << c.max_depth << " " << num(c.arg_vfov) << " " << vec(c.arg_from) << " " << vec(c.arg_at) << " " << vec(c.arg_vup) << " "
	      << num(c.defocus_angle) << " " << num(c.arg_focus) << "\n";
	if (c.background.is_hdri) {
		int im = w.image(c.background.hdri.image);
		w.out << "background hdri " << im << " " << vec(c.background.hdri.rotation) << "\n";
	} else w.out << "background solid " << vec(c.background.color) << "\n";
	std::ofstream f(path);
	if (!f) throw std::runtime_error("cannot write " + path);
	f << w.out.str();
}

HostScene load_graph(const std::string& path, bool include_media) {
	std::ifstream f(path);
	if (!f) throw std::runtime_error("cannot open " + path);
	std::map<int, std::shared_ptr<ImageData>> images;
	std::map<int, std::shared_ptr<Texture>> textures;
	std::map<int, std::shared_ptr<Material>> materials;
	std::map<int, HittablePtr> nodes;  // null = dropped (medium with include_media == false)
	reset_scene_ids();
	auto V = [](std::istringstream& is) { Vec3 v; is >> v.x >> v.y >> v.z; return v; };
	struct { double aspect = 1, conf = 0.95, tol = 0, vfov = 40, defocus = 0, focus = 10; int width = 0; uint32_t batch = 1, max = 0, depth = 1; Vec3 from, at, vup; } cr;
	Background bg;
	int root = -1;
	bool have_header = false, have_camera = false;
	std::string line;
	while (std::getline(f, line)) {
		if (line.empty() || line[0] == '#') continue;
		std::istringstream is(line);
		std::string kw; is >> kw;
		if (kw == "gsgraph") { int v = 0; is >> v; if (v != 1) throw std::runtime_error(path + ": unsupported gsgraph version"); have_header = true; }
		else if (kw == "image") {
			int id, w, h; std::string ty, src; is >> id >> ty >> w >> h; std::getline(is, src);
			src.erase(0, src.find_first_not_of(' '));
			if (src.size() < 4) throw std::runtime_error(path + ": bad image source");
			std::string scheme = src.substr(0, 4), p = src.substr(4);
			std::shared_ptr<ImageData> im;
			if (scheme == "ppm:") im = load_image_u8(p);
			else if (scheme == "hdr:") im = load_hdr(p);
			else if (scheme == "raw:") im = load_raw(p, w, h, ty == "u8" ? GS_IMG_U8_RGB : GS_IMG_F32_RGB);
			else throw std::runtime_error(path + ": unknown image scheme " + scheme);
			if (im->width != w || im->height != h) throw std::runtime_error(path + ": image size mismatch for " + src);
			im->source = src;
			images[id] = im;
		} else if (kw == "texture") {
			int id; std::string ty; is >> id >> ty;
			if (ty == "solid") textures[id] = SolidColorTexture::new_(V(is));
			else if (ty == "checker") { double sc; int e, o; is >> sc >> e >> o; textures[id] = CheckeredTexture::new_(sc, textures.at(e), textures.at(o)); }
			else if (ty == "image") { int im; is >> im; textures[id] = ImageTexture::from_image(images.at(im)); }
			else if (ty == "noise") { double sc; is >> sc; textures[id] = NoiseTexture::new_(sc); }
			else throw std::runtime_error(path + ": texture type " + ty);
		} else if (kw == "material") {
			int id; std::string ty; is >> id >> ty;
			if (ty == "lambertian") { int t; is >> t; materials[id] = Lambertian::from_texture(textures.at(t)); }
			else if (ty == "metal") { Vec3 a = V(is); double fz; is >> fz; materials[id] = Metal::new_(a, fz); }
			else if (ty == "dielectric") { double ir; is >> ir; materials[id] = Dielectric::new_(ir); }
			else if (ty == "light") { int t; is >> t; materials[id] = DiffuseLight::new_(textures.at(t)); }
			else if (ty == "isotropic") { int t; is >> t; materials[id] = Isotropic::new_(textures.at(t)); }
			else throw std::runtime_error(path + ": material type " + ty);
		} else if (kw == "node") {
			int id; std::string ty; is >> id >> ty;
			auto kids = [&]() {
				size_t n; is >> n;
				std::vector<HittablePtr> v;
				for (size_t k = 0; k < n; ++k) { int c; is >> c; auto o = nodes.at(c); if (o) v.push_back(o); }
				return v;
			};
			if (ty == "sphere") { Vec3 c = V(is); double r; int m, pid; is >> r >> m >> pid; auto s = Sphere::new_stationary(c, r, materials.at(m)); s->prim_id = pid; nodes[id] = s; }
			else if (ty == "msphere") { Vec3 c1 = V(is), c2 = V(is); double r; int m, pid; is >> r >> m >> pid; auto s = Sphere::new_moving(c1, c2, r, materials.at(m)); s->prim_id = pid; nodes[id] = s; }
			else if (ty == "quad") { Vec3 q = V(is), u = V(is), v = V(is); int m, pid; is >> m >> pid; auto s = Quad::new_(q, u, v, materials.at(m)); s->prim_id = pid; nodes[id] = s; }
			else if (ty == "tri") { Vec3 a = V(is), b = V(is), c = V(is); int m, pid; is >> m >> pid; auto s = Triangle::new_(a, b, c, materials.at(m)); s->prim_id = pid; nodes[id] = s; }
			else if (ty == "list") { auto l = HittableList::new_(); for (auto& o : kids()) l->add(o); nodes[id] = l; }
			else if (ty == "bvh") { HittableList tmp; for (auto& o : kids()) tmp.add(o); nodes[id] = BVHNode::from_list(tmp); }
			else if (ty == "translate") { int c; is >> c; Vec3 off = V(is); auto o = nodes.at(c); nodes[id] = o ? HittablePtr(Translate::new_(o, off)) : nullptr; }
			else if (ty == "rotate_y") { int c; double deg; is >> c >> deg; auto o = nodes.at(c); nodes[id] = o ? HittablePtr(RotateY::new_(o, deg)) : nullptr; }
			else if (ty == "medium") {
				int c, m, mi; double dens; is >> c >> dens >> m >> mi;
				if (include_media) { auto md = ConstantMedium::new_(nodes.at(c), dens, materials.at(m)); md->medium_index = mi; nodes[id] = md; }
				else nodes[id] = nullptr;
			} else throw std::runtime_error(path + ": node type " + ty);
			if (!is) throw std::runtime_error(path + ": malformed record: " + line);
		} else if (kw == "root") is >> root;
		else if (kw == "camera") {
			is >> cr.aspect >> cr.width >> cr.conf >> cr.tol >> cr.batch >> cr.max >> cr.depth >> cr.vfov;
			cr.from = V(is); cr.at = V(is); cr.vup = V(is); is >> cr.defocus >> cr.focus;
			have_camera = true;
		} else if (kw == "background") {
			std::string ty; is >> ty;
			if (ty == "solid") bg = Background::SOLID(V(is));
			else { int im; is >> im; HDRI h; h.image = images.at(im); h.rotation = V(is); bg = Background::HDRI_(h); }
		} else throw std::runtime_error(path + ": unknown record: " + line);
	}
	if (!have_header || !have_camera || root < 0) throw std::runtime_error(path + ": incomplete graph");
	HostScene s;
	s.name = path;
	s.world = nodes.at(root);
	if (!s.world) throw std::runtime_error(path + ": root dropped");
	s.camera = Camera::new_(cr.aspect, cr.width, SampleSettings{cr.conf, cr.tol, cr.batch, cr.max}, cr.depth, cr.vfov, cr.from, cr.at, cr.vup,
	                        cr.defocus, cr.focus, bg);
	return s;
}

}  // namespace grayshift
