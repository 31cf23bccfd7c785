"""The Rust crate shell (rust/) cannot be compiled here (no Rust toolchain), so its FFI declarations are kept in
lock-step with include/grayshift.h mechanically: every function of the header is declared in rust/src/ffi.rs with the
same number of parameters, and every #[repr(C)] struct lists the fields of its ctypes twin (grayshift_b200/ffi.py, whose
layout tests/test_abi.py checks against the C compiler) in the same order with the corresponding types."""
import ctypes as C
import os
import re

from grayshift_b200 import ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = open(os.path.join(ROOT, "include", "grayshift.h")).read()
RUST = open(os.path.join(ROOT, "rust", "src", "ffi.rs")).read()


def _strip_comments(text):
    return re.sub(r"//[^\n]*", "", re.sub(r"/\*.*?\*/", "", text, flags=re.S))


def _params(arglist):
    arglist = arglist.strip()
    return 0 if arglist in ("", "void") else len([a for a in arglist.split(",") if a.strip()])


def header_functions():
    h = _strip_comments(HEADER)
    return {m.group(1): _params(m.group(2)) for m in re.finditer(r"\b(gs_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h)}


def rust_functions():
    r = _strip_comments(RUST)
    block = r[r.index('extern "C" {'):]
    return {m.group(1): _params(m.group(2)) for m in re.finditer(r"pub fn (gs_[a-z0-9_]+)\s*\(([^)]*)\)", block)}


def rust_structs():
    r = _strip_comments(RUST)
    out = {}
    for m in re.finditer(r"#\[repr\(C\)\][^{]*?pub struct (gs_[a-z_]+)\s*\{(.*?)\}", r, flags=re.S):
        out[m.group(1)] = [(f.group(1), f.group(2).strip()) for f in re.finditer(r"pub (\w+)\s*:\s*([^,\n]+),", m.group(2))]
    return out


# (c_uint64 and c_size_t are one ctypes type on LP64: either Rust spelling has the right size; the names are checked too)
RUST_TYPE = {C.c_double: {"f64"}, C.c_double * 2: {"[f64; 2]"}, C.c_int32: {"i32"}, C.c_uint32: {"u32"}, C.c_uint64: {"u64", "usize"}, ffi.d3: {"[f64; 3]"}}


def test_every_header_function_is_declared_in_the_rust_ffi_with_the_same_arity():
    h, r = header_functions(), rust_functions()
    assert set(h) == set(ffi.CUDA_SYMBOLS), "header and ctypes table differ"
    assert set(h) == set(r), f"missing in ffi.rs: {sorted(set(h) - set(r))}; unknown: {sorted(set(r) - set(h))}"
    for name, n in h.items():
        assert r[name] == n, f"{name}: header takes {n} parameters, ffi.rs declares {r[name]}"
        assert len(ffi.CUDA_SYMBOLS[name][1]) == n, f"{name}: ctypes argtypes out of step"


def test_rust_struct_fields_follow_the_c_layout():
    rs = rust_structs()
    for name in ("gs_sphere", "gs_quad", "gs_triangle", "gs_bvh_node", "gs_medium", "gs_material", "gs_texture", "gs_image",
                 "gs_scene_desc", "gs_camera", "gs_sample_settings", "gs_stats"):
        ct = getattr(ffi, name)
        assert name in rs, f"ffi.rs lacks #[repr(C)] struct {name}"
        fields = rs[name]
        assert [f for f, _ in fields] == [f for f, _ in ct._fields_], f"{name}: field names / order differ"
        for (fname, rtype), (_, ctype) in zip(fields, ct._fields_):
            if ctype in RUST_TYPE:
                assert rtype in RUST_TYPE[ctype], f"{name}.{fname}: {rtype} vs {ctype.__name__}"
            else:  # pointers
                assert rtype.startswith("*const") or rtype.startswith("*mut"), f"{name}.{fname}: expected a pointer, got {rtype}"


def test_rust_constants_match_the_header():
    h = _strip_comments(HEADER)
    for m in re.finditer(r"pub const (GS_[A-Z0-9_]+): \w+ = (\d+);", RUST):
        name, value = m.group(1), int(m.group(2))
        hm = re.search(rf"\b{name}\b\s*=?\s*(\d+)", h)
        assert hm is not None, f"{name} is not in the header"
        assert int(hm.group(1)) == value, f"{name}: {value} in ffi.rs, {hm.group(1)} in the header"


def test_build_rs_compiles_the_same_translation_units_as_the_makefile():
    mk = open(os.path.join(ROOT, "grayshift_b200", "csrc", "Makefile")).read()
    rs = open(os.path.join(ROOT, "rust", "build.rs")).read()
    units = sorted(set(re.findall(r"(\w+\.cu)\b", mk)))
    assert units == sorted(set(re.findall(r'"(\w+\.cu)"', rs))), "build.rs and the Makefile list different .cu files"
    assert "compute_100a,code=sm_100a" in rs and "-fmad=false" in rs


def test_rust_scene_builder_covers_every_primitive_kind_of_the_abi():
    """the flattener twin pushes spheres, quads and triangles with the reference kinds 0 / 1 / 2 of GS_PRIM_REF and hands
    every array of gs_scene_desc to the library"""
    sb = open(os.path.join(ROOT, "rust", "src", "scene_builder.rs")).read()
    for fn in ("push_sphere", "push_quad", "push_triangle", "begin_medium", "end_medium", "finish", "desc"):
        assert re.search(rf"pub fn {fn}\b", sb), f"scene_builder.rs lacks {fn}"
    assert "(1u32 << 30)" in sb and "(2u32 << 30)" in sb
    desc_fields = [f for f, _ in ffi.gs_scene_desc._fields_]
    body = sb[sb.index("gs_scene_desc {"):]
    body = body[:body.index("}\n    }")]
    for f in desc_fields:
        assert re.search(rf"\b{f}:", body), f"FlatScene::desc() does not set gs_scene_desc.{f}"


# the scene-building API main.rs compiles against: module path -> constructors / methods that must exist
RUST_API = {
    "camera.rs": ["pub fn new(", "pub fn render(&self, world: Box<dyn Hittable>, image_file: &mut File) -> std::io::Result<()>", "pub struct SampleSettings",
                  "pub enum Background", "pub struct HDRI", "pub image: radiant::Image", "pub rotation: Vec3"],
    "hittable/hittable.rs": ["pub trait Hittable: Send + Sync", "fn flatten(&self, xf: &Xform, out: &mut SceneBuilder)", "pub fn add(&mut self, object: Box<dyn Hittable>)",
                             "pub fn clear(&mut self)", "impl Hittable for HittableList", "impl Hittable for Translate", "impl Hittable for RotateY",
                             "pub fn new(object: Box<dyn Hittable>, offset: Vec3)", "pub fn new(object: Box<dyn Hittable>, angle: f64)"],
    "hittable/BVH.rs": ["pub fn from_list(list: HittableList) -> Box<Self>", "pub fn construct_tree(", "impl Hittable for BVHNode"],
    "hittable/sphere.rs": ["pub fn new_stationary(center: Vec3, radius: f64, material: Arc<dyn Material>)", "pub fn new_moving(center1: Vec3, center2: Vec3, radius: f64, material: Arc<dyn Material>)"],
    "hittable/quad.rs": ["pub fn new(q: Vec3, u: Vec3, v: Vec3, material: Arc<dyn Material>)", "pub fn cube(point_a: Vec3, point_b: Vec3, material: Arc<dyn Material>) -> HittableList"],
    "hittable/triangle.rs": ["pub fn new(a: Vec3, b: Vec3, c: Vec3, material: Arc<dyn Material>)"],
    "hittable/volume.rs": ["pub fn new(boundary: Box<dyn Hittable>, density: f64, phase_function: Arc<dyn Material>)", "pub fn from_isotropic_color(boundary: Box<dyn Hittable>, density: f64, color: Vec3)"],
    "material.rs": ["pub trait Material: Send + Sync", "pub fn from_color(albedo: Vec3)", "pub fn from_texture(texture: Arc<dyn Texture>)", "pub fn new(albedo: Vec3, fuzz: f64)",
                    "pub fn new(refraction_index: f64)", "pub fn reflectance(cosine: f64, refraction_index: f64) -> f64", "impl Material for DiffuseLight", "impl Material for Isotropic"],
    "texture.rs": ["pub trait Texture: Send + Sync", "pub fn from_rgb(red: f64, green: f64, blue: f64)", "pub fn from_colors(scale: f64, even_color: Vec3, odd_color: Vec3)",
                   "pub fn new<P: AsRef<Path>>(filepath: P) -> Result<Self, ImageError>", "pub fn new(scale: f64)"],
    "util/vec3.rs": ["pub const ZERO: Vec3", "pub const fn new(x: f64, y: f64, z: f64)", "impl Mul<Vec3> for Vec3"],
    "util/util.rs": ["pub fn random_f64(min: f64, max: f64) -> f64", "pub fn random_vector(min: f64, max: f64) -> Vec3"],
    "color.rs": ["pub fn write_color", "pub fn luminance"],
    "AABB.rs": ["pub fn from_corners(a: Vec3, b: Vec3)", "pub fn longest_axis(&self) -> usize"],
    "ray.rs": ["pub struct Ray"], "ONB.rs": ["pub struct OrthonormalBasis"], "hittable/mod.rs": ["pub mod BVH;", "pub mod volume;"], "util/mod.rs": ["pub mod vec3;"],
}


def test_rust_shell_provides_the_scene_building_api_of_main_rs():
    """every module main.rs declares (mod color; mod ray; mod hittable; mod camera; mod util; mod material; mod AABB; mod texture;
    mod ONB;) exists under rust/src with the constructors and signatures main.rs:61-888 calls, so that dropping the reference's
    main.rs next to them is the only step left (uncompiled here: no Rust toolchain)"""
    src = os.path.join(ROOT, "rust", "src")
    for rel, needles in RUST_API.items():
        path = os.path.join(src, rel)
        assert os.path.exists(path), f"rust/src/{rel} is missing"
        text = re.sub(r"\s+", " ", open(path).read())
        for n in needles:
            assert re.sub(r"\s+", " ", n) in text, f"rust/src/{rel} lacks `{n}`"
    # every hittable / material / texture type flattens itself
    for rel, n in (("hittable/sphere.rs", 1), ("hittable/quad.rs", 1), ("hittable/triangle.rs", 1), ("hittable/volume.rs", 1), ("hittable/BVH.rs", 1),
                   ("hittable/hittable.rs", 4), ("material.rs", 6), ("texture.rs", 5)):
        assert open(os.path.join(src, rel)).read().count("fn flatten(") >= n, f"rust/src/{rel}: flatten() impls missing"
    cam = open(os.path.join(src, "camera.rs")).read()
    for call in ("world.flatten(", "builder.finish(", "ffi::Scene::create(", "render_multi(", "write_color("):
        assert call in cam, f"Camera::render lacks {call}"
