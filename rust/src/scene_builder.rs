//! rust/src/scene_builder.rs -- the flattener behind the new `flatten()` trait methods.
//! Rust twin of grayshift_b200/host/flatten.cpp.  NOT COMPILED in this repository's build environment.
//!
//! Usage from the reference's types (patch points, see rust/README.md):
//!   impl Hittable for Sphere       { fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) { out.push_sphere(xf, ...) } }
//!   impl Hittable for Quad         { fn flatten(..) { out.push_quad(xf, self.q, self.u, self.v, &self.material) } }
//!   impl Hittable for Triangle     { fn flatten(..) { out.push_triangle(xf, self.a, self.b, self.c, &self.material) } }
//!   impl Hittable for HittableList { fn flatten(..) { for o in &self.objects { o.flatten(xf, out) } } }
//!   impl Hittable for BVHNode      { fn flatten(..) { self.left.flatten(xf, out); if let Some(r) = &self.right { r.flatten(xf, out) } } }
//!   impl Hittable for Translate    { fn flatten(..) { self.object.flatten(&xf.translated(self.offset), out) } }
//!   impl Hittable for RotateY      { fn flatten(..) { self.object.flatten(&xf.rotated_y(self.sin_theta, self.cos_theta), out) } }
//!   impl Hittable for ConstantMedium { fn flatten(..) { out.begin_medium(xf); self.boundary.flatten(xf, out);
//!                                                        out.end_medium(1.0 / -self.density_neg_inv, &self.phase_function) } }
use crate::camera::ffi::*;
use std::collections::HashMap;

/// p -> R p + t with R a rotation about y: R = [[c,0,s],[0,1,0],[-s,0,c]] (hittable.rs:195-199).
#[derive(Clone, Copy)]
pub struct Xform {
    pub c: f64,
    pub s: f64,
    pub t: [f64; 3],
}

impl Xform {
    pub fn identity() -> Self {
        Xform { c: 1.0, s: 0.0, t: [0.0; 3] }
    }
    pub fn rot(&self, p: [f64; 3]) -> [f64; 3] {
        [self.c * p[0] + self.s * p[2], p[1], -self.s * p[0] + self.c * p[2]]
    }
    pub fn point(&self, p: [f64; 3]) -> [f64; 3] {
        let r = self.rot(p);
        [r[0] + self.t[0], r[1] + self.t[1], r[2] + self.t[2]]
    }
    /// Translate::new(object, offset) inside this transform (hittable.rs:100-103).
    pub fn translated(&self, offset: [f64; 3]) -> Xform {
        let o = self.rot(offset);
        Xform { c: self.c, s: self.s, t: [self.t[0] + o[0], self.t[1] + o[1], self.t[2] + o[2]] }
    }
    /// RotateY::new(object, angle) inside this transform: angles add.
    pub fn rotated_y(&self, sin_theta: f64, cos_theta: f64) -> Xform {
        Xform { c: self.c * cos_theta - self.s * sin_theta, s: self.s * cos_theta + self.c * sin_theta, t: self.t }
    }
    fn is_pure_translation(&self) -> bool {
        self.c == 1.0 && self.s == 0.0
    }
}

pub enum ImagePixels {
    U8(Vec<u8>),
    F32(Vec<f32>),
}

#[derive(Default)]
pub struct SceneBuilder {
    pub spheres: Vec<gs_sphere>,
    pub quads: Vec<gs_quad>,
    pub triangles: Vec<gs_triangle>,
    pub boundary_spheres: Vec<gs_sphere>,
    pub boundary_quads: Vec<gs_quad>,
    pub media: Vec<gs_medium>,
    pub materials: Vec<gs_material>,
    pub textures: Vec<gs_texture>,
    pub images: Vec<(i32, i32, ImagePixels)>,
    material_rows: HashMap<usize, i32>, // Arc::as_ptr -> row: one row per shared Arc (main.rs:741-750)
    texture_rows: HashMap<usize, i32>,
    image_rows: HashMap<usize, i32>,
    bounds: Vec<[f64; 6]>,
    refs: Vec<u32>,
    next_order: i32,
    next_prim_id: i32,
    open_medium: Option<(usize, usize, usize)>, // (sphere_first, quad_first, medium_index) while a boundary is being flattened
    has_noise: bool,
}

impl SceneBuilder {
    pub fn new() -> Self {
        Self::default()
    }

    /// Rows are keyed by the address of the Arc's pointee; `make` is only called for a new row.
    pub fn material_row(&mut self, key: usize, make: impl FnOnce(&mut SceneBuilder) -> gs_material) -> i32 {
        if let Some(&r) = self.material_rows.get(&key) {
            return r;
        }
        let m = make(self);
        self.materials.push(m);
        let r = self.materials.len() as i32 - 1;
        self.material_rows.insert(key, r);
        r
    }
    /// Children of a CheckeredTexture must be pushed (by `make`) before the parent: gs_scene_create checks it.
    pub fn texture_row(&mut self, key: usize, make: impl FnOnce(&mut SceneBuilder) -> gs_texture) -> i32 {
        if let Some(&r) = self.texture_rows.get(&key) {
            return r;
        }
        let t = make(self);
        if t.kind == GS_TEX_NOISE {
            self.has_noise = true;
        }
        self.textures.push(t);
        let r = self.textures.len() as i32 - 1;
        self.texture_rows.insert(key, r);
        r
    }
    pub fn image_row(&mut self, key: usize, width: i32, height: i32, pixels: impl FnOnce() -> ImagePixels) -> i32 {
        if let Some(&r) = self.image_rows.get(&key) {
            return r;
        }
        self.images.push((width, height, pixels()));
        let r = self.images.len() as i32 - 1;
        self.image_rows.insert(key, r);
        r
    }

    pub fn push_sphere(&mut self, xf: &Xform, center: [f64; 3], path: [f64; 3], moving: bool, radius: f64, material: i32) {
        let c = xf.point(center);
        let p = xf.rot(path);
        // uv_rotation: Sphere::hit takes (u, v) from the object-space normal, before RotateY::hit rotates it (hittable.rs:179-211)
        let mut g = gs_sphere { center: c, radius, path: p, material, prim_id: self.next_prim_id, moving: moving as i32, order: 0,
                                uv_rotation: [xf.c, xf.s] };
        self.next_prim_id += 1;
        if self.open_medium.is_some() {
            g.material = 0; // a boundary's own material is never used
            self.boundary_spheres.push(g);
            return;
        }
        g.order = self.next_order;
        self.next_order += 1;
        let r = radius.abs();
        let mut b = [0.0; 6];
        for a in 0..3 {
            let c1 = c[a] + if moving { p[a] } else { 0.0 };
            b[a] = c[a].min(c1) - r;
            b[a + 3] = c[a].max(c1) + r;
        }
        self.bounds.push(b);
        self.refs.push((self.spheres.len() as u32) & 0x3fff_ffff);
        self.spheres.push(g);
    }

    pub fn push_quad(&mut self, xf: &Xform, q: [f64; 3], u: [f64; 3], v: [f64; 3], material: i32) {
        let (q, u, v) = (xf.point(q), xf.rot(u), xf.rot(v));
        let mut g = gs_quad { q, u, v, material, prim_id: self.next_prim_id, order: 0, reserved: 0 };
        self.next_prim_id += 1;
        if self.open_medium.is_some() {
            g.material = 0;
            self.boundary_quads.push(g);
            return;
        }
        g.order = self.next_order;
        self.next_order += 1;
        let mut b = [0.0; 6];
        for a in 0..3 {
            let p = [q[a], q[a] + u[a], q[a] + v[a], q[a] + u[a] + v[a]];
            b[a] = p.iter().cloned().fold(f64::INFINITY, f64::min);
            b[a + 3] = p.iter().cloned().fold(f64::NEG_INFINITY, f64::max);
        }
        self.bounds.push(b);
        self.refs.push((1u32 << 30) | (self.quads.len() as u32));
        self.quads.push(g);
    }

    /// Triangle::new(a, b, c, material) (triangle.rs:20-31): the three vertices go through the instance transform.
    pub fn push_triangle(&mut self, xf: &Xform, a: [f64; 3], b: [f64; 3], c: [f64; 3], material: i32) {
        assert!(self.open_medium.is_none(), "a triangle cannot bound a ConstantMedium (not convex)");
        let (a, b, c) = (xf.point(a), xf.point(b), xf.point(c));
        let g = gs_triangle { a, b, c, material, prim_id: self.next_prim_id, order: self.next_order, reserved: 0 };
        self.next_prim_id += 1;
        self.next_order += 1;
        let mut bb = [0.0; 6];
        for k in 0..3 {
            bb[k] = a[k].min(b[k]).min(c[k]);
            bb[k + 3] = a[k].max(b[k]).max(c[k]);
        }
        self.bounds.push(bb);
        self.refs.push((2u32 << 30) | (self.triangles.len() as u32));
        self.triangles.push(g);
    }

    /// ConstantMedium (volume.rs:10-28): its boundary primitives go to the boundary arrays.
    pub fn begin_medium(&mut self, xf: &Xform) {
        // the medium's fixed normal (1,0,0) (volume.rs:57) would be rotated by an enclosing RotateY
        assert!(xf.is_pure_translation(), "ConstantMedium under RotateY is not supported");
        assert!(self.open_medium.is_none(), "a ConstantMedium cannot bound another medium");
        self.open_medium = Some((self.boundary_spheres.len(), self.boundary_quads.len(), self.media.len()));
    }
    pub fn end_medium(&mut self, density: f64, phase_material: i32) {
        let (s0, q0, index) = self.open_medium.take().expect("end_medium without begin_medium");
        self.media.push(gs_medium {
            sphere_first: s0 as i32,
            sphere_count: (self.boundary_spheres.len() - s0) as i32,
            quad_first: q0 as i32,
            quad_count: (self.boundary_quads.len() - q0) as i32,
            density,
            phase_material,
            medium_index: index as i32,
        });
    }

    /// Builds the binned-SAH BVH and freezes the tables.
    pub fn finish(self, perlin_perm: [u8; 256]) -> FlatScene {
        assert!(!self.refs.is_empty(), "the world holds no surface primitive");
        let (nodes, order) = build_bvh_sah(&self.bounds, 1); // single-primitive leaves: measured +8 % on the final scene
        let prim_refs = order.iter().map(|&k| self.refs[k as usize]).collect();
        FlatScene {
            spheres: self.spheres,
            quads: self.quads,
            triangles: self.triangles,
            boundary_spheres: self.boundary_spheres,
            boundary_quads: self.boundary_quads,
            media: self.media,
            materials: self.materials,
            textures: self.textures,
            images: self.images,
            nodes,
            prim_refs,
            perlin_perm,
            has_noise: self.has_noise,
            image_descs: Vec::new(),
        }
    }
}

/// Owns every array a `gs_scene_desc` points at.
pub struct FlatScene {
    pub spheres: Vec<gs_sphere>,
    pub quads: Vec<gs_quad>,
    pub triangles: Vec<gs_triangle>,
    pub boundary_spheres: Vec<gs_sphere>,
    pub boundary_quads: Vec<gs_quad>,
    pub media: Vec<gs_medium>,
    pub materials: Vec<gs_material>,
    pub textures: Vec<gs_texture>,
    pub images: Vec<(i32, i32, ImagePixels)>,
    pub nodes: Vec<gs_bvh_node>,
    pub prim_refs: Vec<u32>,
    pub perlin_perm: [u8; 256],
    pub has_noise: bool,
    image_descs: Vec<gs_image>,
}

impl FlatScene {
    pub fn desc(&mut self) -> gs_scene_desc {
        self.image_descs = self
            .images
            .iter()
            .map(|(w, h, px)| match px {
                ImagePixels::U8(v) => gs_image { width: *w, height: *h, format: GS_IMG_U8_RGB, reserved: 0, pixels: v.as_ptr() as *const _ },
                ImagePixels::F32(v) => gs_image { width: *w, height: *h, format: GS_IMG_F32_RGB, reserved: 0, pixels: v.as_ptr() as *const _ },
            })
            .collect();
        gs_scene_desc {
            abi_version: GS_ABI_VERSION,
            reserved: 0,
            spheres: self.spheres.as_ptr(),
            n_spheres: self.spheres.len(),
            quads: self.quads.as_ptr(),
            n_quads: self.quads.len(),
            nodes: self.nodes.as_ptr(),
            n_nodes: self.nodes.len(),
            prim_refs: self.prim_refs.as_ptr(),
            n_prim_refs: self.prim_refs.len(),
            boundary_spheres: self.boundary_spheres.as_ptr(),
            n_boundary_spheres: self.boundary_spheres.len(),
            boundary_quads: self.boundary_quads.as_ptr(),
            n_boundary_quads: self.boundary_quads.len(),
            media: self.media.as_ptr(),
            n_media: self.media.len(),
            materials: self.materials.as_ptr(),
            n_materials: self.materials.len(),
            textures: self.textures.as_ptr(),
            n_textures: self.textures.len(),
            images: self.image_descs.as_ptr(),
            n_images: self.image_descs.len(),
            perlin_perm: if self.has_noise { self.perlin_perm.as_ptr() } else { std::ptr::null() },
            triangles: if self.triangles.is_empty() { std::ptr::null() } else { self.triangles.as_ptr() },
            n_triangles: self.triangles.len(),
        }
    }
}

fn area(mn: &[f64; 3], mx: &[f64; 3]) -> f64 {
    let (dx, dy, dz) = (mx[0] - mn[0], mx[1] - mn[1], mx[2] - mn[2]);
    if dx < 0.0 { 0.0 } else { 2.0 * (dx * dy + dy * dz + dz * dx) }
}

/// Binned SAH (16 bins), leaves of <= max_leaf primitives, nodes in pre-order (children after their parent),
/// depth bounded for the device's 64-entry traversal stack.  Same algorithm as build_bvh_sah in flatten.cpp.
pub fn build_bvh_sah(bounds: &[[f64; 6]], max_leaf: usize) -> (Vec<gs_bvh_node>, Vec<u32>) {
    const BINS: usize = 16;
    let n = bounds.len();
    let mut order: Vec<u32> = (0..n as u32).collect();
    let cent: Vec<[f64; 3]> = bounds.iter().map(|b| [0.5 * (b[0] + b[3]), 0.5 * (b[1] + b[4]), 0.5 * (b[2] + b[5])]).collect();
    let mut nodes = vec![gs_bvh_node::default()];
    let mut stack = vec![(0usize, n, 0usize, 0u32)];
    while let Some((lo, hi, node, depth)) = stack.pop() {
        let cnt = hi - lo;
        let (mut mn, mut mx) = ([f64::MAX; 3], [f64::MIN; 3]);
        let (mut cmn, mut cmx) = ([f64::MAX; 3], [f64::MIN; 3]);
        for &k in &order[lo..hi] {
            let (b, c) = (&bounds[k as usize], &cent[k as usize]);
            for a in 0..3 {
                mn[a] = mn[a].min(b[a]);
                mx[a] = mx[a].max(b[a + 3]);
                cmn[a] = cmn[a].min(c[a]);
                cmx[a] = cmx[a].max(c[a]);
            }
        }
        let mut nd = gs_bvh_node { bounds_min: mn, bounds_max: mx, ..Default::default() };
        if cnt == 1 {
            nd.first_ref = lo as i32;
            nd.ref_count = 1;
            nodes[node] = nd;
            continue;
        }
        let mut mid = lo;
        let mut found = false;
        if depth < 32 {
            let mut best = (f64::MAX, usize::MAX, 0usize);
            for axis in 0..3 {
                let ext = cmx[axis] - cmn[axis];
                if !(ext > 0.0) {
                    continue;
                }
                let scale = BINS as f64 / ext;
                let mut bmn = [[f64::MAX; 3]; BINS];
                let mut bmx = [[f64::MIN; 3]; BINS];
                let mut bc = [0usize; BINS];
                for &k in &order[lo..hi] {
                    let bi = (((cent[k as usize][axis] - cmn[axis]) * scale) as usize).min(BINS - 1);
                    let b = &bounds[k as usize];
                    for a in 0..3 {
                        bmn[bi][a] = bmn[bi][a].min(b[a]);
                        bmx[bi][a] = bmx[bi][a].max(b[a + 3]);
                    }
                    bc[bi] += 1;
                }
                let (mut ra, mut rc) = ([0.0; BINS], [0usize; BINS]);
                let (mut amn, mut amx, mut c) = ([f64::MAX; 3], [f64::MIN; 3], 0usize);
                for i in (1..BINS).rev() {
                    for a in 0..3 {
                        amn[a] = amn[a].min(bmn[i][a]);
                        amx[a] = amx[a].max(bmx[i][a]);
                    }
                    c += bc[i];
                    ra[i] = area(&amn, &amx);
                    rc[i] = c;
                }
                let (mut lmn, mut lmx, mut lc) = ([f64::MAX; 3], [f64::MIN; 3], 0usize);
                for i in 1..BINS {
                    for a in 0..3 {
                        lmn[a] = lmn[a].min(bmn[i - 1][a]);
                        lmx[a] = lmx[a].max(bmx[i - 1][a]);
                    }
                    lc += bc[i - 1];
                    if lc == 0 || rc[i] == 0 {
                        continue;
                    }
                    let cost = area(&lmn, &lmx) * lc as f64 + ra[i] * rc[i] as f64;
                    if cost < best.0 {
                        best = (cost, axis, i);
                    }
                }
            }
            if best.1 != usize::MAX {
                let box_area = area(&mn, &mx);
                if cnt > max_leaf || box_area + best.0 < cnt as f64 * box_area {
                    let (axis, bin) = (best.1, best.2);
                    let scale = BINS as f64 / (cmx[axis] - cmn[axis]);
                    let slice = &mut order[lo..hi];
                    let mut i = 0;
                    for j in 0..slice.len() {
                        let bi = (((cent[slice[j] as usize][axis] - cmn[axis]) * scale) as usize).min(BINS - 1);
                        if bi < bin {
                            slice.swap(i, j);
                            i += 1;
                        }
                    }
                    mid = lo + i;
                    found = mid > lo && mid < hi;
                }
            }
        }
        if !found {
            if cnt <= max_leaf {
                nd.first_ref = lo as i32;
                nd.ref_count = cnt as i32;
                nodes[node] = nd;
                continue;
            }
            let e = [cmx[0] - cmn[0], cmx[1] - cmn[1], cmx[2] - cmn[2]];
            let axis = if e[1] > e[0] && e[1] >= e[2] { 1 } else if e[2] > e[0] && e[2] > e[1] { 2 } else { 0 };
            mid = lo + cnt / 2;
            order[lo..hi].select_nth_unstable_by(cnt / 2, |p, q| cent[*p as usize][axis].partial_cmp(&cent[*q as usize][axis]).unwrap());
        }
        nd.left = nodes.len() as i32;
        nodes.push(gs_bvh_node::default());
        nd.right = nodes.len() as i32;
        nodes.push(gs_bvh_node::default());
        nodes[node] = nd;
        stack.push((mid, hi, nd.right as usize, depth + 1));
        stack.push((lo, mid, nd.left as usize, depth + 1));
    }
    (nodes, order)
}
