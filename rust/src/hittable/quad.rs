//! Quad and Quad::cube (reference: hittable/quad.rs:12-80).  Normal, w and d are derived by the library exactly as
//! quad.rs:32-36 / plane.rs:15-18 do; the box is the hull of the two padded diagonals (quad.rs:27-30).
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::hittable::hittable::{Hittable, HittableList};
use crate::material::Material;
use crate::util::vec3::Vec3;
use crate::AABB::AABB;
use std::sync::Arc;

pub struct Quad {
    q: Vec3,
    u: Vec3,
    v: Vec3,
    material: Arc<dyn Material>,
    bbox: AABB,
}

impl Quad {
    pub fn new(q: Vec3, u: Vec3, v: Vec3, material: Arc<dyn Material>) -> Self {
        let bbox = AABB::from_AABB_pair(AABB::from_corners(q, q + u + v), AABB::from_corners(q + u, q + v));
        Quad { q, u, v, material, bbox }
    }

    /// the six faces of the box spanned by two opposite corners, in the reference's order: front, right, back, left, top,
    /// bottom (quad.rs:54-80) -- the order matters for the tie ranks of touching faces
    pub fn cube(point_a: Vec3, point_b: Vec3, material: Arc<dyn Material>) -> HittableList {
        let lo = Vec3::new(point_a.x.min(point_b.x), point_a.y.min(point_b.y), point_a.z.min(point_b.z));
        let hi = Vec3::new(point_a.x.max(point_b.x), point_a.y.max(point_b.y), point_a.z.max(point_b.z));
        let (dx, dy, dz) = (Vec3::new(hi.x - lo.x, 0.0, 0.0), Vec3::new(0.0, hi.y - lo.y, 0.0), Vec3::new(0.0, 0.0, hi.z - lo.z));
        let mut sides = HittableList::new();
        let faces = [
            (Vec3::new(lo.x, lo.y, hi.z), dx, dy),
            (Vec3::new(hi.x, lo.y, hi.z), -dz, dy),
            (Vec3::new(hi.x, lo.y, lo.z), -dx, dy),
            (Vec3::new(lo.x, lo.y, lo.z), dz, dy),
            (Vec3::new(lo.x, hi.y, hi.z), dx, -dz),
            (Vec3::new(lo.x, lo.y, lo.z), dx, dz),
        ];
        for (q, u, v) in faces {
            sides.add(Box::new(Quad::new(q, u, v, material.clone())));
        }
        sides
    }
}

impl Hittable for Quad {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        let row = self.material.flatten(out);
        out.push_quad(xf, self.q.to_array(), self.u.to_array(), self.v.to_array(), row);
    }
}
