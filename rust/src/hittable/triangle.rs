//! Triangle (reference: hittable/triangle.rs:10-31).  The library applies the ray interval to triangles, which the reference's
//! Triangle::hit forgets (include/grayshift.h): the one documented difference.
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::hittable::hittable::Hittable;
use crate::material::Material;
use crate::util::vec3::Vec3;
use crate::AABB::AABB;
use std::sync::Arc;

pub struct Triangle {
    a: Vec3,
    b: Vec3,
    c: Vec3,
    material: Arc<dyn Material>,
    bbox: AABB,
}

impl Triangle {
    pub fn new(a: Vec3, b: Vec3, c: Vec3, material: Arc<dyn Material>) -> Self {
        let lo = Vec3::new(a.x.min(b.x).min(c.x), a.y.min(b.y).min(c.y), a.z.min(b.z).min(c.z));
        let hi = Vec3::new(a.x.max(b.x).max(c.x), a.y.max(b.y).max(c.y), a.z.max(b.z).max(c.z));
        Triangle { a, b, c, material, bbox: AABB::from_corners(lo, hi) }
    }
}

impl Hittable for Triangle {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        let row = self.material.flatten(out);
        out.push_triangle(xf, self.a.to_array(), self.b.to_array(), self.c.to_array(), row);
    }
}
