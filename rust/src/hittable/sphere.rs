//! Sphere (reference: hittable/sphere.rs:11-61): stationary or linearly moving.  The radius keeps its sign (a negative radius
//! turns the normal inward, sphere.rs:91); the box of a moving sphere covers both end positions (sphere.rs:36-39).
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::hittable::hittable::Hittable;
use crate::material::Material;
use crate::util::vec3::Vec3;
use crate::AABB::AABB;
use std::sync::Arc;

pub struct Sphere {
    center_start: Vec3,
    center_path: Vec3,
    is_moving: bool,
    radius: f64,
    material: Arc<dyn Material>,
    bbox: AABB,
}

impl Sphere {
    pub fn new_stationary(center: Vec3, radius: f64, material: Arc<dyn Material>) -> Self {
        let r = Vec3::new(radius, radius, radius);
        Sphere { center_start: center, center_path: Vec3::ZERO, is_moving: false, radius, material, bbox: AABB::from_corners(center - r, center + r) }
    }
    pub fn new_moving(center1: Vec3, center2: Vec3, radius: f64, material: Arc<dyn Material>) -> Self {
        let r = Vec3::new(radius, radius, radius);
        let bbox = AABB::from_AABB_pair(AABB::from_corners(center1 - r, center1 + r), AABB::from_corners(center2 - r, center2 + r));
        Sphere { center_start: center1, center_path: center2 - center1, is_moving: true, radius, material, bbox }
    }
}

impl Hittable for Sphere {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        let row = self.material.flatten(out);
        out.push_sphere(xf, self.center_start.to_array(), self.center_path.to_array(), self.is_moving, self.radius, row);
    }
}
