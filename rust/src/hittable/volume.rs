//! ConstantMedium (reference: hittable/volume.rs:10-28).  On the device a medium is not a BVH leaf: its boundary goes to the
//! boundary arrays, its free-flight draw is taken once per ray before the surface traversal.
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::hittable::hittable::Hittable;
use crate::material::{Isotropic, Material};
use crate::util::vec3::Vec3;
use crate::AABB::AABB;
use std::sync::Arc;

pub struct ConstantMedium {
    boundary: Box<dyn Hittable>,
    density: f64,
    phase_function: Arc<dyn Material>,
}

impl ConstantMedium {
    pub fn new(boundary: Box<dyn Hittable>, density: f64, phase_function: Arc<dyn Material>) -> Self {
        ConstantMedium { boundary, density, phase_function }
    }
    pub fn from_isotropic_color(boundary: Box<dyn Hittable>, density: f64, color: Vec3) -> Self {
        ConstantMedium::new(boundary, density, Arc::new(Isotropic::from_color(color)))
    }
}

impl Hittable for ConstantMedium {
    fn bounding_box(&self) -> AABB {
        self.boundary.bounding_box()
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        // the medium's fixed normal (1,0,0) (volume.rs:57) would be rotated by an enclosing RotateY: not representable
        assert!(xf.c == 1.0 && xf.s == 0.0, "ConstantMedium under RotateY is not supported (rotate the boundary instead)");
        let phase = self.phase_function.flatten(out);
        out.begin_medium(xf);
        self.boundary.flatten(xf, out); // lands in the boundary arrays while the medium is open
        out.end_medium(self.density, phase);
    }
}
