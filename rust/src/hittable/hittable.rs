//! Hittable trait, HittableList, Translate, RotateY (reference: hittable/hittable.rs:10-216).
//! The trait trades `hit` for `flatten`: every node appends itself to a SceneBuilder under the rigid transform accumulated on
//! the way down (Translate / RotateY are baked into world space: t is preserved because the reference never renormalises
//! transformed directions, hittable.rs:107-120, 179-211).
use crate::camera::scene_builder::{SceneBuilder, Xform};
use crate::util::util::deg_to_rad;
use crate::util::vec3::Vec3;
use crate::AABB::AABB;

pub trait Hittable: Send + Sync {
    fn bounding_box(&self) -> AABB;
    /// Append this object's primitives to `out`, in the order the reference's `hit` would visit them.
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder);
}

pub struct HittableList {
    pub objects: Vec<Box<dyn Hittable>>,
    bbox: AABB,
}

impl HittableList {
    pub fn new() -> Self {
        HittableList { objects: Vec::new(), bbox: AABB::EMPTY }
    }
    pub fn clear(&mut self) {
        self.objects.clear();
        self.bbox = AABB::EMPTY;
    }
    pub fn add(&mut self, object: Box<dyn Hittable>) {
        self.bbox = AABB::from_AABB_pair(self.bbox, object.bounding_box());
        self.objects.push(object);
    }
}

impl Hittable for HittableList {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        for o in &self.objects {
            o.flatten(xf, out); // list order = visiting order (hittable.rs:75)
        }
    }
}

pub struct Translate {
    object: Box<dyn Hittable>,
    offset: Vec3,
    bbox: AABB,
}

impl Translate {
    pub fn new(object: Box<dyn Hittable>, offset: Vec3) -> Self {
        let bbox = object.bounding_box() + offset;
        Translate { object, offset, bbox }
    }
}

impl Hittable for Translate {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        self.object.flatten(&xf.translated(self.offset.to_array()), out);
    }
}

pub struct RotateY {
    object: Box<dyn Hittable>,
    sin_theta: f64,
    cos_theta: f64,
    bbox: AABB,
}

impl RotateY {
    /// `angle` in degrees.  The box is the hull of the eight rotated corners (hittable.rs:140-167).
    pub fn new(object: Box<dyn Hittable>, angle: f64) -> Self {
        let radians = deg_to_rad(angle);
        let (sin_theta, cos_theta) = (radians.sin(), radians.cos());
        let b = object.bounding_box();
        let mut lo = Vec3::new(f64::MAX, f64::MAX, f64::MAX);
        let mut hi = Vec3::new(f64::MIN, f64::MIN, f64::MIN);
        for &x in &[b.x.min, b.x.max] {
            for &y in &[b.y.min, b.y.max] {
                for &z in &[b.z.min, b.z.max] {
                    let p = Vec3::new(cos_theta * x + sin_theta * z, y, -sin_theta * x + cos_theta * z);
                    lo = Vec3::new(lo.x.min(p.x), lo.y.min(p.y), lo.z.min(p.z));
                    hi = Vec3::new(hi.x.max(p.x), hi.y.max(p.y), hi.z.max(p.z));
                }
            }
        }
        RotateY { object, sin_theta, cos_theta, bbox: AABB::from_corners(lo, hi) }
    }
}

impl Hittable for RotateY {
    fn bounding_box(&self) -> AABB {
        self.bbox
    }
    fn flatten(&self, xf: &Xform, out: &mut SceneBuilder) {
        self.object.flatten(&xf.rotated_y(self.sin_theta, self.cos_theta), out);
    }
}
