//! OrthonormalBasis (reference: ONB.rs:10-27).  Built per Lambertian scatter on the GPU (gs_shade.cuh `shade`); the host keeps
//! the type so that `mod ONB;` in main.rs resolves.
use crate::util::vec3::Vec3;

pub struct OrthonormalBasis {
    pub u: Vec3,
    pub v: Vec3,
    pub w: Vec3,
}

impl OrthonormalBasis {
    pub fn new(n: Vec3) -> Self {
        let w = n.unit();
        let a = if w.x.abs() > 0.9 { Vec3::new(0.0, 1.0, 0.0) } else { Vec3::new(1.0, 0.0, 0.0) };
        let v = w.cross(a).unit();
        let u = w.cross(v);
        OrthonormalBasis { u, v, w }
    }
    pub fn transform(&self, p: Vec3) -> Vec3 {
        self.u * p.x + self.v * p.y + self.w * p.z
    }
}
