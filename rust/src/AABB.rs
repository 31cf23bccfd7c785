//! AABB (reference: AABB.rs:24-153).  The GPU tests boxes on its own padded fp32 copies; on the host the boxes only decide the
//! SHAPE of the reference's median-split tree (BVH.rs:18-65), i.e. the order in which primitives are visited -- the rank that
//! breaks exact ties between coplanar quads (gs_quad.order).
use crate::util::interval::Interval;
use crate::util::vec3::Vec3;
use std::ops::Add;

#[derive(Clone, Copy, Debug)]
pub struct AABB {
    pub x: Interval,
    pub y: Interval,
    pub z: Interval,
}

impl AABB {
    pub const EMPTY: AABB = AABB { x: Interval::EMPTY, y: Interval::EMPTY, z: Interval::EMPTY };

    pub fn from_corners(a: Vec3, b: Vec3) -> Self {
        let iv = |p: f64, q: f64| if p <= q { Interval::new(p, q) } else { Interval::new(q, p) };
        let mut bb = AABB { x: iv(a.x, b.x), y: iv(a.y, b.y), z: iv(a.z, b.z) };
        bb.pad_to_minimums();
        bb
    }
    #[allow(non_snake_case)]
    pub fn from_AABB_pair(a: AABB, b: AABB) -> Self {
        AABB { x: Interval::from_intervals(a.x, b.x), y: Interval::from_intervals(a.y, b.y), z: Interval::from_intervals(a.z, b.z) }
    }
    pub fn axis(&self, n: usize) -> Interval {
        match n {
            0 => self.x,
            1 => self.y,
            _ => self.z,
        }
    }
    /// strict comparisons: x wins only if it is strictly the longest (AABB.rs:115-121)
    pub fn longest_axis(&self) -> usize {
        if self.x.size() > self.y.size() {
            if self.x.size() > self.z.size() { 0 } else { 2 }
        } else if self.y.size() > self.z.size() {
            1
        } else {
            2
        }
    }
    /// no axis thinner than 1e-4 (AABB.rs:123-128)
    fn pad_to_minimums(&mut self) {
        let delta = 0.0001;
        if self.x.size() < delta { self.x = self.x.expand(delta); }
        if self.y.size() < delta { self.y = self.y.expand(delta); }
        if self.z.size() < delta { self.z = self.z.expand(delta); }
    }
}

impl Add<Vec3> for AABB {
    type Output = AABB;
    fn add(self, o: Vec3) -> AABB {
        AABB {
            x: Interval::new(self.x.min + o.x, self.x.max + o.x),
            y: Interval::new(self.y.min + o.y, self.y.max + o.y),
            z: Interval::new(self.z.min + o.z, self.z.max + o.z),
        }
    }
}
