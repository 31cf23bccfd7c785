//! Interval (reference: util/interval.rs): only what the bounding boxes of the host-side BVH ordering need.
#[derive(Clone, Copy, Debug)]
pub struct Interval {
    pub min: f64,
    pub max: f64,
}

impl Interval {
    pub const EMPTY: Interval = Interval { min: f64::MAX, max: f64::MIN };
    pub fn new(min: f64, max: f64) -> Self {
        Interval { min, max }
    }
    /// the hull of two intervals (interval.rs:17-22)
    pub fn from_intervals(a: Interval, b: Interval) -> Self {
        Interval { min: a.min.min(b.min), max: a.max.max(b.max) }
    }
    pub fn size(&self) -> f64 {
        self.max - self.min
    }
    /// widen by delta, half on each side (interval.rs:41-44)
    pub fn expand(&self, delta: f64) -> Self {
        Interval { min: self.min - delta / 2.0, max: self.max + delta / 2.0 }
    }
}
