//! Scene-construction helpers main.rs imports (reference: util/util.rs:5-16).  The per-path samplers of that file
//! (random_unit_vector, random_cosine_direction, defocus disk) run on the GPU with Philox streams.
use crate::util::vec3::Vec3;

pub fn random_f64(min: f64, max: f64) -> f64 {
    min + (max - min) * fastrand::f64()
}
pub fn random_vector(min: f64, max: f64) -> Vec3 {
    Vec3::new(random_f64(min, max), random_f64(min, max), random_f64(min, max))
}
pub fn deg_to_rad(deg: f64) -> f64 {
    deg / 180.0 * std::f64::consts::PI
}
