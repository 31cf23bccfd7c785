//! Scene graph: the reference's constructors (hittable/*.rs) kept as the scene-BUILDING API; what each node can do now is
//! report its bounding box (for the reference's tree shape) and flatten itself into the device's arrays.  `hit` runs on the GPU.
#![allow(non_snake_case)]
pub mod BVH;
pub mod hittable;
pub mod quad;
pub mod sphere;
pub mod triangle;
pub mod volume;
