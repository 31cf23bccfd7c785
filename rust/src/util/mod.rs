//! util -- the small math layer main.rs builds scenes with (same names and signatures as the reference's util/*.rs).
pub mod interval;
pub mod util;
pub mod vec3;
