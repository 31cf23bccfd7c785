//! Output stage (reference: color.rs:8-33): the PPM text line of a pixel.  The numbers come from the GPU as linear mean
//! radiance (gs_render) and take the reference's own transfer here; `gs_resolve_rgb8` is the same stage on the device.
use crate::util::vec3::Vec3;
use std::io::Write;

fn linear_to_gamma(c: f64) -> f64 {
    if c > 0.0 { c.sqrt() } else { 0.0 }
}

pub fn write_color<W: Write>(out: &mut W, color: Vec3) {
    let byte = |c: f64| (256.0 * linear_to_gamma(c).clamp(0.0, 0.999)) as i32;
    writeln!(out, "{} {} {}", byte(color.x), byte(color.y), byte(color.z)).unwrap();
}

/// 0.144 is the reference's blue weight (sic, color.rs:32); the adaptive sampler on the device uses the same number.
pub fn luminance(color: Vec3) -> f64 {
    0.299 * color.x + 0.587 * color.y + 0.144 * color.z
}
