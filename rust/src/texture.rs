//! Textures (reference: texture.rs:9-131): the constructors main.rs calls, each able to write itself into the device's texture
//! table.  `value_at` runs on the GPU (gs_shade.cuh `texture_value`): solid, 3-D checker over nested textures, nearest-texel
//! image lookup, 7-octave Perlin turbulence marble.
use crate::camera::ffi::{gs_texture, GS_TEX_CHECKER, GS_TEX_IMAGE, GS_TEX_NOISE, GS_TEX_SOLID};
use crate::camera::scene_builder::{ImagePixels, SceneBuilder};
use crate::util::vec3::Vec3;
use image::{DynamicImage, ImageError};
use std::path::Path;
use std::sync::Arc;

pub trait Texture: Send + Sync {
    /// Row of this texture in the device table (one row per shared Arc; a checker's children come first).
    fn flatten(&self, out: &mut SceneBuilder) -> i32;
}

/// the table key of a shared texture / material: the address of the Arc's pointee
pub fn arc_key<T: ?Sized>(p: &T) -> usize {
    p as *const T as *const () as usize
}

pub struct SolidColorTexture {
    albedo: Vec3,
}

impl SolidColorTexture {
    pub fn new(albedo: Vec3) -> Self {
        SolidColorTexture { albedo }
    }
    pub fn from_rgb(red: f64, green: f64, blue: f64) -> Self {
        SolidColorTexture::new(Vec3::new(red, green, blue))
    }
}

impl Texture for SolidColorTexture {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let c = self.albedo;
        out.texture_row(arc_key(self), |_| gs_texture { kind: GS_TEX_SOLID, a: -1, b: -1, reserved: 0, color: c.to_array(), scale: 0.0 })
    }
}

pub struct CheckeredTexture {
    scale: f64, // the constructor's scale; the library keeps 1 / scale like texture.rs:42,50
    even_texture: Arc<dyn Texture>,
    odd_texture: Arc<dyn Texture>,
}

impl CheckeredTexture {
    pub fn new(scale: f64, even_texture: Arc<dyn Texture>, odd_texture: Arc<dyn Texture>) -> Self {
        CheckeredTexture { scale, even_texture, odd_texture }
    }
    pub fn from_colors(scale: f64, even_color: Vec3, odd_color: Vec3) -> Self {
        CheckeredTexture::new(scale, Arc::new(SolidColorTexture::new(even_color)), Arc::new(SolidColorTexture::new(odd_color)))
    }
}

impl Texture for CheckeredTexture {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let scale = self.scale;
        let (even, odd) = (&self.even_texture, &self.odd_texture);
        out.texture_row(arc_key(self), |out| {
            let (a, b) = (even.flatten(out), odd.flatten(out)); // children first: gs_scene_create requires them to precede the parent
            gs_texture { kind: GS_TEX_CHECKER, a, b, reserved: 0, color: [0.0; 3], scale }
        })
    }
}

pub struct ImageTexture {
    image: DynamicImage,
}

impl ImageTexture {
    pub fn new<P: AsRef<Path>>(filepath: P) -> Result<Self, ImageError> {
        Ok(ImageTexture { image: image::open(filepath)? })
    }
}

impl Texture for ImageTexture {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let img = &self.image;
        out.texture_row(arc_key(self), |out| {
            let rgb = img.to_rgb8(); // what get_pixel(i, j) returns channel by channel (texture.rs:89-91), row 0 = top
            let (w, h) = (rgb.width() as i32, rgb.height() as i32);
            let a = out.image_row(arc_key(img), w, h, || ImagePixels::U8(rgb.into_raw()));
            gs_texture { kind: GS_TEX_IMAGE, a, b: -1, reserved: 0, color: [0.0; 3], scale: 0.0 }
        })
    }
}

pub struct NoiseTexture {
    scale: f64,
}

impl NoiseTexture {
    /// Perlin::default() of noise 0.9 (seed 0); its permutation table travels to the device with the scene (Camera::render)
    pub fn new(scale: f64) -> Self {
        NoiseTexture { scale }
    }
}

impl Texture for NoiseTexture {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let scale = self.scale;
        out.texture_row(arc_key(self), |_| gs_texture { kind: GS_TEX_NOISE, a: -1, b: -1, reserved: 0, color: [0.0; 3], scale })
    }
}
