//! Materials (reference: material.rs:11-201): constructors + the row each one writes into the device's material table.
//! `scatter` / `emitted` run on the GPU (gs_shade.cuh `shade`), quirks included: the r2^(1/4) "cosine" lobe, random numbers drawn
//! even when unused, two-sided lights.
use crate::camera::ffi::{gs_material, GS_MAT_DIELECTRIC, GS_MAT_DIFFUSE_LIGHT, GS_MAT_ISOTROPIC, GS_MAT_LAMBERTIAN, GS_MAT_METAL};
use crate::camera::scene_builder::SceneBuilder;
use crate::texture::{arc_key, SolidColorTexture, Texture};
use crate::util::vec3::Vec3;
use std::sync::Arc;

pub trait Material: Send + Sync {
    /// Row of this material in the device table (one row per shared Arc: main.rs puts one material on 1000 spheres).
    fn flatten(&self, out: &mut SceneBuilder) -> i32;
}

fn textured(kind: i32, key: usize, texture: &Arc<dyn Texture>, out: &mut SceneBuilder) -> i32 {
    out.material_row(key, |out| gs_material { kind, texture: texture.flatten(out), albedo: [0.0; 3], param: 0.0 })
}

pub struct Lambertian {
    texture: Arc<dyn Texture>,
}

impl Lambertian {
    pub fn from_color(albedo: Vec3) -> Self {
        Lambertian { texture: Arc::new(SolidColorTexture::new(albedo)) }
    }
    pub fn from_texture(texture: Arc<dyn Texture>) -> Self {
        Lambertian { texture }
    }
}

impl Material for Lambertian {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        textured(GS_MAT_LAMBERTIAN, arc_key(self), &self.texture, out)
    }
}

pub struct Metal {
    albedo: Vec3,
    fuzz: f64, // not clamped, as in the reference (main.rs:688 uses 1.0)
}

impl Metal {
    pub fn new(albedo: Vec3, fuzz: f64) -> Self {
        Metal { albedo, fuzz }
    }
}

impl Material for Metal {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let (albedo, fuzz) = (self.albedo, self.fuzz);
        out.material_row(arc_key(self), |_| gs_material { kind: GS_MAT_METAL, texture: -1, albedo: albedo.to_array(), param: fuzz })
    }
}

pub struct Dielectric {
    refraction_index: f64,
}

impl Dielectric {
    pub fn new(refraction_index: f64) -> Self {
        Dielectric { refraction_index }
    }
    /// Schlick's approximation (material.rs:114-119); host copy for callers that want the number
    pub fn reflectance(cosine: f64, refraction_index: f64) -> f64 {
        let r0 = ((1.0 - refraction_index) / (1.0 + refraction_index)).powi(2);
        r0 + (1.0 - r0) * (1.0 - cosine).powi(5)
    }
}

impl Material for Dielectric {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        let ir = self.refraction_index;
        out.material_row(arc_key(self), |_| gs_material { kind: GS_MAT_DIELECTRIC, texture: -1, albedo: [1.0; 3], param: ir })
    }
}

pub struct DiffuseLight {
    texture: Arc<dyn Texture>,
}

impl DiffuseLight {
    pub fn new(texture: Arc<dyn Texture>) -> Self {
        DiffuseLight { texture }
    }
    pub fn from_color(color: Vec3) -> Self {
        DiffuseLight { texture: Arc::new(SolidColorTexture::new(color)) }
    }
}

impl Material for DiffuseLight {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        textured(GS_MAT_DIFFUSE_LIGHT, arc_key(self), &self.texture, out)
    }
}

pub struct Isotropic {
    texture: Arc<dyn Texture>,
}

impl Isotropic {
    pub fn new(texture: Arc<dyn Texture>) -> Self {
        Isotropic { texture }
    }
    pub fn from_color(color: Vec3) -> Self {
        Isotropic { texture: Arc::new(SolidColorTexture::new(color)) }
    }
}

impl Material for Isotropic {
    fn flatten(&self, out: &mut SceneBuilder) -> i32 {
        textured(GS_MAT_ISOTROPIC, arc_key(self), &self.texture, out)
    }
}
