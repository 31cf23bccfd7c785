//! Camera (reference: camera.rs:17-271): `Camera::new` computes the same private fields from the same arguments
//! (camera.rs:39-98); `Camera::render` keeps its signature (camera.rs:100) and its output -- the PPM text of image.ppm -- but
//! the per-pixel loop (rayon `par_iter` over `sample`, camera.rs:110-114) becomes: flatten the world, hand the arrays to
//! libgrayshift_cuda.so, read the mean radiance back, `write_color` each pixel as before.
//!
//! NOT COMPILED in this repository's build environment (no cargo / rustc): see rust/README.md.
#![allow(non_camel_case_types)]
#[path = "ffi.rs"]
pub mod ffi;
#[path = "scene_builder.rs"]
pub mod scene_builder;

use crate::color::write_color;
use crate::hittable::hittable::Hittable;
use crate::util::util::deg_to_rad;
use crate::util::vec3::Vec3;
use ffi::{gs_camera, gs_sample_settings, GS_BG_HDRI, GS_BG_SOLID};
use scene_builder::{ImagePixels, SceneBuilder, Xform};
use std::fs::File;
use std::io::Write;

pub struct SampleSettings {
    pub confidence: f64, // used directly as the z-value (Q3)
    pub tolerance: f64,
    pub batch_size: u32,
    pub max_samples: u32,
}

pub enum Background {
    SOLID(Vec3),
    HDRI(HDRI),
}

pub struct HDRI {
    pub image: radiant::Image,
    pub rotation: Vec3, // radians (util.rs:67-70)
}

pub struct Camera {
    image_width: i32,
    image_height: i32,
    sample_settings: SampleSettings,
    max_depth: u32,
    center: Vec3,
    starting_pixel_pos: Vec3,
    pixel_delta_u: Vec3,
    pixel_delta_v: Vec3,
    defocus_angle: f64,
    defocus_disk_u: Vec3,
    defocus_disk_v: Vec3,
    background: Background,
}

impl Camera {
    #[allow(clippy::too_many_arguments)]
    pub fn new(aspect_ratio: f64, image_width: i32, sample_settings: SampleSettings, max_depth: u32, v_fov: f64, look_from: Vec3,
               look_at: Vec3, vup: Vec3, defocus_angle: f64, focus_distance: f64, background: Background) -> Self {
        // image height truncates (400 -> 225, 600 -> 337, camera.rs:52)
        let mut image_height = (image_width as f64 / aspect_ratio) as i32;
        if image_height < 1 {
            image_height = 1;
        }
        let theta = deg_to_rad(v_fov);
        let h = (theta / 2.0).tan();
        let viewport_height = 2.0 * h * focus_distance;
        let viewport_width = viewport_height * (image_width as f64 / image_height as f64);
        let w = (look_from - look_at).unit();
        let u = vup.cross(w).unit();
        let v = w.cross(u);
        let viewport_u = u * viewport_width;
        let viewport_v = -v * viewport_height;
        let pixel_delta_u = viewport_u / image_width as f64;
        let pixel_delta_v = viewport_v / image_height as f64;
        let upper_left = look_from - w * focus_distance - viewport_u / 2.0 - viewport_v / 2.0;
        let starting_pixel_pos = upper_left + (pixel_delta_u + pixel_delta_v) * 0.5;
        let defocus_radius = focus_distance * deg_to_rad(defocus_angle / 2.0).tan();
        Camera {
            image_width,
            image_height,
            sample_settings,
            max_depth,
            center: look_from,
            starting_pixel_pos,
            pixel_delta_u,
            pixel_delta_v,
            defocus_angle,
            defocus_disk_u: u * defocus_radius,
            defocus_disk_v: v * defocus_radius,
            background,
        }
    }

    /// Environment variables: GS_DEVICES = number of GPUs to use (default: all), GS_SEED = Philox key (default 1).
    pub fn render(&self, world: Box<dyn Hittable>, image_file: &mut File) -> std::io::Result<()> {
        writeln!(image_file, "P3")?;
        writeln!(image_file, "{} {}", self.image_width, self.image_height)?;
        writeln!(image_file, "255")?;

        // ---- flatten: bake instances, rank primitives in visiting order, dedupe Arc tables, build the SAH tree ----
        let mut builder = SceneBuilder::new();
        world.flatten(&Xform::identity(), &mut builder);
        let mut camera = gs_camera {
            image_width: self.image_width,
            image_height: self.image_height,
            max_depth: self.max_depth,
            background_kind: GS_BG_SOLID,
            center: self.center.to_array(),
            starting_pixel_pos: self.starting_pixel_pos.to_array(),
            pixel_delta_u: self.pixel_delta_u.to_array(),
            pixel_delta_v: self.pixel_delta_v.to_array(),
            defocus_angle: self.defocus_angle,
            defocus_disk_u: self.defocus_disk_u.to_array(),
            defocus_disk_v: self.defocus_disk_v.to_array(),
            background_color: [0.0; 3],
            hdri_image: -1,
            reserved: 0,
            hdri_rotation: [0.0; 3],
        };
        match &self.background {
            Background::SOLID(c) => camera.background_color = c.to_array(),
            Background::HDRI(h) => {
                let (w, hgt) = (h.image.width as i32, h.image.height as i32);
                let texels: Vec<f32> = h.image.data.iter().flat_map(|p| [p.r, p.g, p.b]).collect(); // row 0 = top (camera.rs:265-269)
                camera.background_kind = GS_BG_HDRI;
                camera.hdri_image = builder.image_row(&h.image as *const _ as usize, w, hgt, || ImagePixels::F32(texels));
                camera.hdri_rotation = h.rotation.to_array();
            }
        }
        // the permutation table of Perlin::default() straight from the crate the reference uses (texture.rs:104): nothing restated
        let perm = noise::permutationtable::PermutationTable::new(0);
        let mut table = [0u8; 256];
        for (i, t) in table.iter_mut().enumerate() {
            // hash(&[i]) of a one-element key is values[i] (noise 0.9 permutationtable.rs)
            *t = noise::permutationtable::NoiseHasher::hash(&perm, &[i as isize]) as u8;
        }
        let mut flat = builder.finish(table);
        let desc = flat.desc();

        // ---- render on every GPU the process may use; one call, the library splits samples (or pixels) and sums over NVLink ----
        let settings = gs_sample_settings {
            confidence: self.sample_settings.confidence,
            tolerance: self.sample_settings.tolerance,
            batch_size: self.sample_settings.batch_size,
            max_samples: self.sample_settings.max_samples,
        };
        let seed = std::env::var("GS_SEED").ok().and_then(|s| s.parse().ok()).unwrap_or(1u64);
        let available = unsafe { ffi::gs_device_count() }.max(0) as usize;
        if available == 0 {
            return Err(std::io::Error::new(std::io::ErrorKind::Other, "no CUDA device (there is no CPU fallback)"));
        }
        let wanted = std::env::var("GS_DEVICES").ok().and_then(|s| s.parse::<usize>().ok()).unwrap_or(available).clamp(1, available);
        let scenes: Vec<ffi::Scene> = (0..wanted).map(|d| ffi::Scene::create(&desc, d as i32)).collect::<Result<_, _>>()?;
        let n = (self.image_width as usize) * (self.image_height as usize);
        let mut rgb = vec![0f32; n * 3];
        let stats = ffi::Scene::render_multi(&scenes, &camera, &settings, seed, &mut rgb, None)?;
        log::info!("{} paths, {} rays, {:.1} ms on {} GPU(s) (exchange {:.2} ms)", stats.paths, stats.rays, stats.kernel_ms, wanted, stats.exchange_ms);

        for px in rgb.chunks_exact(3) {
            write_color(image_file, Vec3::new(px[0] as f64, px[1] as f64, px[2] as f64));
        }
        Ok(())
    }
}
