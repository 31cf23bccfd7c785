//! rust/src/ffi.rs -- bindings to include/grayshift.h (libgrayshift_cuda.so), field for field.
//! NOT COMPILED in this repository's build environment (no Rust toolchain); kept in lock-step with the
//! header by hand -- tests/test_abi.py checks the ctypes twin (grayshift_b200/ffi.py) against the C sizes.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

pub const GS_ABI_VERSION: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_sphere {
    pub center: [f64; 3],
    pub radius: f64,
    pub path: [f64; 3],
    pub material: i32,
    pub prim_id: i32,
    pub moving: i32,
    pub order: i32,
    pub uv_rotation: [f64; 2],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_quad {
    pub q: [f64; 3],
    pub u: [f64; 3],
    pub v: [f64; 3],
    pub material: i32,
    pub prim_id: i32,
    pub order: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_triangle {
    pub a: [f64; 3],
    pub b: [f64; 3],
    pub c: [f64; 3],
    pub material: i32,
    pub prim_id: i32,
    pub order: i32,
    pub reserved: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_bvh_node {
    pub bounds_min: [f64; 3],
    pub bounds_max: [f64; 3],
    pub left: i32,
    pub right: i32,
    pub first_ref: i32,
    pub ref_count: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_medium {
    pub sphere_first: i32,
    pub sphere_count: i32,
    pub quad_first: i32,
    pub quad_count: i32,
    pub density: f64,
    pub phase_material: i32,
    pub medium_index: i32,
}

pub const GS_MAT_LAMBERTIAN: i32 = 0;
pub const GS_MAT_METAL: i32 = 1;
pub const GS_MAT_DIELECTRIC: i32 = 2;
pub const GS_MAT_DIFFUSE_LIGHT: i32 = 3;
pub const GS_MAT_ISOTROPIC: i32 = 4;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_material {
    pub kind: i32,
    pub texture: i32,
    pub albedo: [f64; 3],
    pub param: f64,
}

pub const GS_TEX_SOLID: i32 = 0;
pub const GS_TEX_CHECKER: i32 = 1;
pub const GS_TEX_IMAGE: i32 = 2;
pub const GS_TEX_NOISE: i32 = 3;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_texture {
    pub kind: i32,
    pub a: i32,
    pub b: i32,
    pub reserved: i32,
    pub color: [f64; 3],
    pub scale: f64,
}

pub const GS_IMG_U8_RGB: i32 = 0;
pub const GS_IMG_F32_RGB: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct gs_image {
    pub width: i32,
    pub height: i32,
    pub format: i32,
    pub reserved: i32,
    pub pixels: *const c_void,
}

#[repr(C)]
pub struct gs_scene_desc {
    pub abi_version: u32,
    pub reserved: u32,
    pub spheres: *const gs_sphere,
    pub n_spheres: usize,
    pub quads: *const gs_quad,
    pub n_quads: usize,
    pub nodes: *const gs_bvh_node,
    pub n_nodes: usize,
    pub prim_refs: *const u32,
    pub n_prim_refs: usize,
    pub boundary_spheres: *const gs_sphere,
    pub n_boundary_spheres: usize,
    pub boundary_quads: *const gs_quad,
    pub n_boundary_quads: usize,
    pub media: *const gs_medium,
    pub n_media: usize,
    pub materials: *const gs_material,
    pub n_materials: usize,
    pub textures: *const gs_texture,
    pub n_textures: usize,
    pub images: *const gs_image,
    pub n_images: usize,
    pub perlin_perm: *const u8,
    pub triangles: *const gs_triangle,
    pub n_triangles: usize,
}

pub const GS_BG_SOLID: i32 = 0;
pub const GS_BG_HDRI: i32 = 1;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_camera {
    pub image_width: i32,
    pub image_height: i32,
    pub max_depth: u32,
    pub background_kind: i32,
    pub center: [f64; 3],
    pub starting_pixel_pos: [f64; 3],
    pub pixel_delta_u: [f64; 3],
    pub pixel_delta_v: [f64; 3],
    pub defocus_angle: f64,
    pub defocus_disk_u: [f64; 3],
    pub defocus_disk_v: [f64; 3],
    pub background_color: [f64; 3],
    pub hdri_image: i32,
    pub reserved: i32,
    pub hdri_rotation: [f64; 3],
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct gs_sample_settings {
    pub confidence: f64,
    pub tolerance: f64,
    pub batch_size: u32,
    pub max_samples: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default, Debug)]
pub struct gs_stats {
    pub paths: u64,
    pub rays: u64,
    pub tex_fetches: u64,
    pub node_visits: u64,
    pub kernel_launches: u64,
    pub kernel_ms: f64,
    pub total_ms: f64,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub exchange_ms: f64,
}

#[repr(C)]
pub struct gs_scene {
    _private: [u8; 0],
}

#[link(name = "grayshift_cuda")]
extern "C" {
    pub fn gs_last_error() -> *const c_char;
    pub fn gs_abi_version() -> c_int;
    pub fn gs_device_count() -> c_int;
    pub fn gs_scene_create(desc: *const gs_scene_desc, device: c_int, out: *mut *mut gs_scene) -> c_int;
    pub fn gs_scene_destroy(scene: *mut gs_scene);
    pub fn gs_scene_set_option(scene: *mut gs_scene, key: *const c_char, value: i64) -> c_int;
    pub fn gs_scene_get_info(scene: *mut gs_scene, key: *const c_char, value: *mut i64) -> c_int;
    pub fn gs_render_multi(scenes: *const *mut gs_scene, n_scenes: c_int, camera: *const gs_camera, settings: *const gs_sample_settings,
                           seed: u64, out_rgb: *mut f32, out_counts: *mut u32, stats: *mut gs_stats) -> c_int;
    pub fn gs_render(scene: *mut gs_scene, camera: *const gs_camera, settings: *const gs_sample_settings, seed: u64,
                     out_rgb: *mut f32, out_counts: *mut u32, stats: *mut gs_stats) -> c_int;
    pub fn gs_render_accumulate(scene: *mut gs_scene, camera: *const gs_camera, seed: u64, sample_begin: u32, sample_count: u32,
                                d_accum: *mut c_void, cuda_stream: *mut c_void) -> c_int;
    pub fn gs_render_adaptive_accumulate(scene: *mut gs_scene, camera: *const gs_camera, settings: *const gs_sample_settings,
                                         seed: u64, pixel_begin: u32, pixel_stride: u32, d_accum: *mut c_void,
                                         cuda_stream: *mut c_void) -> c_int;
    pub fn gs_resolve(scene: *mut gs_scene, d_accum: *const c_void, width: i32, height: i32, out_rgb: *mut f32,
                      out_counts: *mut u32, cuda_stream: *mut c_void) -> c_int;
    pub fn gs_resolve_rgb8(scene: *mut gs_scene, d_accum: *const c_void, width: i32, height: i32, out_rgb8: *mut u8,
                           out_counts: *mut u32, cuda_stream: *mut c_void) -> c_int;
    pub fn gs_stats_read(scene: *mut gs_scene, stats: *mut gs_stats) -> c_int;
    pub fn gs_tonemap(rgb: *const f32, n_pixels: usize, out_rgb8: *mut u8) -> c_int;
    pub fn gs_trace_closest(scene: *mut gs_scene, rays: *const f64, n: usize, tmin: f64, tmax: f64, prim_id: *mut i32,
                            t: *mut f64) -> c_int;
    pub fn gs_trace_closest_fast(scene: *mut gs_scene, rays: *const f64, n: usize, tmin: f64, tmax: f64, prim_id: *mut i32,
                                 t: *mut f64) -> c_int;
    pub fn gs_camera_rays(scene: *mut gs_scene, camera: *const gs_camera, seed: u64, sample: u32, pixel_begin: i32,
                          pixel_end: i32, rays_out: *mut f64) -> c_int;
    pub fn gs_path_radiance(scene: *mut gs_scene, camera: *const gs_camera, seed: u64, pixel: *const u32, sample: *const u32,
                            n: usize, out_rgb: *mut f32) -> c_int;
}

/// The library's last error message for this thread.
pub fn last_error() -> String {
    unsafe { CStr::from_ptr(gs_last_error()).to_string_lossy().into_owned() }
}

fn check(rc: c_int) -> std::io::Result<()> {
    if rc == 0 {
        Ok(())
    } else {
        Err(std::io::Error::new(std::io::ErrorKind::Other, format!("grayshift_cuda [{rc}]: {}", last_error())))
    }
}

/// RAII owner of a `gs_scene` (gs_scene_create / gs_scene_destroy).
pub struct Scene {
    raw: *mut gs_scene,
}

impl Scene {
    pub fn create(desc: &gs_scene_desc, device: i32) -> std::io::Result<Scene> {
        let mut raw: *mut gs_scene = std::ptr::null_mut();
        check(unsafe { gs_scene_create(desc, device, &mut raw) })?;
        Ok(Scene { raw })
    }

    /// Camera::render's hot part: mean radiance per pixel (W*H*3) and, optionally, per-pixel sample counts.
    pub fn render(&self, camera: &gs_camera, settings: &gs_sample_settings, seed: u64, rgb: &mut [f32],
                  counts: Option<&mut [u32]>) -> std::io::Result<gs_stats> {
        let n = (camera.image_width as usize) * (camera.image_height as usize);
        assert!(rgb.len() >= n * 3);
        let counts_ptr = match counts {
            Some(c) => {
                assert!(c.len() >= n);
                c.as_mut_ptr()
            }
            None => std::ptr::null_mut(),
        };
        let mut stats = gs_stats::default();
        check(unsafe { gs_render(self.raw, camera, settings, seed, rgb.as_mut_ptr(), counts_ptr, &mut stats) })?;
        Ok(stats)
    }

    /// Camera::render on several GPUs in one call (gs_render_multi): `scenes[g]` lives on device g.
    pub fn render_multi(scenes: &[Scene], camera: &gs_camera, settings: &gs_sample_settings, seed: u64, rgb: &mut [f32],
                        counts: Option<&mut [u32]>) -> std::io::Result<gs_stats> {
        let n = (camera.image_width as usize) * (camera.image_height as usize);
        assert!(!scenes.is_empty() && rgb.len() >= n * 3);
        let raws: Vec<*mut gs_scene> = scenes.iter().map(|s| s.raw).collect();
        let counts_ptr = match counts {
            Some(c) => {
                assert!(c.len() >= n);
                c.as_mut_ptr()
            }
            None => std::ptr::null_mut(),
        };
        let mut stats = gs_stats::default();
        check(unsafe { gs_render_multi(raws.as_ptr(), raws.len() as c_int, camera, settings, seed, rgb.as_mut_ptr(), counts_ptr, &mut stats) })?;
        Ok(stats)
    }

    pub fn trace_closest(&self, rays: &[f64], tmin: f64, tmax: f64) -> std::io::Result<(Vec<i32>, Vec<f64>)> {
        let n = rays.len() / 7;
        let (mut ids, mut t) = (vec![0i32; n], vec![0f64; n]);
        check(unsafe { gs_trace_closest(self.raw, rays.as_ptr(), n, tmin, tmax, ids.as_mut_ptr(), t.as_mut_ptr()) })?;
        Ok((ids, t))
    }
}

impl Drop for Scene {
    fn drop(&mut self) {
        unsafe { gs_scene_destroy(self.raw) }
    }
}
