"""Thin Python wrappers over the two C interfaces, named after the reference's types.

HostScene   -- the host layer (scene construction / flatten / BVH build; libgrayshift_host.so)
DeviceScene -- a flattened scene resident on one B200 (libgrayshift_cuda.so): render, trace_closest ...

Rendering entry points mirror `Camera::render` (src/camera.rs:100): they take what the reference's
camera holds (`gs_camera`, `gs_sample_settings`) and return the per-pixel mean radiance the reference's
`Camera::sample` returns (src/camera.rs:125-171), before `write_color` (src/color.rs:8-26).
"""
import ctypes as C

import numpy as np

from . import ffi


class HostScene:
    """A scene graph + camera built by the host layer."""

    def __init__(self, handle):
        if not handle:
            raise ffi.GrayshiftError(1, ffi.host_lib().gsh_last_error().decode())
        self._h = C.c_void_p(handle)

    @classmethod
    def build(cls, name, width=0, spp=0, max_depth=0, scene_seed=1, include_media=True, n_spheres=0, asset_dir=None):
        lib = ffi.host_lib()
        ad = (asset_dir or ffi.ASSET_DIR).encode()
        return cls(lib.gsh_scene_build(name.encode(), width, spp, max_depth, scene_seed, int(include_media), n_spheres, ad))

    @classmethod
    def load(cls, graph_path, include_media=True):
        return cls(ffi.host_lib().gsh_scene_load(str(graph_path).encode(), int(include_media)))

    def dump(self, graph_path):
        ffi.check_host(ffi.host_lib().gsh_scene_dump(self._h, str(graph_path).encode()))

    def flatten(self, max_leaf_size=0, reference_tree=False):
        ffi.check_host(ffi.host_lib().gsh_scene_flatten(self._h, max_leaf_size, int(reference_tree)))
        return self

    def set_sampling(self, confidence=0.95, tolerance=0.0, batch_size=1, max_samples=0, max_depth=0):
        ffi.check_host(ffi.host_lib().gsh_scene_set_sampling(self._h, confidence, tolerance, batch_size, max_samples, max_depth))
        return self

    @property
    def desc(self):
        p = ffi.host_lib().gsh_scene_desc(self._h)
        if not p:
            raise ffi.GrayshiftError(1, ffi.host_lib().gsh_last_error().decode())
        return p

    def camera(self):
        cam, st = ffi.gs_camera(), ffi.gs_sample_settings()
        ffi.check_host(ffi.host_lib().gsh_scene_camera(self._h, C.byref(cam), C.byref(st)))
        return cam, st

    def render_ppm(self, path, device=0, seed=1):
        st = ffi.gs_stats()
        ffi.check_host(ffi.host_lib().gsh_scene_render_ppm(self._h, str(path).encode(), device, seed, C.byref(st)))
        return st

    def close(self):
        if self._h:
            ffi.host_lib().gsh_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fixed_spp(settings):
    """batch_size when the settings give a fixed per-pixel count (max_samples < batch_size: the loop of
    src/camera.rs:135-165 always leaves after its first batch), else None (adaptive)."""
    return settings.batch_size if settings.max_samples < settings.batch_size else None


class DeviceScene:
    """gs_scene: a flattened scene uploaded to one GPU."""

    def __init__(self, host_scene, device=0):
        self._lib = ffi.cuda_lib()
        self.host = host_scene
        self.device = device
        h = C.c_void_p()
        ffi.check(self._lib.gs_scene_create(host_scene.desc, device, C.byref(h)))
        self._h = h
        self.cam, self.settings = host_scene.camera()

    @property
    def width(self):
        return self.cam.image_width

    @property
    def height(self):
        return self.cam.image_height

    def set_option(self, key, value):
        ffi.check(self._lib.gs_scene_set_option(self._h, key.encode(), int(value)))

    def render(self, seed=1, settings=None, want_counts=True):
        """Camera::render without the PPM text -> (rgb float32 [H,W,3], counts uint32 [H,W], gs_stats)."""
        st = settings or self.settings
        n = self.width * self.height
        rgb = np.empty((n, 3), dtype=np.float32)
        counts = np.empty(n, dtype=np.uint32) if want_counts else None
        stats = ffi.gs_stats()
        ffi.check(self._lib.gs_render(self._h, C.byref(self.cam), C.byref(st), seed, rgb.ctypes.data,
                                      counts.ctypes.data if want_counts else None, C.byref(stats)))
        return (rgb.reshape(self.height, self.width, 3), counts.reshape(self.height, self.width) if want_counts else None, stats)

    def render_accumulate(self, d_accum_ptr, seed, sample_begin, sample_count, stream=0):
        """Adds samples [sample_begin, +sample_count) of every pixel into a device float4[W*H] buffer (async)."""
        ffi.check(self._lib.gs_render_accumulate(self._h, C.byref(self.cam), seed, sample_begin, sample_count,
                                                 C.c_void_p(d_accum_ptr), C.c_void_p(stream)))

    def render_adaptive_accumulate(self, d_accum_ptr, seed, pixel_begin=0, pixel_stride=1, settings=None, stream=0):
        st = settings or self.settings
        ffi.check(self._lib.gs_render_adaptive_accumulate(self._h, C.byref(self.cam), C.byref(st), seed, pixel_begin, pixel_stride,
                                                          C.c_void_p(d_accum_ptr), C.c_void_p(stream)))

    def resolve(self, d_accum_ptr, stream=0, out_rgb=None, out_counts=None):
        """pixel_color /= sample_count (camera.rs:167) -> host.  out_rgb / out_counts: caller-owned host arrays (e.g.
        views of pinned memory) to receive the image instead of fresh numpy arrays."""
        n = self.width * self.height
        rgb = np.empty((n, 3), dtype=np.float32) if out_rgb is None else out_rgb.reshape(n, 3)
        counts = np.empty(n, dtype=np.uint32) if out_counts is None else out_counts.reshape(n)
        assert rgb.dtype == np.float32 and counts.dtype == np.uint32 and rgb.flags.c_contiguous and counts.flags.c_contiguous
        ffi.check(self._lib.gs_resolve(self._h, C.c_void_p(d_accum_ptr), self.width, self.height, rgb.ctypes.data, counts.ctypes.data,
                                       C.c_void_p(stream)))
        return rgb.reshape(self.height, self.width, 3), counts.reshape(self.height, self.width)

    def resolve_rgb8(self, d_accum_ptr, stream=0):
        """Device output stage: /count + write_color's transfer (src/color.rs:8-26) -> uint8 [H,W,3] on the host."""
        n = self.width * self.height
        rgb8 = np.empty((n, 3), dtype=np.uint8)
        counts = np.empty(n, dtype=np.uint32)
        ffi.check(self._lib.gs_resolve_rgb8(self._h, C.c_void_p(d_accum_ptr), self.width, self.height, rgb8.ctypes.data, counts.ctypes.data,
                                            C.c_void_p(stream)))
        return rgb8.reshape(self.height, self.width, 3), counts.reshape(self.height, self.width)

    def stats(self):
        st = ffi.gs_stats()
        ffi.check(self._lib.gs_stats_read(self._h, C.byref(st)))
        return st

    def trace_closest(self, rays, tmin=0.001, tmax=1.7976931348623157e308, fast=False):
        """Hittable::hit over the surface primitives: rays [n,7] f64 -> (prim_id int32 [n], t f64 [n]).
        fast=True: answered by the render kernels' build of the traversal (FMA contraction, approximate f64 division)."""
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        ids = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float64)
        fn = self._lib.gs_trace_closest_fast if fast else self._lib.gs_trace_closest
        ffi.check(fn(self._h, rays.ctypes.data, n, tmin, tmax, ids.ctypes.data, t.ctypes.data))
        return ids, t

    def info(self, key):
        v = C.c_int64()
        ffi.check(self._lib.gs_scene_get_info(self._h, key.encode(), C.byref(v)))
        return v.value

    def camera_rays(self, seed, sample, pixel_begin=0, pixel_end=None):
        pixel_end = self.width * self.height if pixel_end is None else pixel_end
        out = np.empty((pixel_end - pixel_begin, 7), dtype=np.float64)
        ffi.check(self._lib.gs_camera_rays(self._h, C.byref(self.cam), seed, sample, pixel_begin, pixel_end, out.ctypes.data))
        return out

    def path_radiance(self, seed, pixels, samples):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        out = np.empty((pixels.shape[0], 3), dtype=np.float32)
        ffi.check(self._lib.gs_path_radiance(self._h, C.byref(self.cam), seed, pixels.ctypes.data, samples.ctypes.data, pixels.shape[0],
                                             out.ctypes.data))
        return out

    def close(self):
        if getattr(self, "_h", None):
            self._lib.gs_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def tonemap(rgb):
    """write_color (src/color.rs:8-26) for a whole frame: float32 [...,3] -> uint8 [...,3]."""
    rgb = np.ascontiguousarray(rgb, dtype=np.float32)
    out = np.empty(rgb.shape, dtype=np.uint8)
    ffi.check(ffi.cuda_lib().gs_tonemap(rgb.ctypes.data, rgb.size // 3, out.ctypes.data))
    return out
