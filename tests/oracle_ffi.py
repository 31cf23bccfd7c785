"""ctypes binding of the CPU oracle (oracle/_build/liboracle*.so).  TEST INFRASTRUCTURE: only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this."""
import ctypes as C
import os
import subprocess

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, "oracle")
BUILD_DIR = os.path.join(ORACLE_DIR, "_build")
ASSET_DIR = os.path.join(REPO, "assets")

_libs = {}


def build():
    subprocess.run(["make", "-C", ORACLE_DIR, "-j4"], check=True, capture_output=True)


def lib(counters=False):
    # counters: False = timing build, True = reference-semantics work counters, "norc" = timing build without the per-hit
    # material reference count (bench.py's cpu_baseline.no_refcount)
    name = "liboracle_norc.so" if counters == "norc" else ("liboracle_counters.so" if counters else "liboracle.so")
    if name not in _libs:
        path = os.path.join(BUILD_DIR, name)
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        L.orc_scene_build.restype = C.c_void_p
        L.orc_scene_build.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_char_p]
        L.orc_scene_load.restype = C.c_void_p
        L.orc_scene_load.argtypes = [C.c_char_p, C.c_int]
        L.orc_scene_dump.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_scene_free.argtypes = [C.c_void_p]
        L.orc_scene_info.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_set_sampling.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_render.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_camera_rays.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
        L.orc_path_radiance.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_perlin_table.argtypes = [C.c_uint32, C.c_void_p]
        L.orc_perlin_get.restype = C.c_double
        L.orc_perlin_get.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_xorshift_words.restype = None
        L.orc_xorshift_words.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_gen_index.restype = C.c_int64
        L.orc_gen_index.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_void_p]
        L.orc_reflectance.restype = C.c_double
        L.orc_reflectance.argtypes = [C.c_double, C.c_double]
        L.orc_luminance.restype = C.c_double
        L.orc_luminance.argtypes = [C.c_double, C.c_double, C.c_double]
        L.orc_color_byte.argtypes = [C.c_double]
        L.orc_scene_rng.restype = C.c_double
        L.orc_scene_rng.argtypes = [C.c_uint64, C.c_int]
        L.orc_texture_value.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.orc_background.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sample_dirs.argtypes = [C.c_int, C.c_uint64, C.c_size_t, C.c_void_p]
        L.orc_scatter.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_load_hdr.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _libs[name] = L
    return _libs[name]


class OracleError(RuntimeError):
    pass


class OracleScene:
    PHILOX, WYRAND = 0, 1

    def __init__(self, handle, L):
        self._L = L
        if not handle:
            raise OracleError(L.orc_last_error().decode())
        self._h = C.c_void_p(handle)
        info = (C.c_int64 * 6)()
        self._check(L.orc_scene_info(self._h, info))
        self.width, self.height, self.prim_count, self.medium_count, self.max_depth, self.batch_size = [int(v) for v in info]

    def _check(self, rc):
        if rc != 0:
            raise OracleError(self._L.orc_last_error().decode())

    @classmethod
    def build(cls, name, width=0, spp=0, max_depth=0, scene_seed=1, include_media=True, n_spheres=0, asset_dir=None, counters=False):
        L = lib(counters)
        return cls(L.orc_scene_build(name.encode(), width, spp, max_depth, scene_seed, int(include_media), n_spheres,
                                     (asset_dir or ASSET_DIR).encode()), L)

    @classmethod
    def load(cls, path, include_media=True, counters=False):
        L = lib(counters)
        return cls(L.orc_scene_load(str(path).encode(), int(include_media)), L)

    def dump(self, path):
        self._check(self._L.orc_scene_dump(self._h, str(path).encode()))

    def set_sampling(self, confidence=0.95, tolerance=0.0, batch_size=1, max_samples=0, max_depth=0):
        self._check(self._L.orc_scene_set_sampling(self._h, confidence, tolerance, batch_size, max_samples, max_depth))
        if max_depth:
            self.max_depth = max_depth
        return self

    def render(self, seed=1, rng=PHILOX, threads=0, sample_begin=0, pixel_begin=0, pixel_end=-1):
        """-> (rgb float64 [H,W,3], counts uint32 [H,W], counters dict, seconds)"""
        n = self.width * self.height
        rgb = np.zeros((n, 3), dtype=np.float64)
        counts = np.zeros(n, dtype=np.uint32)
        ctr = (C.c_uint64 * 5)()
        sec = C.c_double()
        self._check(self._L.orc_render(self._h, seed, rng, threads, sample_begin, pixel_begin, pixel_end, rgb.ctypes.data, counts.ctypes.data,
                                       ctr, C.byref(sec)))
        names = ["rays", "box_tests", "prim_tests", "tex_fetches", "paths"]
        return (rgb.reshape(self.height, self.width, 3), counts.reshape(self.height, self.width), dict(zip(names, [int(v) for v in ctr])),
                sec.value)

    def trace_closest(self, rays, tmin=0.001, tmax=1.7976931348623157e308, extra=False):
        rays = np.ascontiguousarray(rays, dtype=np.float64)
        n = rays.shape[0]
        ids = np.empty(n, dtype=np.int32)
        t = np.empty(n, dtype=np.float64)
        ex = np.empty((n, 8), dtype=np.float64) if extra else None
        self._check(self._L.orc_trace_closest(self._h, rays.ctypes.data, n, tmin, tmax, ids.ctypes.data, t.ctypes.data,
                                              ex.ctypes.data if extra else None))
        return (ids, t, ex) if extra else (ids, t)

    def camera_rays(self, seed, sample, pixel_begin=0, pixel_end=None):
        pixel_end = self.width * self.height if pixel_end is None else pixel_end
        out = np.empty((pixel_end - pixel_begin, 7), dtype=np.float64)
        self._check(self._L.orc_camera_rays(self._h, seed, sample, pixel_begin, pixel_end, out.ctypes.data))
        return out

    def path_radiance(self, seed, pixels, samples):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint32)
        samples = np.ascontiguousarray(samples, dtype=np.uint32)
        out = np.empty((pixels.shape[0], 3), dtype=np.float64)
        self._check(self._L.orc_path_radiance(self._h, seed, pixels.ctypes.data, samples.ctypes.data, pixels.shape[0], out.ctypes.data))
        return out

    def texture_value(self, tex, u, v, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        out = np.empty(3, dtype=np.float64)
        self._check(self._L.orc_texture_value(self._h, tex, u, v, p.ctypes.data, out.ctypes.data))
        return out

    def background(self, direction):
        d = np.ascontiguousarray(direction, dtype=np.float64)
        out = np.empty(3, dtype=np.float64)
        self._check(self._L.orc_background(self._h, d.ctypes.data, out.ctypes.data))
        return out

    def scatter(self, mat, seed, pixel, sample, ray, p, n, u=0.0, v=0.0):
        """-> (scattered bool, direction[3], attenuation[3], emitted[3])"""
        inp = np.array(list(ray) + list(p) + list(n) + [u, v], dtype=np.float64)
        out = np.empty(10, dtype=np.float64)
        self._check(self._L.orc_scatter(self._h, mat, seed, pixel, sample, inp.ctypes.data, out.ctypes.data))
        return bool(out[0]), out[1:4].copy(), out[4:7].copy(), out[7:10].copy()

    def close(self):
        if getattr(self, "_h", None):
            self._L.orc_scene_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_triangle_interval(respect, counters=False):
    """False = the reference's Triangle::hit (ignores ray_t, Q14); True = the order-independent rule the device uses."""
    lib(counters).orc_set_triangle_interval(int(bool(respect)))


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().orc_philox(c, k, o)
    return [int(v) for v in o]


def perlin_table(seed=0):
    out = np.empty(256, dtype=np.uint8)
    lib().orc_perlin_table(seed, out.ctypes.data)
    return out


def sample_dirs(which, seed, n):
    out = np.empty((n, 3), dtype=np.float64)
    if lib().orc_sample_dirs(which, seed, n, out.ctypes.data) != 0:
        raise OracleError(lib().orc_last_error().decode())
    return out


def load_hdr(path):
    w, h = C.c_int(), C.c_int()
    L = lib()
    if L.orc_load_hdr(str(path).encode(), C.byref(w), C.byref(h), None) != 0:
        raise OracleError(L.orc_last_error().decode())
    out = np.empty((h.value, w.value, 3), dtype=np.float32)
    L.orc_load_hdr(str(path).encode(), C.byref(w), C.byref(h), out.ctypes.data)
    return out
