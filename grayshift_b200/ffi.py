"""ctypes declarations of include/grayshift.h and include/grayshift_host.h.

The Python layer is plumbing for tests and bench.py (PyTorch supplies device memory, streams and
torch.distributed); the product is the C ABI.  There is no fallback: a missing shared library raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
REPO_ROOT = os.path.dirname(_HERE)
ASSET_DIR = os.path.join(REPO_ROOT, "assets")

GS_ABI_VERSION = 2
GS_OK, GS_ERR_INVALID, GS_ERR_NO_DEVICE, GS_ERR_CUDA, GS_ERR_NCCL, GS_ERR_UNSUPPORTED = range(6)

d3 = C.c_double * 3


class gs_sphere(C.Structure):
    _fields_ = [("center", d3), ("radius", C.c_double), ("path", d3), ("material", C.c_int32), ("prim_id", C.c_int32),
                ("moving", C.c_int32), ("order", C.c_int32), ("uv_rotation", C.c_double * 2)]


class gs_quad(C.Structure):
    _fields_ = [("q", d3), ("u", d3), ("v", d3), ("material", C.c_int32), ("prim_id", C.c_int32), ("order", C.c_int32),
                ("reserved", C.c_int32)]


class gs_triangle(C.Structure):
    _fields_ = [("a", d3), ("b", d3), ("c", d3), ("material", C.c_int32), ("prim_id", C.c_int32), ("order", C.c_int32),
                ("reserved", C.c_int32)]


class gs_bvh_node(C.Structure):
    _fields_ = [("bounds_min", d3), ("bounds_max", d3), ("left", C.c_int32), ("right", C.c_int32), ("first_ref", C.c_int32),
                ("ref_count", C.c_int32)]


class gs_medium(C.Structure):
    _fields_ = [("sphere_first", C.c_int32), ("sphere_count", C.c_int32), ("quad_first", C.c_int32), ("quad_count", C.c_int32),
                ("density", C.c_double), ("phase_material", C.c_int32), ("medium_index", C.c_int32)]


class gs_material(C.Structure):
    _fields_ = [("kind", C.c_int32), ("texture", C.c_int32), ("albedo", d3), ("param", C.c_double)]


class gs_texture(C.Structure):
    _fields_ = [("kind", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("reserved", C.c_int32), ("color", d3), ("scale", C.c_double)]


class gs_image(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("format", C.c_int32), ("reserved", C.c_int32), ("pixels", C.c_void_p)]


class gs_scene_desc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("reserved", C.c_uint32),
                ("spheres", C.POINTER(gs_sphere)), ("n_spheres", C.c_size_t),
                ("quads", C.POINTER(gs_quad)), ("n_quads", C.c_size_t),
                ("nodes", C.POINTER(gs_bvh_node)), ("n_nodes", C.c_size_t),
                ("prim_refs", C.POINTER(C.c_uint32)), ("n_prim_refs", C.c_size_t),
                ("boundary_spheres", C.POINTER(gs_sphere)), ("n_boundary_spheres", C.c_size_t),
                ("boundary_quads", C.POINTER(gs_quad)), ("n_boundary_quads", C.c_size_t),
                ("media", C.POINTER(gs_medium)), ("n_media", C.c_size_t),
                ("materials", C.POINTER(gs_material)), ("n_materials", C.c_size_t),
                ("textures", C.POINTER(gs_texture)), ("n_textures", C.c_size_t),
                ("images", C.POINTER(gs_image)), ("n_images", C.c_size_t),
                ("perlin_perm", C.POINTER(C.c_uint8)),
                ("triangles", C.POINTER(gs_triangle)), ("n_triangles", C.c_size_t)]


class gs_camera(C.Structure):
    _fields_ = [("image_width", C.c_int32), ("image_height", C.c_int32), ("max_depth", C.c_uint32), ("background_kind", C.c_int32),
                ("center", d3), ("starting_pixel_pos", d3), ("pixel_delta_u", d3), ("pixel_delta_v", d3),
                ("defocus_angle", C.c_double), ("defocus_disk_u", d3), ("defocus_disk_v", d3), ("background_color", d3),
                ("hdri_image", C.c_int32), ("reserved", C.c_int32), ("hdri_rotation", d3)]


class gs_sample_settings(C.Structure):
    _fields_ = [("confidence", C.c_double), ("tolerance", C.c_double), ("batch_size", C.c_uint32), ("max_samples", C.c_uint32)]


class gs_stats(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("rays", C.c_uint64), ("tex_fetches", C.c_uint64), ("node_visits", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64), ("exchange_ms", C.c_double)]


# every symbol include/grayshift.h declares (tests check the library exports all of them)
CUDA_SYMBOLS = {
    "gs_last_error": (C.c_char_p, []),
    "gs_abi_version": (C.c_int, []),
    "gs_device_count": (C.c_int, []),
    "gs_scene_create": (C.c_int, [C.POINTER(gs_scene_desc), C.c_int, C.POINTER(C.c_void_p)]),
    "gs_scene_destroy": (None, [C.c_void_p]),
    "gs_render": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.POINTER(gs_sample_settings), C.c_uint64, C.c_void_p, C.c_void_p,
                            C.POINTER(gs_stats)]),
    "gs_render_multi": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(gs_camera), C.POINTER(gs_sample_settings), C.c_uint64, C.c_void_p,
                                  C.c_void_p, C.POINTER(gs_stats)]),
    "gs_render_accumulate": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]),
    "gs_render_adaptive_accumulate": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.POINTER(gs_sample_settings), C.c_uint64, C.c_uint32,
                                                C.c_uint32, C.c_void_p, C.c_void_p]),
    "gs_scene_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "gs_scene_get_info": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(C.c_int64)]),
    "gs_resolve": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gs_resolve_rgb8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "gs_stats_read": (C.c_int, [C.c_void_p, C.POINTER(gs_stats)]),
    "gs_tonemap": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "gs_trace_closest": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "gs_trace_closest_fast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "gs_camera_rays": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.c_uint64, C.c_uint32, C.c_int32, C.c_int32, C.c_void_p]),
    "gs_path_radiance": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
}

HOST_SYMBOLS = {
    "gsh_last_error": (C.c_char_p, []),
    "gsh_scene_build": (C.c_void_p, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int, C.c_int, C.c_char_p]),
    "gsh_scene_load": (C.c_void_p, [C.c_char_p, C.c_int]),
    "gsh_scene_dump": (C.c_int, [C.c_void_p, C.c_char_p]),
    "gsh_scene_free": (None, [C.c_void_p]),
    "gsh_scene_flatten": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "gsh_scene_desc": (C.POINTER(gs_scene_desc), [C.c_void_p]),
    "gsh_scene_camera": (C.c_int, [C.c_void_p, C.POINTER(gs_camera), C.POINTER(gs_sample_settings)]),
    "gsh_scene_set_sampling": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_uint32, C.c_uint32, C.c_uint32]),
    "gsh_scene_render_ppm": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.c_uint64, C.POINTER(gs_stats)]),
    "gsh_bvh_build": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(gs_bvh_node), C.POINTER(C.c_size_t), C.c_void_p]),
    "gsh_perlin_table": (None, [C.c_uint32, C.c_void_p]),
    "gsh_scene_rng": (C.c_double, [C.c_uint64, C.c_int]),
}

_cuda = None
_host = None


def _bind(lib, table):
    for name, (res, args) in table.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    return lib


def cuda_lib():
    """libgrayshift_cuda.so (the product).  Raises if it has not been built: there is no fallback."""
    global _cuda
    if _cuda is None:
        # GS_CUDA_LIB: A/B runs of differently tuned builds of the same library (scripts/build_variants.sh); still no fallback
        path = os.environ.get("GS_CUDA_LIB") or os.path.join(LIB_DIR, "libgrayshift_cuda.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _cuda = _bind(C.CDLL(path, mode=C.RTLD_GLOBAL), CUDA_SYMBOLS)
        if _cuda.gs_abi_version() != GS_ABI_VERSION:
            raise RuntimeError("libgrayshift_cuda.so ABI version mismatch")
    return _cuda


def host_lib():
    """libgrayshift_host.so (scene API mirror, flatten, BVH build)."""
    global _host
    if _host is None:
        cuda_lib()  # dependency, loaded globally first
        path = os.path.join(LIB_DIR, "libgrayshift_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _host = _bind(C.CDLL(path), HOST_SYMBOLS)
    return _host


class GrayshiftError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[{code}] {message}")
        self.code = code


def check(rc):
    if rc != 0:
        raise GrayshiftError(rc, cuda_lib().gs_last_error().decode())


def check_host(rc):
    if rc != 0:
        raise GrayshiftError(rc, host_lib().gsh_last_error().decode())
