"""grayshift_b200 -- B200-native (sm_100a) implementation of grayshift's per-pixel path-tracing hot path.

Layout:
  csrc/   hand-written CUDA kernels + the C ABI (include/grayshift.h) -> lib/libgrayshift_cuda.so
  host/   C++ mirror of the reference's Rust scene API, flatten, SAH BVH build -> lib/libgrayshift_host.so
  ffi.py / api.py   ctypes plumbing for tests and bench.py

Importing the package never touches the CPU oracle (oracle/ is test infrastructure).
"""
from .api import DeviceScene, HostScene, fixed_spp, tonemap  # noqa: F401
from .ffi import GrayshiftError, cuda_lib, host_lib  # noqa: F401

__all__ = ["DeviceScene", "HostScene", "GrayshiftError", "cuda_lib", "host_lib", "fixed_spp", "tonemap"]
