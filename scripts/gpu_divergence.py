"""Reads the lane-occupancy counters of the GS_DIVERGENCE_STATS diagnostic build of the megakernel (built by
scripts/build_variants.sh divstats "-DGS_DIVERGENCE_STATS", selected with GS_CUDA_LIB).
usage: GS_CUDA_LIB=.../variants/divstats/libgrayshift_cuda.so python scripts/gpu_divergence.py [cfg ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402
from grayshift_b200 import ffi  # noqa: E402

lib = ffi.cuda_lib()
lib.gs_debug_counters.restype = C.c_int
lib.gs_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong)]
SIZES = {"c4": (800, 16), "c1": (400, 100), "c3": (600, 32), "c5": (960, 4), "c2": (1280, 16)}
for cfg in sys.argv[1:] or ["c4", "c1", "c3", "c5"]:
    width, spp = SIZES[cfg]
    dev = gs.DeviceScene(gs.HostScene.build(cfg, width=width, spp=spp), device=0)
    out = (C.c_ulonglong * 16)()
    dev.render(seed=1, want_counts=False)
    lib.gs_debug_counters(dev._h, out)   # discard: first render includes warm-up effects on nothing, but keeps the protocol simple
    _, _, st = dev.render(seed=2, want_counts=False)
    # gs_render already read and reset slots 0..3; the diagnostic slots are still there
    lib.gs_debug_counters(dev._h, out)
    c = list(out)
    it, step, wait, done, lr, ls, lq, ld, pb, alive = c[4:14]
    print(f"{cfg}: rays {st.rays} paths {st.paths} | phase-B entries {pb}, lanes with a ray at entry {alive / max(pb, 1):.1f}")
    print(f"   node loop: {it} warp iterations = {it / max(st.rays, 1) * 32:.1f} lane-slots/ray; per iteration: stepping {step / it:.2f}, waiting on a leaf {wait / it:.2f}, "
          f"finished/no ray {done / it:.2f}; node steps per ray {step / max(st.rays, 1):.2f}")
    print(f"   leaf rounds: {lr} = {lr / max(st.rays, 1) * 32:.1f} lane-slots/ray; per round: spheres {ls / lr:.2f}, quads {lq / lr:.2f}, idle {ld / lr:.2f}; "
          f"leaf tests per ray {(ls + lq) / max(st.rays, 1):.2f}", flush=True)
    dev.close()
