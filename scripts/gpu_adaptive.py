"""Throughput of the adaptive-sampling kernel (Camera::sample with its batch loop, camera.rs:125-171) on the
reference's own scenes, which all use it.  usage: python scripts/gpu_adaptive.py [scene ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

for name in sys.argv[1:] or ["scene9", "scene0", "scene6", "scene11", "scene2"]:
    dev = gs.DeviceScene(gs.HostScene.build(name), device=0)
    best = None
    for rep in range(3):
        rgb, counts, st = dev.render(seed=1 + rep)
        if rep and (best is None or st.kernel_ms < best[0]):
            best = (st.kernel_ms, st.paths, st.rays, float(counts.mean()), int(counts.min()), int(counts.max()))
    s = dev.settings
    print(f"{name}: {dev.width}x{dev.height} batch {s.batch_size} max {s.max_samples} tol {s.tolerance} | {best[0]:.1f} ms, "
          f"{best[1] / best[0] / 1e3:.1f} Mpaths/s, {best[2] / best[0] / 1e3:.1f} Mrays/s, spp mean {best[3]:.1f} min {best[4]} max {best[5]}", flush=True)
    dev.close()
