"""Small renders of every kernel family for compute-sanitizer (run as: compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py).
Fixed-count megakernel (media, Perlin, image textures, triangles), adaptive batch rounds, wavefront pipeline, parity hooks."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grayshift_b200 as gs  # noqa: E402

for name, kw in (("c4", dict(width=48, spp=4)), ("c1", dict(width=64, spp=2)), ("c2", dict(width=64, spp=2)), ("scene7", dict(width=40, spp=2)),
                 ("scene10", dict(width=40, spp=2)), ("scene6", dict(width=24)), ("scene11", dict(width=32))):
    dev = gs.DeviceScene(gs.HostScene.build(name, **kw), device=0)
    rgb, counts, st = dev.render(seed=1)
    assert np.all(np.isfinite(rgb))
    if name in ("c4", "c1"):
        dev.set_option("wavefront", 1)
        dev.set_option("pool_capacity", 2048)
        dev.render(seed=1)
        dev.set_option("wavefront", 0)
        dev.set_option("count_nodes", 1)
        dev.render(seed=1)
    rays = dev.camera_rays(seed=3, sample=0)[:2000]
    dev.trace_closest(rays)
    dev.path_radiance(seed=2, pixels=np.arange(64, dtype=np.uint32), samples=np.zeros(64, dtype=np.uint32))
    print(name, "ok", st.paths, st.rays, flush=True)
    dev.close()
# round 2: the eight-wide tree (forced on small scenes, chosen by size on a sphere cloud), the adaptive sampler's chains of rounds
# and the tail that renders batches ahead (4 and 8 warps per pixel, both trees), the sorted wavefront queues
only = os.environ.get("GS_SANITIZE_ONLY", "")
for name, kw, width in (("c1", dict(width=64, spp=2), 8), ("c4", dict(width=48, spp=2), 8), ("c5", dict(width=64, spp=2, n_spheres=70000), 0),
                        ("scene3", dict(width=96), 2), ("scene3", dict(width=200), 2), ("scene1", dict(width=96), 8), ("scene6", dict(width=32), 2)):
    if only and name not in only.split(","):
        continue
    dev = gs.DeviceScene(gs.HostScene.build(name, **kw), device=0)
    if width:
        dev.set_option("bvh_width", width)
    for chains, ahead in ((1, 1), (2, 1), (4, 0)):
        dev.set_option("adaptive_chains", chains)
        dev.set_option("tail_ahead", ahead)
        rgb, counts, st = dev.render(seed=2)
        assert np.all(np.isfinite(rgb))
    if name == "c4":
        dev.set_option("bvh_width", 2)
        dev.set_option("wavefront", 1)
        dev.set_option("wavefront_sort", 1)
        dev.set_option("pool_capacity", 2048)
        dev.render(seed=1)
    rays = dev.camera_rays(seed=3, sample=0)[:2000]
    dev.trace_closest(rays)
    dev.trace_closest(rays, fast=True)
    print(name, kw, "width", dev.info("bvh_width"), "ok", st.paths, st.rays, flush=True)
    dev.close()
print("done")
