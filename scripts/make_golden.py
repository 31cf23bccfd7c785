"""Generates tests/golden/*.npz: frozen outputs of the CPU oracle (oracle/, the C++ restatement of the reference) on fixed
inputs.  The reference itself cannot run here (Rust, no toolchain) and holds no tests or fixtures, so these vectors pin
the ORACLE against drift -- they are not outputs of the reference (parity unpinned).  The CPU suite
checks that today's oracle still reproduces them; the GPU suite checks the CUDA path against them without needing the
oracle at run time.   usage: python scripts/make_golden.py   (rewrites tests/golden/)"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ffi as orc  # noqa: E402
from graphs import random_scene  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
N = 2048


def ray_batch(ora, n_total, seed):
    """camera rays + random secondary rays leaving their first hit points (the recipe of tests/test_gpu_parity.py)"""
    rng = np.random.default_rng(seed)
    cams = []
    while sum(c.shape[0] for c in cams) < n_total // 2:
        cams.append(ora.camera_rays(seed=seed, sample=len(cams)))
    cam = np.concatenate(cams)
    cam = cam[rng.permutation(cam.shape[0])[: n_total // 2]]
    ids, t, ex = ora.trace_closest(cam, extra=True)
    p = ex[ids >= 0, 0:3]
    p = p[rng.integers(0, p.shape[0], n_total - cam.shape[0])]
    d = rng.normal(size=p.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[::3] *= rng.uniform(0.1, 30.0, size=(d[::3].shape[0], 1))
    return np.concatenate([cam, np.concatenate([p, d, rng.random((p.shape[0], 1))], 1)])


SCENES = {"c1": dict(width=100), "c3": dict(width=100), "c4": dict(width=100), "scene5": dict(width=100), "scene4": dict(width=64)}
for name, kw in SCENES.items():
    ora = orc.OracleScene.build(name, spp=1, include_media=False, **kw)  # closest SURFACE hit: ConstantMedium::hit draws random numbers
    rays = ray_batch(ora, N, seed=7)
    ids, t = ora.trace_closest(rays)
    cam = ora.camera_rays(seed=5, sample=3)[:256]
    np.savez_compressed(os.path.join(OUT, f"{name}_closest.npz"), rays=rays, ids=ids, t=t, camera_rays_seed5_sample3=cam,
                        build=np.array([kw["width"]]))
    print(f"{name}: {N} rays, {float((ids >= 0).mean()):.2f} hit, {len(np.unique(ids))} distinct primitives")

# a random scene graph (instances, quads, cubes, spheres): the graph file travels with the vectors
with tempfile.TemporaryDirectory() as tmp:
    path = os.path.join(tmp, "g.gsg")
    random_scene(path, seed=11)
    ora = orc.OracleScene.load(path)
    rays = ray_batch(ora, N, seed=9)
    ids, t = ora.trace_closest(rays)
    np.savez_compressed(os.path.join(OUT, "graph11_closest.npz"), rays=rays, ids=ids, t=t, graph=np.frombuffer(open(path, "rb").read(), dtype=np.uint8))
    print(f"graph11: {float((ids >= 0).mean()):.2f} hit")

# one small image and a handful of whole paths under the shared Philox streams (Cornell box: no media)
ora = orc.OracleScene.build("c3", width=48, spp=16)
img, counts, ctr, _ = ora.render(seed=3, rng=orc.OracleScene.PHILOX)
rng = np.random.default_rng(1)
px = rng.integers(0, 48 * 48, 512).astype(np.uint32)
sm = rng.integers(0, 64, 512).astype(np.uint32)
rad = ora.path_radiance(seed=3, pixels=px, samples=sm)
np.savez_compressed(os.path.join(OUT, "c3_image_48x48_16spp_seed3.npz"), rgb=img.astype(np.float32), counts=counts, rays=np.array([ctr["rays"]]),
                    path_pixels=px, path_samples=sm, path_radiance=rad)
print("c3 image: mean", float(img.mean()), "rays", ctr["rays"])
