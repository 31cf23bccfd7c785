"""Reference-semantics work counts per ray: AABB::hit calls, primitive hit calls and texel
fetches per world.hit, measured by the oracle's counting build on the reference-built BVH of each config.
Frozen into bench_constants.json (the roofline's algorithmic bytes per ray)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_ffi as orc  # noqa: E402

CASES = {"c1": dict(width=400, spp=8), "c2": dict(width=640, spp=4), "c3": dict(width=300, spp=8), "c4": dict(width=400, spp=8),
         "c5": dict(width=480, spp=1, n_spheres=1000000)}
out = {}
for name, kw in CASES.items():
    s = orc.OracleScene.build(name, counters=True, **kw)
    _, _, c, sec = s.render(seed=1, rng=orc.OracleScene.WYRAND)
    rays = c["rays"]
    rec = {"box_per_ray": c["box_tests"] / rays, "prim_per_ray": c["prim_tests"] / rays, "tex_per_ray": c["tex_fetches"] / rays,
           "rays_per_path": rays / c["paths"], "sample": f"{s.width}x{s.height}x{kw['spp']}spp, {rays} rays"}
    rec["bytes_per_ray"] = 32 * rec["box_per_ray"] + 32 * rec["prim_per_ray"] + 16 * rec["tex_per_ray"]
    rec["bytes_per_path"] = rec["rays_per_path"] * rec["bytes_per_ray"] + 12
    out[name] = rec
    print(name, json.dumps(rec), flush=True)
json.dump(out, open(os.path.join(ROOT, "bench_constants.json"), "w"), indent=1)
