"""First-contact GPU probe: parity + throughput of every config at reduced sample counts.
Run on a B200 box:  python scripts/gpu_probe.py [--quick]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import grayshift_b200 as gs  # noqa: E402
import oracle_ffi as orc  # noqa: E402

quick = "--quick" in sys.argv


def secondary_rays(ora, rays, rng):
    """uniformly random rays leaving the first-hit points of `rays`"""
    ids, t, ex = ora.trace_closest(rays, extra=True)
    hit = ids >= 0
    p = ex[hit, 0:3]
    d = rng.normal(size=p.shape)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    out = np.concatenate([p, d, rng.random((p.shape[0], 1))], axis=1)
    return out


def parity(name, **kw):
    host = gs.HostScene.build(name, include_media=False, **kw)
    dev = gs.DeviceScene(host)
    ora = orc.OracleScene.build(name, include_media=False, **kw)
    rng = np.random.default_rng(1)
    n_px = dev.width * dev.height
    pix = rng.choice(n_px, size=min(n_px, 1 << 16), replace=False)
    cam = dev.camera_rays(seed=5, sample=0)
    ocam = ora.camera_rays(seed=5, sample=0)
    cam_err = float(np.abs(cam - ocam).max())
    rays = cam[pix]
    sec = secondary_rays(ora, rays, rng)
    allr = np.concatenate([rays, sec])
    t0 = time.time(); oid, ot = ora.trace_closest(allr); t_or = time.time() - t0
    t0 = time.time(); gid, gt = dev.trace_closest(allr); t_gpu = time.time() - t0
    mism = int((gid != oid).sum())
    hit = (oid >= 0) & (gid == oid)
    rel = float(np.max(np.abs(gt[hit] - ot[hit]) / np.abs(ot[hit]))) if hit.any() else 0.0
    print(json.dumps({"parity": name, "rays": int(allr.shape[0]), "id_mismatch": mism, "max_rel_t": rel, "camera_ray_max_abs_err": cam_err,
                      "hit_frac": float((oid >= 0).mean()), "oracle_s": round(t_or, 3), "gpu_s": round(t_gpu, 3)}), flush=True)
    if mism:
        bad = np.nonzero(gid != oid)[0][:5]
        for b in bad:
            print("   mismatch", int(b), "gpu", int(gid[b]), float(gt[b]), "oracle", int(oid[b]), float(ot[b]), flush=True)


def image(name, spp, **kw):
    host = gs.HostScene.build(name, spp=spp, **kw)
    dev = gs.DeviceScene(host)
    rgb, counts, st = dev.render(seed=1)   # warm-up (module load, clocks)
    rgb, counts, st = dev.render(seed=1)
    res = {"render": name, "w": dev.width, "h": dev.height, "spp": spp, "paths": int(st.paths), "rays": int(st.rays),
           "kernel_ms": round(st.kernel_ms, 3), "total_ms": round(st.total_ms, 3),
           "Mpaths_s": round(st.paths / st.kernel_ms / 1e3, 2), "Mrays_s": round(st.rays / st.kernel_ms / 1e3, 2),
           "rays_per_path": round(st.rays / st.paths, 3), "mean": [round(float(v), 5) for v in rgb.reshape(-1, 3).mean(0)]}
    print(json.dumps(res), flush=True)
    return dev, rgb


def rmse_check(name, spp, **kw):
    dev, rgb = image(name, spp, **kw)
    ora = orc.OracleScene.build(name, spp=spp, **kw)
    a, _, ca, sec = ora.render(seed=1, rng=orc.OracleScene.PHILOX)
    b, _, _, _ = ora.render(seed=2, rng=orc.OracleScene.PHILOX)
    tm = lambda x: np.clip(np.sqrt(np.maximum(x, 0)), 0, 0.999)
    noise = float(np.sqrt(np.mean((tm(a) - tm(b)) ** 2)))
    err = float(np.sqrt(np.mean((tm(rgb.astype(np.float64)) - tm(b)) ** 2)))
    same = float(np.sqrt(np.mean((tm(rgb.astype(np.float64)) - tm(a)) ** 2)))
    print(json.dumps({"rmse": name, "spp": spp, "oracle_AB": round(noise, 5), "gpu_vs_B": round(err, 5), "gpu_vs_A_same_seed": round(same, 5),
                      "mean_gpu": float(rgb.mean()), "mean_oracle_A": float(a.mean()), "mean_oracle_B": float(b.mean()),
                      "oracle_s": round(sec, 2), "oracle_rays": ca["rays"]}), flush=True)


if __name__ == "__main__":
    print("devices", gs.cuda_lib().gs_device_count(), flush=True)
    parity("c3", width=200)
    parity("c1")
    parity("c4", width=200)
    rmse_check("c3", 32, width=120)
    rmse_check("c1", 32, width=200)
    rmse_check("c2", 16, width=320)
    rmse_check("c4", 32, width=160)
    rmse_check("scene7", 32, width=120)
    rmse_check("scene3", 32, width=200)
    if not quick:
        image("c1", 100)
        image("c2", 64)
        image("c3", 64)
        image("c4", 64)
        image("c5", 4, n_spheres=1000000, width=1920)
