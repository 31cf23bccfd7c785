"""GPU parity tests, round 2: the gaps the round-1 review named.

  * the RENDER kernels' own build of the traversal (FMA contraction, approximate f64 division / square root) on the 2^20-ray
    batches (gs_trace_closest_fast), for both trees;
  * the eight-wide compressed tree against the oracle and against the binary tree;
  * image gates at the BASELINE.json sizes (C1 400x225x100, C2 1280x720x32, C3 600x600x64, C4 800x800x64) and on the real
    10^6-sphere scene of config 5;
  * the constant medium on the device against Beer-Lambert in closed form, thin and thick (volume.rs:32-63);
  * the adaptive sampler's per-pixel sample counts on the Book-2 final scene with main.rs's own settings;
  * the hollow-glass (negative radius) sphere, sphere (u, v) under RotateY;
  * gs_render_multi (one call, several GPUs).
"""
import ctypes as C
import os

import numpy as np
import pytest

import grayshift_b200 as gs
import oracle_ffi as orc
from conftest import tonemap_np
from graphs import Graph
from grayshift_b200 import ffi
from test_gpu_parity import REL_T, assert_hits_equal, image_gate, ray_batch

pytestmark = pytest.mark.gpu


# ---- closest hits: render-kernel arithmetic, eight-wide tree ---------------------------------------------------------------
@pytest.mark.parametrize("width", [2, 8])
@pytest.mark.parametrize("name,kw", [("c1", {}), ("c3", {}), ("c4", {"width": 400}), ("scene5", {"width": 400}), ("c5", {"n_spheres": 200000, "width": 480})])
def test_render_build_of_the_traversal_picks_the_oracles_primitives(name, kw, width):
    """gs_trace_closest_fast = the closest-hit batch kernel compiled in the render kernels' translation unit: what the timed
    code really computes.  Same 2^20 rays as the parity build's test; the parity build must stay bit-exact, the render build is
    allowed a handful of grazing rays (reported) and must keep t within 1e-5 relative on the others."""
    host = gs.HostScene.build(name, include_media=False, **kw)
    dev = gs.DeviceScene(host)
    dev.set_option("bvh_width", width)
    assert dev.info("bvh_width") == width
    ora = orc.OracleScene.build(name, include_media=False, **kw)
    rays = ray_batch(ora, 1 << 20)
    oid, ot = ora.trace_closest(rays)
    # parity build on this tree: bit-exact ids
    gid, gt = dev.trace_closest(rays)
    assert np.array_equal(gid, oid), f"parity build, width {width}: {int((gid != oid).sum())} ids differ"
    hit = oid >= 0
    assert np.all(np.abs(gt[hit] - ot[hit]) <= REL_T * np.abs(ot[hit]))
    # render build
    fid, ft = dev.trace_closest(rays, fast=True)
    bad = np.nonzero(fid != oid)[0]
    print(f"{name} width {width}: render-build id mismatches {bad.size} of {rays.shape[0]}")
    assert bad.size <= 4, f"render build: {bad.size} ids differ, e.g. rays {bad[:5]}: gpu {fid[bad[:5]]} oracle {oid[bad[:5]]}"
    same = (fid == oid) & hit
    rel = np.abs(ft[same] - ot[same]) / np.abs(ot[same])
    assert rel.max(initial=0.0) <= REL_T, f"render build: max relative t error {rel.max():.3e}"
    # a mismatch is only acceptable as a near-tie: the render build's t must sit within 1e-9 relative of the oracle's closest t
    for k in bad:
        assert oid[k] >= 0 and fid[k] >= 0 and abs(ft[k] - ot[k]) <= 1e-9 * abs(ot[k]), f"ray {k}: gpu ({fid[k]}, {ft[k]}) oracle ({oid[k]}, {ot[k]})"


def test_closest_hit_on_the_million_sphere_scene():
    """config 5 as BASELINE.json states it: 10^6 spheres (the eight-wide tree is the default at this size)"""
    kw = dict(n_spheres=1000000, width=480)
    host = gs.HostScene.build("c5", include_media=False, **kw)
    dev = gs.DeviceScene(host)
    assert dev.info("bvh_width") == 8 and dev.info("nodes8") > 100000
    ora = orc.OracleScene.build("c5", include_media=False, **kw)
    rays = ray_batch(ora, 1 << 18)
    hit_frac, _ = assert_hits_equal(dev, ora, rays)
    assert hit_frac > 0.3
    fid, _ = dev.trace_closest(rays, fast=True)
    oid, _ = ora.trace_closest(rays)
    assert int((fid != oid).sum()) <= 2
    dev.set_option("bvh_width", 2)
    assert_hits_equal(dev, ora, rays[: 1 << 16])


@pytest.mark.parametrize("name,kw", [("c4", {"width": 200, "spp": 16}), ("c1", {"width": 200, "spp": 16}), ("scene7", {"width": 120, "spp": 16})])
def test_both_trees_render_the_same_paths(name, kw):
    """Same Philox streams, same closest hits whatever the tree: the two images differ by float summation order only
    (media scenes included: the medium draws do not depend on the traversal)."""
    host = gs.HostScene.build(name, **kw)
    dev = gs.DeviceScene(host)
    dev.set_option("bvh_width", 2)
    a, ca, sa = dev.render(seed=5)
    dev.set_option("bvh_width", 8)
    b, cb, sb = dev.render(seed=5)
    assert np.array_equal(ca, cb) and sa.paths == sb.paths
    # (config 1 and scene 7 trace exactly the same rays with both kernels; on config 4 a few paths in 10^4 part ways at a decision
    # boundary -- the two kernel variants are compiled separately and round a handful of fp32 shading operations differently --
    # and a path that parts ways runs to another depth: 3e-5 of the rays with IEEE fp32 square roots, 2.3e-4 with the approximate ones)
    assert abs(int(sa.rays) - int(sb.rays)) <= 5e-4 * int(sa.rays)
    same = np.isclose(a, b, rtol=1e-3, atol=1e-5).all(axis=2).mean()
    assert same > 0.998, f"only {same:.4f} of the pixels agree between the binary and the eight-wide tree"


# ---- images at the BASELINE.json sizes ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,spp,size", [("c1", 100, (400, 225)), ("c2", 32, (1280, 720)), ("c3", 64, (600, 600)), ("c4", 64, (800, 800))])
def test_image_gate_at_the_baseline_size(name, spp, size):
    """RMSE(gpu, oracle B) <= 1.25 x RMSE(oracle A, oracle B) in the reference's output domain and in linear radiance, mean signed
    error within 3 sigma / sqrt(N), at the full resolution of the config (C1 with its full 100 spp)."""
    host = gs.HostScene.build(name, spp=spp)
    cam, _ = host.camera()
    assert (cam.image_width, cam.image_height) == size
    err, noise, psnr = image_gate(name, spp)
    print(f"{name} {size[0]}x{size[1]} {spp} spp: RMSE(gpu, oracle) {err:.4f}, RMSE(oracle, oracle) {noise:.4f}, PSNR {psnr:.1f} dB")


def test_image_gate_on_the_million_sphere_scene():
    err, noise, psnr = image_gate("c5", 4, width=480, n_spheres=1000000)
    print(f"c5 10^6 spheres 480x270 4 spp: RMSE(gpu, oracle) {err:.4f}, RMSE(oracle, oracle) {noise:.4f}, PSNR {psnr:.1f} dB")


# ---- constant medium vs Beer-Lambert --------------------------------------------------------------------------------------------
def _transmittance_scene(tmp_path, boundary, density, look_from, look_at, vfov, width, spp):
    """white background, a medium whose 'phase function' is a black light (emits 0, never scatters: the path ends with 0) and one
    tiny far-away sphere so that the surface list is not empty: a pixel's mean is the fraction of paths without a medium event"""
    g = Graph()
    black = g.light_color([0, 0, 0])
    far = g.sphere([1e5, 1e5, 1e5], 1e-3, g.lambertian_color([0.5, 0.5, 0.5]))
    med = g.medium(boundary(g), density, black)
    g.background_solid([1, 1, 1])
    g.camera(width=width, batch=spp, max_samples=0, tolerance=0.0, depth=4, vfov=vfov, look_from=look_from, look_at=look_at)
    return g.write(str(tmp_path / "m.gsg"), g.list([far, med]))


def _pixel_center_rays(dev):
    """camera rays through the pixel centres in f64 (Camera::new's pixel00 + i du + j dv, camera.rs:69-74)"""
    c = dev.cam
    i, j = np.meshgrid(np.arange(c.image_width), np.arange(c.image_height))
    p = (np.array(c.starting_pixel_pos)[None, None, :] + i[..., None] * np.array(c.pixel_delta_u)[None, None, :]
         + j[..., None] * np.array(c.pixel_delta_v)[None, None, :])
    o = np.array(c.center)
    return o, p - o


def test_thin_medium_matches_beer_lambert(tmp_path):
    """the world fog of main.rs:728-738 (rho = 1e-4, r = 5000) seen from inside, off centre: transmittance along each pixel's ray
    is exp(-rho * distance to the boundary) -- 0.55 ... 0.67 here -- to 0.3 % on 10^7 paths, pixel by pixel within its noise"""
    rho, r = 1e-4, 5000.0
    path = _transmittance_scene(tmp_path, lambda g: g.sphere([0, 0, 0], r, g.dielectric(1.5)), rho, look_from=(800, -300, 500),
                                look_at=(0, 100, -4000), vfov=60.0, width=100, spp=1000)
    dev = gs.DeviceScene(gs.HostScene.load(path))
    rgb, counts, st = dev.render(seed=3)
    assert st.paths == 100 * 100 * 1000
    o, d = _pixel_center_rays(dev)
    dn = d / np.linalg.norm(d, axis=2, keepdims=True)
    b = (dn * o).sum(axis=2)
    dist = -b + np.sqrt(b * b - (o @ o - r * r))  # exit distance from inside the sphere
    expect = np.exp(-rho * dist)
    got = rgb[..., 0].astype(np.float64)
    assert 0.5 < expect.min() < expect.max() < 0.7
    assert abs(got.mean() / expect.mean() - 1.0) < 3e-3, (got.mean(), expect.mean())
    z = (got - expect) / np.sqrt(expect * (1 - expect) / 1000.0)
    assert abs(z.mean()) < 0.05 and 0.9 < z.std() < 1.1, (z.mean(), z.std())
    ora = orc.OracleScene.load(path)
    ora.set_sampling(batch_size=200, max_samples=0)
    a, _, _, _ = ora.render(seed=5, rng=orc.OracleScene.PHILOX)
    assert abs(a[..., 0].mean() / expect.mean() - 1.0) < 5e-3  # the oracle obeys the same law (fewer samples)


def test_thick_medium_matches_beer_lambert(tmp_path):
    """rho = 0.2 (the fog ball of main.rs:714-726) in a box 10 thick (cube boundary: the quad path of ConstantMedium::hit), viewed
    nearly head-on: transmittance exp(-rho * 10 / cos) ~ 0.13, to 0.3 % on 10^7 paths"""
    rho, L = 0.2, 10.0
    path = _transmittance_scene(tmp_path, lambda g: g.cube([-500, -500, -L / 2], [500, 500, L / 2], g.dielectric(1.5)), rho,
                                look_from=(30, 20, 400), look_at=(0, 0, 0), vfov=20.0, width=100, spp=1000)
    dev = gs.DeviceScene(gs.HostScene.load(path))
    rgb, counts, st = dev.render(seed=4)
    o, d = _pixel_center_rays(dev)
    cos = np.abs(d[..., 2]) / np.linalg.norm(d, axis=2)
    expect = np.exp(-rho * L / cos)
    got = rgb[..., 0].astype(np.float64)
    assert 0.10 < expect.min() < expect.max() < 0.14
    assert abs(got.mean() / expect.mean() - 1.0) < 3e-3, (got.mean(), expect.mean())
    z = (got - expect) / np.sqrt(expect * (1 - expect) / 1000.0)
    assert abs(z.mean()) < 0.05 and 0.9 < z.std() < 1.1, (z.mean(), z.std())


def test_fog_sphere_seen_from_outside_matches_beer_lambert(tmp_path):
    """the single-sphere boundary from outside, near the limb where the chord is ~10 long (both roots of one quadratic)"""
    rho, r = 0.2, 70.0
    bimp = np.sqrt(r * r - 25.0)  # impact parameter of a chord of length 10
    path = _transmittance_scene(tmp_path, lambda g: g.sphere([0, 0, 0], r, g.dielectric(1.5)), rho, look_from=(bimp, 0, 5000),
                                look_at=(bimp, 0, 0), vfov=0.002, width=64, spp=2500)
    dev = gs.DeviceScene(gs.HostScene.load(path))
    rgb, _, _ = dev.render(seed=6)
    o, d = _pixel_center_rays(dev)
    dn = d / np.linalg.norm(d, axis=2, keepdims=True)
    b = (dn * o).sum(axis=2)
    disc = b * b - (o @ o - r * r)
    chord = np.where(disc > 0, 2 * np.sqrt(np.maximum(disc, 0)), 0.0)
    expect = np.exp(-rho * chord)
    got = rgb[..., 0].astype(np.float64)
    assert abs(got.mean() / expect.mean() - 1.0) < 3e-3, (got.mean(), expect.mean())
    z = (got - expect) / np.sqrt(np.maximum(expect * (1 - expect), 1e-9) / 2500.0)
    assert abs(z.mean()) < 0.06 and 0.88 < z.std() < 1.12, (z.mean(), z.std())


# ---- adaptive sampler on the Book-2 final scene -----------------------------------------------------------------------------------
def test_adaptive_counts_on_the_final_scene_follow_the_oracle():
    """scene 8 = final_scene(400, 50) with main.rs:764-781's own SampleSettings {.95, .25, 32, 1000}: the device's per-pixel sample
    counts against the oracle's under the shared Philox streams.  Media draw their free-flight numbers at different points of the
    stream in the two programs, so the counts agree statistically, not pixel by pixel: same distribution over the
    batch counts, same mean, and pixel-wise agreement where the stop test is not marginal."""
    kw = dict(width=160)
    host = gs.HostScene.build("scene8", **kw)
    dev = gs.DeviceScene(host)
    ora = orc.OracleScene.build("scene8", **kw)
    rgb, counts, st = dev.render(seed=7)
    a, ca, _, _ = ora.render(seed=7, rng=orc.OracleScene.PHILOX)
    assert counts.min() >= 32 and counts.max() <= 1024 and np.all(counts % 32 == 0) and np.all(ca % 32 == 0)
    assert st.paths == counts.sum()
    assert abs(counts.mean() - ca.mean()) < 0.03 * ca.mean(), (counts.mean(), ca.mean())
    hg, ho = np.bincount(counts.ravel() // 32, minlength=33) / counts.size, np.bincount(ca.ravel() // 32, minlength=33) / ca.size
    assert np.abs(hg - ho).sum() < 0.06, f"batch-count histograms differ by {np.abs(hg - ho).sum():.3f} in total variation x2"
    assert (np.abs(counts.astype(np.int64) - ca.astype(np.int64)) <= 32).mean() > 0.65   # (0.74 measured: the media decorrelate the paths)
    err = np.sqrt(np.mean((tonemap_np(rgb.astype(np.float64)) - tonemap_np(a)) ** 2))
    b, _, _, _ = ora.render(seed=8, rng=orc.OracleScene.PHILOX)
    noise = np.sqrt(np.mean((tonemap_np(b) - tonemap_np(a)) ** 2))
    assert err <= 1.25 * noise, (err, noise)


# ---- negative radius, sphere (u, v) under RotateY ---------------------------------------------------------------------------------
def test_hollow_glass_sphere_keeps_the_signed_radius(tmp_path):
    """Sphere::hit divides by the SIGNED radius (sphere.rs:91): a negative radius turns the normal inward and with it front_face
    and the refraction ratio -- the classic hollow-glass bubble.  Whole paths against the oracle under shared streams."""
    g = Graph()
    glass, ground = g.dielectric(1.5), g.lambertian(g.checker(0.5, g.solid([0.2, 0.3, 0.1]), g.solid([0.9, 0.9, 0.9])))
    items = [g.sphere([0, -100.5, -1], 100, ground), g.sphere([0, 0, -1], 0.5, g.lambertian_color([0.1, 0.2, 0.5])),
             g.sphere([-1.1, 0, -1], 0.5, glass), g.sphere([-1.1, 0, -1], -0.4, glass), g.sphere([1.1, 0, -1], 0.5, g.metal([0.8, 0.6, 0.2], 0.0))]
    g.camera(aspect=16 / 9, width=160, batch=16, depth=20, vfov=40.0, look_from=(-2, 2, 1), look_at=(0, 0, -1))
    p = g.write(str(tmp_path / "hollow.gsg"), g.bvh(items))
    host, ora = gs.HostScene.load(p), orc.OracleScene.load(p)
    dev = gs.DeviceScene(host)
    assert_hits_equal(dev, ora, ray_batch(ora, 1 << 16, seed=3))
    rng = np.random.default_rng(2)
    n = 40000
    px = rng.integers(0, dev.width * dev.height, n).astype(np.uint32)
    sm = rng.integers(0, 64, n).astype(np.uint32)
    gpu = dev.path_radiance(seed=4, pixels=px, samples=sm).astype(np.float64)
    ref = ora.path_radiance(seed=4, pixels=px, samples=sm)
    close = np.all(np.abs(gpu - ref) <= 2e-3 * np.maximum(np.abs(ref), 1e-3) + 1e-5, axis=1)
    assert close.mean() >= 0.985, f"only {close.mean():.4f} of the paths agree"
    assert abs(gpu.mean() - ref.mean()) <= 0.02 * ref.mean()


def test_image_textured_sphere_under_rotate_y(tmp_path):
    """Sphere::hit computes (u, v) from the object-space normal; RotateY::hit only turns the normal afterwards (hittable.rs:179-211):
    a rotated textured globe shows the texture rotated WITH it.  Same-seed paths against the oracle."""
    rng = np.random.default_rng(5)
    w, h = 64, 32
    raw = tmp_path / "tex.raw"
    rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8).tofile(raw)
    g = Graph()
    tex = g.image_texture(g.image("u8", w, h, f"raw:{raw}"))
    globe = g.translate(g.rotate_y(g.sphere([0, 0, 0], 2.0, g.lambertian(tex)), 63.0), [0.5, 0.2, -1.0])
    plain = g.sphere([-4.5, 0, 0], 1.5, g.lambertian(tex))
    spun = g.rotate_y(g.translate(g.rotate_y(g.sphere([0, 0, 0], 1.0, g.light(tex)), -40.0), [4.0, 0.5, 0.0]), 25.0)
    g.camera(aspect=2.0, width=160, batch=8, depth=6, vfov=35.0, look_from=(0, 1, 14), look_at=(0, 0, 0))
    p = g.write(str(tmp_path / "globe.gsg"), g.bvh([globe, plain, spun]))
    host, ora = gs.HostScene.load(p), orc.OracleScene.load(p)
    dev = gs.DeviceScene(host)
    n = 30000
    px = rng.integers(0, dev.width * dev.height, n).astype(np.uint32)
    sm = rng.integers(0, 32, n).astype(np.uint32)
    gpu = dev.path_radiance(seed=9, pixels=px, samples=sm).astype(np.float64)
    ref = ora.path_radiance(seed=9, pixels=px, samples=sm)
    close = np.all(np.abs(gpu - ref) <= 2e-3 * np.maximum(np.abs(ref), 1e-3) + 1e-5, axis=1)
    assert close.mean() >= 0.99, f"only {close.mean():.4f} of the paths agree"


# ---- several GPUs in one call -------------------------------------------------------------------------------------------------
def _render_multi(host, devices, seed, settings=None):
    lib = ffi.cuda_lib()
    scenes = [gs.DeviceScene(host, device=d) for d in devices]
    cam, st = scenes[0].cam, settings or scenes[0].settings
    n = cam.image_width * cam.image_height
    rgb = np.empty((n, 3), dtype=np.float32)
    counts = np.empty(n, dtype=np.uint32)
    stats = ffi.gs_stats()
    handles = (C.c_void_p * len(scenes))(*[s._h for s in scenes])
    ffi.check(lib.gs_render_multi(handles, len(scenes), C.byref(cam), C.byref(st), seed, rgb.ctypes.data, counts.ctypes.data, C.byref(stats)))
    for s in scenes:
        s.close()
    return rgb.reshape(cam.image_height, cam.image_width, 3), counts.reshape(cam.image_height, cam.image_width), stats


@pytest.mark.parametrize("name,kw", [("c4", {"width": 200, "spp": 32}), ("scene6", {"width": 72})])
def test_render_multi_on_one_device_equals_render(name, kw):
    host = gs.HostScene.build(name, **kw)
    a, ca, sa = gs.DeviceScene(host).render(seed=3)
    b, cb, sb = _render_multi(host, [0], seed=3)
    assert np.array_equal(ca, cb) and sa.paths == sb.paths
    assert np.allclose(a, b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,kw", [("c4", {"width": 200, "spp": 33}), ("scene6", {"width": 72})])
def test_render_multi_splits_over_every_gpu_of_the_box(name, kw):
    """fixed count: sample slices per device (33 does not divide evenly), adaptive: interleaved pixels; the first device sums the
    peers' accumulators through peer-mapped memory while it resolves.  Same samples as one GPU: counts identical, radiance equal
    up to float summation order."""
    import torch

    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("one GPU on this box")
    host = gs.HostScene.build(name, **kw)
    a, ca, sa = gs.DeviceScene(host).render(seed=3)
    for g in sorted({2, ndev}):
        b, cb, sb = _render_multi(host, list(range(g)), seed=3)
        assert np.array_equal(ca, cb), f"{g} GPUs: sample counts differ"
        assert sa.paths == sb.paths and abs(int(sa.rays) - int(sb.rays)) <= 1e-4 * int(sa.rays)
        assert np.allclose(a, b, rtol=2e-4, atol=1e-5), f"{g} GPUs: image differs"
        assert sb.exchange_ms > 0.0


@pytest.mark.parametrize("name,kw,width", [("scene6", {"width": 96}, 2), ("scene3", {"width": 160}, 2), ("scene1", {"width": 200}, 2),
                                           ("scene1", {"width": 200}, 8), ("scene11", {"width": 120}, 2)])
def test_adaptive_tail_rendering_batches_ahead_keeps_the_sequential_result(name, kw, width):
    """Few pixels left: the tail kernel renders several of a pixel's batches at once and replays Camera::sample's loop over them
    in order (camera.rs:137-164).  Batch b is the same global sample indices whoever renders it, so counts are identical to the
    one-batch-at-a-time tail and the image differs only by the order of float additions; dropped batches do not show in the stats."""
    host = gs.HostScene.build(name, **kw)
    dev = gs.DeviceScene(host)
    dev.set_option("bvh_width", width)
    dev.set_option("tail_ahead", 0)
    a, ca, sa = dev.render(seed=9)
    dev.set_option("tail_ahead", 1)
    b, cb, sb = dev.render(seed=9)
    assert np.array_equal(ca, cb)
    assert sa.paths == sb.paths == ca.sum() and sa.rays == sb.rays
    assert np.allclose(a, b, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("name,kw", [("scene6", {"width": 96}), ("scene8", {"width": 120}), ("scene1", {"width": 200})])
def test_adaptive_chains_render_the_same_image(name, kw):
    """The stop test is per pixel (camera.rs:125-171), so the pixels can be dealt out to independent chains of rounds on streams of
    their own (the launches of one chain fill the SMs another leaves idle as it ends).  Same samples per pixel whatever the
    number of chains: identical counts and statistics, radiance up to the order of float additions."""
    host = gs.HostScene.build(name, **kw)
    dev = gs.DeviceScene(host)
    dev.set_option("adaptive_chains", 1)
    a, ca, sa = dev.render(seed=5)
    for k in (2, 3, 4):
        dev.set_option("adaptive_chains", k)
        b, cb, sb = dev.render(seed=5)
        assert np.array_equal(ca, cb), k
        assert sa.paths == sb.paths and sa.rays == sb.rays
        assert np.allclose(a, b, rtol=1e-4, atol=1e-6)
        assert sb.kernel_launches > sa.kernel_launches


# ---- NoiseTextures evaluated by the whole warp (noise_turb_coop, gs_shade.cuh) -------------------------------------------------
@pytest.mark.parametrize("name,kw", [("scene3", {"width": 96}), ("scene5", {"width": 96}), ("c4", {"width": 120})])
def test_noise_textures_evaluated_by_the_warp_match_the_per_lane_evaluation(name, kw):
    """The megakernel defers NoiseTexture::value_at (texture.rs:107-130) to the end of its shading phase and serves the pending
    lanes with all 32 (up to eight of them; more evaluate on their own); the per-path hook evaluates it inline, lane by lane.
    Same Philox streams, so pixel means agree up to float rounding: scene 3 / 5 are all marble (warps full of pending lanes),
    config 4 has one marble sphere (one or two pending lanes at a time)."""
    spp = 8
    host = gs.HostScene.build(name, spp=spp, **kw)
    dev = gs.DeviceScene(host)
    rgb, counts, st = dev.render(seed=9)
    n = dev.width * dev.height
    assert np.all(counts == spp)
    px = np.repeat(np.arange(n, dtype=np.uint32), spp)
    sm = np.tile(np.arange(spp, dtype=np.uint32), n)
    per_path = dev.path_radiance(seed=9, pixels=px, samples=sm).astype(np.float64).reshape(n, spp, 3).mean(axis=1)
    got = rgb.reshape(n, 3).astype(np.float64)
    same = np.isclose(got, per_path, rtol=1e-3, atol=1e-5).all(axis=1).mean()
    assert same > 0.998, f"only {same:.4f} of the pixels agree between the megakernel and the per-path hook"
    assert abs(got.mean() - per_path.mean()) <= 1e-4 * per_path.mean()


def test_deferred_noise_on_lights_checkers_and_media(tmp_path):
    """The three ways a NoiseTexture reaches the megakernel's deferred evaluation besides a Lambertian albedo: as a DiffuseLight's
    emission (added, not multiplied, and the path ends there), nested in a CheckeredTexture, and as the albedo of a medium's
    Isotropic phase function.  Megakernel image against the per-path hook (inline evaluation), then both against the oracle."""
    g = Graph()
    marble = g.noise(4.0)
    lamp = g.quad([-2, 3.5, -2], [4, 0, 0], [0, 0, 4], g.light(marble))
    floor = g.quad([-30, -1, -30], [60, 0, 0], [0, 0, 60], g.lambertian(g.checker(0.8, marble, g.solid([0.8, 0.3, 0.2]))))
    ball = g.sphere([0, 0.2, 0], 1.2, g.lambertian(marble))
    smoke = g.medium(g.sphere([2.6, 0.3, 0.5], 1.0, g.lambertian_color([1, 1, 1])), 0.9, g.isotropic(marble))
    g.camera(aspect=1.5, width=96, batch=8, depth=12, vfov=45.0, look_from=(1, 2.5, 7), look_at=(0.5, 0.3, 0))
    g.background_solid([0.25, 0.3, 0.4])
    p = g.write(str(tmp_path / "noise.gsg"), g.bvh([lamp, floor, ball, smoke]))
    host, ora = gs.HostScene.load(p), orc.OracleScene.load(p)
    dev = gs.DeviceScene(host)
    spp = 8
    rgb, counts, _ = dev.render(seed=6)
    n = dev.width * dev.height
    assert np.all(counts == spp)
    px = np.repeat(np.arange(n, dtype=np.uint32), spp)
    sm = np.tile(np.arange(spp, dtype=np.uint32), n)
    per_path = dev.path_radiance(seed=6, pixels=px, samples=sm).astype(np.float64)
    got = rgb.reshape(n, 3).astype(np.float64)
    want = per_path.reshape(n, spp, 3).mean(axis=1)
    same = np.isclose(got, want, rtol=1e-3, atol=1e-5).all(axis=1).mean()
    assert same > 0.998, f"only {same:.4f} of the pixels agree between the megakernel and the per-path hook"
    # media scenes are compared with the oracle statistically (the device draws a medium's free flight before the traversal)
    a, _, _, _ = ora.render(seed=7, rng=orc.OracleScene.PHILOX)
    b, _, _, _ = ora.render(seed=8, rng=orc.OracleScene.PHILOX)
    err = np.sqrt(np.mean((tonemap_np(rgb.astype(np.float64)) - tonemap_np(a)) ** 2))
    noise = np.sqrt(np.mean((tonemap_np(b) - tonemap_np(a)) ** 2))
    assert err <= 1.25 * noise, (err, noise)
