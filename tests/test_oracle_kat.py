"""Known-answer tests that pin the CPU oracle.

The reference ships no tests or golden vectors and cannot be built here (no Rust toolchain), so the
oracle is pinned against answers that do not depend on it: closed-form geometry, published test
vectors (Philox4x32-10 from Random123), an independent decoder (OpenCV for the Radiance file), the
constants and branch rules written in the reference source, and brute force.  Each test names the
reference lines it pins (paths relative to /root/reference/src).
"""
import math
import os

import ctypes as C

import numpy as np
import pytest

import oracle_ffi as orc
from graphs import Graph, random_scene

ASSETS = orc.ASSET_DIR


def scene_from(tmp_path, build):
    g = Graph()
    root = build(g)
    p = g.write(str(tmp_path / "s.gsg"), root)
    return orc.OracleScene.load(p)


# ---- RNG -----------------------------------------------------------------------------------------
def test_philox_random123_vectors():
    # Random123 kat_vectors, philox4x32 10 rounds
    assert orc.philox([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert orc.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert orc.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_scene_rng_is_splitmix64():
    # SplitMix64 reference outputs for seed 1234567: 6457827717110365317, 3203168211198807973
    L = orc.lib()
    assert L.orc_scene_rng(1234567, 0) == (6457827717110365317 >> 11) / 2.0**53
    assert L.orc_scene_rng(1234567, 1) == (3203168211198807973 >> 11) / 2.0**53


# ---- sphere.rs:64-106, hittable.rs:25-43 -----------------------------------------------------------
def test_sphere_hit_known_answer(tmp_path):
    s = scene_from(tmp_path, lambda g: g.bvh([g.sphere([0, 0, 0], 1.0, g.lambertian_color([0.5, 0.5, 0.5]))]))
    ids, t, ex = s.trace_closest(np.array([[0, 0, -5, 0, 0, 1, 0.0]]), extra=True)
    assert ids[0] == 0 and t[0] == 4.0
    np.testing.assert_allclose(ex[0, 0:3], [0, 0, -1])
    np.testing.assert_allclose(ex[0, 3:6], [0, 0, -1])          # front face: normal opposes the ray
    np.testing.assert_allclose(ex[0, 6:8], [0.75, 0.5], atol=1e-15)  # sphere_uv: phi = atan2(1, 0) + pi = 3pi/2, theta = pi/2
    # unnormalised direction: t scales by 1/|d| (Q6)
    ids, t = s.trace_closest(np.array([[0, 0, -5, 0, 0, 2, 0.0]]))
    assert t[0] == 2.0
    # from inside: second root, normal flipped
    ids, t, ex = s.trace_closest(np.array([[0, 0, 0, 1, 0, 0, 0.0]]), extra=True)
    assert t[0] == 1.0
    np.testing.assert_allclose(ex[0, 3:6], [-1, 0, 0])
    # strict interval (interval.rs:33-35): a hit at exactly t_min or t_max is rejected
    ids, t = s.trace_closest(np.array([[0, 0, -5, 0, 0, 1, 0.0]]), tmin=4.0, tmax=5.0)
    assert ids[0] == 0 and t[0] == 6.0 or ids[0] == -1  # near root rejected; far root 6 > tmax -> miss
    assert ids[0] == -1
    ids, t = s.trace_closest(np.array([[0, 0, -5, 0, 0, 1, 0.0]]), tmin=0.001, tmax=4.0)
    assert ids[0] == -1


def test_moving_sphere_center(tmp_path):  # sphere.rs:51-53
    s = scene_from(tmp_path, lambda g: g.bvh([g.moving_sphere([0, 0, 0], [2, 0, 0], 0.5, g.lambertian_color([1, 1, 1]))]))
    rays = np.array([[1.0, 0, -5, 0, 0, 1, 0.5], [1.0, 0, -5, 0, 0, 1, 0.0], [2.0, 0, -5, 0, 0, 1, 1.0]])
    ids, t = s.trace_closest(rays)
    assert list(ids) == [0, -1, 0]
    assert t[0] == 4.5 and t[2] == 4.5


# ---- quad.rs:25-38, 84-109; plane.rs:15-32 -----------------------------------------------------------
def test_quad_hit_known_answer(tmp_path):
    # a bare HittableList root: no AABB test in front of the quad (an axis-parallel ray lying exactly in a box
    # face is rejected by AABB::hit's NaN handling, AABB.rs:62-75, which would mask the quad's own edge rule)
    s = scene_from(tmp_path, lambda g: g.list([g.quad([0, 0, 0], [2, 0, 0], [0, 2, 0], g.lambertian_color([1, 1, 1]))]))
    rays = np.array([
        [0.5, 1.5, 5, 0, 0, -1, 0],    # interior: alpha .25, beta .75
        [0.0, 0.0, 5, 0, 0, -1, 0],    # corner q: alpha = beta = 0 -> inclusive [0,1] accepts
        [2.0, 2.0, 5, 0, 0, -1, 0],    # far corner: alpha = beta = 1 -> accepted
        [2.0000001, 1, 5, 0, 0, -1, 0],  # just outside
        [0.5, 0.5, 5, 1, 0, 0, 0],     # parallel: |denominator| < 1e-8
    ])
    ids, t, ex = s.trace_closest(rays, extra=True)
    assert list(ids) == [0, 0, 0, -1, -1]
    assert t[0] == 5.0
    np.testing.assert_allclose(ex[0, 6:8], [0.25, 0.75], atol=1e-15)
    np.testing.assert_allclose(ex[1, 6:8], [0, 0], atol=0)
    np.testing.assert_allclose(ex[2, 6:8], [1, 1], atol=0)
    # normal = unit(u x v) = +z; the ray travels -z so it is a front face and the normal stays +z
    np.testing.assert_allclose(ex[0, 3:6], [0, 0, 1])
    # inclusive interval (interval.rs:29-31): t == t_max is accepted for quads
    ids, t = s.trace_closest(rays[:1], tmin=0.001, tmax=5.0)
    assert ids[0] == 0


def test_cube_is_six_quads_in_reference_order(tmp_path):  # quad.rs:54-80
    s = scene_from(tmp_path, lambda g: g.bvh([g.cube([0, 0, 0], [1, 2, 3], g.lambertian_color([1, 1, 1]))]))
    c = np.array([0.5, 1.0, 1.5])
    dirs = {(0, 0, 1): 0, (1, 0, 0): 1, (0, 0, -1): 2, (-1, 0, 0): 3, (0, 1, 0): 4, (0, -1, 0): 5}  # face hit from outside
    for d, face in dirs.items():
        d = np.array(d, float)
        o = c + 10 * d
        ids, t = s.trace_closest(np.array([[*o, *(-d), 0.0]]))
        assert ids[0] == face


# ---- AABB.rs:24-48, 123-128; interval.rs:41-44 ---------------------------------------------------------
def test_flat_quad_is_still_hit_through_its_padded_box(tmp_path):
    # a quad in the plane z = 0 has a zero-thickness box; from_corners pads it to +-5e-5 so BVH traversal reaches it
    s = scene_from(tmp_path, lambda g: g.bvh([g.quad([0, 0, 0], [1, 0, 0], [0, 1, 0], g.lambertian_color([1, 1, 1])),
                                               g.sphere([10, 0, 0], 1, g.lambertian_color([1, 1, 1])),
                                               g.sphere([-10, 0, 0], 1, g.lambertian_color([1, 1, 1]))]))
    ids, t = s.trace_closest(np.array([[0.5, 0.5, 1, 0, 0, -1, 0.0]]))
    assert ids[0] == 0 and t[0] == 1.0


# ---- BVH.rs:18-90 + hittable.rs:70-91: the tree returns the brute-force closest hit ---------------------
def brute_force(prims, rays, tmin, tmax):
    """closest hit over (kind, params) with the reference's per-primitive rules, numpy f64"""
    n = rays.shape[0]
    best_t = np.full(n, tmax)
    best_id = np.full(n, -1)
    o, d, time = rays[:, 0:3], rays[:, 3:6], rays[:, 6]
    for pid, (kind, p) in enumerate(prims):
        if kind == "sphere":
            c1, c2, r = p
            center = c1 + time[:, None] * (c2 - c1)
            oc = center - o
            a = (d * d).sum(1); h = (d * oc).sum(1); c = (oc * oc).sum(1) - r * r
            disc = h * h - a * c
            ok = disc >= 0
            sq = np.sqrt(np.where(ok, disc, 0))
            t1, t2 = (h - sq) / a, (h + sq) / a
            use1 = ok & (tmin < t1) & (t1 < best_t)
            use2 = ok & ~use1 & (tmin < t2) & (t2 < best_t)
            t = np.where(use1, t1, t2)
            upd = use1 | use2
        else:
            q, u, v = p
            nrm = np.cross(u, v); normal = nrm / np.linalg.norm(nrm); w = nrm / nrm.dot(nrm)
            den = d @ normal
            ok = np.abs(den) >= 1e-8
            t = (normal.dot(q) - o @ normal) / np.where(ok, den, 1)
            ok &= (tmin <= t) & (t <= best_t)
            P = o + d * t[:, None] - q
            al = np.cross(P, v) @ w; be = np.cross(u, P) @ w
            upd = ok & (0 <= al) & (al <= 1) & (0 <= be) & (be <= 1)
        best_t = np.where(upd, t, best_t)
        best_id = np.where(upd, pid, best_id)
    return best_id, best_t


def test_bvh_matches_brute_force(tmp_path):
    rng = np.random.default_rng(3)
    prims, g = [], Graph()
    m = g.lambertian_color([1, 1, 1])
    items = []
    for i in range(60):
        c = (rng.random(3) - 0.5) * 20
        if i % 3 == 0:
            u, v = rng.normal(size=3) * 2, rng.normal(size=3) * 2
            prims.append(("quad", (c, u, v))); items.append(g.quad(c, u, v, m))
        elif i % 7 == 0:
            c2 = c + rng.normal(size=3); r = 0.3 + rng.random()
            prims.append(("sphere", (c, c2, r))); items.append(g.moving_sphere(c, c2, r, m))
        else:
            r = 0.3 + 1.5 * rng.random()
            prims.append(("sphere", (c, c, r))); items.append(g.sphere(c, r, m))
    p = g.write(str(tmp_path / "b.gsg"), g.bvh(items))
    s = orc.OracleScene.load(p)
    o = (rng.random((4000, 3)) - 0.5) * 40
    d = -o + rng.normal(size=(4000, 3)) * 6  # aimed roughly at the cloud
    rays = np.concatenate([o, d, rng.random((4000, 1))], 1)
    ids, t = s.trace_closest(rays)
    bid, bt = brute_force(prims, rays, 0.001, 1.7976931348623157e308)
    assert np.array_equal(ids, bid)
    hit = ids >= 0
    assert hit.sum() > 500
    np.testing.assert_allclose(t[hit], bt[hit], rtol=1e-12)


# ---- hittable.rs:93-216: instances agree with hand-transformed geometry ----------------------------------
def test_translate_rotate_known_answer(tmp_path):
    def build(g):
        m = g.lambertian_color([1, 1, 1])
        return g.bvh([g.translate(g.rotate_y(g.quad([-1, -1, 0], [2, 0, 0], [0, 2, 0], m), 90.0), [5, 0, 0])])
    s = scene_from(tmp_path, build)
    # RotateY(90): object +z maps to world +x, so the quad (normal +z at the origin) faces +x at x = 5
    ids, t, ex = s.trace_closest(np.array([[10, 0.25, 0.5, -1, 0, 0, 0.0]]), extra=True)
    assert ids[0] == 0
    assert abs(t[0] - 5.0) < 1e-12
    np.testing.assert_allclose(ex[0, 0:3], [5, 0.25, 0.5], atol=1e-12)
    np.testing.assert_allclose(ex[0, 3:6], [1, 0, 0], atol=1e-12)


# ---- material.rs -------------------------------------------------------------------------------------
def test_reflectance_schlick():  # material.rs:114-119
    L = orc.lib()
    assert abs(L.orc_reflectance(1.0, 1.5) - 0.04) < 1e-15
    assert abs(L.orc_reflectance(0.0, 1.5) - 1.0) < 1e-15
    assert abs(L.orc_reflectance(1.0, 1 / 1.5) - 0.04) < 1e-15


def test_fake_cosine_lobe_moments():  # util.rs:48-60 (quirk Q1): cos = sqrt(1-s) / sqrt(sqrt(s) + 1 - s), s ~ U
    d = orc.sample_dirs(0, seed=11, n=200000)
    cos = d[:, 2] / np.linalg.norm(d, axis=1)
    s = (np.arange(2_000_000) + 0.5) / 2_000_000
    expect = np.mean(np.sqrt(1 - s) / np.sqrt(np.sqrt(s) + 1 - s))
    assert abs(expect - 0.612) < 2e-3           # 0.612, not the 2/3 of a true cosine lobe
    assert abs(cos.mean() - expect) < 3e-3
    phi = np.arctan2(d[:, 1], d[:, 0])
    assert abs(np.mean(np.cos(phi))) < 5e-3 and abs(np.mean(np.sin(phi))) < 5e-3


def test_random_unit_vector_is_uniform_on_the_sphere():  # util.rs:18-29
    d = orc.sample_dirs(1, seed=5, n=100000)
    np.testing.assert_allclose(np.linalg.norm(d, axis=1), 1.0, atol=1e-12)
    assert np.all(np.abs(d.mean(0)) < 0.01)
    np.testing.assert_allclose((d * d).mean(0), [1 / 3] * 3, atol=0.01)


def _one_material_scene(tmp_path, mat_builder):
    holder = {}

    def build(g):
        holder["mat"] = mat_builder(g)
        return g.bvh([g.sphere([0, 0, 0], 1, holder["mat"])])
    s = scene_from(tmp_path, build)
    return s, holder["mat"]


def test_metal_mirror_and_absorption(tmp_path):  # material.rs:87-102
    s, m = _one_material_scene(tmp_path, lambda g: g.metal([0.8, 0.7, 0.6], 0.0))
    ray = [0, 0, 0, 1, -1, 0, 0.0]
    ok, d, att, em = s.scatter(m, 1, 0, 0, ray, p=[0, 0, 0], n=[0, 1, 0])
    assert ok
    np.testing.assert_allclose(d, np.array([1, 1, 0]) / math.sqrt(2), atol=1e-15)  # unit(reflect), fuzz 0
    np.testing.assert_allclose(att, [0.8, 0.7, 0.6])
    assert np.all(em == 0)
    s2, m2 = _one_material_scene(tmp_path, lambda g: g.metal([1, 1, 1], 5.0))
    absorbed = sum(not s2.scatter(m2, 1, 0, k, ray, p=[0, 0, 0], n=[0, 1, 0])[0] for k in range(200))
    assert 20 < absorbed < 180  # large fuzz pushes many directions below the surface -> None


def test_dielectric_total_internal_reflection_and_refraction(tmp_path):  # material.rs:122-148, vec3.rs:57-62
    s, m = _one_material_scene(tmp_path, lambda g: g.dielectric(1.5))
    # back face (inside glass), grazing: ri = 1.5, sin > 1/1.5 -> must reflect whatever the random draw is
    d_in = np.array([math.sin(1.2), -math.cos(1.2), 0.0])
    for k in range(20):
        ok, d, att, _ = s.scatter(m, 1, 0, k, [0, 0, 0, *d_in, 0.0], p=[0, 0, 0], n=[0, -1, 0])  # outward normal -y => ray is leaving: back face
        assert ok and np.all(att == 1)
        np.testing.assert_allclose(d, [d_in[0], -d_in[1], 0], atol=1e-15)
    # front face at normal incidence: refracted ray continues straight unless Schlick (4 %) reflects
    outs = [s.scatter(m, 1, 0, k, [0, 2, 0, 0, -1, 0, 0.0], p=[0, 1, 0], n=[0, 1, 0])[1] for k in range(400)]
    straight = sum(np.allclose(d, [0, -1, 0]) for d in outs)
    mirrored = sum(np.allclose(d, [0, 1, 0]) for d in outs)
    assert straight + mirrored == 400 and 3 <= mirrored <= 40


def test_diffuse_light_is_two_sided_and_ends_the_path(tmp_path):  # material.rs:165-169 (Q8)
    s, m = _one_material_scene(tmp_path, lambda g: g.light_color([4, 5, 6]))
    for n in ([0, 1, 0], [0, -1, 0]):
        ok, d, att, em = s.scatter(m, 1, 0, 0, [0, 2, 0, 0, -1, 0, 0.0], p=[0, 1, 0], n=n)
        assert not ok
        np.testing.assert_allclose(em, [4, 5, 6])


# ---- texture.rs -------------------------------------------------------------------------------------
def test_checker_parity_including_negative_cells(tmp_path):  # texture.rs:57-70: Rust % keeps the sign, -1 is odd
    def build(g):
        t = g.checker(1.0, g.solid([1, 0, 0]), g.solid([0, 0, 1]))
        return g.bvh([g.sphere([0, 0, 0], 1, g.lambertian(t))])
    s = scene_from(tmp_path, build)
    tex = 2  # solid, solid, checker in creation order
    even, odd = [1, 0, 0], [0, 0, 1]
    np.testing.assert_allclose(s.texture_value(tex, 0, 0, [0.5, 0.5, 0.5]), even)    # 0+0+0
    np.testing.assert_allclose(s.texture_value(tex, 0, 0, [1.5, 0.5, 0.5]), odd)     # 1
    np.testing.assert_allclose(s.texture_value(tex, 0, 0, [-0.5, 0.5, 0.5]), odd)    # -1 % 2 == -1
    np.testing.assert_allclose(s.texture_value(tex, 0, 0, [-0.5, -0.5, 0.5]), even)  # -2
    np.testing.assert_allclose(s.texture_value(tex, 0, 0, [-0.5, -0.5, -0.5]), odd)  # -3


def test_image_texture_nearest_texel_v_flip(tmp_path):  # texture.rs:84-94
    from PIL import Image

    img = np.asarray(Image.open(os.path.join(ASSETS, "earthmap.jpg")).convert("RGB"))
    s = orc.OracleScene.build("scene2")
    H, W = img.shape[:2]
    rng = np.random.default_rng(0)
    for _ in range(200):
        u, v = rng.random(), rng.random()
        i, j = int(u * W), int((1 - v) * H)
        np.testing.assert_allclose(s.texture_value(0, u, v, [0, 0, 0]), img[min(j, H - 1), min(i, W - 1)] / 255.0, atol=1e-15)
    # Q13: u == 1 / v == 0 index one past the edge in the reference (panic); the oracle clamps
    np.testing.assert_allclose(s.texture_value(0, 1.0, 0.0, [0, 0, 0]), img[H - 1, W - 1] / 255.0)


def test_perlin_table_and_noise_properties():
    tab = orc.perlin_table(0)
    assert sorted(tab.tolist()) == list(range(256))          # a permutation
    assert not np.array_equal(tab, np.arange(256))
    L = orc.lib()
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(4000, 3)) * 20
    vals = np.array([L.orc_perlin_get(*p) for p in pts])
    assert np.all(np.abs(vals) <= 1.0) and vals.std() > 0.1 and abs(vals.mean()) < 0.05
    # gradient noise vanishes on the integer lattice, and is continuous
    for p in rng.integers(-50, 50, size=(50, 3)):
        assert L.orc_perlin_get(float(p[0]), float(p[1]), float(p[2])) == 0.0
    p = np.array([3.3, -7.1, 12.9])
    assert abs(L.orc_perlin_get(*p) - L.orc_perlin_get(*(p + 1e-9))) < 1e-7


def test_xorshift_matches_the_rand_xorshift_crate_vectors():
    """rand_xorshift's own `test_xorshift_true_values`: seed bytes 16..1, nine next_u32 and then nine next_u64
    (low word first).  This pins the generator behind noise 0.9's PermutationTable::new (texture.rs:104)."""
    seed = np.arange(16, 0, -1, dtype=np.uint8)
    out = np.empty(27, dtype=np.uint32)
    orc.lib().orc_xorshift_words(seed.ctypes.data, 27, out.ctypes.data)
    assert out[:9].tolist() == [2081028795, 620940381, 269070770, 16943764, 854422573, 29242889, 1550291885, 1227154591, 271695242]
    u64 = [(int(out[9 + 2 * k + 1]) << 32) | int(out[9 + 2 * k]) for k in range(9)]
    assert u64 == [9247529084182843387, 8321512596129439293, 14104136531997710878, 6848554330849612046, 343577296533772213,
                   17828467390962600268, 9847333257685787782, 7717352744383350108, 1133407547287910111]


def _pcg32_words(state, n, stream=11634580027462260723):
    """rand_pcg::Pcg32::new(state, stream) -- the generator behind rand's `crate::test::rng(seed)`."""
    M, mult = (1 << 64) - 1, 6364136223846793005
    inc = ((stream << 1) | 1) & M
    st = (state + inc) & M
    st = (st * mult + inc) & M
    out = []
    for _ in range(n):
        old, st = st, (st * mult + inc) & M
        rot, xsh = old >> 59, (((old >> 18) ^ old) >> 27) & 0xffffffff
        out.append(((xsh >> rot) | (xsh << ((32 - rot) & 31))) & 0xffffffff)
    return np.array(out, dtype=np.uint32)


def test_gen_index_arithmetic():
    """rand 0.8 `gen_index` = UniformInt<u32>::sample_single(0, ubound): widening multiply, accept when the low word is
    <= (range << lz) - 1.  Checked (i) against the definition in exact integer arithmetic on random words, (ii) for
    unbiasedness of the accepted values, (iii) against the one value of rand's `value_stability_slice` that could be
    restated with confidence: `chars.choose(&mut rng(413))` over 14 letters gives 'l' (index 11)."""
    L = orc.lib()
    rng = np.random.default_rng(5)
    used = C.c_size_t()
    for ub in (2, 3, 14, 100, 255, 256, 1000):
        words = rng.integers(0, 1 << 32, 64, dtype=np.uint64).astype(np.uint32)
        got = L.orc_gen_index(words.ctypes.data, words.size, ub, C.byref(used))
        zone = ((ub << (32 - ub.bit_length())) & 0xffffffff) - 1
        k = next(i for i, w in enumerate(words.tolist()) if ((w * ub) & 0xffffffff) <= zone)
        assert used.value == k + 1 and got == (int(words[k]) * ub) >> 32
    words = _pcg32_words(413, 8)
    assert "abcdefghijklmn"[L.orc_gen_index(words.ctypes.data, words.size, 14, C.byref(used))] == "l"


def test_perlin_gradient_cases_follow_noise_0_9():
    """gradient_dot_v of noise 0.9's core::perlin::perlin_3d: hash & 15 -> 12 edge vectors, 12..15 repeating
    (1,1,0), (-1,1,0), (0,-1,1), (0,-1,-1).  Probed through Perlin::get: inside one lattice cell, close to a corner,
    the value is dominated by that corner's gradient dotted with the offset."""
    tab = orc.perlin_table(0).astype(np.int64)
    L = orc.lib()
    grads = [(1, 1, 0), (-1, 1, 0), (1, -1, 0), (-1, -1, 0), (1, 0, 1), (-1, 0, 1), (1, 0, -1), (-1, 0, -1), (0, 1, 1), (0, -1, 1), (0, 1, -1),
             (0, -1, -1), (1, 1, 0), (-1, 1, 0), (0, -1, 1), (0, -1, -1)]
    seen = set()
    eps = 1e-4
    for x in range(-8, 8):
        for y in range(-8, 8):
            for z in range(-8, 8):
                h = int(tab[tab[tab[x & 255] ^ (y & 255)] ^ (z & 255)]) & 15
                seen.add(h)
                g = np.array(grads[h], dtype=np.float64)
                for d in ((eps, 0, 0), (0, eps, 0), (0, 0, eps)):
                    v = L.orc_perlin_get(x + d[0], y + d[1], z + d[2])
                    expect = 1.1547005383792515 * float(g @ np.array(d))
                    assert abs(v - expect) < 5e-3 * eps + 1e-12, (x, y, z, h, d, v, expect)
    assert seen == set(range(16))


def test_noise_texture_formula(tmp_path):  # texture.rs:107-130
    def build(g):
        return g.bvh([g.sphere([0, 0, 0], 1, g.lambertian(g.noise(4.0)))])
    s = scene_from(tmp_path, build)
    L = orc.lib()
    p = np.array([0.37, 1.91, -2.45])
    turb = abs(sum(L.orc_perlin_get(*(p * 2**i)) / 2**i for i in range(7)))
    expect = 0.5 * (1 + math.sin(4.0 * p[2] + 10 * turb))
    np.testing.assert_allclose(s.texture_value(0, 0, 0, p), [expect] * 3, rtol=1e-12)


# ---- camera.rs ---------------------------------------------------------------------------------------
def test_hdr_decoder_matches_opencv():
    cv2 = pytest.importorskip("cv2")
    ref = cv2.imread(os.path.join(ASSETS, "airport.hdr"), cv2.IMREAD_UNCHANGED)[:, :, ::-1]  # BGR -> RGB
    mine = orc.load_hdr(os.path.join(ASSETS, "airport.hdr"))
    assert mine.shape == (512, 1024, 3) == ref.shape
    np.testing.assert_array_equal(mine, ref.astype(np.float32))
    assert abs(float(mine.max()) - 3808.0) < 1.0  # max radiance 3808


def test_hdri_sample_angles():  # camera.rs:257-270, util.rs:67-86
    s = orc.OracleScene.build("scene11", width=64)   # rotation (pi/2, pi, 0)
    hdr = orc.load_hdr(os.path.join(ASSETS, "airport.hdr"))
    H, W = hdr.shape[:2]
    rx, ry, rz = math.pi / 2, math.pi, 0.0
    sx, cx, sy, cy, sz, cz = math.sin(rx), math.cos(rx), math.sin(ry), math.cos(ry), math.sin(rz), math.cos(rz)
    M = np.array([[cy * cz, cx * sz + sx * sy * cz, sx * sz - cx * sy * cz],
                  [-cy * sz, cx * cz - sx * sy * sz, sx * cz + cx * sy * sz],
                  [sy, -sx * cy, cx * cy]])
    rng = np.random.default_rng(4)
    for d in rng.normal(size=(300, 3)):
        r = M @ d
        r /= np.linalg.norm(r)
        u = 0.5 + math.atan2(r[1], r[0]) / (2 * math.pi)
        v = 0.5 - math.asin(r[2]) / math.pi
        x, y = int(u * W) % W, int(v * H) % H
        np.testing.assert_array_equal(s.background(d), hdr[y, x].astype(np.float64))


def test_camera_geometry_cornell():  # camera.rs:39-98
    s = orc.OracleScene.build("c3", width=600)
    assert (s.width, s.height) == (600, 600)
    rays = s.camera_rays(seed=1, sample=0)
    o, d = rays[:, :3], rays[:, 3:6]
    np.testing.assert_array_equal(o, np.tile([278.0, 278.0, -800.0], (rays.shape[0], 1)))  # no defocus: origin is the centre
    # vfov 40, focus 10: the viewport is 2*10*tan(20deg) tall at distance 10; pixels jitter within +-0.5
    half = 10 * math.tan(math.radians(20))
    assert np.all(np.abs(d[:, 2] - 10.0) < 1e-9)
    assert d[:, 0].min() >= -half - 1e-9 and d[:, 0].max() <= half + 1e-9
    assert d[:, 1].min() >= -half - 1e-9 and d[:, 1].max() <= half + 1e-9
    # +x points left in image space here (camera looks down +z with vup +y): pixel 0 is top-left => largest y, x = +half...
    px0 = d[0]
    assert px0[1] > half * 0.99 and abs(abs(px0[0]) - half) < half * 0.01
    assert np.all(rays[:, 6] >= 0) and np.all(rays[:, 6] < 1)


def test_image_height_truncates():  # camera.rs:52
    assert orc.OracleScene.build("c1").height == 225            # 400 / (16/9) = 225.0
    assert orc.OracleScene.build("scene0").height == 337        # 600 / (16/9) = 337.5 -> 337
    assert orc.OracleScene.build("c2", width=1280).height == 720


def test_defocus_disk_origins():  # camera.rs:223-226, util.rs:36-46
    s = orc.OracleScene.build("c1")
    rays = s.camera_rays(seed=2, sample=3)
    off = rays[:, :3] - np.array([13.0, 2.0, 3.0])
    radius = 10 * math.tan(math.radians(0.3))
    r = np.linalg.norm(off, axis=1)
    assert r.max() < radius and r.max() > 0.9 * radius
    w = np.array([13.0, 2.0, 3.0]); w /= np.linalg.norm(w)
    assert np.all(np.abs(off @ w) < 1e-12)  # the disk is perpendicular to the view direction
    # every ray still passes through its pixel sample on the focus plane: origin + direction
    target = rays[:, :3] + rays[:, 3:6]
    assert np.all(np.abs((target - np.array([13.0, 2.0, 3.0])) @ w + 10.0) < 1e-9)


def test_luminance_and_write_color():  # color.rs:8-33
    L = orc.lib()
    assert L.orc_luminance(1, 1, 1) == 0.299 + 0.587 + 0.144      # 0.144 (sic), Q5
    assert L.orc_color_byte(0.0) == 0 and L.orc_color_byte(-1.0) == 0
    assert L.orc_color_byte(1.0) == 255 and L.orc_color_byte(100.0) == 255   # clamp 0.999 * 256 = 255.7
    assert L.orc_color_byte(0.25) == 128                                    # sqrt gamma


# ---- Camera::sample, camera.rs:125-171 ------------------------------------------------------------------
def test_fixed_spp_mapping_and_adaptive_rules(tmp_path):
    def build(g):
        g.camera(width=8, batch=5, max_samples=0, tolerance=0.0, depth=4)
        g.background_solid([0.5, 0.5, 0.5])
        return g.bvh([g.sphere([0, 0, 0], 1.0, g.lambertian_color([0.5, 0.5, 0.5]))])
    s = scene_from(tmp_path, build)
    _, counts, ctr, _ = s.render()
    assert np.all(counts == 5) and ctr["paths"] == 5 * 64           # tol 0, max 0: exactly one batch
    # With max_samples >= batch_size the loop runs until n > max (Q4) wherever the luminance varies.  Where it
    # is constant (pure background) `sq_sum - sum*sum/n` rounds to a tiny negative number now and then, and
    # `negative < 0` stops the pixel early even at tolerance 0 -- so only max_samples < batch_size is a true
    # fixed count, which is what the configs use.
    s.set_sampling(tolerance=0.0, batch_size=5, max_samples=12)
    c = s.render()[1]
    assert c[4, 4] == 15 and set(np.unique(c)) <= {5, 10, 15}         # n > max: 5, 10, 15
    s.set_sampling(tolerance=0.0, batch_size=32, max_samples=1000)
    c = s.render()[1]
    assert c[4, 4] == 1024 and c.max() == 1024                        # 32 * 32 = 1024
    # constant-radiance pixels (pure background) have zero variance: 0 < mean^2 tol^2 stops after one batch;
    # black pixels (mean 0) never satisfy 0 < 0 and run to the cap
    s.set_sampling(tolerance=0.25, batch_size=4, max_samples=20)
    rgb, counts, _, _ = s.render()
    corner = counts[0, 0]
    assert corner == 4 and np.allclose(rgb[0, 0], 0.5)
    assert counts.max() <= 24 and counts[4, 4] >= 4


def test_black_pixels_run_to_the_cap(tmp_path):
    def build(g):
        g.camera(width=4, batch=4, max_samples=10, tolerance=0.5, depth=3)
        g.background_solid([0, 0, 0])
        return g.bvh([g.sphere([0, 0, 0], 1.0, g.lambertian_color([0.5, 0.5, 0.5]))])
    s = scene_from(tmp_path, build)
    rgb, counts, _, _ = s.render()
    assert np.all(rgb == 0) and np.all(counts == 12)


def test_ray_color_closed_form_furnace(tmp_path):
    # a Lambertian sphere of albedo a inside a uniform white environment: every path returns a^k for k bounces
    # before escaping; from a convex sphere every scattered ray escapes, so hits return exactly a and misses 1
    def build(g):
        g.camera(width=16, batch=16, depth=10, look_from=(0, 0, 5), vfov=30.0)
        g.background_solid([1, 1, 1])
        return g.bvh([g.sphere([0, 0, 0], 1.0, g.lambertian_color([0.3, 0.6, 0.9]))])
    s = scene_from(tmp_path, build)
    rgb, _, _, _ = s.render()
    np.testing.assert_allclose(rgb[8, 8], [0.3, 0.6, 0.9], rtol=1e-12)
    np.testing.assert_allclose(rgb[0, 0], [1, 1, 1], rtol=1e-12)


def test_constant_medium_transmittance(tmp_path):  # volume.rs:32-63
    # absorbing (black isotropic) slab-like sphere of radius R, density rho, seen through its centre against a
    # white background: transmitted fraction = exp(-rho * 2R)
    rho, R = 0.35, 1.5

    def build(g):
        g.camera(width=3, batch=40000, depth=5, look_from=(0, 0, 50), vfov=0.5)
        g.background_solid([1, 1, 1])
        boundary = g.sphere([0, 0, 0], R, g.dielectric(1.5))
        return g.bvh([g.medium(boundary, rho, g.isotropic(g.solid([0, 0, 0])))])
    s = scene_from(tmp_path, build)
    rgb, _, _, _ = s.render(threads=8)
    expect = math.exp(-rho * 2 * R)
    assert abs(rgb[1, 1, 0] - expect) < 4 * math.sqrt(expect * (1 - expect) / 40000) + 2e-3


def test_counters_build_counts_reference_work():
    s = orc.OracleScene.build("c3", width=32, spp=4, counters=True)
    _, _, ctr, _ = s.render()
    assert ctr["paths"] == 32 * 32 * 4 and ctr["rays"] >= ctr["paths"]
    assert ctr["box_tests"] > ctr["rays"] and ctr["prim_tests"] > ctr["rays"]


def test_random_scene_graph_roundtrip(tmp_path):
    p = str(tmp_path / "r.gsg")
    random_scene(p, seed=9)
    s = orc.OracleScene.load(p)
    s.dump(str(tmp_path / "r2.gsg"))
    s2 = orc.OracleScene.load(str(tmp_path / "r2.gsg"))
    rays = s.camera_rays(1, 0)
    a, b = s.trace_closest(rays), s2.trace_closest(rays)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


# ---- triangle.rs:10-73 (quirk Q14) ------------------------------------------------------------------------
def test_triangle_known_answers_and_the_ignored_interval(tmp_path):
    def build(g):
        m = g.lambertian_color([1, 1, 1])
        # counter-clockwise seen from +z ... with the reference's edge naming (edge_1 = c - a, edge_2 = b - a) the
        # FRONT face (det > 0) is the one whose normal (b - a) x (c - a) points AWAY from the viewer
        near = g.triangle([0, 0, 0], [0, 2, 0], [2, 0, 0], m)        # prim 0, z = 0, normal (0,0,-4)
        far = g.triangle([0, 0, -3], [0, 2, -3], [2, 0, -3], m)      # prim 1, z = -3
        return g.list([near, far])
    s = scene_from(tmp_path, build)
    ray = np.array([[0.5, 0.5, 5, 0, 0, -1, 0.0]])
    try:
        # the reference: the later triangle replaces the closer hit because Triangle::hit never looks at ray_t
        ids, t, ex = s.trace_closest(ray, extra=True)
        assert ids[0] == 1 and t[0] == 8.0
        np.testing.assert_allclose(ex[0, 6:8], [0.25, 0.25])               # u along c - a, v along b - a
        np.testing.assert_allclose(ex[0, 3:6], [0, 0, 4])                  # unnormalised (b-a)x(c-a) = (0,0,-4), flipped against the ray
        # ... and a triangle BEHIND the origin is hit too
        ids, t = s.trace_closest(np.array([[0.5, 0.5, -1, 0, 0, -1, 0.0]]))
        assert ids[0] == 1 and t[0] == 2.0
        ids, t = s.trace_closest(np.array([[0.5, 0.5, -5, 0, 0, -1, 0.0]]))
        assert ids[0] == 1 and t[0] == -2.0                                # negative t accepted (Q14)
        # back faces are culled: the same triangles seen from below
        ids, t = s.trace_closest(np.array([[0.5, 0.5, -9, 0, 0, 1, 0.0]]))
        assert ids[0] == -1
        # with the interval applied (the device's rule) the nearest one wins and hits behind the origin vanish
        orc.set_triangle_interval(True)
        ids, t = s.trace_closest(ray)
        assert ids[0] == 0 and t[0] == 5.0
        ids, t = s.trace_closest(np.array([[0.5, 0.5, -5, 0, 0, -1, 0.0]]))
        assert ids[0] == -1
    finally:
        orc.set_triangle_interval(False)


def test_the_reference_triangle_scene_is_insensitive_to_the_interval_bug():
    """main.rs:836-888 (`triangles`): three triangles forming an open corner.  No camera ray and no scattered ray
    meets a triangle outside its interval, so the reference's Triangle::hit and the interval-respecting variant the
    device implements return the same hit for every ray."""
    s = orc.OracleScene.build("scene10", width=200)
    rng = np.random.default_rng(2)
    cam = s.camera_rays(seed=1, sample=0)
    ids, t, ex = s.trace_closest(cam, extra=True)
    assert (ids >= 0).mean() > 0.05   # the three triangles cover ~6 % of the view
    p = np.repeat(ex[ids >= 0, 0:3], 4, axis=0)
    n = np.repeat(ex[ids >= 0, 3:6], 4, axis=0)
    d = rng.normal(size=p.shape)
    # scattered rays as the scene's Lambertian materials produce them: in the hemisphere of the normal that faces the
    # incoming ray (a ray sent THROUGH its own triangle would re-hit it at t = 0 in the reference: t_min is ignored)
    d = np.where(((d * n).sum(1) < 0)[:, None], -d, d)
    rays = np.concatenate([cam, np.concatenate([p, d, np.zeros((p.shape[0], 1))], 1)])
    a = s.trace_closest(rays)
    try:
        orc.set_triangle_interval(True)
        b = s.trace_closest(rays)
    finally:
        orc.set_triangle_interval(False)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
