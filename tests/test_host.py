"""Host-layer logic on the CPU: scene builders agree with the oracle's, flatten() bakes instances and
ranks primitives correctly (checked by a numpy brute force over the flattened arrays against the
oracle's tree walk), and the SAH BVH is a valid tree.  No GPU needed."""
import ctypes as C
import os
import sys
import filecmp

import numpy as np
import pytest

import grayshift_b200 as gs
import oracle_ffi as orc
from grayshift_b200 import ffi
from graphs import random_scene

SCENES = ["c1", "c2", "c3", "c4", "scene0", "scene1", "scene2", "scene3", "scene4", "scene5", "scene6", "scene7", "scene8", "scene9", "scene10", "scene11"]


@pytest.mark.parametrize("name", SCENES)
def test_host_and_oracle_build_identical_scene_graphs(tmp_path, name):
    """Two independently written builders (grayshift_b200/host/scenes.cpp, oracle/scenes.cpp) of main.rs's
    scenes, same seeded construction RNG: the canonical graph dumps must be byte-identical."""
    a, b = str(tmp_path / "host.gsg"), str(tmp_path / "oracle.gsg")
    gs.HostScene.build(name).dump(a)
    orc.OracleScene.build(name).dump(b)
    assert filecmp.cmp(a, b, shallow=False)


def test_scene_seed_changes_the_scene(tmp_path):
    a, b = str(tmp_path / "a.gsg"), str(tmp_path / "b.gsg")
    gs.HostScene.build("c1", scene_seed=1).dump(a)
    gs.HostScene.build("c1", scene_seed=2).dump(b)
    assert not filecmp.cmp(a, b, shallow=False)
    assert ffi.host_lib().gsh_scene_rng(7, 3) == orc.lib().orc_scene_rng(7, 3)


def test_graph_roundtrip(tmp_path):
    a, b = str(tmp_path / "a.gsg"), str(tmp_path / "b.gsg")
    gs.HostScene.build("c4").dump(a)
    gs.HostScene.load(a).dump(b)
    assert filecmp.cmp(a, b, shallow=False)


def arrays(desc):
    d = desc.contents
    sph = np.ctypeslib.as_array(C.cast(d.spheres, C.POINTER(C.c_double)), shape=(d.n_spheres, 11)).copy() if d.n_spheres else np.zeros((0, 11))
    sph_i = np.ctypeslib.as_array(C.cast(d.spheres, C.POINTER(C.c_int32)), shape=(d.n_spheres, 22))[:, 14:18].copy() if d.n_spheres else np.zeros((0, 4), int)
    sph = sph[:, :7]
    quad = np.ctypeslib.as_array(C.cast(d.quads, C.POINTER(C.c_double)), shape=(d.n_quads, 11))[:, :9].copy() if d.n_quads else np.zeros((0, 9))
    quad_i = np.ctypeslib.as_array(C.cast(d.quads, C.POINTER(C.c_int32)), shape=(d.n_quads, 22))[:, 18:22].copy() if d.n_quads else np.zeros((0, 4), int)
    nodes = [(list(d.nodes[i].bounds_min), list(d.nodes[i].bounds_max), d.nodes[i].left, d.nodes[i].right, d.nodes[i].first_ref, d.nodes[i].ref_count)
             for i in range(d.n_nodes)]
    refs = np.ctypeslib.as_array(d.prim_refs, shape=(d.n_prim_refs,)).copy()
    return sph, sph_i, quad, quad_i, nodes, refs


def brute_force_flat(sph, sph_i, quad, quad_i, rays, tmin=0.001, tmax=1.7976931348623157e308):
    """Closest hit over the flattened world-space primitives, visiting them in `order` rank with the
    reference's acceptance rules (spheres t < closest, quads t <= closest): HittableList::hit semantics."""
    n = rays.shape[0]
    best_t = np.full(n, tmax)
    best_id = np.full(n, -1)
    o, d, time = rays[:, 0:3], rays[:, 3:6], rays[:, 6]
    items = [(int(sph_i[k, 3]), 0, k) for k in range(sph.shape[0])] + [(int(quad_i[k, 2]), 1, k) for k in range(quad.shape[0])]
    a = (d * d).sum(1)
    with np.errstate(all="ignore"):
        for _, kind, k in sorted(items):
            if kind == 0:
                c, r, path, (mat, pid, moving, _) = sph[k, 0:3], sph[k, 3], sph[k, 4:7], sph_i[k]
                center = c + (time[:, None] * path if moving else 0.0)
                oc = center - o
                h = (d * oc).sum(1); cc = (oc * oc).sum(1) - r * r
                disc = h * h - a * cc
                ok = disc >= 0
                sq = np.sqrt(np.where(ok, disc, 0))
                t1, t2 = (h - sq) / a, (h + sq) / a
                use1 = ok & (tmin < t1) & (t1 < best_t)
                use2 = ok & ~use1 & (tmin < t2) & (t2 < best_t)
                t = np.where(use1, t1, t2); upd = use1 | use2
            else:
                q, u, v, (mat, pid, _, _) = quad[k, 0:3], quad[k, 3:6], quad[k, 6:9], quad_i[k]
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


def mixed_rays(ora, n_cam=3000, seed=1):
    rng = np.random.default_rng(seed)
    cam = ora.camera_rays(seed=5, sample=0)
    cam = cam[rng.choice(cam.shape[0], size=min(n_cam, cam.shape[0]), replace=False)]
    ids, t, ex = ora.trace_closest(cam, extra=True)
    p = ex[ids >= 0, 0:3]
    dd = rng.normal(size=p.shape)
    sec = np.concatenate([p, dd / np.linalg.norm(dd, axis=1, keepdims=True), rng.random((p.shape[0], 1))], 1)
    return np.concatenate([cam, sec])


@pytest.mark.parametrize("name,kw", [("c3", {}), ("c4", {"width": 200}), ("c1", {})])
def test_flattened_primitives_reproduce_the_oracle_hits(name, kw):
    """Instance baking (Translate/RotateY -> world space), primitive ids and the tie-break ranks, with no GPU
    in the loop: brute force over gs_scene_desc == the oracle's BVH/list/transform walk, ids exact."""
    host = gs.HostScene.build(name, include_media=False, **kw)
    ora = orc.OracleScene.build(name, include_media=False, **kw)
    sph, sph_i, quad, quad_i, _, _ = arrays(host.desc)
    rays = mixed_rays(ora, n_cam=1500 if name != "c3" else 3000)
    oid, ot = ora.trace_closest(rays)
    bid, bt = brute_force_flat(sph, sph_i, quad, quad_i, rays)
    assert np.array_equal(bid, oid), f"{int((bid != oid).sum())} id mismatches"
    hit = oid >= 0
    assert hit.mean() > 0.3
    np.testing.assert_allclose(bt[hit], ot[hit], rtol=1e-9)


def test_exact_ties_follow_the_reference_order(tmp_path):
    """Coplanar overlapping quads tie exactly in t.  The reference keeps the one it visits LAST (Plane::hit
    accepts t <= closest, plane.rs:27), in HittableList order (hittable.rs:75) or BVH left-then-right order
    (BVH.rs:71-79, children sorted by box minimum): gs_quad.order must reproduce that in both containers."""
    from graphs import Graph

    def scene(container, flip):
        g = Graph()
        m = g.lambertian_color([1, 1, 1])
        a = g.quad([0, 0, 0], [4, 0, 0], [0, 4, 0], m)          # prim 0
        b = g.quad([1, 1, 0], [4, 0, 0], [0, 4, 0], m)          # prim 1, overlaps a on [1,4]^2
        c = g.quad([-9, 0, 3], [1, 0, 0], [0, 1, 0], m)         # prim 2, elsewhere
        kids = [b, a, c] if flip else [a, b, c]
        root = g.list(kids) if container == "list" else g.bvh(kids)
        g.camera(width=8, look_from=(2, 2, 20))
        return g.write(str(tmp_path / f"{container}{flip}.gsg"), root)

    rng = np.random.default_rng(0)
    n = 500
    o = np.stack([rng.uniform(1.2, 3.8, n), rng.uniform(1.2, 3.8, n), np.full(n, 5.0)], 1)   # straight down onto the overlap
    rays = np.concatenate([o, np.tile([0.0, 0.0, -1.0], (n, 1)), np.zeros((n, 1))], 1)
    winners = set()
    for container in ("list", "bvh"):
        for flip in (False, True):
            p = scene(container, flip)
            host, ora = gs.HostScene.load(p), orc.OracleScene.load(p)
            sph, sph_i, quad, quad_i, _, _ = arrays(host.desc)
            oid, ot = ora.trace_closest(rays)
            bid, bt = brute_force_flat(sph, sph_i, quad, quad_i, rays)
            assert np.array_equal(bid, oid) and np.all(ot == 5.0)
            assert len(set(oid.tolist())) == 1
            winners.add((container, flip, int(oid[0])))
    # list order decides in a list; in a BVH the stable sort by box minimum does, whatever the insertion order
    assert ("list", False, 1) in winners and ("list", True, 0) in winners
    assert {w[2] for w in winners if w[0] == "bvh"} == {1}


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scene_graphs_flatten_correctly(tmp_path, seed):
    p = str(tmp_path / "r.gsg")
    random_scene(p, seed=seed)
    host, ora = gs.HostScene.load(p), orc.OracleScene.load(p)
    sph, sph_i, quad, quad_i, _, _ = arrays(host.desc)
    rays = mixed_rays(ora, n_cam=2500, seed=seed)
    oid, ot = ora.trace_closest(rays)
    bid, bt = brute_force_flat(sph, sph_i, quad, quad_i, rays)
    assert np.array_equal(bid, oid)
    hit = oid >= 0
    np.testing.assert_allclose(bt[hit], ot[hit], rtol=1e-9)


@pytest.mark.parametrize("name,kw,ref_tree", [("c4", {}, False), ("c1", {}, False), ("c3", {}, True), ("c5", {"n_spheres": 20000}, False)])
def test_bvh_is_a_valid_tree_over_all_primitives(name, kw, ref_tree):
    host = gs.HostScene.build(name, **kw).flatten(reference_tree=ref_tree)
    sph, sph_i, quad, quad_i, nodes, refs = arrays(host.desc)
    seen = np.zeros(len(refs), bool)
    prim_seen = set()

    def prim_box(ref):
        kind, idx = ref >> 30, ref & ((1 << 30) - 1)
        if kind == 0:
            c0 = sph[idx, 0:3]; c1 = c0 + (sph[idx, 4:7] if sph_i[idx, 2] else 0)
            r = abs(sph[idx, 3])
            return np.minimum(c0, c1) - r, np.maximum(c0, c1) + r
        q, u, v = quad[idx, 0:3], quad[idx, 3:6], quad[idx, 6:9]
        pts = np.stack([q, q + u, q + v, q + u + v])
        return pts.min(0), pts.max(0)

    def walk(i, depth):
        mn, mx, left, right, first, count = nodes[i]
        mn, mx = np.array(mn), np.array(mx)
        if count > 0:
            assert count <= 15
            for k in range(first, first + count):
                assert not seen[k]
                seen[k] = True
                prim_seen.add(int(refs[k]))
                pmn, pmx = prim_box(int(refs[k]))
                assert np.all(pmn >= mn - 1e-9) and np.all(pmx <= mx + 1e-9)
            return depth
        assert left > i and right > i            # children after parents (what gs_scene_create requires)
        for c in (left, right):
            cmn, cmx = np.array(nodes[c][0]), np.array(nodes[c][1])
            assert np.all(cmn >= mn - 1e-9) and np.all(cmx <= mx + 1e-9)
        return max(walk(left, depth + 1), walk(right, depth + 1))

    import sys
    sys.setrecursionlimit(10000)
    depth = walk(0, 0)
    assert seen.all() and len(prim_seen) == sph.shape[0] + quad.shape[0]
    assert depth < 60
    n = len(refs)
    assert depth <= 3 * int(np.ceil(np.log2(max(n, 2)))) + 8


def test_sah_tree_is_cheaper_than_the_reference_median_split():
    """SAH cost (sum of area(node)/area(root) over nodes + leaf prims) of the device tree vs the reference's."""
    def cost(host):
        _, _, _, _, nodes, _ = arrays(host.desc)
        def area(n):
            e = np.array(n[1]) - np.array(n[0])
            return 2 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0])
        root = area(nodes[0])
        return sum(area(n) / root * (1 if n[5] == 0 else n[5]) for n in nodes)
    sah = cost(gs.HostScene.build("c4").flatten())
    ref = cost(gs.HostScene.build("c4").flatten(reference_tree=True))
    assert sah < ref


def test_gsh_bvh_build_entry_point():
    rng = np.random.default_rng(0)
    n = 500
    c = rng.random((n, 3)) * 10
    b = np.concatenate([c - 0.1, c + 0.1], 1).astype(np.float64)
    nodes = (ffi.gs_bvh_node * (2 * n))()
    order = np.empty(n, dtype=np.uint32)
    nn = C.c_size_t()
    ffi.check_host(ffi.host_lib().gsh_bvh_build(b.ctypes.data, n, 4, nodes, C.byref(nn), order.ctypes.data))
    assert sorted(order.tolist()) == list(range(n)) and 0 < nn.value <= 2 * n - 1
    leaves = sum(nodes[i].ref_count for i in range(nn.value))
    assert leaves == n


def _bvh(bounds, threads):
    """gsh_bvh_build with GSH_BUILD_THREADS set -> (nodes as python tuples, order)"""
    n = bounds.shape[0]
    nodes = (ffi.gs_bvh_node * (2 * n))()
    order = np.empty(n, dtype=np.uint32)
    nn = C.c_size_t()
    old = os.environ.get("GSH_BUILD_THREADS")
    os.environ["GSH_BUILD_THREADS"] = str(threads)
    try:
        ffi.check_host(ffi.host_lib().gsh_bvh_build(bounds.ctypes.data, n, 1, nodes, C.byref(nn), order.ctypes.data))
    finally:
        if old is None:
            del os.environ["GSH_BUILD_THREADS"]
        else:
            os.environ["GSH_BUILD_THREADS"] = old
    return [nodes[i] for i in range(nn.value)], order


def _canonical(nodes, order, i=0):
    """the tree as nested tuples, independent of how the nodes are numbered"""
    nd = nodes[i]
    box = (tuple(nd.bounds_min), tuple(nd.bounds_max))
    if nd.ref_count > 0:
        return (box, tuple(sorted(int(order[nd.first_ref + k]) for k in range(nd.ref_count))))
    assert nd.left > i and nd.right > i   # gs_scene_create requires children after their parent
    return (box, _canonical(nodes, order, nd.left), _canonical(nodes, order, nd.right))


def test_threaded_bvh_build_gives_the_same_tree():
    """subtrees built by several threads (scenes of >= 50 000 primitives; forced here on a small one): same splits, same
    boxes, same leaves as the single-threaded build, and the same node array on every run"""
    sys.setrecursionlimit(10000)
    rng = np.random.default_rng(3)
    n = 6000
    c = rng.random((n, 3)) * np.array([40.0, 5.0, 20.0])
    r = 0.05 + 0.4 * rng.random((n, 1))
    b = np.ascontiguousarray(np.concatenate([c - r, c + r], 1), dtype=np.float64)
    seq_nodes, seq_order = _bvh(b, 1)
    par_nodes, par_order = _bvh(b, 4)
    assert len(seq_nodes) == len(par_nodes) == 2 * n - 1
    assert _canonical(seq_nodes, seq_order) == _canonical(par_nodes, par_order)
    again_nodes, again_order = _bvh(b, 4)
    assert np.array_equal(par_order, again_order)
    assert all((a.left, a.right, a.first_ref, a.ref_count) == (x.left, x.right, x.first_ref, x.ref_count) for a, x in zip(par_nodes, again_nodes))


def test_camera_abi_matches_the_oracle_camera():
    host = gs.HostScene.build("c1")
    cam, st = host.camera()
    ora = orc.OracleScene.build("c1")
    assert (cam.image_width, cam.image_height, cam.max_depth) == (ora.width, ora.height, ora.max_depth) == (400, 225, 50)
    assert (st.batch_size, st.max_samples, st.tolerance) == (100, 0, 0.0)      # fixed-spp mapping of config c1
    assert gs.fixed_spp(st) == 100
    assert cam.background_kind == 1 and list(cam.hdri_rotation) == [0.0, -90.0, 90.0]  # radians, passed verbatim (Q12)
    # ray through the centre of pixel (i, j) = pixel00 + i du + j dv; compare with the oracle's jittered rays (|jitter| <= .5 px)
    rays = ora.camera_rays(seed=1, sample=0)
    p00, du, dv = np.array(cam.starting_pixel_pos), np.array(cam.pixel_delta_u), np.array(cam.pixel_delta_v)
    for n in (0, 399, 400 * 100 + 7, 400 * 225 - 1):
        i, j = n % 400, n // 400
        target = rays[n, 0:3] + rays[n, 3:6]
        off = target - (p00 + i * du + j * dv)
        assert abs(off @ du) / (du @ du) <= 0.5 + 1e-9 and abs(off @ dv) / (dv @ dv) <= 0.5 + 1e-9
    adaptive = gs.HostScene.build("scene0").camera()[1]
    assert (adaptive.confidence, adaptive.tolerance, adaptive.batch_size, adaptive.max_samples) == (0.95, 0.05, 64, 200)
    assert gs.fixed_spp(adaptive) is None


def test_perlin_table_matches_between_host_and_oracle():
    out = np.empty(256, dtype=np.uint8)
    ffi.host_lib().gsh_perlin_table(0, out.ctypes.data)
    assert np.array_equal(out, orc.perlin_table(0))
    ffi.host_lib().gsh_perlin_table(1234, out.ctypes.data)
    assert np.array_equal(out, orc.perlin_table(1234))


def test_media_are_kept_out_of_the_bvh():
    host = gs.HostScene.build("c4")
    d = host.desc.contents
    assert d.n_media == 2 and d.n_boundary_spheres == 2 and d.n_boundary_quads == 0
    assert d.n_spheres == 1005 and d.n_quads == 2401  # 1000 balls + moving, glass, metal, earth, noise; 400 cubes + light
    by_index = {d.media[k].medium_index: d.media[k] for k in range(2)}  # emitted in visiting order, indexed by construction order
    assert by_index[0].density == 0.2 and by_index[1].density == 0.0001
    assert d.materials[by_index[0].phase_material].kind == 0  # a Lambertian phase function (main.rs:720, Q10)
    s7 = gs.HostScene.build("scene7").desc.contents
    assert s7.n_media == 2 and s7.n_boundary_quads == 12 and s7.n_quads == 6
    assert s7.materials[s7.media[0].phase_material].kind == 4  # Isotropic
