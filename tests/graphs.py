"""Tiny authoring helper for the "gsgraph 1" scene-graph text format (grayshift_b200/host/graph_io.cpp,
oracle/graph_io.cpp) so tests can hand the SAME arbitrary scene to the host layer and to the oracle."""
import numpy as np


def _f(v):
    return repr(float(v))


def _v(v):
    return " ".join(_f(x) for x in v)


class Graph:
    def __init__(self):
        self.lines = ["gsgraph 1"]
        self.n_img = self.n_tex = self.n_mat = self.n_node = 0
        self.n_prim = 0
        self.n_medium = 0
        self.camera_line = None
        self.background_line = "background solid 0.7 0.8 1.0"
        self.root = None

    # images / textures / materials
    def image(self, kind, w, h, source):
        self.lines.append(f"image {self.n_img} {kind} {w} {h} {source}")
        self.n_img += 1
        return self.n_img - 1

    def _tex(self, rec):
        self.lines.append(f"texture {self.n_tex} {rec}")
        self.n_tex += 1
        return self.n_tex - 1

    def solid(self, c):
        return self._tex("solid " + _v(c))

    def checker(self, scale, even, odd):
        return self._tex(f"checker {_f(scale)} {even} {odd}")

    def image_texture(self, image):
        return self._tex(f"image {image}")

    def noise(self, scale):
        return self._tex(f"noise {_f(scale)}")

    def _mat(self, rec):
        self.lines.append(f"material {self.n_mat} {rec}")
        self.n_mat += 1
        return self.n_mat - 1

    def lambertian(self, tex):
        return self._mat(f"lambertian {tex}")

    def lambertian_color(self, c):
        return self.lambertian(self.solid(c))

    def metal(self, albedo, fuzz):
        return self._mat(f"metal {_v(albedo)} {_f(fuzz)}")

    def dielectric(self, ir):
        return self._mat(f"dielectric {_f(ir)}")

    def light(self, tex):
        return self._mat(f"light {tex}")

    def light_color(self, c):
        return self.light(self.solid(c))

    def isotropic(self, tex):
        return self._mat(f"isotropic {tex}")

    # nodes
    def _node(self, rec):
        self.lines.append(f"node {self.n_node} {rec}")
        self.n_node += 1
        return self.n_node - 1

    def _prim(self):
        self.n_prim += 1
        return self.n_prim - 1

    def sphere(self, c, r, mat):
        return self._node(f"sphere {_v(c)} {_f(r)} {mat} {self._prim()}")

    def moving_sphere(self, c1, c2, r, mat):
        return self._node(f"msphere {_v(c1)} {_v(c2)} {_f(r)} {mat} {self._prim()}")

    def quad(self, q, u, v, mat):
        return self._node(f"quad {_v(q)} {_v(u)} {_v(v)} {mat} {self._prim()}")

    def triangle(self, a, b, c, mat):
        return self._node(f"tri {_v(a)} {_v(b)} {_v(c)} {mat} {self._prim()}")

    def cube(self, a, b, mat):
        """Quad::cube, src/hittable/quad.rs:54-80"""
        a, b = np.asarray(a, float), np.asarray(b, float)
        mn, mx = np.minimum(a, b), np.maximum(a, b)
        dx, dy, dz = np.array([mx[0] - mn[0], 0, 0]), np.array([0, mx[1] - mn[1], 0]), np.array([0, 0, mx[2] - mn[2]])
        sides = [self.quad([mn[0], mn[1], mx[2]], dx, dy, mat), self.quad([mx[0], mn[1], mx[2]], -dz, dy, mat),
                 self.quad([mx[0], mn[1], mn[2]], -dx, dy, mat), self.quad([mn[0], mn[1], mn[2]], dz, dy, mat),
                 self.quad([mn[0], mx[1], mx[2]], dx, -dz, mat), self.quad([mn[0], mn[1], mn[2]], dx, dz, mat)]
        return self.list(sides)

    def list(self, ids):
        return self._node(f"list {len(ids)} " + " ".join(str(i) for i in ids))

    def bvh(self, ids):
        return self._node(f"bvh {len(ids)} " + " ".join(str(i) for i in ids))

    def translate(self, child, off):
        return self._node(f"translate {child} {_v(off)}")

    def rotate_y(self, child, deg):
        return self._node(f"rotate_y {child} {_f(deg)}")

    def medium(self, boundary, density, phase_mat):
        self.n_medium += 1
        return self._node(f"medium {boundary} {_f(density)} {phase_mat} {self.n_medium - 1}")

    # camera
    def camera(self, aspect=1.0, width=64, confidence=0.95, tolerance=0.0, batch=8, max_samples=0, depth=10, vfov=40.0,
               look_from=(0, 0, 10), look_at=(0, 0, 0), vup=(0, 1, 0), defocus=0.0, focus=10.0):
        self.camera_line = (f"camera {_f(aspect)} {width} {_f(confidence)} {_f(tolerance)} {batch} {max_samples} {depth} {_f(vfov)} "
                            f"{_v(look_from)} {_v(look_at)} {_v(vup)} {_f(defocus)} {_f(focus)}")

    def background_solid(self, c):
        self.background_line = "background solid " + _v(c)

    def background_hdri(self, image, rot):
        self.background_line = f"background hdri {image} {_v(rot)}"

    def write(self, path, root):
        if self.camera_line is None:
            self.camera()
        text = "\n".join(self.lines + [f"root {root}", self.camera_line, self.background_line]) + "\n"
        with open(path, "w") as f:
            f.write(text)
        return path


def random_scene(path, seed, n_spheres=40, n_quads=20, n_cubes=6, with_transforms=True, extent=10.0, **camera):
    """A random nest of lists / BVHs / translated+rotated instances over spheres, quads and cubes."""
    rng = np.random.default_rng(seed)
    g = Graph()
    mats = [g.lambertian_color(rng.random(3)), g.metal(0.5 + 0.5 * rng.random(3), 0.3 * rng.random()), g.dielectric(1.5),
            g.lambertian(g.checker(0.7, g.solid(rng.random(3)), g.solid(rng.random(3)))), g.light_color([4, 4, 4])]
    items = []
    for _ in range(n_spheres):
        c = (rng.random(3) - 0.5) * 2 * extent
        if rng.random() < 0.15:
            items.append(g.moving_sphere(c, c + rng.normal(size=3) * 0.5, 0.2 + rng.random(), int(rng.integers(len(mats)))))
        else:
            items.append(g.sphere(c, 0.2 + 1.5 * rng.random(), int(rng.integers(len(mats)))))
    for _ in range(n_quads):
        q = (rng.random(3) - 0.5) * 2 * extent
        items.append(g.quad(q, rng.normal(size=3) * 2, rng.normal(size=3) * 2, int(rng.integers(len(mats)))))
    for _ in range(n_cubes):
        a = (rng.random(3) - 0.5) * 2 * extent
        cube = g.cube(a, a + 0.5 + 2 * rng.random(3), int(rng.integers(len(mats))))
        if with_transforms:
            cube = g.translate(g.rotate_y(cube, float(rng.uniform(-90, 90))), rng.normal(size=3) * 2)
        items.append(cube)
    rng.shuffle(items)
    third = len(items) // 3
    sub_bvh = g.bvh(items[:third])
    if with_transforms:
        sub_bvh = g.translate(g.rotate_y(sub_bvh, float(rng.uniform(-180, 180))), rng.normal(size=3))
    sub_list = g.list(items[third:2 * third])
    root = g.bvh([sub_bvh, sub_list] + items[2 * third:])
    cam = dict(width=64, batch=4, depth=8, look_from=(0, 2, 3 * extent), vfov=50.0)
    cam.update(camera)
    g.camera(**cam)
    g.write(path, root)
    return g
