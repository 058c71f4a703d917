"""The five BASELINE.json configs, restated from the reference's example scripts with
the host mirror of rpt's API (rpt_b200.api).  Resolution / spp / bounces come from
BASELINE.json; where it is silent, from the example file.

    sphere   examples/sphere.rs:4-34
    cornell  examples/cornell.rs:14-88
    teapot   examples/teapot.rs:10-34
    dragon   examples/dragon.rs:32-75   (mesh: examples/pegasus.zip subdivided to 801 104 triangles, see
                                         pegasus_proxy; `dragon-knot` keeps round 1's procedural tube as a second case)
    glass    examples/glass.rs:27-49    (HDRI: synthetic stand-in, see synthetic_hdri)

and, for the two-level kd-trees and the MonomialSurface, not BASELINE configs:

    fractal_spheres  examples/fractal_spheres.rs:3-77
    fractal_teapots  examples/fractal_teapots.rs:8-88
    monomial_glass   examples/monomial_glass.rs:25-88  (synthetic HDRI)
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

from .api import (Camera, Environment, Hdri, KdTree, Light, Material, Mesh, Object, Scene, cube, hex_color,
                  monomial_surface, plane, polygon, sphere, vec3)

_ASSETS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "assets")


@dataclass
class Config:
    name: str
    scene: Scene
    camera: Camera
    width: int
    height: int
    spp: int
    max_bounces: int
    note: str = ""

    def nominal_segments(self) -> int:
        return self.width * self.height * self.spp * (self.max_bounces + 1)


def sphere_scene() -> Config:
    scene = Scene()
    scene.add(Object(sphere()))  # default red material
    scene.add(Object(plane(vec3(0.0, 1.0, 0.0), -1.0)).material(Material.diffuse(hex_color(0xAAAAAA))))
    scene.add(Light.Object(
        Object(sphere().scale(vec3(2.0, 2.0, 2.0)).translate(vec3(0.0, 12.0, 0.0)))
        .material(Material.light(hex_color(0xFFFFFF), 40.0))))
    camera = Camera.look_at(vec3(-2.5, 4.0, 6.5), vec3(0.0, -0.25, 0.0), vec3(0.0, 1.0, 0.0), math.pi / 4)
    return Config("sphere", scene, camera, 960, 540, 100, 2, "examples/sphere.rs verbatim")


def cornell_scene() -> Config:
    scene = Scene()
    camera = Camera(eye=vec3(278.0, 273.0, -800.0), direction=vec3(0.0, 0.0, 1.0), up=vec3(0.0, 1.0, 0.0), fov=0.686)
    white = Material.diffuse(hex_color(0xAAAAAA))
    red = Material.diffuse(hex_color(0xBC0000))
    green = Material.diffuse(hex_color(0x00BC00))
    light_mtl = Material.light(hex_color(0xFFFEFA), 100.0)  # 6500 K
    floor = polygon([vec3(0.0, 0.0, 0.0), vec3(0.0, 0.0, 559.2), vec3(556.0, 0.0, 559.2), vec3(556.0, 0.0, 0.0)])
    ceiling = polygon([vec3(0.0, 548.9, 0.0), vec3(556.0, 548.9, 0.0), vec3(556.0, 548.9, 559.2),
                       vec3(0.0, 548.9, 559.2)])
    light_rect = polygon([vec3(343.0, 548.8, 227.0), vec3(343.0, 548.8, 332.0), vec3(213.0, 548.8, 332.0),
                          vec3(213.0, 548.8, 227.0)])
    back_wall = polygon([vec3(0.0, 0.0, 559.2), vec3(0.0, 548.9, 559.2), vec3(556.0, 548.9, 559.2),
                         vec3(556.0, 0.0, 559.2)])
    right_wall = polygon([vec3(0.0, 0.0, 0.0), vec3(0.0, 548.9, 0.0), vec3(0.0, 548.9, 559.2), vec3(0.0, 0.0, 559.2)])
    left_wall = polygon([vec3(556.0, 0.0, 0.0), vec3(556.0, 0.0, 559.2), vec3(556.0, 548.9, 559.2),
                         vec3(556.0, 548.9, 0.0)])
    large_box = (cube().scale(vec3(165.0, 330.0, 165.0)).rotate_y(2.0 * math.pi * (-253.0 / 360.0))
                 .translate(vec3(368.0, 165.0, 351.0)))
    small_box = (cube().scale(vec3(165.0, 165.0, 165.0)).rotate_y(2.0 * math.pi * (-197.0 / 360.0))
                 .translate(vec3(185.0, 82.5, 169.0)))
    scene.add(Object(floor).material(white))
    scene.add(Object(ceiling).material(white))
    scene.add(Object(back_wall).material(white))
    scene.add(Object(left_wall).material(red))
    scene.add(Object(right_wall).material(green))
    scene.add(Object(large_box).material(white))
    scene.add(Object(small_box).material(white))
    scene.add(Light.Object(Object(light_rect).material(light_mtl)))
    return Config("cornell", scene, camera, 800, 800, 512, 6,
                  "examples/cornell.rs geometry; 800x800x512 spp, max_bounces 6 per BASELINE.json")


def teapot_triangles() -> np.ndarray:
    return np.load(os.path.join(_ASSETS, "teapot_tris.npz"))["tris"]


def teapot_scene() -> Config:
    scene = Scene()
    teapot = Mesh(teapot_triangles())
    scene.add(Object(teapot.scale(vec3(0.5, 0.5, 0.5)).translate(vec3(0.0, -1.0, 0.0)))
              .material(Material.metallic_(hex_color(0xFF0000), 0.4)))
    scene.add(Object(plane(vec3(0.0, 1.0, 0.0), -1.0)).material(Material.diffuse(hex_color(0xAAAAAA))))
    scene.add(Light.Ambient(vec3(0.02, 0.02, 0.02)))
    scene.add(Light.Point(vec3(60.0, 60.0, 60.0), vec3(0.0, 5.0, 5.0)))
    return Config("teapot", scene, Camera.default(), 1920, 1080, 256, 0,
                  "examples/teapot.rs; teapot.obj has 2256 triangles; max_bounces 0 is the example's default")


def pegasus_indexed():
    """examples/pegasus.zip as parsed by load_obj (tools/make_assets.py): vertices, vertex normals, faces."""
    z = np.load(os.path.join(_ASSETS, "pegasus_indexed.npz"))
    return z["verts"], z["norms"], z["faces"]


def _unit(n: np.ndarray) -> np.ndarray:
    """normalize rows; a zero row (pegasus.obj has a few zero vertex normals) stays zero"""
    l = np.linalg.norm(n, axis=1, keepdims=True)
    return np.divide(n, l, out=np.zeros_like(n), where=l > 0)


def subdivide4(v: np.ndarray, n: np.ndarray, f: np.ndarray, alpha: float = 0.75):
    """1 -> 4: a vertex on every edge (shared by the two faces of the edge, so the mesh stays closed), moved off
    the flat midpoint by Phong tessellation -- the mean of its projections onto the tangent planes of the edge's
    two ends, blended in by `alpha` -- so the finer mesh is a smoother surface, not four coplanar pieces."""
    e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = np.sort(e, axis=1).astype(np.int64)
    uniq, inv = np.unique(key[:, 0] * len(v) + key[:, 1], return_inverse=True)
    a, b = (uniq // len(v)).astype(np.int64), (uniq % len(v)).astype(np.int64)
    m = 0.5 * (v[a] + v[b])
    na, nb = _unit(n[a]), _unit(n[b])
    pa = m - np.sum((m - v[a]) * na, axis=1, keepdims=True) * na
    pb = m - np.sum((m - v[b]) * nb, axis=1, keepdims=True) * nb
    mid = (1.0 - alpha) * m + alpha * 0.5 * (pa + pb)
    v2 = np.concatenate([v, mid])
    n2 = np.concatenate([n, _unit(na + nb)])
    nf = len(f)
    m01, m12, m20 = (len(v) + inv[0:nf]), (len(v) + inv[nf:2 * nf]), (len(v) + inv[2 * nf:3 * nf])
    i0, i1, i2 = f[:, 0].astype(np.int64), f[:, 1].astype(np.int64), f[:, 2].astype(np.int64)
    f2 = np.stack([np.stack([i0, m01, m20], 1), np.stack([m01, i1, m12], 1), np.stack([m20, m12, i2], 1),
                   np.stack([m01, m12, m20], 1)], axis=1).reshape(-1, 3)
    return v2, n2, f2


def subdivide2(v: np.ndarray, n: np.ndarray, f: np.ndarray) -> np.ndarray:
    """1 -> 2: every triangle is cut from the (linear) midpoint of its longest edge to the opposite corner.
    Returns (2 * len(f), 18) rows v1 v2 v3 n1 n2 n3, the input of Mesh::new."""
    p = v[f]  # (F, 3, 3)
    q = n[f]
    l = np.stack([np.sum((p[:, 1] - p[:, 0]) ** 2, 1), np.sum((p[:, 2] - p[:, 1]) ** 2, 1), np.sum((p[:, 0] - p[:, 2]) ** 2, 1)], 1)
    k = np.argmax(l, axis=1)  # edge k joins corner k and k + 1
    r = np.arange(len(f))
    ia, ib, ic = k, (k + 1) % 3, (k + 2) % 3
    A, B, Cc = p[r, ia], p[r, ib], p[r, ic]
    nA, nB, nC = q[r, ia], q[r, ib], q[r, ic]
    Mid = 0.5 * (A + B)
    nM = _unit(_unit(nA) + _unit(nB))
    t1 = np.concatenate([A, Mid, Cc, nA, nM, nC], axis=1)  # (a, m, c) and (m, b, c): the winding of (a, b, c)
    t2 = np.concatenate([Mid, B, Cc, nM, nB, nC], axis=1)
    return np.ascontiguousarray(np.stack([t1, t2], axis=1).reshape(-1, 18))


def pegasus_proxy(level: int = 2) -> np.ndarray:
    """The offline stand-in SURVEY 8(d) fixes for the Stanford dragon (an HTTP download in the reference,
    examples/dragon.rs:11-14; 871 414 triangles): the scanned statue of examples/pegasus.zip (100 138 triangles,
    edge lengths 4e-4 .. 6e-2: thin legs and wings, concavities) subdivided 1 -> 4 then 1 -> 2 = 801 104
    triangles, and put where the G3D dragon.obj sits -- longest dimension 0.7, centred in x and z, resting on
    y = -1/3.4 so that examples/dragon.rs's `scale 3.4` stands it on the plane y = -1.  Every table labels
    it "dragon-proxy".  level 0 = the 100 138 original triangles, 1 = 400 552, 2 = 801 104."""
    v, n, f = pegasus_indexed()
    lo, hi = v.min(0), v.max(0)
    s = 0.7 / float((hi - lo).max())
    c = 0.5 * (lo + hi)
    v = (v - np.array([c[0], lo[1], c[2]])) * s
    v[:, 1] += -1.0 / 3.4
    if level >= 1:
        v, n, f = subdivide4(v, n, f)
    if level >= 2:
        return subdivide2(v, n, f)
    return np.ascontiguousarray(np.concatenate([v[f].reshape(-1, 9), n[f].reshape(-1, 9)], axis=1))


def knot_proxy(n_u: int = 1320, n_v: int = 330, seed: int = 7) -> np.ndarray:
    """Round 1's stand-in, kept as a second case: a closed, bumpy (2,3) torus-knot tube with n_u*n_v*2
    triangles (871 200 by default), smooth vertex normals, scaled into the dragon's bounding box.  Uniform
    triangle size, no thin features: the easiest input a mesh of that size can be for a kd-tree or a BVH."""
    rng = np.random.default_rng(seed)
    u = np.linspace(0.0, 2.0 * np.pi, n_u, endpoint=False)
    v = np.linspace(0.0, 2.0 * np.pi, n_v, endpoint=False)
    p, q = 2.0, 3.0
    # centre curve of the knot and a parallel-transport-free frame (Frenet is fine: no inflections)
    r = np.cos(q * u) + 2.0
    c = np.stack([r * np.cos(p * u), -np.sin(q * u), r * np.sin(p * u)], axis=1)
    du = 1e-4
    r2 = np.cos(q * (u + du)) + 2.0
    c2 = np.stack([r2 * np.cos(p * (u + du)), -np.sin(q * (u + du)), r2 * np.sin(p * (u + du))], axis=1)
    t = c2 - c
    t /= np.linalg.norm(t, axis=1, keepdims=True)
    up = np.array([0.0, 1.0, 0.0])
    b = np.cross(t, up)
    b /= np.linalg.norm(b, axis=1, keepdims=True)
    n = np.cross(b, t)
    uu, vv = np.meshgrid(u, v, indexing="ij")
    # scales ("bumps"): a few seeded octaves of periodic displacement of the tube radius
    disp = np.zeros_like(uu)
    for k in range(6):
        fu, fv = int(rng.integers(3, 90)), int(rng.integers(1, 24))
        ph = rng.uniform(0, 2 * np.pi, 2)
        disp += (0.5 ** (k * 0.5)) * np.sin(fu * uu + ph[0]) * np.sin(fv * vv + ph[1])
    rad = 0.42 * (1.0 + 0.11 * disp)
    pos = (c[:, None, :] + rad[..., None] * (np.cos(vv)[..., None] * n[:, None, :] + np.sin(vv)[..., None] * b[:, None, :]))
    # fit into the dragon's box: x in [-0.5, 0.5], y_min = -1/3.4
    lo, hi = pos.reshape(-1, 3).min(0), pos.reshape(-1, 3).max(0)
    s = 1.0 / (hi[0] - lo[0])
    pos = (pos - (lo + hi) / 2.0) * s
    pos[..., 1] += (-1.0 / 3.4) - pos[..., 1].min()
    i0 = np.arange(n_u)[:, None]
    j0 = np.arange(n_v)[None, :]
    i1, j1 = (i0 + 1) % n_u, (j0 + 1) % n_v
    a, bq, cq, dq = pos[i0, j0], pos[i1, j0], pos[i1, j1], pos[i0, j1]
    # smooth normals: area-weighted face normals accumulated on the grid vertices
    fn1 = np.cross(bq - a, cq - a)
    fn2 = np.cross(cq - a, dq - a)
    vn = np.zeros_like(pos)
    for (ii, jj), f in (((i0, j0), fn1 + fn2), ((i1, j0), fn1), ((i1, j1), fn1 + fn2), ((i0, j1), fn2)):
        np.add.at(vn, (np.broadcast_to(ii, fn1.shape[:2]), np.broadcast_to(jj, fn1.shape[:2])), f)
    vn /= np.linalg.norm(vn, axis=2, keepdims=True)
    na, nb, nc, nd = vn[i0, j0], vn[i1, j0], vn[i1, j1], vn[i0, j1]
    # wind the triangles (and flip the accumulated normals) so that both point out of the tube
    t1 = np.concatenate([a, cq, bq, -na, -nc, -nb], axis=2)
    t2 = np.concatenate([a, dq, cq, -na, -nd, -nc], axis=2)
    tris = np.stack([t1, t2], axis=2).reshape(-1, 18)
    return np.ascontiguousarray(tris)


_DRAGON_CACHE = {}


def dragon_mesh(level: int = 2, knot=None) -> Mesh:
    """knot = (n_u, n_v) selects round 1's procedural tube instead of the pegasus-derived mesh"""
    key = ("knot",) + tuple(knot) if knot else ("pegasus", level)
    if key not in _DRAGON_CACHE:
        path = os.environ.get("RPT_DRAGON_OBJ")
        if path and os.path.exists(path) and not knot:
            from .api import load_obj
            _DRAGON_CACHE[key] = load_obj(path)
        else:
            _DRAGON_CACHE[key] = Mesh(knot_proxy(*knot) if knot else pegasus_proxy(level))
    return _DRAGON_CACHE[key]


def dragon_scene(level: int = 2, knot=None) -> Config:
    scene = Scene()
    dragon = dragon_mesh(level, knot)
    scene.add(Object(dragon.scale(vec3(3.4, 3.4, 3.4)).rotate_y(math.pi / 2))
              .material(Material.specular(hex_color(0xB7CA79), 0.1)))
    scene.add(Object(plane(vec3(0.0, 1.0, 0.0), -1.0)).material(Material.diffuse(hex_color(0xAAAAAA))))
    scene.add(Light.Ambient(vec3(0.01, 0.01, 0.01)))
    scene.add(Light.Object(Object(sphere().scale(vec3(2.0, 2.0, 2.0)).translate(vec3(0.0, 20.0, 3.0)))
                           .material(Material.light(vec3(1.0, 1.0, 1.0), 160.0))))
    scene.add(Light.Object(Object(sphere().scale(vec3(0.05, 0.05, 0.05)).translate(vec3(-1.0, 0.71, 0.0)))
                           .material(Material.light(hex_color(0xFFAAAA), 400.0))))
    camera = Camera.look_at(vec3(-2.5, 4.0, 6.5), vec3(0.0, 0.0, 0.0), vec3(0.0, 1.0, 0.0), math.pi / 6)
    real = bool(os.environ.get("RPT_DRAGON_OBJ")) and not knot
    name = "dragon" if real else "dragon-knot" if knot else "dragon-proxy"
    mesh = "dragon.obj" if real else "procedural torus-knot tube" if knot else "examples/pegasus.zip subdivided (level %d)" % level
    return Config(name, scene, camera, 1920, 1080, 1024, 2, "examples/dragon.rs layout; mesh = %s (%d triangles)" % (mesh, len(dragon)))


def dragon_knot_scene(n_u: int = 1320, n_v: int = 330) -> Config:
    return dragon_scene(knot=(n_u, n_v))


def synthetic_hdri(width: int = 2048, height: int = 1024, seed: int = 11) -> Hdri:
    """Stand-in for ballroom_2k.hdr (examples/glass.rs:30 fetches it over HTTPS): a smooth
    vertical gradient 0.05..1.0 plus 8 Gaussian lamps of peak radiance 50 at seeded
    directions, generated identically for the oracle and the GPU."""
    rng = np.random.default_rng(seed)
    y = (np.arange(height) + 0.5) / height
    x = (np.arange(width) + 0.5) / width
    base = 1.0 - 0.95 * y  # bright zenith, dim nadir
    img = np.repeat(base[:, None, None], width, axis=1) * np.array([0.9, 0.95, 1.0])[None, None, :]
    img = np.repeat(img, 1, axis=2)
    for _ in range(8):
        cx, cy = rng.uniform(0, 1), rng.uniform(0.1, 0.6)
        sx, sy = rng.uniform(0.004, 0.02), rng.uniform(0.004, 0.02)
        tint = rng.uniform(0.7, 1.0, 3)
        dxw = np.minimum(np.abs(x - cx), 1.0 - np.abs(x - cx))  # wrap in azimuth
        g = np.exp(-0.5 * (dxw[None, :] / sx) ** 2 - 0.5 * ((y[:, None] - cy) / sy) ** 2)
        img = img + 50.0 * g[..., None] * tint[None, None, :]
    return Hdri(width, height, img.reshape(-1, 3))


def glass_scene(hdri_width: int = 2048, hdri_height: int = 1024) -> Config:
    scene = Scene()
    scene.environment = Environment.Hdri(synthetic_hdri(hdri_width, hdri_height))
    scene.add(Object(sphere().translate(vec3(1.1, 0.0, 0.0))).material(Material.metallic_(hex_color(0xFFFFFF), 0.0001)))
    scene.add(Object(sphere().translate(vec3(-1.1, 0.0, 0.0))).material(Material.clear(1.5, 0.0001)))
    return Config("glass", scene, Camera.default(), 1920, 1080, 4096, 12,
                  "examples/glass.rs; synthetic HDRI; 1920x1080x4096 spp, max_bounces 12 per BASELINE.json")


FRACTAL_COLORS = [0x264653, 0x2A9D8F, 0xE9C46A, 0xF4A261, 0xE76F51]


def _fractal(levels: int, make):
    """gen() of examples/fractal_spheres.rs:3-32 / fractal_teapots.rs:8-41: one shape at p with radius
    rad, then five (six at the root) children at distance 7/5 rad with radius 2/5 rad, skipping the
    direction that leads back; shapes are grouped by depth."""
    groups = [[] for _ in range(levels)]

    def gen(p, rad, depth, last_dir):
        groups[depth].append(make(p, rad))
        if depth == levels - 1:
            return
        disp = rad * 7.0 / 5.0
        dx = [disp, -disp, 0.0, 0.0, 0.0, 0.0]
        dy = [0.0, 0.0, disp, -disp, 0.0, 0.0]
        dz = [0.0, 0.0, 0.0, 0.0, disp, -disp]
        for i in range(6):
            if last_dir is None or i != (last_dir ^ 1):
                gen(p + vec3(dx[i], dy[i], dz[i]), rad * 2.0 / 5.0, depth + 1, i)

    gen(vec3(0.0, 0.0, 0.0), 1.0, 0, None)
    return groups


def _fractal_scene(name: str, groups, levels: int, note: str) -> Config:
    scene = Scene()
    for i, group in enumerate(groups):
        scene.add(Object(KdTree(group)).material(Material.specular(hex_color(FRACTAL_COLORS[i]), 0.25)))
    scene.add(Object(plane(vec3(0.0, 0.0, 1.0), -6.0)).material(Material.diffuse(hex_color(0xFFCCCC))))
    scene.add(Light.Ambient(vec3(0.02, 0.02, 0.02)))
    d = vec3(0.0, -0.65, -1.0)
    scene.add(Light.Directional(vec3(0.6, 0.6, 0.6), d / np.linalg.norm(d)))
    scene.add(Light.Point(vec3(100.0, 100.0, 100.0), vec3(0.0, 5.0, 5.0)))
    direction = vec3(-0.285714, -0.5, -1.0)
    up = vec3(0.0, 1.0, -0.5)
    camera = Camera(eye=vec3(2.0, 3.5, 7.0), direction=direction / np.linalg.norm(direction),
                    up=up / np.linalg.norm(up), fov=math.pi / 6)
    return Config(name, scene, camera, 800, 600, 1, 0, note)


def fractal_spheres_scene(levels: int = 5) -> Config:
    groups = _fractal(levels, lambda p, rad: sphere().scale(vec3(rad, rad, rad)).translate(p))
    return _fractal_scene("fractal_spheres", groups, levels,
                          "examples/fractal_spheres.rs: one KdTree<Box<dyn Bounded>> of spheres per level (1, 6, 30, 150, 750)")


def fractal_teapots_scene(levels: int = 5) -> Config:
    teapot = Mesh(teapot_triangles())  # Arc<Mesh>: every instance shares this one kd-tree

    def make(p, rad):
        return teapot.scale(vec3(0.5, 0.5, 0.5)).scale(vec3(rad, rad, rad)).translate(p)

    return _fractal_scene("fractal_teapots", _fractal(levels, make), levels,
                          "examples/fractal_teapots.rs: kd-trees of transformed instances of one teapot kd-tree")


def monomial_glass_scene(hdri_width: int = 512, hdri_height: int = 256) -> Config:
    scene = Scene()
    scene.environment = Environment.Hdri(synthetic_hdri(hdri_width, hdri_height))
    scene.add(Object(monomial_surface(2.0, 4.0).translate(vec3(0.0, -1.0, 0.0)))
              .material(Material.metallic_(hex_color(0xFFFFFF), 0.0001)))
    scene.add(Object(cube().rotate_y(math.pi / 6.0).scale(vec3(0.5, 0.3, 0.4)).translate(vec3(0.4, -0.8, 4.0)))
              .material(Material.specular(hex_color(0xFF00FF), 0.5)))
    scene.add(Object(sphere().scale(vec3(0.5, 0.5, 0.5)).translate(vec3(1.5, -0.5, 1.0)))
              .material(Material.specular(hex_color(0x0000FF), 0.1)))
    scene.add(Object(sphere().scale(vec3(0.5, 0.5, 0.5)).translate(vec3(-1.5, -0.5, 1.0)))
              .material(Material.specular(hex_color(0x00FF00), 0.1)))
    scene.add(Object(plane(vec3(0.0, 1.0, 0.0), -1.0)).material(Material.specular(hex_color(0xAAAAAA), 0.5)))
    scene.add(Light.Ambient(vec3(0.01, 0.01, 0.01)))
    scene.add(Light.Point(vec3(100.0, 100.0, 100.0), vec3(0.0, 5.0, 5.0)))
    return Config("monomial_glass", scene, Camera.default(), 800, 600, 100, 1,
                  "examples/monomial_glass.rs; synthetic HDRI")


EXTRA_CONFIGS = {
    "fractal_spheres": fractal_spheres_scene,
    "fractal_teapots": fractal_teapots_scene,
    "monomial_glass": monomial_glass_scene,
    "dragon_knot": dragon_knot_scene,
}

CONFIGS = {
    "sphere": sphere_scene,
    "cornell": cornell_scene,
    "teapot": teapot_scene,
    "dragon": dragon_scene,
    "glass": glass_scene,
}
