"""CPU tests that pin the oracle on SURVEY 8f row N4: kd-trees over whole shapes
(KdTree<Box<dyn Bounded>>, src/kdtree.rs:99-223 as used by examples/fractal_spheres.rs and
examples/fractal_teapots.rs) and MonomialSurface (src/shape/monomial_surface.rs:13-187).

Nothing in the reference's tests touches these (its one monomial test covers closest_point, which is
not on the render path), so -- as for the rest of the path -- the oracle is pinned by independent
restatements: brute force over the children, numpy bounding boxes, numpy quartic roots."""
import math

import numpy as np
import pytest

from rpt_b200 import api, scenes
from rpt_b200 import _capi as capi
from tests import util


def small_scene(monomials: bool = False):
    """A bit of everything a KdTree<Box<dyn Bounded>> may hold, > 16 children so the tree really splits.
    MonomialSurface children only on request: its intersect is not a pure closest-hit query (which root
    it finds depends on the t_min the tree passes down, monomial_surface.rs:49-74), so with them the
    reference's result legitimately depends on the tree and brute force is not a valid check."""
    rng = np.random.default_rng(5)
    tea = api.Mesh(scenes.teapot_triangles())
    kids = []
    for i in range(40):
        p = rng.uniform(-4, 4, 3)
        s = rng.uniform(0.2, 0.7)
        base = [api.sphere(), api.cube(), tea, api.monomial_surface(1.5, 4.0) if monomials else api.sphere()][i % 4]
        kids.append(base.scale(api.vec3(s, s * rng.uniform(0.5, 1.5), s)).rotate_y(rng.uniform(0, 6.28)).translate(p))
    kids.append(api.sphere())                      # bare (untransformed) children too
    kids.append(api.cube())
    scene = api.Scene()
    scene.add(api.Object(api.KdTree(kids)).material(api.Material.specular(api.hex_color(0x2A9D8F), 0.25)))
    scene.add(api.Object(api.KdTree(kids[:5]).scale(api.vec3(0.5, 0.5, 0.5)).translate(api.vec3(0.0, 6.0, 0.0))))  # Transformed<KdTree>, one leaf
    scene.add(api.Object(api.plane(api.vec3(0.0, 1.0, 0.0), -5.0)))
    scene.add(api.Light.Point(api.vec3(50.0, 50.0, 50.0), api.vec3(0.0, 8.0, 3.0)))
    return scene, kids


def rays_into(rng, n, radius=9.0, target=4.0):
    o = util.random_unit(rng, n) * radius
    d = util.normalize(rng.uniform(-target, target, (n, 3)) - o)
    return np.concatenate([o, d], axis=1)


def test_group_tree_equals_brute_force(orc):
    scene, _ = small_scene()
    flat = api.FlatScene(scene)
    kd, bf = orc.OracleScene(flat), orc.OracleScene(flat, brute_force=True)
    rng = np.random.default_rng(1)
    rays = np.concatenate([rays_into(rng, 30000), util.interior_rays([-4, -4, -4], [4, 7, 4], 20000, rng)])
    t0, o0, n0, s0 = kd.closest_hit(rays)
    t1, o1, n1, s1 = bf.closest_hit(rays)
    np.testing.assert_array_equal(o0, o1)
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(n0, n1)
    assert s0["object_tests"] < s1["object_tests"]          # the tree prunes children
    assert ((o0 == 0).mean() > 0.2) and (o0 == 1).any() and (o0 == 2).any()


@pytest.mark.parametrize("name", ["fractal_spheres", "fractal_teapots"])
def test_fractal_examples_tree_equals_brute_force(orc, name):
    cfg = scenes.fractal_spheres_scene(5) if name == "fractal_spheres" else scenes.fractal_teapots_scene(3)
    flat = api.FlatScene(cfg.scene)
    assert flat.desc.ngroups == len(cfg.scene.objects) - 1
    assert [int(flat.groups[i].nchildren) for i in range(3)] == [1, 6, 30]
    if name == "fractal_teapots":
        assert flat.desc.nmeshes == 1                         # Arc<Mesh>: one teapot, 37 instances
    kd, bf = orc.OracleScene(flat), orc.OracleScene(flat, brute_force=True)
    rng = np.random.default_rng(2)
    rays = util.camera_rays(cfg.camera, 20000 if name == "fractal_spheres" else 3000, rng, spread=0.3)
    t0, o0, n0, _ = kd.closest_hit(rays)
    t1, o1, n1, _ = bf.closest_hit(rays)
    np.testing.assert_array_equal(o0, o1)
    np.testing.assert_array_equal(t0, t1)
    np.testing.assert_array_equal(n0, n1)
    assert len(np.unique(o0)) >= 4                            # several levels + the back plane are seen


def test_bounding_boxes_follow_the_reference_rules(orc):
    """Sphere [-1,1]^3, Cube [-.5,.5]^3, MonomialSurface (-1,0,-1)..(1,h,1), Mesh = its vertices' box,
    Transformed = box of the 8 transformed corners (src/shape.rs:153-175), KdTree = merge; Plane: none."""
    tea = api.Mesh(scenes.teapot_triangles(), build=False)
    m = api._translate([1.0, 2.0, 3.0]) @ api._rotate(0.7, (0.3, 1.0, 0.2)) @ api._scale([2.0, 0.5, 1.5])
    bases = {
        "sphere": (api.sphere(), [-1, -1, -1], [1, 1, 1]),
        "cube": (api.cube(), [-0.5, -0.5, -0.5], [0.5, 0.5, 0.5]),
        "monomial": (api.monomial_surface(2.5, 4.0), [-1, 0, -1], [1, 2.5, 1]),
        "mesh": (tea, tea.triangles[:, :9].reshape(-1, 3).min(0), tea.triangles[:, :9].reshape(-1, 3).max(0)),
    }
    scene = api.Scene()
    want = []
    for shape, lo, hi in bases.values():
        lo, hi = np.asarray(lo, float), np.asarray(hi, float)
        scene.add(api.Object(shape))
        want.append((lo, hi))
        scene.add(api.Object(api.Transformed(shape, m)))
        corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        w = (m @ corners.T).T[:, :3]
        want.append((w.min(0), w.max(0)))
    kids = [api.sphere().translate(api.vec3(3.0, 0.0, 0.0)), api.cube().translate(api.vec3(0.0, -2.0, 0.0))]
    scene.add(api.Object(api.KdTree(kids)))
    want.append((np.array([-0.5, -2.5, -1.0]), np.array([4.0, 1.0, 1.0])))
    scene.add(api.Object(api.plane(api.vec3(0.0, 1.0, 0.0), 0.0)))
    flat = api.FlatScene(scene)
    for i, (lo, hi) in enumerate(want):
        got = orc.shape_bounds(flat, i)
        np.testing.assert_allclose(got[0], lo, rtol=0, atol=1e-12)
        np.testing.assert_allclose(got[1], hi, rtol=0, atol=1e-12)
    assert orc.shape_bounds(flat, len(want)) is None


def test_box_tree_builder_matches_oracle(orc):
    """The library's `construct` over boxes (rptb_build_kdtree_boxes) emits the oracle's tree."""
    import ctypes as C
    rng = np.random.default_rng(9)
    for n in (1, 15, 16, 200, 5000):
        lo = rng.uniform(-10, 10, (n, 3))
        boxes = np.concatenate([lo, lo + rng.uniform(0.01, 2.0, (n, 3)) ** 2], axis=1)
        nodes, refs, depth, max_leaf = orc.build_kdtree(boxes, boxes=True)
        out = capi.KdTreeOut()
        capi.check(capi.lib().rptb_build_kdtree_boxes(boxes.ctypes.data_as(capi.c_double_p), n, C.byref(out)), "build")
        try:
            assert int(out.nnodes) == len(nodes) and int(out.nrefs) == len(refs)
            got = np.frombuffer(C.string_at(out.nodes, C.sizeof(capi.KdNode) * len(nodes)), dtype=orc.KDNODE_DTYPE)
            for f in ("split", "kind", "left", "right", "first_ref", "num_refs"):
                np.testing.assert_array_equal(got[f], nodes[f])
            np.testing.assert_array_equal(np.ctypeslib.as_array(out.refs, shape=(len(refs),)), refs)
            assert (int(out.depth), int(out.max_leaf)) == (depth, max_leaf)
        finally:
            capi.lib().rptb_free_kdtree(C.byref(out))


# ------------------------------------------------------------------ MonomialSurface -------
def monomial_scene(height=2.0):
    scene = api.Scene()
    scene.add(api.Object(api.monomial_surface(height, 4.0)))
    return scene


def quartic_roots(o, d, height):
    """All real t with  o.y + t d.y = height * ((o.x + t d.x)^2 + (o.z + t d.z)^2)^2."""
    c0 = o[0] ** 2 + o[2] ** 2
    c1 = 2 * (o[0] * d[0] + o[2] * d[2])
    c2 = d[0] ** 2 + d[2] ** 2
    q = np.polynomial.polynomial.polypow([c0, c1, c2], 2) * height
    q[0] -= o[1]
    q[1] -= d[1]
    r = np.polynomial.polynomial.polyroots(q)
    return np.sort(r[np.abs(r.imag) < 1e-9].real)


def test_monomial_hits_lie_on_the_surface_and_match_quartic_roots(orc):
    h = 2.0
    flat = api.FlatScene(monomial_scene(h))
    sc = orc.OracleScene(flat)
    rng = np.random.default_rng(4)
    o = rng.uniform(-2.5, 2.5, (20000, 3)) + np.array([0.0, 1.0, 0.0])
    tgt = np.stack([rng.uniform(-0.9, 0.9, 20000), rng.uniform(0.0, h, 20000), rng.uniform(-0.9, 0.9, 20000)], axis=1)
    d = util.normalize(tgt - o)
    rays = np.concatenate([o, d], axis=1)
    t, obj, nrm, _ = sc.closest_hit(rays)
    hit = obj == 0
    assert 0.5 < hit.mean() < 1.0
    p = o[hit] + t[hit, None] * d[hit]
    r2 = p[:, 0] ** 2 + p[:, 2] ** 2
    np.testing.assert_allclose(p[:, 1], h * r2 ** 2, atol=1e-9)            # on the surface (60 bisections)
    assert (r2 <= 1.0).all() and (t[hit] > 0).all()                          # inside the rim, in front
    # normal = +-normalize(4 h x r^2, -1, 4 h z r^2), turned against the ray (monomial_surface.rs:93-103)
    g = np.stack([4 * h * p[:, 0] * r2, -np.ones(len(p)), 4 * h * p[:, 2] * r2], axis=1)
    g = util.normalize(g)
    g = np.where(((g * d[hit]).sum(1) > 0)[:, None], -g, g)
    np.testing.assert_allclose(nrm[hit], g, atol=1e-9)
    # every reported t is a root of the quartic; for rays that start above the bowl (dist > 0) and cross the
    # surface exactly once in front, it is that crossing
    idx = np.nonzero(hit)[0][:1500]
    checked = 0
    for i in idx:
        roots = quartic_roots(o[i], d[i], h)
        assert np.abs(roots - t[i]).min() < 1e-7 * max(1.0, abs(t[i]))
        front = roots[roots > 1e-9]
        if o[i, 1] - h * (o[i, 0] ** 2 + o[i, 2] ** 2) ** 2 > 0 and len(front) == 1:
            assert abs(front[0] - t[i]) < 1e-7 * max(1.0, t[i])
            checked += 1
    assert checked > 100
    # misses: rays pointing away from the bounding box never hit
    away = np.concatenate([o, -d], axis=1)[np.abs(o).max(1) > 1.2 + h]
    assert (sc.closest_hit(away)[1] == -1).all()


def test_monomial_sample_is_on_the_rim_with_constant_pdf(orc):
    """MonomialSurface::sample draws from UnitCircle -- the rim x^2 + z^2 = 1 only -- flips the normal
    with probability 1/2 and returns pdf 1 / (2 * 6.3406654362) (monomial_surface.rs:107-122)."""
    scene = api.Scene()
    scene.add(api.Object(api.plane(api.vec3(0.0, 1.0, 0.0), -1.0)))
    scene.add(api.Light.Object(api.Object(api.monomial_surface(2.0, 4.0)).material(api.Material.light(api.vec3(1, 1, 1), 1.0))))
    flat = api.FlatScene(scene)
    sc = orc.OracleScene(flat)
    pos = np.tile(np.array([[0.0, 5.0, 0.0]]), (4000, 1))
    inten, wi, dist = sc.illuminate(0, pos, seed=3)
    v = pos + wi * dist[:, None]
    np.testing.assert_allclose(v[:, 0] ** 2 + v[:, 2] ** 2, 1.0, atol=1e-9)
    np.testing.assert_allclose(v[:, 1], 2.0, atol=1e-9)
    lit = inten[:, 0] > 0
    assert 0.4 < lit.mean() < 0.6                       # the coin flip on the normal
    # intensity = emittance * color * cos / len^2 / pdf, with the rim normal (8x, -1, 8z)/sqrt(65) or its negation
    n = util.normalize(np.stack([8 * v[:, 0], -np.ones(len(v)), 8 * v[:, 2]], axis=1))
    disp = v - pos
    cos = np.abs((disp * n).sum(1)) / dist
    np.testing.assert_allclose(inten[lit, 0], (cos / dist ** 2 * 2 * 6.3406654362)[lit], rtol=1e-9)


def test_monomial_glass_scene_renders_finite(orc):
    cfg = scenes.monomial_glass_scene(128, 64)
    flat = api.FlatScene(cfg.scene)
    sc = orc.OracleScene(flat)
    r = api.Renderer(cfg.scene, cfg.camera).width(80).height(60).max_bounces(1)
    img, st = sc.render(cfg.camera, r.params(8))
    assert np.isfinite(img).all() and img.min() >= 0.0 and img.mean() > 0.05
    assert st["env_lookups"] > 0 and st["segments"] > 80 * 60 * 8
