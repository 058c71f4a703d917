"""The device geometry code checked against the oracle WITHOUT a GPU.

tests/hostemu compiles the functions the CUDA kernels inline (closest_hit, kd_intersect, group_intersect,
monomial_intersect, finalize_hit, shape_sample, illuminate -- rpt_b200/csrc/geometry.cuh, shading.cuh) for the
host and runs them over the arrays rptb_scene_create would upload.  With Real = double every operation is
the oracle's, so hits, normals and traversal counters must agree exactly; with Real = float within the
f32 tolerances the GPU parity tests use.  This is test infrastructure (it cannot say anything about the
kernels' scheduling or fast-math); the `-m gpu` tests through the C ABI remain the gate."""
import numpy as np
import pytest

from rpt_b200 import api, scenes
from rpt_b200 import _capi as capi
from tests import util
from tests.hostemu import emu
from tests.test_oracle_instancing import rays_into, small_scene

F_TREE, F_TRANSP, F_HDRI, F_SMALL, F_GROUP, F_MONO = 1, 2, 4, 8, 16, 32


def scene_rays(name):
    rng = np.random.default_rng(11)
    if name == "mixed":
        scene, kids = small_scene(monomials=True)
        scene.add(api.Light.Object(api.Object(api.monomial_surface(2.0, 4.0).scale(api.vec3(1.5, 1.0, 0.8))
                                              .translate(api.vec3(0.0, 7.0, 0.0))).material(api.Material.light(api.vec3(1, 0.9, 0.8), 30.0))))
        scene.add(api.Light.Object(api.Object(api.KdTree(kids[:7]).rotate_y(0.4).translate(api.vec3(0.0, 9.0, 0.0)))
                                   .material(api.Material.light(api.vec3(1, 1, 1), 10.0))))
        rays = np.concatenate([rays_into(rng, 30000), util.interior_rays([-4, -4, -4], [4, 7, 4], 15000, rng)])
        return scene, rays, F_TREE | F_GROUP | F_MONO
    cfg, want = {
        "cornell": (scenes.cornell_scene, 0),
        "teapot": (scenes.teapot_scene, F_TREE | F_SMALL),
        "glass": (lambda: scenes.glass_scene(64, 32), F_TRANSP | F_HDRI | F_SMALL),
        "fractal_spheres": (lambda: scenes.fractal_spheres_scene(4), F_GROUP),
        "fractal_teapots": (lambda: scenes.fractal_teapots_scene(3), F_GROUP | F_TREE),
        "monomial_glass": (lambda: scenes.monomial_glass_scene(64, 32), F_MONO | F_HDRI),
    }[name]
    cfg = cfg()
    rays = util.camera_rays(cfg.camera, 20000, rng, spread=0.4)
    return cfg.scene, rays, want


NAMES = ["cornell", "teapot", "glass", "fractal_spheres", "fractal_teapots", "monomial_glass", "mixed"]


@pytest.mark.parametrize("name", NAMES)
def test_device_closest_hit_f64_is_the_oracles(orc, name):
    scene, rays, want_features = scene_rays(name)
    flat = api.FlatScene(scene)
    e, o = emu.EmuScene(flat), orc.OracleScene(flat)
    assert e.features == want_features          # the kernel variant rptb_render_samples would pick
    t0, o0, n0, s0 = o.closest_hit(rays)
    t1, o1, n1, s1 = e.closest_hit(rays, precision=capi.PRECISION_F64)
    np.testing.assert_array_equal(o1, o0)
    np.testing.assert_array_equal(t1, t0)
    np.testing.assert_array_equal(n1, n0)
    assert s1["node_visits"] == s0["node_visits"] and s1["tri_tests"] == s0["tri_tests"]   # the same leaves, the same trees
    assert (o0 >= 0).mean() > 0.3


@pytest.mark.parametrize("name", NAMES)
def test_device_closest_hit_f32_within_tolerance(orc, name):
    scene, rays, _ = scene_rays(name)
    flat = api.FlatScene(scene)
    e, o = emu.EmuScene(flat), orc.OracleScene(flat)
    t0, o0, n0, _ = o.closest_hit(rays)
    t1, o1, n1, _ = e.closest_hit(rays, precision=capi.PRECISION_F32)
    same = o1 == o0
    assert same.mean() > 0.999
    hit = same & (o0 >= 0)
    rel = np.abs(t1[hit] - t0[hit]) / np.abs(t0[hit])
    assert np.median(rel) <= 2e-7 and np.quantile(rel, 0.99) <= 1e-5 and np.quantile(rel, 0.999) <= 1e-3
    assert np.quantile(np.abs(n1[hit] - n0[hit]).max(1), 0.99) < 1e-3


def test_device_light_sampling_of_monomial_and_group_lights(orc):
    """Light::illuminate over MonomialSurface::sample and KdTree::sample (uniform child, pdf / num), with the
    oracle's random stream: identical draws -> identical samples in f64."""
    scene, _, _ = scene_rays("mixed")
    flat = api.FlatScene(scene)
    e, o = emu.EmuScene(flat), orc.OracleScene(flat)
    pos = np.random.default_rng(2).uniform(-3, 3, (3000, 3))
    for light in (1, 2):
        a = o.illuminate(light, pos, seed=5)
        b = e.illuminate(light, pos, seed=5, precision=capi.PRECISION_F64)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(y, x)
        assert np.isfinite(a[0]).all() and 0.3 < (a[0][:, 0] > 0).mean() < 0.9
        c = e.illuminate(light, pos, seed=5, precision=capi.PRECISION_F32)
        # f32 decides rejection loops / child picks on 32-bit words: almost always the same sample
        same = np.abs(c[2] - a[2]) < 1e-3 * a[2]
        assert same.mean() > 0.995
        scale = np.abs(a[0][same]).max(1) + 1e-30
        assert np.quantile(np.abs(c[0][same] - a[0][same]).max(1) / scale, 0.99) < 1e-3


def test_flattener_rejects_what_the_device_cannot_hold():
    tea = api.Mesh(scenes.teapot_triangles(), build=False)
    with pytest.raises(TypeError):
        api.KdTree([api.plane(api.vec3(0, 1, 0), 0.0)])
    with pytest.raises(TypeError):
        api.KdTree([api.KdTree([api.sphere()])])
    scene = api.Scene()
    scene.add(api.Object(api.KdTree([api.sphere(), tea.translate(api.vec3(1, 0, 0))])))
    flat = api.FlatScene(scene)
    flat.groups[0].children[0].kind = capi.SHAPE_PLANE
    with pytest.raises(ValueError, match="not Bounded"):
        emu.EmuScene(flat)
    flat.groups[0].children[0].kind = capi.SHAPE_GROUP
    with pytest.raises(ValueError, match="not supported"):
        emu.EmuScene(flat)
    flat.groups[0].children[0].kind = capi.SHAPE_SPHERE
    flat.groups[0].children[1].mesh = 7
    with pytest.raises(ValueError, match="mesh 7 out of range"):
        emu.EmuScene(flat)
    flat.groups[0].children[1].mesh = 0
    flat.objects[0].mesh = 3
    with pytest.raises(ValueError, match="group 3 out of range"):
        emu.EmuScene(flat)
    # the product boundary reports the same as a status + message, before it even looks for a device
    import ctypes as C
    h = C.c_void_p()
    rc = capi.lib().rptb_scene_create(C.byref(flat.desc), 0, C.byref(h))
    assert rc == -1 and b"group 3 out of range" in capi.lib().rptb_last_error() and not h


# ---------------------------------------------------------------- the f32 path's own BVH ----------------
def _bvh_pair(scene):
    return (emu.EmuScene(api.FlatScene(scene, accel=capi.ACCEL_BVH)), emu.EmuScene(api.FlatScene(scene, accel=capi.ACCEL_KDTREE)))


@pytest.mark.parametrize("name", ["teapot", "dragon_small"])
def test_bvh_finds_the_hits_of_the_reference_tree(orc, name):
    """rptb_accel BVH (bvhbuild.cpp + bvh_intersect): a different structure over the same triangles and the same
    triangle test -> the f32 hits are those of the kd-tree traversal, to the bit (ties on shared edges aside),
    for a fraction of the node visits and triangle tests."""
    cfg = scenes.teapot_scene() if name == "teapot" else scenes.dragon_scene(330, 82)
    eb, ek = _bvh_pair(cfg.scene)
    assert (eb.features & 64) and not (ek.features & 64)
    chk = eb.bvh_check(0)
    ntris = len(cfg.scene.objects[0].shape.shape.triangles)
    assert chk["violations"] == 0 and chk["distinct"] == ntris and chk["max_leaf"] <= 4 and chk["depth"] < 94
    assert ek.bvh_check(0) is None
    rng = np.random.default_rng(5)
    rays = np.concatenate([util.camera_rays(cfg.camera, 30000, rng, spread=0.4), util.interior_rays([-2, -0.99, -2], [2, 1.5, 2], 15000, rng)])
    tb, ob, nb, sb = eb.closest_hit(rays, precision=capi.PRECISION_F32)
    tk, ok, nk, sk = ek.closest_hit(rays, precision=capi.PRECISION_F32)
    same = (tb == tk) & (ob == ok)
    assert same.mean() >= 0.9999
    assert np.abs(nb[same] - nk[same]).max() <= 1e-6
    assert sb["node_visits"] < 0.5 * sk["node_visits"] and sb["tri_tests"] < 0.1 * sk["tri_tests"]
    # and against the oracle, the usual f32 tolerances
    t0, o0, n0, _ = orc.OracleScene(api.FlatScene(cfg.scene)).closest_hit(rays)
    agree = ob == o0
    assert agree.mean() > 0.9999
    hit = agree & (o0 >= 0)
    rel = np.abs(tb[hit] - t0[hit]) / np.abs(t0[hit])
    assert np.median(rel) <= 2e-7 and np.quantile(rel, 0.99) <= 1e-5 and np.quantile(rel, 0.999) <= 1e-4
    # f64 never uses it
    tb64, ob64, _, s64 = eb.closest_hit(rays, precision=capi.PRECISION_F64)
    np.testing.assert_array_equal(tb64, t0)
    assert s64["tri_tests"] == sk["tri_tests"] or s64["tri_tests"] > sb["tri_tests"]


def test_bvh_builder_on_awkward_meshes():
    rng = np.random.default_rng(8)

    def mesh_scene(tris):
        scene = api.Scene()
        scene.add(api.Object(api.Mesh(np.asarray(tris))))
        return scene

    def check(tris, nrays=4000):
        scene = mesh_scene(tris)
        eb, ek = _bvh_pair(scene)
        if not (ek.features & 1):          # the reference tree is one leaf: no BVH is built, the leaf is scanned
            assert eb.bvh_check(0) is None and not (eb.features & 64)
            return None
        chk = eb.bvh_check(0)
        assert chk["violations"] == 0 and chk["distinct"] == len(tris) and chk["max_leaf"] <= 4 and chk["depth"] < 94
        lo, hi = np.asarray(tris)[:, :9].reshape(-1, 3).min(0), np.asarray(tris)[:, :9].reshape(-1, 3).max(0)
        c, r = (lo + hi) / 2, np.linalg.norm(hi - lo) + 1e-3
        o = c + util.random_unit(rng, nrays) * r
        tgt = lo + rng.uniform(0, 1, (nrays, 3)) * (hi - lo)
        rays = np.concatenate([o, util.normalize(tgt - o)], axis=1)
        tb, ob, _, _ = eb.closest_hit(rays, precision=capi.PRECISION_F32)
        tk, ok, _, _ = ek.closest_hit(rays, precision=capi.PRECISION_F32)
        assert ((tb == tk) & (ob == ok)).mean() >= 0.999
        return chk

    # 64 copies of one triangle: the kd builder gives up (one leaf) -> no BVH
    assert check(np.tile(api.Triangle.from_vertices([0, 0, 0], [1, 0, 0], [0, 1, 0]), (64, 1))) is None
    # a 40 x 40 grid of quads (3 200 triangles, many coplanar, zero-thickness boxes)
    g = []
    for i in range(40):
        for j in range(40):
            a, b, c_, d = [i, 0, j], [i + 1, 0, j], [i + 1, 0, j + 1], [i, 0, j + 1]
            g += [api.Triangle.from_vertices(a, b, c_), api.Triangle.from_vertices(a, c_, d)]
    assert check(np.stack(g))["nodes"] > 500
    # sizes spanning six decades along one axis (SAH peels them off one by one -> a deep, thin tree)
    s = []
    for k in range(60):
        x = 1.5 ** k * 1e-3
        s.append(api.Triangle.from_vertices([x, 0, 0], [x * 1.4, 0, 0], [x, x * 0.4, x * 0.1]))
    chk = check(np.stack(s))
    assert chk is None or chk["depth"] <= 60
    # random soup
    c = rng.uniform(-1, 1, (5000, 1, 3))
    v = c + rng.normal(0, 0.03, (5000, 3, 3))
    assert check(np.stack([api.Triangle.from_vertices(*t) for t in v]))["nodes"] > 1000


def test_accel_field_is_validated():
    flat = api.FlatScene(scenes.teapot_scene().scene)
    flat.desc.accel = 7
    with pytest.raises(ValueError, match="bad accel"):
        emu.EmuScene(flat)


def test_bvh_under_a_kd_tree_of_shapes(orc):
    """rptb_closest_hit without counters on a scene that has both a kd-tree of shapes and BVH meshes runs
    closest_hit<F_EVERY | F_BVH>: the instances of examples/fractal_teapots are entered through the group's
    kd-tree and then traversed through the shared teapot's BVH.  Same hits as through its kd-tree."""
    cfg = scenes.fractal_teapots_scene(3)
    eb, ek = _bvh_pair(cfg.scene)
    assert (eb.features & (16 | 64)) == (16 | 64) and eb.bvh_check(0)["violations"] == 0
    rng = np.random.default_rng(12)
    rays = np.concatenate([util.camera_rays(cfg.camera, 30000, rng, spread=0.35), util.interior_rays([-2.5] * 3, [2.5] * 3, 15000, rng)])
    tb, ob, nb, sb = eb.closest_hit(rays, precision=capi.PRECISION_F32)
    tk, ok, nk, sk = ek.closest_hit(rays, precision=capi.PRECISION_F32)
    same = (tb == tk) & (ob == ok)
    assert same.mean() >= 0.9999 and np.abs(nb[same] - nk[same]).max() <= 1e-6
    assert sb["tri_tests"] < 0.2 * sk["tri_tests"]
    t0, o0, _, _ = orc.OracleScene(api.FlatScene(cfg.scene)).closest_hit(rays)
    assert (ob == o0).mean() > 0.9995
