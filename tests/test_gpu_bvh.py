"""GPU tests (-m gpu) of rptb_accel = BVH: the f32 product path traversing meshes through the library's own
SAH BVH instead of the reference-shaped kd-tree.  The closest hit of a ray does not depend on the structure
(same triangles, same triangle test), so the BVH scene must reproduce the kd-tree scene's hits and images --
bit for bit except where two triangles tie in t on a shared edge -- and, like it, agree with the oracle within
the f32 tolerances of tests/test_gpu_parity.py."""
import ctypes as C

import numpy as np
import pytest

from rpt_b200 import _capi as capi
from rpt_b200 import api, scenes
from tests import util
from tests.test_gpu_parity import F32, F64, _gpu_render, _oracle_render, _rays_for

pytestmark = pytest.mark.gpu
SMALL = {"teapot": (96, 54, 16, 0), "dragon": (96, 54, 8, 2)}


@pytest.fixture(scope="module")
def pairs(gpu_ok):
    cache = {}

    def get(name):
        if name not in cache:
            cfg = scenes.teapot_scene() if name == "teapot" else scenes.dragon_scene(level=0)
            kd = api.DeviceScene(api.FlatScene(cfg.scene, accel=capi.ACCEL_KDTREE))
            bvh = api.DeviceScene(api.FlatScene(cfg.scene, accel=capi.ACCEL_BVH))
            cache[name] = (cfg, kd, bvh)
        return cache[name]

    yield get
    for _, kd, bvh in cache.values():
        kd.close()
        bvh.close()


def _render(cfg, ds, w, h, spp, mb, seed, engine=capi.ENGINE_AUTO, shard=(0, 1), stats=0, precision=F32):
    r = api.Renderer(cfg.scene, cfg.camera).width(w).height(h).max_bounces(mb).seed(seed).engine(engine).precision(precision)
    p = r.params(spp, 0, shard[0], shard[1], collect_stats=stats)
    cam = cfg.camera.to_c()
    out = np.empty((w * h, 3))
    st = capi.Stats()
    capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), out.ctypes.data_as(capi.c_double_p), C.byref(st)),
               "rptb_render_samples")
    return out, st.as_dict()


@pytest.mark.parametrize("name", ["teapot", "dragon"])
def test_bvh_closest_hits_are_the_kd_trees(orc, pairs, name):
    cfg, kd, bvh = pairs(name)
    rng = np.random.default_rng(13)
    rays = _rays_for(name, cfg, rng, 100000)
    tk, ok, nk = kd.closest_hit(rays, precision=F32)
    tb, ob, nb = bvh.closest_hit(rays, precision=F32)
    same = (tb == tk) & (ob == ok)
    assert same.mean() >= 0.9999, same.mean()
    assert np.abs(nb[same] - nk[same]).max() <= 1e-6
    # with counters each scene counts the structure it traverses: same answers as without, its own counters
    t1, o1, n1, s1 = kd.closest_hit(rays, precision=F32, want_stats=True)
    t2, o2, n2, s2 = bvh.closest_hit(rays, precision=F32, want_stats=True)
    np.testing.assert_array_equal(t1, tk)
    np.testing.assert_array_equal(t2, tb)
    assert s1["node_visits"] > 0 and s1["tri_tests"] > 0 and s1["bvh_node_visits"] == 0 and s1["bvh_tri_tests"] == 0
    assert s2["node_visits"] == 0 and s2["tri_tests"] == 0
    assert 0 < s2["bvh_node_visits"] < 0.5 * s1["node_visits"] and 0 < s2["bvh_tri_tests"] < 0.1 * s1["tri_tests"]
    # the f64 gate ignores the BVH
    t3, o3, _ = bvh.closest_hit(rays[:20000], precision=F64)
    t0, o0, _, _ = orc.OracleScene(bvh.flat).closest_hit(rays[:20000])
    np.testing.assert_array_equal(t3, t0)
    assert (o3 == o0).all()
    hit = (ob[:20000] == o0) & (o0 >= 0)
    rel = np.abs(tb[:20000][hit] - t0[hit]) / np.abs(t0[hit])
    assert (ob[:20000] == o0).mean() >= 0.9999 and np.median(rel) <= 2e-7 and np.quantile(rel, 0.999) <= 1e-4


@pytest.mark.parametrize("engine", [capi.ENGINE_MEGAKERNEL, capi.ENGINE_WAVEFRONT])
@pytest.mark.parametrize("name", ["teapot", "dragon"])
def test_bvh_images_match_the_kd_tree_images(pairs, name, engine):
    cfg, kd, bvh = pairs(name)
    w, h, spp, mb = SMALL[name]
    a, sa = _render(cfg, kd, w, h, spp, mb, 3, engine)
    b, sb = _render(cfg, bvh, w, h, spp, mb, 3, engine)
    assert sa["engine"] == sb["engine"] == engine
    assert sa["segments"] == sb["segments"] and sa["rays"] == sb["rays"]
    rel = np.abs(a - b).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-6)
    assert (rel <= 1e-6).mean() >= 0.999, (rel <= 1e-6).mean()      # same hits, same random streams -> same radiance
    assert np.isfinite(b).all()


def test_bvh_render_parity_with_the_oracle(orc, pairs):
    cfg, kd, bvh = pairs("teapot")
    w, h, spp, mb = 96, 54, 16, 2
    ref, st0 = _oracle_render(orc, cfg, bvh.flat, w, h, spp, mb, 1)
    ref2, _ = _oracle_render(orc, cfg, bvh.flat, w, h, spp, mb, 2)
    cl = lambda x: np.clip(x, 0.0, 1.0)
    noise = util.rmse(cl(ref), cl(ref2))
    for engine in (capi.ENGINE_MEGAKERNEL, capi.ENGINE_WAVEFRONT):
        g, st = _render(cfg, bvh, w, h, spp, mb, 1, engine)
        assert util.rmse(cl(g), cl(ref)) <= 0.25 * noise
        assert abs(cl(g).mean() - cl(ref).mean()) <= 5e-3 * cl(ref).mean()
        assert 0.85 * st0["segments"] <= st["segments"] <= 1.001 * st0["segments"]
    g64, _ = _render(cfg, bvh, w, h, spp, mb, 1, precision=F64)            # the gate: kd-tree, literal
    rel = np.abs(g64 - ref) / np.maximum(np.abs(ref), 1e-6)
    assert (rel.max(axis=1) < 1e-9).mean() >= 0.98


def test_bvh_shards_and_stats_pass(pairs):
    cfg, kd, bvh = pairs("dragon")
    w, h, spp, mb = SMALL["dragon"]
    for engine in (capi.ENGINE_MEGAKERNEL, capi.ENGINE_WAVEFRONT):
        full, _ = _render(cfg, bvh, w, h, spp, mb, 5, engine)
        parts = [_render(cfg, bvh, w, h, spp, mb, 5, engine, shard=(i, 2))[0] for i in range(2)]
        np.testing.assert_array_equal(parts[0] + parts[1], full)
    # the counting pass walks the reference-shaped trees (its counters are SURVEY 8d's algorithmic work):
    # same counters as the kd scene, image equal to the kd scene's counting pass
    a, sa = _render(cfg, kd, w, h, spp, mb, 5, capi.ENGINE_MEGAKERNEL, stats=2)
    b, sb = _render(cfg, bvh, w, h, spp, mb, 5, capi.ENGINE_MEGAKERNEL, stats=2)
    np.testing.assert_array_equal(a, b)
    assert sa["node_visits"] == sb["node_visits"] > 0 and sa["tri_tests"] == sb["tri_tests"] > 0
    assert sb["bvh_node_visits"] == 0
    # collect_stats = 1 counts what the product path traverses on THIS scene: the BVH (rptb_stats::bvh_*), and the
    # image is the BVH image (another instantiation of the same code: equal up to FMA contraction)
    c, sc = _render(cfg, bvh, w, h, spp, mb, 5, capi.ENGINE_MEGAKERNEL, stats=1)
    plain, sp = _render(cfg, bvh, w, h, spp, mb, 5, capi.ENGINE_MEGAKERNEL)
    assert sc["node_visits"] == 0 and sc["segments"] == sp["segments"] and sc["rays"] == sp["rays"]
    assert 0 < sc["bvh_node_visits"] < 0.5 * sb["node_visits"] and 0 < sc["bvh_tri_tests"] < 0.1 * sb["tri_tests"]
    rel = np.abs(c - plain) / np.maximum(np.abs(plain), 1e-4)
    assert np.quantile(rel.max(axis=1), 0.99) < 1e-4


def test_bvh_device_bytes_and_env_override(pairs, monkeypatch):
    cfg, kd, bvh = pairs("teapot")
    assert bvh.device_bytes() > kd.device_bytes()                      # the BVH arrays come on top of the kd arrays
    monkeypatch.setenv("RPTB_ACCEL", "bvh")
    with api.DeviceScene(api.FlatScene(cfg.scene)) as ds:              # accel AUTO + environment
        assert ds.device_bytes() == bvh.device_bytes()
    monkeypatch.setenv("RPTB_ACCEL", "kdtree")
    with api.DeviceScene(api.FlatScene(cfg.scene)) as ds:
        assert ds.device_bytes() == kd.device_bytes()


def test_bvh_axis_aligned_directional_light(gpu_ok):
    """A Directional light along an axis gives every shadow ray an exactly zero direction component -- the case
    slab_rcp (geometry.cuh) exists for: the BVH must cast the teapot's shadow exactly where the kd-tree does."""
    cfg = scenes.teapot_scene()
    scene = api.Scene()
    for o in cfg.scene.objects:
        scene.add(o)
    scene.add(api.Light.Directional(api.vec3(0.9, 0.9, 0.9), api.vec3(0.0, -1.0, 0.0)))
    scene.add(api.Light.Directional(api.vec3(0.2, 0.1, 0.1), api.vec3(1.0, 0.0, 0.0)))
    camera = api.Camera.look_at(api.vec3(0.0, 6.0, 0.001), api.vec3(0.0, -1.0, 0.0), api.vec3(0.0, 0.0, -1.0), 0.8)
    w, h, spp = 128, 96, 4
    imgs = {}
    for accel in (capi.ACCEL_KDTREE, capi.ACCEL_BVH):
        with api.DeviceScene(api.FlatScene(scene, accel=accel)) as ds:
            cfg2 = scenes.Config("teapot_dir", scene, camera, w, h, spp, 0)
            imgs[accel], st = _render(cfg2, ds, w, h, spp, 0, 2, capi.ENGINE_MEGAKERNEL)
            imgs[(accel, "wf")], _ = _render(cfg2, ds, w, h, spp, 0, 2, capi.ENGINE_WAVEFRONT)
    a, b = imgs[capi.ACCEL_KDTREE], imgs[capi.ACCEL_BVH]
    rel = np.abs(a - b).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-6)
    assert (rel <= 1e-5).mean() >= 0.999, (rel <= 1e-5).mean()
    lum = a.sum(axis=1)
    assert (lum < 0.25 * np.median(lum)).mean() > 0.02          # there is a shadow to get wrong
    rel = np.abs(imgs[(capi.ACCEL_KDTREE, "wf")] - imgs[(capi.ACCEL_BVH, "wf")]).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-6)
    assert (rel <= 1e-5).mean() >= 0.999
