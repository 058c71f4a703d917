"""GPU edge cases of the hot path (-m gpu): scenes and parameters at the corners of what the
reference accepts, each checked against the CPU oracle on the same Philox streams."""
import ctypes as C
import math

import numpy as np
import pytest

from rpt_b200 import _capi as capi
from rpt_b200 import api
from tests import util

pytestmark = pytest.mark.gpu
F32, F64 = capi.PRECISION_F32, capi.PRECISION_F64


def _both(orc, scene, camera, w, h, spp, mb, seed=1, precision=F64, ev=0.0, engine=capi.ENGINE_AUTO):
    flat = api.FlatScene(scene)
    r = api.Renderer(scene, camera).width(w).height(h).max_bounces(mb).seed(seed).precision(precision) \
        .exposure_value(ev).engine(engine)
    ref, st0 = orc.OracleScene(flat).render(camera, r.params(spp))
    buf = api.Buffer(w, h)
    r.sample(spp, buf)
    got, st1 = buf.batches[0], r.last_stats
    r.close()
    return ref, got, st0, st1


def _close(ref, got, frac=0.98, rtol=1e-9):
    rel = np.abs(got - ref) / np.maximum(np.abs(ref), 1e-9)
    return (rel.max(axis=1) < rtol).mean() >= frac


def test_empty_scene_is_the_environment(orc, gpu_ok):
    scene = api.Scene()
    scene.environment = api.Environment.Color(api.vec3(0.1, 0.2, 0.3))
    for prec in (F64, F32):
        ref, got, st0, st1 = _both(orc, scene, api.Camera.default(), 33, 17, 3, 4, precision=prec, ev=-1.0)
        np.testing.assert_allclose(got, np.tile([0.05, 0.1, 0.15], (33 * 17, 1)), rtol=1e-6)
        assert st1["segments"] == st0["segments"] == 33 * 17 * 3


def test_one_pixel_image_and_many_bounces(orc, gpu_ok):
    """1x1 image (max(width,height) = 1 in the pixel mapping), max_bounces 40 (> 16: the second
    kernel instantiation) inside a closed emissive sphere."""
    scene = api.Scene()
    scene.add(api.Object(api.sphere().scale(api.vec3(5.0, 5.0, 5.0))).material(api.Material.light(api.vec3(0.5, 0.6, 0.7), 0.2)))
    cam = api.Camera(eye=api.vec3(0, 0, 0), direction=api.vec3(0, 0, -1), up=api.vec3(0, 1, 0), fov=1.0)
    ref, got, st0, st1 = _both(orc, scene, cam, 1, 1, 16, 40)
    # seen from inside, the sphere's outward normal faces away: bsdf == 0, yet the reference keeps
    # bouncing with weight 0 (sample_f still returns directions) -- the f64 gate does the same
    assert st0["segments"] == st1["segments"] > 16
    np.testing.assert_allclose(got, ref, rtol=1e-9)
    # the f32 path recognises the dead vertex and stops at once: same radiance, 16 segments
    ref, got32, st0, st32 = _both(orc, scene, cam, 1, 1, 16, 40, precision=F32)
    np.testing.assert_allclose(got32, ref, rtol=1e-5)
    assert st32["segments"] == 16


def test_directional_point_ambient_lights_and_transformed_plane(orc, gpu_ok):
    scene = api.Scene()
    # a plane given through Transformed<Plane>: tilted 20 degrees about z, lifted
    scene.add(api.Object(api.plane(api.vec3(0, 1, 0), 0.0).rotate_z(math.radians(20)).translate(api.vec3(0, -1.0, 0)))
              .material(api.Material.specular(api.hex_color(0x88AACC), 0.3)))
    scene.add(api.Object(api.cube().scale(api.vec3(1.0, 2.0, 0.5)).rotate_y(0.7).translate(api.vec3(0.5, 0.2, 0)))
              .material(api.Material.metallic_(api.hex_color(0xD4AF37), 0.25)))
    scene.add(api.Light.Ambient(api.vec3(0.03, 0.03, 0.04)))
    scene.add(api.Light.Directional(api.vec3(1.5, 1.4, 1.2), api.vec3(-0.3, -1.0, -0.2)))
    scene.add(api.Light.Point(api.vec3(20, 10, 10), api.vec3(-3, 4, 2)))
    scene.add(api.Light.Ambient(api.vec3(0.0, 0.01, 0.0)))  # a trailing ambient light: order of additions
    cam = api.Camera.look_at(api.vec3(0, 2, 8), api.vec3(0, 0, 0), api.vec3(0, 1, 0), 0.6)
    ref, got, st0, st1 = _both(orc, scene, cam, 64, 40, 8, 3)
    assert _close(ref, got)
    assert st0["rays"] == st1["rays"]
    ref, got32, _, _ = _both(orc, scene, cam, 64, 40, 8, 3, precision=F32)
    assert util.rmse(np.clip(got32, 0, 1), np.clip(ref, 0, 1)) < 2e-3
    # and through the wavefront schedule
    _, got_wf, _, _ = _both(orc, scene, cam, 64, 40, 8, 3, precision=F32, engine=capi.ENGINE_WAVEFRONT)
    rel = np.abs(got_wf - got32) / np.maximum(np.abs(got32), 1e-4)
    assert np.quantile(rel.max(axis=1), 0.99) < 1e-4


def test_depth_of_field_camera(orc, gpu_ok):
    """aperture > 0 draws a UnitDisc sample per camera ray (src/camera.rs:70-76)."""
    scene = api.Scene()
    scene.add(api.Object(api.sphere().translate(api.vec3(0, 0, 0))).material(api.Material.diffuse(api.hex_color(0xCC4444))))
    scene.add(api.Object(api.sphere().scale(api.vec3(0.5, 0.5, 0.5)).translate(api.vec3(1.5, 0, -3)))
              .material(api.Material.diffuse(api.hex_color(0x44CC44))))
    scene.add(api.Light.Point(api.vec3(50, 50, 50), api.vec3(3, 5, 5)))
    scene.environment = api.Environment.Color(api.vec3(0.2, 0.2, 0.3))
    cam = api.Camera.look_at(api.vec3(0, 1, 6), api.vec3(0, 0, 0), api.vec3(0, 1, 0), 0.5).focus(api.vec3(0, 0, 0), 0.3)
    ref, got, st0, st1 = _both(orc, scene, cam, 48, 32, 16, 1)
    assert _close(ref, got)
    assert st0["segments"] == st1["segments"]


def test_cube_and_mesh_object_lights(orc, gpu_ok):
    """Light::Object over a cube (face pick + 2 uniforms) and over a multi-triangle mesh
    (uniform triangle pick + rejection), both transformed (pdf area rescale, src/shape.rs:139-150)."""
    scene = api.Scene()
    scene.add(api.Object(api.plane(api.vec3(0, 1, 0), -1.0)).material(api.Material.diffuse(api.hex_color(0xAAAAAA))))
    scene.add(api.Object(api.sphere()).material(api.Material.specular(api.hex_color(0x3366CC), 0.2)))
    scene.add(api.Light.Object(api.Object(api.cube().scale(api.vec3(1.0, 0.2, 2.0)).rotate_z(0.4).translate(api.vec3(-3, 3, 0)))
                               .material(api.Material.light(api.vec3(1, 0.9, 0.8), 15.0))))
    fan = api.polygon([api.vec3(0, 0, 0), api.vec3(1, 0, 0), api.vec3(1.5, 0, 1), api.vec3(0.5, 0, 1.8), api.vec3(-0.5, 0, 1)])
    scene.add(api.Light.Object(api.Object(fan.rotate_x(math.pi).scale(api.vec3(1.5, 1.0, 1.5)).translate(api.vec3(2, 4, -1)))
                               .material(api.Material.light(api.vec3(0.8, 0.9, 1.0), 25.0))))
    cam = api.Camera.look_at(api.vec3(0, 2, 7), api.vec3(0, 0, 0), api.vec3(0, 1, 0), 0.7)
    ref, got, st0, st1 = _both(orc, scene, cam, 48, 32, 16, 2)
    assert _close(ref, got, frac=0.97)
    ref2, _, _, _ = _both(orc, scene, cam, 48, 32, 16, 2, seed=2)
    ref, got32, _, _ = _both(orc, scene, cam, 48, 32, 16, 2, precision=F32)
    noise = util.rmse(np.clip(ref, 0, 1), np.clip(ref2, 0, 1))
    assert util.rmse(np.clip(got32, 0, 1), np.clip(ref, 0, 1)) <= 0.25 * noise


def test_transparent_tinted_glass_with_lights(orc, gpu_ok):
    """Transmissive material lit by a sampled light: the signed cosine of the direct term
    (src/renderer.rs:199) makes back-side contributions negative -- the quirk is kept."""
    scene = api.Scene()
    scene.add(api.Object(api.sphere()).material(api.Material.transparent_(api.vec3(0.9, 0.6, 0.4), 1.4, 0.3)))
    scene.add(api.Object(api.plane(api.vec3(0, 1, 0), -1.0)).material(api.Material.diffuse(api.hex_color(0xBBBBBB))))
    scene.add(api.Light.Object(api.Object(api.sphere().scale(api.vec3(0.7, 0.7, 0.7)).translate(api.vec3(2, 3, 2)))
                               .material(api.Material.light(api.vec3(1, 1, 1), 30.0))))
    cam = api.Camera.look_at(api.vec3(0, 1.5, 6), api.vec3(0, 0, 0), api.vec3(0, 1, 0), 0.6)
    ref, got, st0, st1 = _both(orc, scene, cam, 48, 32, 16, 5)
    assert _close(ref, got, frac=0.97)
    assert (ref < 0).any() == (got < 0).any()
    ref2, _, _, _ = _both(orc, scene, cam, 48, 32, 16, 5, seed=2)
    ref, got32, _, _ = _both(orc, scene, cam, 48, 32, 16, 5, precision=F32)
    noise = util.rmse(ref, ref2)
    assert util.rmse(got32, ref) <= 0.3 * noise


def test_many_objects_uniform_loop(orc, gpu_ok):
    """40 spheres + 12 cubes: the linear object scan of get_closest_hit at a non-toy count."""
    rng = np.random.default_rng(4)
    scene = api.Scene()
    scene.add(api.Object(api.plane(api.vec3(0, 1, 0), -1.0)).material(api.Material.diffuse(api.hex_color(0x999999))))
    for i in range(40):
        c = rng.uniform(-4, 4, 3) * np.array([1, 0.3, 1])
        s = rng.uniform(0.2, 0.6)
        scene.add(api.Object(api.sphere().scale(api.vec3(s, s, s)).translate(c))
                  .material(api.Material.specular(rng.uniform(0.2, 0.9, 3), float(rng.uniform(0.1, 0.9)))))
    for i in range(12):
        c = rng.uniform(-4, 4, 3) * np.array([1, 0.2, 1])
        scene.add(api.Object(api.cube().scale(rng.uniform(0.3, 0.8, 3)).rotate_y(float(rng.uniform(0, 3))).translate(c))
                  .material(api.Material.metallic_(rng.uniform(0.3, 0.9, 3), 0.3)))
    scene.add(api.Light.Object(api.Object(api.sphere().scale(api.vec3(2, 2, 2)).translate(api.vec3(0, 10, 0)))
                               .material(api.Material.light(api.vec3(1, 1, 1), 30.0))))
    cam = api.Camera.look_at(api.vec3(0, 4, 10), api.vec3(0, 0, 0), api.vec3(0, 1, 0), 0.7)
    ref, got, st0, st1 = _both(orc, scene, cam, 64, 36, 8, 2)
    assert _close(ref, got, frac=0.97)
    assert abs(st0["rays"] - st1["rays"]) <= 1e-3 * st0["rays"]
    rays = util.camera_rays(cam, 20000, rng, spread=0.5)
    flat = api.FlatScene(scene)
    t0, o0, n0, _ = orc.OracleScene(flat).closest_hit(rays)
    with api.DeviceScene(flat) as ds:
        t1, o1, n1 = ds.closest_hit(rays, precision=F64)
    assert (o0 == o1).all()
    np.testing.assert_array_equal(t0, t1)


def test_hdri_lookup_and_mirror_ball(orc, gpu_ok):
    """Environment::Hdri bilinear lookup incl. the clamped last row/column."""
    rng = np.random.default_rng(6)
    hd = api.Hdri(16, 8, rng.uniform(0, 2, (16 * 8, 3)))
    scene = api.Scene()
    scene.environment = api.Environment.Hdri(hd)
    scene.add(api.Object(api.sphere()).material(api.Material.metallic_(api.vec3(1, 1, 1), 0.05)))
    cam = api.Camera.default()
    ref, got, st0, st1 = _both(orc, scene, cam, 40, 30, 8, 3)
    assert _close(ref, got)
    assert st1["env_lookups"] == st0["env_lookups"] > 0
    ref, got32, _, _ = _both(orc, scene, cam, 40, 30, 8, 3, precision=F32)
    rel = np.abs(got32 - ref) / np.maximum(np.abs(ref), 1e-3)
    assert np.quantile(rel.max(axis=1), 0.95) < 1e-2
