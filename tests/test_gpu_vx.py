"""The optional vertex-at-once schedule of the f32 megakernel (RPTB_VX=1, integrator_vx.cuh) against the default slot
schedule on real warps: ragged image sizes (warps with fewer than 32 lanes at the image's edge), shadow + bounce rays
of several lanes compacted into one work list.  The environment variable is read once per process, so each engine
renders in its own interpreter."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import ctypes as C, json, sys
import numpy as np
sys.path.insert(0, %r)
from rpt_b200 import api, scenes, _capi as capi
out = {}
for name, (w, h, spp, mb) in {"dragon": (99, 54, 8, 2), "teapot": (101, 57, 8, 0), "cornell": (75, 61, 70, 6), "glass": (64, 36, 16, 12)}.items():
    cfg = scenes.dragon_scene(level=0) if name == "dragon" else scenes.glass_scene(128, 64) if name == "glass" else scenes.CONFIGS[name]()
    ds = api.DeviceScene(api.FlatScene(cfg.scene))
    r = api.Renderer(cfg.scene, cfg.camera).width(w).height(h).max_bounces(mb).seed(5)
    p = r.params(spp)
    cam = cfg.camera.to_c()
    img = np.empty((w * h, 3)); st = capi.Stats()
    capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), img.ctypes.data_as(capi.c_double_p), C.byref(st)), "render")
    out[name] = {"img": img.tolist(), "segments": int(st.segments), "rays": int(st.rays)}
    ds.close()
json.dump(out, open(sys.argv[1], "w"))
"""


def test_vertex_at_once_engine_matches_the_slot_engine(gpu_ok, tmp_path):
    res = {}
    for vx in ("0", "1"):
        f = tmp_path / ("vx%s.json" % vx)
        subprocess.check_call([sys.executable, "-c", CHILD % ROOT, str(f)], env=dict(os.environ, RPTB_VX=vx))
        res[vx] = json.load(open(f))
    for name in res["0"]:
        a, b = np.array(res["0"][name]["img"]), np.array(res["1"][name]["img"])
        assert res["0"][name]["segments"] == res["1"][name]["segments"], name
        assert res["0"][name]["rays"] == res["1"][name]["rays"], name
        rel = np.abs(a - b).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-4)
        assert (rel > 1e-5).mean() <= 1e-3, (name, float((rel > 1e-5).mean()))   # same per-path operations; two compilations of them
