"""Diagnose the vertex-at-once engine against the slot engine on the GPU: python tools/gpu_vxdiag.py (spawns itself per engine)."""
import json, os, subprocess, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

def child():
    from rpt_b200 import api, scenes, _capi as capi
    import ctypes as C
    out = {}
    for case in ("full", "mb0", "mb1", "one_light", "no_lights", "kd"):
        cfg = scenes.dragon_scene(level=0)
        mb = {"mb0": 0, "mb1": 1}.get(case, 2)
        if case == "one_light":
            cfg.scene.lights = cfg.scene.lights[:2]
        if case == "no_lights":
            cfg.scene.lights = cfg.scene.lights[:1]
        accel = capi.ACCEL_KDTREE if case == "kd" else capi.ACCEL_BVH
        ds = api.DeviceScene(api.FlatScene(cfg.scene, accel=accel))
        r = api.Renderer(cfg.scene, cfg.camera).width(96).height(54).max_bounces(mb).seed(5)
        p = r.params(8)
        cam = cfg.camera.to_c()
        img = np.empty((96 * 54, 3)); st = capi.Stats()
        capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), img.ctypes.data_as(capi.c_double_p), C.byref(st)), "render")
        out[case] = {"img": img.tolist(), "segments": int(st.segments), "rays": int(st.rays)}
        ds.close()
    json.dump(out, open(sys.argv[2], "w"))

if len(sys.argv) > 1 and sys.argv[1] == "child":
    child()
else:
    res = {}
    for vx in ("0", "1"):
        f = "/tmp/vxdiag_%s.json" % vx
        subprocess.check_call([sys.executable, __file__, "child", f], env=dict(os.environ, RPTB_VX=vx))
        res[vx] = json.load(open(f))
    for case in res["0"]:
        a, b = np.array(res["0"][case]["img"]), np.array(res["1"][case]["img"])
        rel = np.abs(a - b).max(axis=1) / np.maximum(np.abs(a).max(axis=1), 1e-4)
        bad = np.nonzero(rel > 1e-4)[0]
        print(case, "segments", res["0"][case]["segments"], res["1"][case]["segments"], "rays", res["0"][case]["rays"], res["1"][case]["rays"],
              "pixels off", len(bad), "of", len(rel), "first", [(int(i % 96), int(i // 96)) for i in bad[:12]])
