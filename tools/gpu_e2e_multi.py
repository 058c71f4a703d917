"""Where the end-to-end time of rptb_render_samples on a multi-device handle goes: python tools/gpu_e2e_multi.py [ndev ...]"""
import ctypes as C, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import api, scenes, _capi as capi

cfg = scenes.cornell_scene()
flat = api.FlatScene(cfg.scene)
cam = cfg.camera.to_c()
out = np.empty((cfg.width * cfg.height, 3))
ndevs = [int(a) for a in sys.argv[1:]] or [1, 2]
for nd in ndevs:
    for spp in (64, 512):
        r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
        p = r.params(spp * nd)
        rows = []
        for rep in range(4):
            t0 = time.perf_counter()
            ds = api.DeviceScene(flat, list(range(nd)) if nd > 1 else 0)
            t1 = time.perf_counter()
            st = capi.Stats()
            capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), out.ctypes.data_as(capi.c_double_p), C.byref(st)), "render")
            t2 = time.perf_counter()
            capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), out.ctypes.data_as(capi.c_double_p), C.byref(st)), "render")
            t3 = time.perf_counter()
            ds.close()
            t4 = time.perf_counter()
            rows.append((1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), 1e3 * (t4 - t3), st.gpu_ms))
        print("ndev %d spp/gpu %d: create %.1f ms, first render %.1f ms, second render %.1f ms, destroy %.1f ms, gpu_ms %.1f (last of 4 reps; first rep: %s)"
              % ((nd, spp) + rows[-1] + (", ".join("%.1f" % v for v in rows[0]),)), flush=True)
