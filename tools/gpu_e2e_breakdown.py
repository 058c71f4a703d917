"""Where the end-to-end call spends its time beyond the kernel (Cornell, 1 GPU)."""
import sys, os, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 512
use_torch = len(sys.argv) > 2
if use_torch:
    import torch
    torch.cuda.set_device(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
cfg = scenes.cornell_scene()
flat = api.FlatScene(cfg.scene)
lib = capi.lib()
r = api.Renderer(cfg.scene, cfg.camera).width(800).height(800).max_bounces(6).seed(1)
cam = cfg.camera.to_c()
out = np.empty((640000, 3))
for it in range(4):
    t0 = time.perf_counter()
    ds = api.DeviceScene(flat, 0)
    t1 = time.perf_counter()
    p = r.params(spp)
    st = capi.Stats()
    lib.rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), out.ctypes.data_as(capi.c_double_p), C.byref(st))
    t2 = time.perf_counter()
    lib.rptb_render_samples(ds.handle, C.byref(cam), C.byref(p), out.ctypes.data_as(capi.c_double_p), None)
    t3 = time.perf_counter()
    ds.close()
    t4 = time.perf_counter()
    print("iter %d: create %.2f ms | render+stats %.2f ms (kernel %.2f) | render, stats=NULL %.2f ms | destroy %.2f ms" % (
        it, 1e3 * (t1 - t0), 1e3 * (t2 - t1), st.gpu_ms, 1e3 * (t3 - t2), 1e3 * (t4 - t3)), flush=True)
