"""CPU check of the Rng<float> buffering variants (rng.cuh, RPTB_FIFO_MODE): host-emulation builds of each mode under
build/emu_modes/ must give the same f32 images bit for bit -- they buffer the same stream differently.
    python tools/emu_fifo_modes.py <mode>      (0 = the default build of tests/hostemu)"""
import sys, os, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mode = sys.argv[1]
from hostemu import emu
if mode != "0":
    emu.LIB_PATH = os.path.join(ROOT, "build", "emu_modes", "libhostemu_m%s.so" % mode)
    emu.subprocess.check_call = lambda *a, **k: 0
from rpt_b200 import api, _capi as capi
import util
out = {}
for name in ("cornell", "glass", "sphere"):
    cfg = util.golden_config(name)
    e = emu.EmuScene(api.FlatScene(cfg.scene))
    r = api.Renderer(cfg.scene, cfg.camera).width(32).height(24).max_bounces(cfg.max_bounces).seed(5)
    g, st, _ = e.render(cfg.camera, r.precision(capi.PRECISION_F32).params(6))
    out[name] = [float(np.nansum(g)), int(st["segments"]), int(st["rays"])]
print(mode, json.dumps(out))
