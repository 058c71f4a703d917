"""Quick megakernel throughput of named configs: python tools/gpu_quick.py teapot glass ..."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi
spps = {"sphere": 100, "cornell": 64, "teapot": 64, "dragon": 8, "glass": 64}
for name in sys.argv[1:]:
    cfg = scenes.CONFIGS[name]()
    r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1).engine(capi.ENGINE_MEGAKERNEL)
    buf = api.Buffer(cfg.width, cfg.height); r.sample(2, buf); r._next_sample = 0
    best = 0
    for _ in range(2):
        buf = api.Buffer(cfg.width, cfg.height); r._next_sample = 0; r.sample(spps[name], buf)
        st = r.last_stats; best = max(best, st["segments"] / st["gpu_ms"] / 1e3)
    print(name, "Msamples/s %.1f" % best, flush=True)
    r.close()
