"""Full-resolution finite-ness check of every config (f32), a few spp."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api
for name, spp in [("sphere", 100), ("cornell", 16), ("teapot", 16), ("dragon", 8), ("glass", 32)]:
    cfg = scenes.CONFIGS[name]()
    r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
    buf = api.Buffer(cfg.width, cfg.height)
    r.sample(spp, buf)
    img = buf.batches[0]
    bad = ~np.isfinite(img).all(axis=1)
    print(name, "non-finite pixels:", int(bad.sum()), "mean", float(np.nanmean(img)), "max", float(np.nanmax(img)), "Msamples/s %.1f" % (r.last_stats["segments"] / r.last_stats["gpu_ms"] / 1e3), flush=True)
    r.close()
