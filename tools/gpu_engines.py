"""Megakernel vs wavefront throughput per config (1 GPU)."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi
which = sys.argv[1:] or ["sphere", "cornell", "teapot", "dragon", "glass"]
spps = {"sphere": 64, "cornell": 32, "teapot": 32, "dragon": 8, "glass": 32}
for name in which:
    cfg = scenes.CONFIGS[name]()
    imgs = {}
    for eng, label in ((capi.ENGINE_MEGAKERNEL, "megakernel"), (capi.ENGINE_WAVEFRONT, "wavefront")):
        r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1).engine(eng)
        buf = api.Buffer(cfg.width, cfg.height); r.sample(1, buf); r._next_sample = 0
        buf = api.Buffer(cfg.width, cfg.height); r.sample(spps[name], buf)
        st = r.last_stats; imgs[label] = buf.batches[0]
        print(name, label, "gpu_ms %.1f" % st["gpu_ms"], "Msamples/s %.1f" % (st["segments"] / st["gpu_ms"] / 1e3), "Mrays/s %.1f" % (st["rays"] / st["gpu_ms"] / 1e3), "launches", st["launches"], "mean %.6f" % imgs[label].mean(), flush=True)
        r.close()
    d = np.abs(imgs["megakernel"] - imgs["wavefront"])
    print(name, "max abs diff", d.max(), "finite", np.isfinite(imgs["wavefront"]).all(), flush=True)
