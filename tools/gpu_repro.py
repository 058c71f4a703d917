"""Tiny renders of a kd-tree-of-shapes scene, one (precision, max_bounces, stats) case per process argument,
for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/gpu_repro.py f32 1 0"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import _capi as capi, api, scenes  # noqa: E402


def main():
    prec = capi.PRECISION_F64 if sys.argv[1] == "f64" else capi.PRECISION_F32
    mb, stats = int(sys.argv[2]), int(sys.argv[3])
    name = sys.argv[4] if len(sys.argv) > 4 else "fractal_spheres"
    cfg = scenes.fractal_spheres_scene(3) if name == "fractal_spheres" else scenes.fractal_teapots_scene(2)
    r = api.Renderer(cfg.scene, cfg.camera).width(32).height(24).max_bounces(mb).seed(1).precision(prec)
    buf = api.Buffer(32, 24)
    r.sample(2, buf, collect_stats=stats)
    img = buf.batches[0]
    print("ok", sys.argv[1:], float(img.mean()), bool(np.isfinite(img).all()), r.last_stats["segments"], flush=True)
    r.close()


if __name__ == "__main__":
    main()
