"""Per-config throughput table on one GPU, reduced spp: the five BASELINE configs, or with `extra` as the
first argument the three row-N4 example scenes (kd-trees over whole shapes, MonomialSurface)."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi

def main():
    extra = len(sys.argv) > 1 and sys.argv[1] == "extra"
    spps = {"sphere": 100, "cornell": 64, "teapot": 64, "dragon": 32, "glass": 64,
            "fractal_spheres": 64, "fractal_teapots": 32, "monomial_glass": 100}
    names = ["fractal_spheres", "fractal_teapots", "monomial_glass"] if extra else ["sphere", "cornell", "teapot", "dragon", "glass"]
    rows = []
    for name in names:
        t0 = time.time()
        cfg = (scenes.EXTRA_CONFIGS if extra else scenes.CONFIGS)[name]()
        build_s = time.time() - t0
        r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
        t0 = time.time(); r.device_scene(); upload_s = time.time() - t0
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(2, buf)  # warm-up
        r._next_sample = 0
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(spps[name], buf, collect_stats=1)
        st = dict(r.last_stats)
        r._next_sample = 0
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(spps[name], buf, collect_stats=0)
        st2 = r.last_stats
        row = {"config": cfg.name, "res": "%dx%d" % (cfg.width, cfg.height), "spp": spps[name], "mb": cfg.max_bounces,
               "gpu_ms": st2["gpu_ms"], "Msamples_s": st2["segments"] / st2["gpu_ms"] / 1e3, "Mrays_s": st2["rays"] / st2["gpu_ms"] / 1e3,
               "segments": st2["segments"], "rays": st2["rays"], "node_visits_per_ray": st["node_visits"] / max(st["rays"], 1),
               "tri_tests_per_ray": st["tri_tests"] / max(st["rays"], 1), "scene_build_s": build_s, "upload_s": upload_s,
               "device_bytes": r.device_scene().device_bytes(), "image_mean": float(buf.batches[0].mean())}
        rows.append(row)
        print(json.dumps(row), flush=True)
        r.close()
    json.dump(rows, open("gpurun_out/scenes_table_extra.json" if extra else "gpurun_out/scenes_table.json", "w"), indent=1)

if __name__ == "__main__":
    main()
