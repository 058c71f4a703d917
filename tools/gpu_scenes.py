"""Per-config throughput table (all five BASELINE configs) on one GPU, reduced spp."""
import sys, os, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi

def main():
    spps = {"sphere": 100, "cornell": 64, "teapot": 64, "dragon": 32, "glass": 64}
    rows = []
    for name in ["sphere", "cornell", "teapot", "dragon", "glass"]:
        t0 = time.time()
        cfg = scenes.CONFIGS[name]()
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
    json.dump(rows, open("gpurun_out/scenes_table.json", "w"), indent=1)

if __name__ == "__main__":
    main()
