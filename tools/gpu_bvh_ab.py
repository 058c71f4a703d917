"""A/B of library builds on one GPU (reduced spp, 2 repeats each, best of): tools/gpu_bvh_ab.py <tag> [config ...]
(default: the mesh configs teapot, dragon, dragon_knot, fractal_teapots).  The library under test is chosen with RPTB_LIB;
prints one JSON line per config and writes gpurun_out/bvh_ab_<tag>.json."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api

def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "main"
    spps = {"teapot": 64, "dragon": 32, "dragon_knot": 32, "fractal_teapots": 32, "sphere": 100, "cornell": 64, "glass": 64,
            "fractal_spheres": 64, "monomial_glass": 100}
    names = sys.argv[2:] or ["teapot", "dragon", "dragon_knot", "fractal_teapots"]
    rows = []
    for name in names:
        spp = spps[name]
        cfg = (scenes.CONFIGS.get(name) or scenes.EXTRA_CONFIGS[name])()
        r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
        r.device_scene()
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(2, buf)
        best = None
        for rep in range(2):
            r._next_sample = 0
            buf = api.Buffer(cfg.width, cfg.height)
            r.sample(spp, buf, collect_stats=0)
            st = r.last_stats
            if best is None or st["gpu_ms"] < best["gpu_ms"]:
                best = dict(st)
        r._next_sample = 0
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(min(spp, 8), buf, collect_stats=2)
        sc = r.last_stats
        row = {"tag": tag, "config": name, "spp": spp, "gpu_ms": best["gpu_ms"], "Msamples_s": best["segments"] / best["gpu_ms"] / 1e3,
               "bvh_nodes_per_ray": sc["bvh_node_visits"] / max(sc["rays"], 1), "bvh_tris_per_ray": sc["bvh_tri_tests"] / max(sc["rays"], 1),
               "image_mean": float(buf.batches[0].mean())}
        rows.append(row)
        print(json.dumps(row), flush=True)
        r.close()
    json.dump(rows, open("gpurun_out/bvh_ab_%s.json" % tag, "w"), indent=1)

if __name__ == "__main__":
    main()
