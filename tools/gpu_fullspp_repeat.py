"""Is the spread of the one-launch full-size renders (bench.py's `secondary`) per box, per process or per launch?
Renders glass at 4 096 spp three times and the dragon proxy at 1 024 spp twice in THIS process; run it in two processes."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api

def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "p"
    for name, reps in (("glass", 3), ("dragon", 2)):
        cfg = scenes.CONFIGS[name]()
        r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
        r.device_scene()
        buf = api.Buffer(cfg.width, cfg.height)
        r.sample(max(1, cfg.spp // 64), buf)
        for rep in range(reps):
            r._next_sample = 0
            buf = api.Buffer(cfg.width, cfg.height)
            r.sample(cfg.spp, buf, collect_stats=0)
            st = r.last_stats
            print(json.dumps({"proc": tag, "config": name, "spp": cfg.spp, "rep": rep, "gpu_ms": st["gpu_ms"],
                              "Msamples_s": st["segments"] / st["gpu_ms"] / 1e3}), flush=True)
        r.close()

if __name__ == "__main__":
    main()
