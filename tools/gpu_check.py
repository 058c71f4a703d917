"""Ad-hoc GPU bring-up check (not a test): parity vs the oracle + a first timing."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rpt_b200 import scenes, api, _capi as capi
from oracle import oracle_py as orc

def rays_for(cfg, n, seed=0):
    rng = np.random.default_rng(seed)
    cam = cfg.camera
    o = np.tile(cam.eye, (n, 1)) + rng.normal(0, 0.01, (n, 3))
    right = np.cross(cam.direction, cam.up)
    d = cam.direction[None, :] * (1 / np.tan(cam.fov / 2)) + rng.uniform(-1, 1, (n, 1)) * right[None, :] + rng.uniform(-0.6, 0.6, (n, 1)) * cam.up[None, :]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d], axis=1)

def main():
    out = {}
    for name in ["sphere", "cornell", "teapot"]:
        cfg = scenes.CONFIGS[name]()
        flat = api.FlatScene(cfg.scene)
        osc = orc.OracleScene(flat)
        ds = api.DeviceScene(flat)
        rays = rays_for(cfg, 200000)
        t0, o0, n0, _ = osc.closest_hit(rays)
        for prec in (capi.PRECISION_F64, capi.PRECISION_F32):
            t1, o1, n1, st = ds.closest_hit(rays, precision=prec, want_stats=True)
            same = (o0 == o1)
            hit = same & (o0 >= 0)
            rel = np.abs(t1[hit] - t0[hit]) / np.maximum(np.abs(t0[hit]), 1e-30)
            nerr = np.abs(n1[hit] - n0[hit]).max() if hit.any() else 0
            print(name, "prec", prec, "obj agree %.6f" % same.mean(), "max rel dt %.3g" % (rel.max() if rel.size else 0), "p99.99 %.3g" % (np.quantile(rel, 0.9999) if rel.size else 0), "max dn %.3g" % nerr, st)
        # render parity
        w, h = (96, 54) if name != "cornell" else (64, 64)
        r = api.Renderer(cfg.scene, cfg.camera).width(w).height(h).max_bounces(cfg.max_bounces).seed(1)
        p = r.params(32)
        img0, st0 = osc.render(cfg.camera, p)
        for prec in (capi.PRECISION_F64, capi.PRECISION_F32):
            r.precision(prec)
            buf = api.Buffer(w, h)
            r._next_sample = 0
            r.sample(32, buf, collect_stats=1)
            img1 = buf.batches[0]
            d = np.abs(img1 - img0)
            rel = d / np.maximum(np.abs(img0), 1e-3)
            print(name, "render prec", prec, "mean", img1.mean(0), "oracle mean", img0.mean(0), "max abs %.3g" % d.max(), "median rel %.3g" % np.median(rel), "frac rel>1e-3: %.4f" % (rel > 1e-3).mean(), "rmse %.4g" % np.sqrt((d**2).mean()))
            print("   stats gpu", r.last_stats, "oracle", {k: st0[k] for k in ('segments','rays','node_visits','tri_tests')})
        r.close(); ds.close()
    # timing: cornell
    cfg = scenes.cornell_scene()
    r = api.Renderer(cfg.scene, cfg.camera).width(800).height(800).max_bounces(6).seed(1)
    for spp in (8, 64):
        buf = api.Buffer(800, 800)
        t = time.time(); r.sample(spp, buf); dt = time.time() - t
        st = r.last_stats
        print("cornell 800x800 spp", spp, "gpu_ms %.2f" % st['gpu_ms'], "wall %.3f" % dt, "Mseg/s %.1f" % (st['segments'] / st['gpu_ms'] / 1e3), st)

if __name__ == "__main__":
    main()
