#!/usr/bin/env python
"""bench.py -- headline benchmark of the path-tracing hot path (BASELINE.json metric:
Msamples/s = path segments, i.e. trace_ray invocations, per second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path
                                                             # (rpt-restated C++ oracle: the
                                                             # reference is Rust, no rustc here)

A "step" is one pass of Renderer::sample over one batch: the BASELINE configs[1] workload,
Cornell box 800x800, 512 spp per GPU, max_bounces 6 (at N GPUs the image gets 512*N spp and
every GPU renders 1/N of the pixel tiles: per-GPU work is fixed => weak scaling), followed
by the single all-reduce of the float3 buffer when N > 1.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Msamples/s (path segments = trace_ray invocations per second)"
UNIT = "Msamples/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cornell",
                    choices=["sphere", "cornell", "teapot", "dragon", "glass",
                             "fractal_spheres", "fractal_teapots", "monomial_glass"],  # the last three: exploration only
                    help="default = the BASELINE configs[1] workload; anything else is for exploration / profiling")
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per GPU (exploration only)")
    ap.add_argument("--engine", default="auto", choices=["auto", "megakernel", "wavefront"], help="rptb_engine (exploration only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def workload(name: str, spp_override: int = 0):
    from rpt_b200 import scenes

    cfg = (scenes.CONFIGS[name] if name in scenes.CONFIGS else scenes.EXTRA_CONFIGS[name])()
    if spp_override:
        cfg.spp = spp_override
    return cfg


def config_dict(cfg, world: int, extra=None):
    d = {
        "workload": "%s %dx%d, %d spp per GPU (%d total), max_bounces %d; %s" % (
            cfg.name, cfg.width, cfg.height, cfg.spp, cfg.spp * world, cfg.max_bounces, cfg.note),
        "scene": cfg.name,
        "width": cfg.width,
        "height": cfg.height,
        "spp_per_gpu": cfg.spp,
        "spp_total": cfg.spp * world,
        "max_bounces": cfg.max_bounces,
        "parallelism": "pixel tiles 16x8 round-robin over %d GPU(s); one all-reduce(sum) of the float3 buffer" % world,
        "rng": "Philox4x32-10 keyed (seed=1, pixel, sample)",
        "cache": "L2 flushed between timed steps (256 MiB memset); the scene itself is < 1 MiB and cache-resident by nature",
    }
    if extra:
        d.update(extra)
    return d


# ------------------------------------------------------------------ clocks ------------
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, ngpus: int):
        self.ngpus = ngpus
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                idx = int(f[0])
                if idx >= self.ngpus:
                    continue
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------ roofline ----------
def algorithmic_bytes(stats: dict, pixels: int) -> float:
    """SURVEY 8(d) per-unit figures with this repo's device layout:
    64 B object record per Shape::intersect dispatch, 8 B per kd node visited, 4 + 48 B per
    triangle test (leaf ref + packed triangle), 36 B vertex normals per mesh hit, 32 B
    material per segment (<= one fetch), 64 B (4 texels x 16 B) per HDRI lookup, 12 B per
    pixel written."""
    return (64.0 * stats["object_tests"] + 8.0 * stats["node_visits"] + 52.0 * stats["tri_tests"]
            + 64.0 * stats.get("bvh_node_visits", 0) + 52.0 * stats.get("bvh_tri_tests", 0)
            + 36.0 * stats["mesh_hits"] + 32.0 * stats["segments"] + 64.0 * stats["env_lookups"] + 12.0 * pixels)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload_name: str, spp: int, engine: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed ncu capture (profiles/ncu_traffic.json); megakernel launches are scaled from the
    capture's spp to the workload's (their DRAM traffic is per-sample local-memory traffic)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(path)).get(workload_name)
        if not isinstance(rec, dict):
            return None
        b = float(rec["dram_bytes_per_launch"])
        if (workload_name == "dragon-proxy") != (engine == 2):
            return None  # the dragon record is a wavefront trace launch; no capture of the megakernel on that workload yet
        return b if engine == 2 else b * spp / float(rec["capture_spp"])
    except Exception:
        return None


# ------------------------------------------------------------------ CPU arm -----------
def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_rate(cfg, budget_s: float = 15.0):
    """Time the CPU restatement of the same workload on all host threads, on a bounded
    sample: whole-resolution renders at reduced spp (the rate is spp-independent)."""
    from oracle import oracle_py as orc
    from rpt_b200 import api

    flat = api.FlatScene(cfg.scene)
    osc = orc.OracleScene(flat)
    r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
    cores = host_threads()
    t0 = time.perf_counter()
    _, st = osc.render(cfg.camera, r.params(1), nthreads=cores)
    t1 = time.perf_counter() - t0
    spp = 1
    segs, secs = st["segments"], t1
    if t1 < budget_s / 2:
        spp = int(max(1, min(96, math.floor(budget_s / max(t1, 1e-3)) - 1)))
        t0 = time.perf_counter()
        _, st = osc.render(cfg.camera, r.params(spp, first_sample=1), nthreads=cores)
        secs = time.perf_counter() - t0
        segs = st["segments"]
    return {
        "value": segs / secs / 1e6,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": "%s %dx%d, %d of %d spp, max_bounces %d: %d segments in %.2f s (rpt-restated C++ f64 oracle, "
                  "OpenMP over rows; rpt itself is Rust and cannot be built here)" % (
                      cfg.name, cfg.width, cfg.height, spp, cfg.spp, cfg.max_bounces, segs, secs),
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import oracle_py as orc
    from rpt_b200 import api

    cfg = workload(args.workload, args.spp)
    flat = api.FlatScene(cfg.scene)
    osc = orc.OracleScene(flat)
    r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces).seed(1)
    cores = host_threads()  # torchrun exports OMP_NUM_THREADS=1: ask for the cores explicitly
    sample_spp = 1  # one sample per pixel of the full-resolution image per step
    for i in range(args.warmup):
        osc.render(cfg.camera, r.params(sample_spp, first_sample=i), nthreads=cores)
    segs, t = 0, 0.0
    for i in range(args.steps):
        t0 = time.perf_counter()
        _, st = osc.render(cfg.camera, r.params(sample_spp, first_sample=args.warmup + i), nthreads=cores)
        t += time.perf_counter() - t0
        segs += st["segments"]
    value = segs / t / 1e6
    sample = "%s %dx%d, %d of %d spp per step, max_bounces %d (rpt-restated C++ f64 oracle on %d host threads)" % (
        cfg.name, cfg.width, cfg.height, sample_spp, cfg.spp, cfg.max_bounces, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(cfg, args.gpus, {"note_reference": "CPU path; n_gpus is echoed, no GPU is used"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ------------------------------------------------------------------ GPU arm -----------
def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def run_native(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import numpy as np
    import torch
    import torch.distributed as dist

    from rpt_b200 import _capi as capi
    from rpt_b200 import api
    from rpt_b200.distributed import render_shard_device

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if capi.lib().rptb_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device and rpt_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = workload(args.workload, args.spp)
    spp_total = cfg.spp * world
    r = api.Renderer(cfg.scene, cfg.camera).width(cfg.width).height(cfg.height).max_bounces(cfg.max_bounces) \
        .seed(1).device(local).engine({"auto": capi.ENGINE_AUTO, "megakernel": capi.ENGINE_MEGAKERNEL, "wavefront": capi.ENGINE_WAVEFRONT}[args.engine])
    flat = api.FlatScene(cfg.scene)
    npix = cfg.width * cfg.height
    stream = torch.cuda.Stream(dev)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()

    with torch.cuda.stream(stream):
        out = torch.empty(npix * 3, dtype=torch.float32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        raw = stream.cuda_stream

        def step(ev_mid=None):
            render_shard_device(r, spp_total, out, rank, world, 0, raw)
            if ev_mid is not None:
                ev_mid.record()
            if world > 1:
                dist.all_reduce(out)

        # un-timed statistics pass: exact segment count of one step + traversal counters
        st = capi.Stats()
        render_shard_device(r, spp_total, out, rank, world, 0, raw, stats=st, collect_stats=1)
        mine = st.as_dict()
        keys = ["segments", "rays", "node_visits", "tri_tests", "mesh_hits", "env_lookups", "object_tests",
                "bvh_node_visits", "bvh_tri_tests"]
        tot = torch.tensor([mine[k] for k in keys], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        total = dict(zip(keys, [int(v) for v in tot.tolist()]))
        launches_per_step = int(st.launches)
        engine = int(st.engine)

        for _ in range(W):
            step()
        stream.synchronize()
        barrier()
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
        sampler = ClockSampler(world) if rank == 0 else None
        if sampler:
            sampler.start()
        t_wall = time.perf_counter()
        for i in range(K):
            flush.zero_()  # evict L2 between timed steps (outside the event pair)
            ev[i][0].record()
            step(ev[i][1])
            ev[i][2].record()
        stream.synchronize()
        barrier()
        t_wall = time.perf_counter() - t_wall
        clocks = sampler.stop() if sampler else None
        step_ms = sum(e[0].elapsed_time(e[2]) for e in ev)
        kern_ms = sum(e[0].elapsed_time(e[1]) for e in ev)
        times = torch.tensor([step_ms, kern_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        step_ms, kern_ms = [float(v) for v in times.tolist()]
        image_mean = float(out.mean().item())

        # ---- e2e: the public call with HOST buffers, copies inside the timed region -----
        e2e = None
        if not args.no_e2e:
            host = torch.empty(npix * 3, dtype=torch.float32).pin_memory()
            host64 = np.empty((npix, 3), np.float64)
            cam = cfg.camera.to_c()

            def e2e_step():
                ds = api.DeviceScene(flat, local)  # H2D: the flattened scene (host arrays -> HBM)
                try:
                    p = r.params(spp_total, 0, rank, world)
                    if world == 1:
                        # the reference-facing C-ABI call: host double buffer out (D2H inside)
                        capi.check(capi.lib().rptb_render_samples(ds.handle, C.byref(cam), C.byref(p),
                                                                  host64.ctypes.data_as(capi.c_double_p), None),
                                   "rptb_render_samples")
                    else:
                        capi.check(capi.lib().rptb_render_samples_device(ds.handle, C.byref(cam), C.byref(p),
                                                                         C.c_void_p(out.data_ptr()), C.c_void_p(raw), None),
                                   "rptb_render_samples_device")
                        dist.all_reduce(out)
                        host.copy_(out, non_blocking=True)  # D2H of the assembled image
                        stream.synchronize()
                finally:
                    ds.close()

            e2e_step()  # warm-up
            stream.synchronize()
            barrier()
            t0 = time.perf_counter()
            for _ in range(K):
                e2e_step()
            stream.synchronize()
            barrier()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e_s = float(te.item())
            e2e = {
                "value": total["segments"] * K / e2e_s / 1e6, "unit": UNIT,
                "h2d_bytes_per_step": int(flat.host_bytes() + C.sizeof(capi.Camera) + C.sizeof(capi.RenderParams)) * world,
                "d2h_bytes_per_step": int(npix * 3 * 4),
                "ms_per_step": 1e3 * e2e_s / K,
                "call": "rptb_scene_create + rptb_render_samples(host double* out) + rptb_scene_destroy per step" if world == 1
                        else "per rank: rptb_scene_create + rptb_render_samples_device + NCCL all-reduce + D2H to pinned host + destroy",
                "timer": "host perf_counter between synchronize+barrier, max over ranks",
                "h2d_source": "pageable host arrays (cudaMemcpy inside rptb_scene_create)",
            }

    if rank == 0:
        value = total["segments"] * K / (step_ms / 1e3) / 1e6
        peak, peak_src = measured_peaks()
        # dominant kernel: render_kernel<float,16,false>; one launch per step per GPU
        bytes_all = algorithmic_bytes(total, npix)  # summed over ranks (each writes its own pixels)
        kern_s = kern_ms / 1e3 / K
        achieved = bytes_all / world / kern_s / 1e9  # per GPU
        issue = None
        if clocks and clocks.get("sm_mhz"):
            issue = {"note": "this path is instruction-issue bound, not HBM bound: the whole scene is cache resident",
                     "segments_per_sm_clock": total["segments"] / world / kern_s / (148 * clocks["sm_mhz"] * 1e6)}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(cfg, world),
            "segments_per_step": total["segments"], "rays_per_step": total["rays"],
            "kernel_ms_per_step": kern_ms / K, "wall_s_timed_region": t_wall, "image_mean": image_mean,
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches_per_step * K,
            "roofline": {
                "bound": "hbm",
                "kernel": "rptb::render_kernel<float,16,false,FEAT> (megakernel: one launch per step, + chunk resolve)" if engine != 2 else
                          "rptb::wf_trace_kernel<false> (+ wf_shade_kernel; wavefront engine: the duration is the whole step's kernels)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(cfg.name, spp_total, engine), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_all / world,
                "bytes_model": "64*object_tests + 8*node_visits + 52*tri_tests + 64*bvh_node_visits + 52*bvh_tri_tests + 36*mesh_hits + 32*segments + 64*env_lookups + 12*pixels (counters of the structure that rendered the step)",
                "counters_per_step": total,
                "secondary": issue,
            },
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = oracle_rate(cfg)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    r.close()
    return 0


def main():
    args = parse_args()
    # Exactly ONE line may reach stdout.  Libraries write there too (NCCL prints its version banner
    # to stdout when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the JSON
    # line goes to the saved descriptor.
    if args.impl == "native" and args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun exactly like the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(free_port()), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_native(args)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


if __name__ == "__main__":
    sys.exit(main())
