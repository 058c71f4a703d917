"""Turn the ncu captures of tools/make_profiles.sh (gpurun_out/<round>_*) into the committed
summaries under profiles/: key metrics, stall reasons, launch-time shares, per-source-line hot
spots, and profiles/ncu_traffic.json (dram bytes per launch, read by bench.py).

    python tools/summarize_profiles.py r01
"""
import csv, io, json, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
G = os.path.join(ROOT, "gpurun_out")

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "launch__local_memory_size" if False else "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2]


def summarize(rep, title, obj, kern, note):
    hdr, units, vals = raw(rep)
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    lines = ["# %s" % title, "", note, "", "Capture: `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replay: use shares, not absolutes).", "",
             "| metric | value | unit |", "|---|---|---|"]
    lines.append("| kernel | `%s` | |" % d.get("Kernel Name", ("?", ""))[0])
    for k in KEYS:
        if k in d:
            lines.append("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
    lines += ["", "## Warp stall reasons (average warps stalled per issue-active cycle)", "", "| reason | value |", "|---|---|"]
    st = [(float(v[0].replace(",", "")), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    for v, h in sorted(st, reverse=True)[:10]:
        lines.append("| %s | %.3f |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    tool = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, obj, kern, "25"], capture_output=True, text=True).stdout
    lines += ["", "## Source-level hot spots (ncu SASS counters joined with nvdisasm line info, tools/ncu_lines.py)", "", "```", tool.rstrip(), "```", ""]
    dram = None
    try:
        def tobytes(v, u):
            x = float(v.replace(",", ""))
            return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        dram = tobytes(*d["dram__bytes_read.sum"]) + tobytes(*d["dram__bytes_write.sum"])
    except Exception:
        pass
    return "\n".join(lines), dram


def launches(csvpath, title):
    rows = [r for r in csv.reader(open(csvpath)) if len(r) > 10 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0][:90]
        t = float(r[-1].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    lines = ["# %s" % title, "", "`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch device time; shares only).", "",
             "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| `%s` | %d | %.3f | %.1f %% |" % (k, v[0], v[1] / 1e6, 100 * v[1] / tot))
    return "\n".join(lines)


def main():
    os.makedirs(OUT, exist_ok=True)
    traffic = {}
    obj = os.path.join(ROOT, "build", "obj", "kernels_f32.o")
    jobs = [("%s_render_kernel_cornell" % R, "Cornell 800x800 (32 spp capture) -- rptb::render_kernel<float,16,false,0>",
             "_ZN4rptb13render_kernelIfLi16ELb0ELi0EEEvNS_9SceneViewIT_EENS_10RenderArgsIS2_EE", "cornell",
             "Dominant kernel of the bench workload (BASELINE configs[1]); one launch per step per GPU."),
            ("%s_wf_trace_dragon" % R, "dragon-proxy 1920x1080 (4 spp capture) -- rptb::wf_trace_kernel<false>",
             "_ZN4rptb15wf_trace_kernelILb0EEEvNS_9SceneViewIfEENS_9WfBuffersEPKjPNS_14DeviceCountersE", "dragon-proxy",
             "Dominant kernel of the wavefront engine (one launch per path-vertex step)."),
            ("%s_render_kernel_fractal_teapots" % R, "examples/fractal_teapots 800x600 (8 spp capture) -- rptb::render_kernel<float,16,false,F_EVERY>",
             "_ZN4rptb13render_kernelIfLi16ELb0ELi55EEEvNS_9SceneViewIT_EENS_10RenderArgsIS2_EE", "fractal_teapots",
             "Row N4: kd-trees over 781 instances of one teapot kd-tree (tools/ncu_ext.sh); not a BASELINE config.")]
    for stem, title, kern, wl, note in jobs:
        rep = os.path.join(G, stem + ".ncu-rep")
        if not os.path.exists(rep):
            print("missing", rep)
            continue
        txt, dram = summarize(rep, title, obj, kern, note)
        open(os.path.join(OUT, stem + ".md"), "w").write(txt + "\n")
        if dram is not None:
            traffic[wl] = dram
        print("wrote", stem + ".md", "dram bytes/launch", dram)
    for wl in ("cornell", "dragon"):
        p = os.path.join(G, "%s_launches_%s.csv" % (R, wl))
        if os.path.exists(p):
            open(os.path.join(OUT, "%s_launches_%s.md" % (R, wl)), "w").write(launches(p, "Launch list: bench.py --workload %s" % wl) + "\n")
            print("wrote launches", wl)
    tpath = os.path.join(OUT, "ncu_traffic.json")
    old = json.load(open(tpath)) if os.path.exists(tpath) else {}
    note = old.get("_note", {})
    spp = {"cornell": 32, "dragon-proxy": 4, "fractal_teapots": 8}
    for k, v in traffic.items():
        old[k] = {"dram_bytes_per_launch": v, "capture_spp": spp[k]}
        note[k] = ("dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel in the %s capture, "
                   "taken at capture_spp samples per pixel (ncu replays a launch ~40 times).  The megakernel's traffic is "
                   "its local-memory level stack and register spills and scales with spp; bench.py scales it to the "
                   "workload's spp.  A wavefront trace launch covers one path-vertex step and does not depend on spp." % R)
    old["_note"] = note
    json.dump(old, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
