"""Summarise one `ncu --set full` capture as markdown: key metrics, stall reasons and (when the object file the
capture ran is still in build/obj) the per-source-line hot spots of tools/ncu_lines.py.

    python tools/ncu_summary.py <report.ncu-rep> "<title>" [<object file the capture ran | - >] [> profiles/rNN_<what>.md]

The source-line section needs the VERY object file the captured kernel came from (its SASS offsets are the join key):
pass the copy saved next to the capture, or `-` to leave the section out.
"""
import csv, io, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_active.avg",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return rows[0], rows[1], rows[2]


def main():
    rep, title = sys.argv[1], sys.argv[2]
    hdr, units, vals = raw(rep)
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("# %s\n" % title)
    print("Capture: `ncu --set full --clock-control none --import-source on` (cold-cache, serialised replay: use shares, not absolutes).\n")
    print("| metric | value | unit |\n|---|---|---|")
    print("| kernel | `%s` | |" % d.get("Kernel Name", ("?", ""))[0])
    for k in KEYS:
        if k in d:
            print("| %s | %s | %s |" % (k, d[k][0], d[k][1]))
    print("\n## Warp stall reasons (average warps stalled per issue-active cycle)\n\n| reason | value |\n|---|---|")
    st = [(float(v[0].replace(",", "")), h) for h, v in d.items() if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    for v, h in sorted(st, reverse=True)[:10]:
        print("| %s | %.3f |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    # the raw page has the demangled short name only: rebuild the mangled one for render_kernel<R, MAXD, STATS, FEAT>
    import re
    obj = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "build", "obj", "kernels_f32.o")
    name = d.get("Kernel Name", ("", ""))[0]
    mangled = None
    m = re.match(r"void render_kernel<(float|double), (\d+), (\d+), (\d+)>", name)
    if m and obj != "-":
        mangled = "_ZN4rptb13render_kernelI%sLi%sELb%sELi%sEEEvNS_9SceneViewIT_EENS_10RenderArgsIS2_EE" % (
            "f" if m.group(1) == "float" else "d", m.group(2), m.group(3), m.group(4))
    m = re.match(r"void render_kernel_vx<(\d+), (\d+)>", name)
    if m and obj != "-":
        mangled = "_ZN4rptb16render_kernel_vxILb%sELi%sEEEvNS_9SceneViewIfEENS_10RenderArgsIfEE" % (m.group(1), m.group(2))
    if mangled and os.path.exists(obj):
        tool = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, obj, mangled, "30"], capture_output=True, text=True)
        print("\n## Source-level hot spots (ncu SASS counters joined with nvdisasm line info, tools/ncu_lines.py)\n\n```")
        print((tool.stdout or tool.stderr).rstrip())
        print("```")


if __name__ == "__main__":
    main()
