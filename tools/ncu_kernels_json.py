"""Record the per-launch figures bench.py's roofline uses from an `ncu --set full` capture of one launch of the dominant
kernel, taken under `python bench.py --workload <w> --spp <s> --steps 1 --warmup 1 ...` (whose own JSON line, in the log,
says how many segments that launch traced):

    python tools/ncu_kernels_json.py <workload name as bench prints it> <report.ncu-rep> <bench log> [<workload> <rep> <log> ...]

-> profiles/ncu_kernels.json: {workload: {dram_bytes, inst, thread_inst, segments, capture_spp, kernel, gpu_time_ms, source}}
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_kernels.json")


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    return {h: (v, u) for h, u, v in zip(rows[0], rows[1], rows[2])}


def num(d, key):
    v, u = d[key]
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}.get(u, 1)


def main():
    db = json.load(open(OUT)) if os.path.exists(OUT) else {}
    a = sys.argv[1:]
    for name, rep, log in zip(a[0::3], a[1::3], a[2::3]):
        d = raw(rep)
        line = [json.loads(l) for l in open(log) if l.startswith("{") and '"segments_per_step"' in l][-1]
        db[name] = {
            "kernel": d["Kernel Name"][0],
            "dram_bytes": num(d, "dram__bytes_read.sum") + num(d, "dram__bytes_write.sum"),
            "inst": num(d, "smsp__inst_executed.sum"),
            "thread_inst": num(d, "thread_inst_executed"),
            "segments": line["segments_per_step"],
            "capture_spp": line["config"]["spp_per_gpu"],
            "gpu_time_ms": num(d, "gpu__time_duration.sum"),
            "lts_hit_rate_pct": num(d, "lts__t_sector_hit_rate.pct"),
            "source": os.path.basename(rep),
        }
        print(name, db[name])
    db["_note"] = ("ONE launch of the dominant kernel per workload under `ncu --set full --clock-control none` (cold cache, replayed); "
                   "per-segment ratios (thread_inst / segments, dram_bytes / segments) are what bench.py uses, scaled to the timed step's segments")
    json.dump(db, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
