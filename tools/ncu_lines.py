"""Join an ncu SASS source page (per-instruction counters) with nvdisasm -gi line info and
aggregate per CUDA source line.  Usage:
    python tools/ncu_lines.py <report.ncu-rep> <object.o> <mangled kernel name> [top N]
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
asm = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# instruction -> (file, line) in order for the kernel's text section
lines = []
inside = False
cur = ("?", 0)
fresh_group = True  # nvdisasm prints one //## line per inline level, innermost first
for ln in asm:
    if ln.startswith("\t.section\t.text."):
        inside = (".text." + kern + ",") in ln
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh_group:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh_group = False
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
        fresh_group = True
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hdr = rows[1]
ie, te, ws = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = rows[2:]
base = int(data[0][0], 16)
byoff = {off: (fl, text) for off, fl, text in lines}
agg = collections.defaultdict(lambda: [0, 0, 0])
fagg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
miss = 0
for r in data:
    off = int(r[0], 16) - base
    fl, _ = byoff.get(off, (("?", 0), ""))
    if off not in byoff:
        miss += 1
    v = (int(r[ie]), int(r[te]), int(r[ws]))
    for k in range(3):
        agg[fl][k] += v[k]
        fagg[fl[0]][k] += v[k]
        tot[k] += v[k]
print("instructions %d  thread-inst/inst %.2f  stall samples %d  (unmatched rows %d)" % (tot[0], tot[1] / max(tot[0], 1), tot[2], miss))
print("\n== by file ==")
for f, v in sorted(fagg.items(), key=lambda kv: -kv[1][0]):
    print("%-18s inst %6.2f%%  lanes %5.1f  stalls %6.2f%%" % (f, 100.0 * v[0] / tot[0], v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1)))
print("\n== top lines by warp instructions ==")
for fl, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s:%-4d inst %5.2f%%  lanes %5.1f  stalls %5.2f%%" % (fl[0], fl[1], 100.0 * v[0] / tot[0], v[1] / max(v[0], 1), 100.0 * v[2] / max(tot[2], 1)))
