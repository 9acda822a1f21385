"""Per-source-line stall samples of one kernel: joins `ncu --page source --csv --print-source sass` (per SASS instruction) with the
line table nvdisasm prints for the same object.  python scripts/ncu_lines.py report.ncu-rep object.o kernel-name-fragment [top]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict
rep, obj, frag = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
base = int(data[0][0], 16)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
kname = rows[0][1]
mang = None
start = None
for i, ln in enumerate(sass):
    if ln.startswith(".text.") and frag in ln:
        start = i; break
assert start is not None, "kernel not found in the object"
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.") or sass[i].startswith(".section"))
line_of = {}; cur = None
for ln in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m: line_of[int(m.group(1), 16)] = (cur, m.group(2))
agg = defaultdict(lambda: defaultdict(int))
keys = ["# Samples", "Instructions Executed", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_barrier", "stall_mio", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_no_inst", "stall_lg", "L1 Wavefronts Shared Excessive"]
tot = 0
for r in data:
    off = int(r[0], 16) - base
    ln = line_of.get(off, (None, ""))[0]
    for k in keys:
        try: agg[ln][k] += int(float(r[ix[k]]))
        except Exception: pass
    tot += int(r[ix["# Samples"]] or 0)
print("kernel:", kname[:100]); print("total samples", tot, "instructions", sum(a["Instructions Executed"] for a in agg.values()))
src_cache = {}
def src(ln):
    if not ln: return ""
    f = [p for p in ("jmbayes_b200/csrc/" + ln[0],) if os.path.exists(p)]
    if not f: return ""
    if f[0] not in src_cache: src_cache[f[0]] = open(f[0]).read().split("\n")
    return src_cache[f[0]][ln[1] - 1].strip()[:100]
for ln, a in sorted(agg.items(), key=lambda x: -x[1]["# Samples"])[:top]:
    print(f"{str(ln):32s} smp {a['# Samples']:5d} ({100*a['# Samples']/tot:4.1f}%) inst {a['Instructions Executed']:9d} long {a['stall_long_sb']:5d} wait {a['stall_wait']:5d} short {a['stall_short_sb']:4d} bar {a['stall_barrier']:4d} mio {a['stall_mio']:3d} math {a['stall_math']:3d} br {a["stall_branch_resolving"]:3d} noi {a["stall_no_inst"]:4d} xs {a['L1 Wavefronts Shared Excessive']:7d} | {src(ln)}")
