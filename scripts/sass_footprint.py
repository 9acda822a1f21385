"""Static code size of one kernel by source line (instruction-cache footprint): python scripts/sass_footprint.py object.o kernel-fragment [top]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
obj, frag = sys.argv[1:3]; top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, ln in enumerate(sass) if ln.startswith(".text.") and frag in ln)
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith(".text.") or sass[i].startswith(".section"))
c = Counter(); cur = None; n = 0
for ln in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln): c[cur] += 1; n += 1
print("instructions", n, "=", n * 16 // 1024, "KB")
for k, v in c.most_common(top):
    src = ""
    p = "jmbayes_b200/csrc/" + k[0] if k else ""
    if k and os.path.exists(p): src = open(p).read().split("\n")[k[1] - 1].strip()[:90]
    print(f"{str(k):34s} {v:5d}  {src}")
