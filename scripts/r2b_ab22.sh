#!/bin/bash
# Round 2 (second session), GPU call 22: occupancy bound of the draw kernel (A/B).
set -u
mkdir -p gpurun_out
L=jmbayes_b200
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s22_$tag.json 2> gpurun_out/r2b_s22_$tag.err; }
run def_1 JMB_X=0
run prep6 JMB_LIB=$L/libjmb200_prep6.so
run prep8 JMB_LIB=$L/libjmb200_prep8.so
run isig JMB_LIB=$L/libjmb200_isig.so
run def_2 JMB_X=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s22_*.json")):
    l = json.loads(open(f).read().strip().splitlines()[-1])
    print(f.split("/")[-1][8:-5].ljust(8), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
PY
