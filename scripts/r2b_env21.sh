#!/bin/bash
# Round 2 (second session), GPU call 21: environment-only sweep on the final build (waves, cost weights, staging variant).
set -u
mkdir -p gpurun_out
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s21_$tag.json 2> gpurun_out/r2b_s21_$tag.err; }
run def_1 JMB_X=0
run waves2 JMB_ROWS_WAVES=2
run r10v10 JMB_LPT_ROW=0.10 JMB_LPT_VIS=0.10
run r20v02 JMB_LPT_ROW=0.20 JMB_LPT_VIS=0.02
run q8 JMB_ROWS_EARLY=3
run def_2 JMB_X=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s21_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(10), l["roofline"]["kernel"].ljust(34), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1], "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
