#!/bin/bash
# Round 2 (second session), GPU call 15: same-box A/B of the library before / after the pre-generated-draws change (switch off and on).
set -u
mkdir -p gpurun_out
L=jmbayes_b200
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s15_$tag.json 2> gpurun_out/r2b_s15_$tag.err; }
run prev_1 JMB_LIB=$L/libjmb200_prev.so
run off_1 JMB_RNG_PREGEN=0
run on_1 JMB_X=0
run prev_2 JMB_LIB=$L/libjmb200_prev.so
run off_2 JMB_RNG_PREGEN=0
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_prev JMB_LIB=$L/libjmb200_prev.so
run c2_off JMB_RNG_PREGEN=0
run c2_on JMB_X=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s15_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(10), l["roofline"]["kernel"].ljust(40), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1], "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
