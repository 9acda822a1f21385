#!/bin/bash
# Round 2 (second session), GPU call 19: lane-parallel log p(b) (A/B against the same build without it), draw-chunk length.
set -u
mkdir -p gpurun_out
L=jmbayes_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s19_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s19_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s19_$tag.json 2> gpurun_out/r2b_s19_$tag.err; }
run pbl_1 JMB_X=0
run nopbl_1 JMB_LIB=$L/libjmb200_nopbl.so
run pbl_2 JMB_X=0
run nopbl_2 JMB_LIB=$L/libjmb200_nopbl.so
run chunk25 JMB_RNG_CHUNK=25
run chunk50 JMB_RNG_CHUNK=50
run chunk5 JMB_RNG_CHUNK=5
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_pbl JMB_X=0
run c3_nopbl JMB_LIB=$L/libjmb200_nopbl.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s19_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(10), l["roofline"]["kernel"].ljust(34), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1], "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
