#!/bin/bash
# Round 2 (second session), GPU call 14: pre-generated draws on the finalize launches (A/B), parity, launch list.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py tests/test_gpu_rcpp_shim.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s14_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s14_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s14_$tag.json 2> gpurun_out/r2b_s14_$tag.err; }
run on_1 JMB_X=0
run off_1 JMB_RNG_PREGEN=0
run on_2 JMB_X=0
run off_2 JMB_RNG_PREGEN=0
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_on JMB_X=0
run c3_off JMB_RNG_PREGEN=0
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_on JMB_X=0
run c2_off JMB_RNG_PREGEN=0
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_on JMB_X=0
run ms_off JMB_RNG_PREGEN=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s14_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(10), l["roofline"]["kernel"].ljust(40), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1], "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
