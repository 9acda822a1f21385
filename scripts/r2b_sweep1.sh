#!/bin/bash
# Round 2 (second session), GPU call 1: cost-balanced step lists (LPT) and the pipelined variant (JMB_ROWS_EARLY=2).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s1_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s1_pytest.log
export JMB_ROWS_DEBUG=1
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { # tag, env...
  tag=$1; shift
  env "$@" $B > gpurun_out/r2b_s1_$tag.json 2> gpurun_out/r2b_s1_$tag.err
}
run old JMB_ROWS_LPT=0
run lpt JMB_ROWS_LPT=1
run lpt_r08 JMB_LPT_ROW=0.08
run lpt_r20 JMB_LPT_ROW=0.20
run lpt_v0 JMB_LPT_VIS=0.0
run lpt_v10 JMB_LPT_VIS=0.10
run p8 JMB_ROWS_EARLY=2
run p8_nolpt JMB_ROWS_EARLY=2 JMB_ROWS_LPT=0
run p8_vt1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run p8_r20 JMB_ROWS_EARLY=2 JMB_LPT_ROW=0.20
run p8_r08 JMB_ROWS_EARLY=2 JMB_LPT_ROW=0.08
run p4 JMB_ROWS_EARLY=2 JMB_ROWS_G=4
run p16 JMB_ROWS_EARLY=2 JMB_ROWS_G=16
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_def JMB_X=0
run c3_p16 JMB_ROWS_EARLY=2
run c3_p8 JMB_ROWS_EARLY=2 JMB_ROWS_G=8
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_def JMB_X=0
run c2_p4 JMB_ROWS_EARLY=2
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_def JMB_X=0
run ms_p8 JMB_ROWS_EARLY=2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s1_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM frac=%.3f e2e=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["roofline"]["frac"], l["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-300:])
PY
