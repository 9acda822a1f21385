#!/bin/bash
# Round 2 (second session), GPU call 7: division-free logit rows, lane-parallel censoring terms, variant 3 without the meta stall.
set -u
mkdir -p gpurun_out
export JMB_ROWS_DEBUG=1
L=jmbayes_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s7_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s7_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s7_$tag.json 2> gpurun_out/r2b_s7_$tag.err; }
run def JMB_X=0
run e1 JMB_ROWS_EARLY=1
run q8 JMB_ROWS_EARLY=3
run e0 JMB_ROWS_EARLY=0
run p8 JMB_ROWS_EARLY=2
run nologit_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_nologit.so
run nologit_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_nologit.so
run noptbl_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_noptbl.so
run noptbl_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_noptbl.so
run q8_r08 JMB_ROWS_EARLY=3 JMB_LPT_ROW=0.08
run q8_r20 JMB_ROWS_EARLY=3 JMB_LPT_ROW=0.20
run q8_r30 JMB_ROWS_EARLY=3 JMB_LPT_ROW=0.30
run q8_v10 JMB_ROWS_EARLY=3 JMB_LPT_VIS=0.10
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_def JMB_X=0
run c3_q8 JMB_ROWS_EARLY=3
run c3_q16 JMB_ROWS_EARLY=3 JMB_ROWS_G=16
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_def JMB_X=0
run c2_q8 JMB_ROWS_EARLY=3
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_def JMB_X=0
run ms_q16 JMB_ROWS_EARLY=3
run ms_p16v2 JMB_ROWS_VT=2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s7_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][7:-5].ljust(14), l["roofline"]["kernel"].ljust(46), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1][7:-5].ljust(14), "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
JMB_ROWS_EARLY=3 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s7_q8 $B > gpurun_out/r2b_ncu8.log 2>&1
JMB_ROWS_EARLY=1 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s7_e1 $B > gpurun_out/r2b_ncu9.log 2>&1
ls -la gpurun_out/r2b_s7*.ncu-rep
