#!/bin/bash
# Round 2 (second session), GPU call 6: variant 3 (cp.async by the lanes into the per-group stage), ILP switches again, more warps per SM.
set -u
mkdir -p gpurun_out
export JMB_ROWS_DEBUG=1
L=jmbayes_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s6_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s6_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s6_$tag.json 2> gpurun_out/r2b_s6_$tag.err; }
run e1 JMB_ROWS_EARLY=1
run q8 JMB_ROWS_EARLY=3
run q4 JMB_ROWS_EARLY=3 JMB_ROWS_G=4
run q16 JMB_ROWS_EARLY=3 JMB_ROWS_G=16
run vis2_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_vis2.so
run vis2_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_vis2.so
run ptb3_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_ptb3.so
run ptb3_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_ptb3.so
run t288_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_t288.so
run t288_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_t288.so
run t320_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_t320.so
run t320_q8 JMB_ROWS_EARLY=3 JMB_LIB=$L/libjmb200_t320.so
run t320_e0 JMB_ROWS_EARLY=0 JMB_LIB=$L/libjmb200_t320.so
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_e1_8 JMB_ROWS_EARLY=1 JMB_ROWS_G=8
run c3_q8 JMB_ROWS_EARLY=3 JMB_ROWS_G=8
run c3_q16 JMB_ROWS_EARLY=3 JMB_ROWS_G=16
run c3_q4 JMB_ROWS_EARLY=3 JMB_ROWS_G=4
run c3_t320_q8 JMB_ROWS_EARLY=3 JMB_ROWS_G=8 JMB_LIB=$L/libjmb200_t320.so
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=8
run c2_q8 JMB_ROWS_EARLY=3 JMB_ROWS_G=8
run c2_q4 JMB_ROWS_EARLY=3 JMB_ROWS_G=4
run c2_q16 JMB_ROWS_EARLY=3 JMB_ROWS_G=16
run c2_t320_q8 JMB_ROWS_EARLY=3 JMB_ROWS_G=8 JMB_LIB=$L/libjmb200_t320.so
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_p16v1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1 JMB_ROWS_G=16
run ms_q8 JMB_ROWS_EARLY=3 JMB_ROWS_G=8
run ms_q16 JMB_ROWS_EARLY=3 JMB_ROWS_G=16
run ms_e1_16 JMB_ROWS_EARLY=1 JMB_ROWS_G=16
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s6_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][7:-5].ljust(14), l["roofline"]["kernel"].ljust(46), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1][7:-5].ljust(14), "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
JMB_ROWS_EARLY=3 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s6_q8 $B > gpurun_out/r2b_ncu7.log 2>&1
ls -la gpurun_out/r2b_s6*.ncu-rep
