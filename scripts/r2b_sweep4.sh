#!/bin/bash
# Round 2 (second session), GPU call 4: slimmed kernels (cold code out of line, one copy of the censoring code), variants.
set -u
mkdir -p gpurun_out
export JMB_ROWS_DEBUG=1
L=jmbayes_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s4_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s4_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s4_$tag.json 2> gpurun_out/r2b_s4_$tag.err; }
run e0 JMB_ROWS_EARLY=0
run e1 JMB_ROWS_EARLY=1
run e1_g4 JMB_ROWS_EARLY=1 JMB_ROWS_G=4
run e1_g16 JMB_ROWS_EARLY=1 JMB_ROWS_G=16
run p8v1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run p8v4 JMB_ROWS_EARLY=2 JMB_ROWS_VT=4
run p4v4 JMB_ROWS_EARLY=2 JMB_ROWS_VT=4 JMB_ROWS_G=4
run p16v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=16
run t224_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_t224.so
run t224_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_LIB=$L/libjmb200_t224.so
run t192_e1 JMB_ROWS_EARLY=1 JMB_LIB=$L/libjmb200_t192.so
run t192_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_LIB=$L/libjmb200_t192.so
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_e1_16 JMB_ROWS_EARLY=1
run c3_e1_8 JMB_ROWS_EARLY=1 JMB_ROWS_G=8
run c3_p16v1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run c3_p16v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run c3_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=8
run c3_p8v4 JMB_ROWS_EARLY=2 JMB_ROWS_VT=4 JMB_ROWS_G=8
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_e0_4 JMB_ROWS_EARLY=0
run c2_e1_8 JMB_ROWS_EARLY=1 JMB_ROWS_G=8
run c2_p4v4 JMB_ROWS_EARLY=2 JMB_ROWS_VT=4
run c2_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=8
run c2_p8v4 JMB_ROWS_EARLY=2 JMB_ROWS_VT=4 JMB_ROWS_G=8
run c2_p16v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=16
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_e0_8 JMB_ROWS_EARLY=0
run ms_e1_8 JMB_ROWS_EARLY=1
run ms_p8v2 JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run ms_p16v1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1 JMB_ROWS_G=16
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s4_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][7:-5].ljust(14), l["roofline"]["kernel"].ljust(46), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1][7:-5].ljust(14), "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
JMB_ROWS_EARLY=1 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s4_e1 $B > gpurun_out/r2b_ncu5.log 2>&1
JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s4_p8v2 $B > gpurun_out/r2b_ncu6.log 2>&1
ls -la gpurun_out/r2b_s4*.ncu-rep
