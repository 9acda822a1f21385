#!/bin/bash
# Round 2 (second session), GPU call 2: paired 16-byte ring copies (L1-free), thread-count variants, L1 carve-out test, ncu.
set -u
mkdir -p gpurun_out
export JMB_ROWS_DEBUG=1
L=jmbayes_b200
JMB_LIB=$L/libjmb200_pair.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s2_pytest_pair.log 2>&1
echo "parity (pair): exit $?"; tail -3 gpurun_out/r2b_s2_pytest_pair.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s2_$tag.json 2> gpurun_out/r2b_s2_$tag.err; }
run base_e0 JMB_X=0
run base_e0_co100 JMB_ROWS_CARVEOUT=100
run base_p8v1 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run base_p8v1_co100 JMB_ROWS_EARLY=2 JMB_ROWS_VT=1 JMB_ROWS_CARVEOUT=100
run noreread_p8v1 JMB_LIB=$L/libjmb200_noreread.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run pair_e0 JMB_LIB=$L/libjmb200_pair.so
run pair_e0_co100 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_CARVEOUT=100
run pair_p8v1 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run pair_p8v2 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run pair_p8v0 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=0
run pair_p4v2 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=2 JMB_ROWS_G=4
run pair_e1 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=1
run pair224_e0 JMB_LIB=$L/libjmb200_pair224.so
run pair224_p8v1 JMB_LIB=$L/libjmb200_pair224.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run pair224_p8v2 JMB_LIB=$L/libjmb200_pair224.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run pair160x3_e0 JMB_LIB=$L/libjmb200_pair160x3.so
run pair160x3_p8v1 JMB_LIB=$L/libjmb200_pair160x3.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run t224_e0 JMB_LIB=$L/libjmb200_t224.so
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_pair_def JMB_LIB=$L/libjmb200_pair.so
run c3_pair_p16v1 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
run c3_pair_p16v2 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
run c3_pair_e0_8 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=0 JMB_ROWS_G=8
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_pair_def JMB_LIB=$L/libjmb200_pair.so
run c2_pair_p4v2 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=2
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_pair_def JMB_LIB=$L/libjmb200_pair.so
run ms_pair_p8v1 JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s2_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f, "FAILED", e, open(f.replace(".json", ".err")).read()[-300:])
PY
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
JMB_LIB=$L/libjmb200_pair.so JMB_ROWS_EARLY=2 JMB_ROWS_VT=1 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_pair_p8v1 $B > gpurun_out/r2b_ncu1.log 2>&1
JMB_LIB=$L/libjmb200_pair.so ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_pair_e0 $B > gpurun_out/r2b_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
