#!/bin/bash
# Round 2 (second session), GPU call 8: split-polynomial exponentials, LDS visit path, slimmer finalize kernel; launch list.
set -u
mkdir -p gpurun_out
export JMB_ROWS_DEBUG=1
L=jmbayes_b200
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2b_s8_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s8_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s8_$tag.json 2> gpurun_out/r2b_s8_$tag.err; }
run def JMB_X=0
run e1 JMB_ROWS_EARLY=1
run nosplit JMB_LIB=$L/libjmb200_nosplit.so
run nopdl JMB_PDL=0
B="python bench.py --workload config3 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c3_def JMB_X=0
B="python bench.py --workload config2 --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run c2_def JMB_X=0
run c2_q8 JMB_ROWS_EARLY=3
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run ms_def JMB_X=0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s8_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][7:-5].ljust(14), l["roofline"]["kernel"].ljust(46), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f.split("/")[-1][7:-5].ljust(14), "FAILED", open(f.replace(".json", ".err")).read().strip().splitlines()[-1][:150])
PY
python bench.py --steps 120 --warmup 5 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2b_s8_pre_ncu.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_s8_launches.csv python bench.py --steps 120 --warmup 5 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2b_ncu10.log 2>&1
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s8_def $B > gpurun_out/r2b_ncu11.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jmb_finalize -s 6 -c 1 -o gpurun_out/r2b_s8_fin $B > gpurun_out/r2b_ncu12.log 2>&1
ls -la gpurun_out/r2b_s8*.ncu-rep
