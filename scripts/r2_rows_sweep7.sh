#!/bin/bash
# Round 2, GPU call 8: interval-row skipping; default bench line; ncu capture of the default config-4 kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider -x -k "not survfit" > gpurun_out/r2_skip_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2_skip_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
env JMB_KERNEL=tma $B > gpurun_out/r2_s7_tma.json 2> gpurun_out/r2_s7_tma.err
for G in 4 8 16; do for E in 0 1; do
  env JMB_ROWS_G=$G JMB_ROWS_EARLY=$E $B > gpurun_out/r2_s7_G${G}_e$E.json 2> gpurun_out/r2_s7_G${G}_e$E.err
done; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s7_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM frac=%.3f" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
python bench.py > gpurun_out/r2_bench_default_v2.json 2> gpurun_out/r2_bench_default_v2.err; echo "default bench exit $?"; tail -c 400 gpurun_out/r2_bench_default_v2.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_s20_v2.json 2> gpurun_out/r2_bench_s20_v2.err; echo "s20 bench exit $?"
ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_v3_default \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2_ncu_v3.log 2>&1
JMB_ROWS_G=16 JMB_ROWS_EARLY=1 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_v3_G16e \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2_ncu_v3b.log 2>&1
