#!/bin/bash
# Round 2, GPU call 3: split-phase L2 prefetch + joint exponentials in the row-strided kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_rows2_pytest.log 2>&1
echo "parity subset: exit $?"; tail -3 gpurun_out/r2_rows2_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
for G in 4 8; do for PF in 2 0; do
  JMB_ROWS_G=$G JMB_ROWS_PREFETCH=$PF $B > gpurun_out/r2_s2_c4_G${G}_pf$PF.json 2> gpurun_out/r2_s2_c4_G${G}_pf$PF.err
done; done
JMB_ROWS_G=16 $B > gpurun_out/r2_s2_c4_G16_pf2.json 2> gpurun_out/r2_s2_c4_G16_pf2.err
JMB_ROWS_G=2 $B > gpurun_out/r2_s2_c4_G2_pf2.json 2> gpurun_out/r2_s2_c4_G2_pf2.err
for G in 4 8; do
  JMB_ROWS_G=$G $B --workload config3 > gpurun_out/r2_s2_c3_G$G.json 2> gpurun_out/r2_s2_c3_G$G.err
  JMB_ROWS_G=$G $B --workload config2 > gpurun_out/r2_s2_c2_G$G.json 2> gpurun_out/r2_s2_c2_G$G.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s2_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, l["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
python bench.py > gpurun_out/r2_bench_default_v1.json 2> gpurun_out/r2_bench_default_v1.err; echo "default bench exit $?"; tail -c 600 gpurun_out/r2_bench_default_v1.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_s20_v1.json 2> gpurun_out/r2_bench_s20_v1.err; echo "s20 bench exit $?"
JMB_ROWS_G=4 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_rows2_G4 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2_ncu_rows2_G4.log 2>&1
