#!/bin/bash
# Round 2, GPU call 6: row-strided kernel v2 (parameter pairs, cp.async ring, bulk-copied early stage).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py tests/test_pbc2_config1.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_v2_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2_v2_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { local tag=$1; shift; env "$@" $B > gpurun_out/r2_s5_$tag.json 2> gpurun_out/r2_s5_$tag.err; }
run tma JMB_KERNEL=tma
for G in 8 4 16; do run v2_G$G JMB_ROWS_G=$G; done
for G in 8 4; do
  run early0_G$G JMB_ROWS_G=$G JMB_LIB=jmbayes_b200/libjmb200_early0.so
  run ring3_G$G JMB_ROWS_G=$G JMB_LIB=jmbayes_b200/libjmb200_ring3.so
  run ne3_G$G JMB_ROWS_G=$G JMB_LIB=jmbayes_b200/libjmb200_ne3.so
done
for W in config3 config2; do for G in 8 4 16; do
  env JMB_ROWS_G=$G $B --workload $W > gpurun_out/r2_s5_${W}_G$G.json 2> gpurun_out/r2_s5_${W}_G$G.err
done; env JMB_KERNEL=tma $B --workload $W > gpurun_out/r2_s5_${W}_tma.json 2> gpurun_out/r2_s5_${W}_tma.err; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s5_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM frac=%.3f" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
JMB_ROWS_G=8 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_v2_G8 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2_ncu_v2_G8.log 2>&1
