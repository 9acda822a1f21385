#!/bin/bash
# Round 2, GPU call 5: cp.async lane-private ring for the hazard rows (2 / 3 stages), with and without the joint exponentials.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { local tag=$1; shift; env "$@" $B > gpurun_out/r2_s4_$tag.json 2> gpurun_out/r2_s4_$tag.err; }
JMB_LIB=jmbayes_b200/libjmb200_ring3.so python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -x -k "static_and_generic or chain_matches" > gpurun_out/r2_ring_pytest.log 2>&1
echo "ring parity: exit $?"; tail -3 gpurun_out/r2_ring_pytest.log
run tma JMB_KERNEL=tma
for G in 4 8; do
  for V in ring3 ring2 ring3ne ring2ne ring3ld; do
    run ${V}_G${G} JMB_ROWS_G=$G JMB_LIB=jmbayes_b200/libjmb200_$V.so
  done
  run ring3_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0 JMB_LIB=jmbayes_b200/libjmb200_ring3.so
done
run ring3_G16 JMB_ROWS_G=16 JMB_LIB=jmbayes_b200/libjmb200_ring3.so
run ring3_G2 JMB_ROWS_G=2 JMB_LIB=jmbayes_b200/libjmb200_ring3.so
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s4_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f, "FAILED", e)
PY
JMB_ROWS_G=4 JMB_LIB=jmbayes_b200/libjmb200_ring3.so ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_ring3_G4 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2_ncu_ring3_G4.log 2>&1
