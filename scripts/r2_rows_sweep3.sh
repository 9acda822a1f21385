#!/bin/bash
# Round 2, GPU call 4: load-hint / per-row prefetch / exp3 A-B builds of the row-strided kernel, TMA kernel as the calibration.
set -u
mkdir -p gpurun_out
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { # tag, env...
  local tag=$1; shift
  env "$@" $B > gpurun_out/r2_s3_$tag.json 2> gpurun_out/r2_s3_$tag.err
}
run tma JMB_KERNEL=tma
for G in 4 8; do
  run base_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0
  run noexp3_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0 JMB_LIB=jmbayes_b200/libjmb200_noexp3.so
  run ldhint_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0 JMB_LIB=jmbayes_b200/libjmb200_ldhint.so
  run pfrow2_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0 JMB_LIB=jmbayes_b200/libjmb200_pfrow2.so
  run ldpf_G${G}_pf0 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=0 JMB_LIB=jmbayes_b200/libjmb200_ldpf.so
  run ldhint_G${G}_pf2 JMB_ROWS_G=$G JMB_ROWS_PREFETCH=2 JMB_LIB=jmbayes_b200/libjmb200_ldhint.so
done
run tma2 JMB_KERNEL=tma
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s3_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f, "FAILED", e)
PY
