#!/bin/bash
# Round 2, GPU call 7: (shape, group width, early stage) table of the row-strided kernel; compiled Rcpp glue; full GPU suite.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rcpp_shim.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_shim_pytest.log 2>&1
echo "rcpp glue: exit $?"; tail -4 gpurun_out/r2_shim_pytest.log
python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_rcpp_shim.py > gpurun_out/r2_pytest_gpu2.log 2>&1
echo "gpu suite: exit $?"; tail -4 gpurun_out/r2_pytest_gpu2.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
for W in config4_shard config3 config2; do
  env JMB_KERNEL=tma $B --workload $W > gpurun_out/r2_s6_${W}_tma.json 2> gpurun_out/r2_s6_${W}_tma.err
  for G in 4 8 16; do for E in 0 1; do
    env JMB_ROWS_G=$G JMB_ROWS_EARLY=$E $B --workload $W > gpurun_out/r2_s6_${W}_G${G}_e$E.json 2> gpurun_out/r2_s6_${W}_G${G}_e$E.err
  done; done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_s6_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM frac=%.3f" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
