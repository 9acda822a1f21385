#!/bin/bash
# Round 2 (second session), GPU call 18: last build - parity subset (every kernel variant), the 1M-subject cohort on one GPU, the driver's command.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_multistate.py tests/test_reference_golden.py tests/test_gpu_rcpp_shim.py tests/test_gpu_fullsize.py -q -m gpu -p no:cacheprovider > gpurun_out/r2b_s18_pytest.log 2>&1
echo "parity: exit $?"; tail -3 gpurun_out/r2b_s18_pytest.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_s18_bench_s20.json 2> gpurun_out/r2b_s18_bench_s20.err; echo "s20 exit $?"
python bench.py --n-per-gpu 1000000 --steps 100 --warmup 5 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2b_s18_bench_1m.json 2> gpurun_out/r2b_s18_bench_1m.err; echo "1M bench exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s18_bench_*.json")):
    l = json.loads(open(f).read().strip().splitlines()[-1]); r = l["roofline"]
    print(f.split("/")[-1][8:-5].ljust(12), r["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f" % (r["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, r["frac"]), l["clocks"])
PY
