#!/bin/bash
# full GPU suite + default bench lines of the current build
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu3.log 2>&1
echo "gpu suite: exit $?"; tail -5 gpurun_out/r2_pytest_gpu3.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2_smoke.log
python bench.py --workload multistate > gpurun_out/r2_bench_multistate_v3.json 2> gpurun_out/r2_bench_multistate_v3.err; echo "ms bench $?"
python bench.py --workload config3 --no-cpu-baseline > gpurun_out/r2_bench_config3.json 2> gpurun_out/r2_bench_config3.err; echo "c3 bench $?"
python bench.py --workload config2 --no-cpu-baseline > gpurun_out/r2_bench_config2.json 2> gpurun_out/r2_bench_config2.err; echo "c2 bench $?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref arm $?"
python - <<'PY'
import json, glob
for f in ["gpurun_out/r2_bench_multistate_v3.json", "gpurun_out/r2_bench_config3.json", "gpurun_out/r2_bench_config2.json", "gpurun_out/r2_bench_reference.json"]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l.get("roofline", {}).get("kernel"), "value=%.2fM ms=%.4f e2e=%.2fM" % (l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6), l.get("cpu_baseline", {}).get("cores"))
    except Exception as e:
        print(f, "FAILED", e)
PY
