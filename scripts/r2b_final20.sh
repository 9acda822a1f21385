#!/bin/bash
# Round 2 (second session), GPU call 20: committed end state - full GPU suite, smoke, the driver's command.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2b_s20_pytest.log 2>&1
echo "gpu suite: exit $?"; tail -3 gpurun_out/r2b_s20_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_s20_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2b_s20_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_s20_bench_s20.json 2> gpurun_out/r2b_s20_bench_s20.err; echo "s20 exit $?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2b_s20_bench_s20.json").read().strip().splitlines()[-1]); r = l["roofline"]
print(r["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f" % (r["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, r["frac"]), l["clocks"], l["cpu_baseline"]["value"])
PY
