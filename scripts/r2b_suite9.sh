#!/bin/bash
# Round 2 (second session), GPU call 9: full GPU suite, smoke, E1 / E3 A/B, default bench lines, 1M subjects on one GPU, reference arm.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2b_s9_pytest.log 2>&1
echo "gpu suite: exit $?"; tail -4 gpurun_out/r2b_s9_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_s9_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2b_s9_smoke.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0 --no-parity"
run() { tag=$1; shift; env "$@" $B > gpurun_out/r2b_s9_$tag.json 2> gpurun_out/r2b_s9_$tag.err; }
for i in 1 2 3; do run e1_$i JMB_ROWS_EARLY=1; run q8_$i JMB_ROWS_EARLY=3; done
python bench.py > gpurun_out/r2b_s9_bench_default.json 2> gpurun_out/r2b_s9_bench_default.err; echo "default bench exit $?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_s9_bench_s20.json 2> gpurun_out/r2b_s9_bench_s20.err; echo "s20 bench exit $?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2b_s9_bench_reference.json 2> gpurun_out/r2b_s9_bench_reference.err; echo "reference arm exit $?"
python bench.py --n-per-gpu 1000000 --steps 100 --warmup 5 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2b_s9_bench_1m.json 2> gpurun_out/r2b_s9_bench_1m.err; echo "1M bench exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s9_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        r = l.get("roofline", {})
        print(f.split("/")[-1][7:-5].ljust(16), str(r.get("kernel")).ljust(34), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f" % (r.get("kernel_ms", 0), l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, r.get("frac", 0)))
    except Exception as e:
        print(f.split("/")[-1][7:-5].ljust(16), "FAILED", e)
PY
