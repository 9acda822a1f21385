#!/bin/bash
# Round 2 (second session), GPU call 17 (call 16 again: its output exceeded what gpurun copies back): final build - full GPU suite, smoke, bench lines of every MCMC workload, ncu captures.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2b_s17_pytest.log 2>&1
echo "gpu suite: exit $?"; tail -3 gpurun_out/r2b_s17_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_s17_smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/r2b_s17_smoke.log
python bench.py > gpurun_out/r2b_s17_bench_default.json 2> gpurun_out/r2b_s17_bench_default.err; echo "default bench exit $?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_s17_bench_s20.json 2> gpurun_out/r2b_s17_bench_s20.err; echo "s20 bench exit $?"
for w in config3 config2 multistate; do
  python bench.py --workload $w --no-cpu-baseline > gpurun_out/r2b_s17_bench_$w.json 2> gpurun_out/r2b_s17_bench_$w.err; echo "$w exit $?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s17_bench_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        r = l.get("roofline", {})
        print(f.split("/")[-1][8:-5].ljust(18), str(r.get("kernel")).ljust(44), "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f multi=%.1fM" % (r.get("kernel_ms", 0), l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, r.get("frac", 0), (l.get("multi_chain") or {}).get("value", 0) / 1e6))
    except Exception as e:
        print(f.split("/")[-1], "FAILED", e)
PY
B="--steps 5 --warmup 3 --no-cpu-baseline --chains 0 --no-parity"
ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2b_s17_c4 python bench.py $B > gpurun_out/r2b_ncu17a.log 2>&1
for w in config3 config2 multistate; do
  ncu --set full --clock-control none -k regex:jmb_step_rows -s 6 -c 1 -o /tmp/r2b_s17_$w python bench.py --workload $w $B > gpurun_out/r2b_ncu17_$w.log 2>&1
  ncu -i /tmp/r2b_s17_$w.ncu-rep --page raw --csv > gpurun_out/r2b_s17_${w}_raw.csv 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_s17_launches.csv python bench.py --steps 100 --warmup 5 --no-cpu-baseline --chains 0 --no-parity > gpurun_out/r2b_ncu17b.log 2>&1
ls -la gpurun_out/r2b_s17*.ncu-rep
