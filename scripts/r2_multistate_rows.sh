#!/bin/bash
# Round 2, GPU call 10: specialised multi-state kernel (jmb_step_rows<SpecMultiState, G, EARLY>): tests, bench sweep, ncu.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multistate.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_ms_pytest.log 2>&1
echo "multistate + parity tests: exit $?"; tail -3 gpurun_out/r2_ms_pytest.log
B="python bench.py --workload multistate --steps 300 --warmup 10 --no-cpu-baseline"
env JMB_KERNEL=generic $B > gpurun_out/r2_ms_generic.json 2> gpurun_out/r2_ms_generic.err
for G in 4 8 16; do for E in 0 1; do
  env JMB_ROWS_G=$G JMB_ROWS_EARLY=$E $B > gpurun_out/r2_ms_G${G}_e$E.json 2> gpurun_out/r2_ms_G${G}_e$E.err
done; done
python bench.py --workload multistate > gpurun_out/r2_bench_multistate_rows.json 2> gpurun_out/r2_bench_multistate_rows.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ms_*.json")) + ["gpurun_out/r2_bench_multistate_rows.json"]:
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM e2e=%.1fM frac=%.3f" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6, l["e2e"]["value"] / 1e6, l["roofline"]["frac"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_multistate_rows_launches.csv \
    python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_rows_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 5 -c 1 -o gpurun_out/r2_multistate_rows \
    python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_rows_full.log 2>&1
