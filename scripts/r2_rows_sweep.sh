#!/bin/bash
# Round 2, GPU call 2: parity of the row-strided step kernel on the device, then its group width / prefetch sweep on the
# config-4 shard against the round-1 TMA-staged kernel, and one full ncu capture.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_reference_golden.py tests/test_pbc2_config1.py tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider -x > gpurun_out/r2_rows_pytest.log 2>&1
echo "parity subset: exit $?"; tail -3 gpurun_out/r2_rows_pytest.log
B="python bench.py --steps 300 --warmup 10 --no-cpu-baseline --chains 0"
for G in 4 8 16 2; do
  JMB_ROWS_G=$G $B > gpurun_out/r2_sweep_c4_G$G.json 2> gpurun_out/r2_sweep_c4_G$G.err
done
JMB_KERNEL=tma $B > gpurun_out/r2_sweep_c4_tma.json 2> gpurun_out/r2_sweep_c4_tma.err
JMB_ROWS_G=4 JMB_ROWS_PREFETCH=0 $B > gpurun_out/r2_sweep_c4_G4_nopf.json 2> gpurun_out/r2_sweep_c4_G4_nopf.err
JMB_ROWS_G=8 JMB_ROWS_PREFETCH=0 $B > gpurun_out/r2_sweep_c4_G8_nopf.json 2> gpurun_out/r2_sweep_c4_G8_nopf.err
for G in 4 8; do
  JMB_ROWS_G=$G $B --workload config3 > gpurun_out/r2_sweep_c3_G$G.json 2> gpurun_out/r2_sweep_c3_G$G.err
  JMB_ROWS_G=$G $B --workload config2 > gpurun_out/r2_sweep_c2_G$G.json 2> gpurun_out/r2_sweep_c2_G$G.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_sweep_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], l["roofline"]["kernel"], "kernel_ms=%.4f ms_per_step=%.4f value=%.1fM" % (l["roofline"]["kernel_ms"], l["ms_per_step"], l["value"] / 1e6))
    except Exception as e:
        print(f, "FAILED", e)
PY
JMB_ROWS_G=4 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_rows_G4 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 > gpurun_out/r2_ncu_rows_G4.log 2>&1
JMB_ROWS_G=8 ncu --set full --clock-control none --import-source on -k regex:jmb_step_rows -s 6 -c 1 -o gpurun_out/r2_rows_G8 \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --chains 0 > gpurun_out/r2_ncu_rows_G8.log 2>&1
