#!/bin/bash
# First gpurun call of the next round: everything that was written after round 1's GPU minutes were spent gets its first
# run on a B200 (multi-state path, accuracy sums), then the measurements of the multi-state workload.
#   gpurun --timeout 1500 -- 'bash scripts/r2_first_gpu_call.sh'
# Outputs land in gpurun_out/ (copy what should be judged into profiles/).
set -u
mkdir -p gpurun_out
python -m pytest tests/test_zz_gpu_multistate.py tests/test_zz_gpu_accuracy.py -q --runxfail -x > gpurun_out/r2_new_paths_pytest.log 2>&1
echo "new paths: exit $?" | tee -a gpurun_out/r2_new_paths_pytest.log
python -m pytest tests -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu suite: exit $?" | tee -a gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
python bench.py --workload multistate > gpurun_out/r2_bench_multistate.json 2> gpurun_out/r2_bench_multistate.err
tail -c 1500 gpurun_out/r2_bench_multistate.json
# launch list + one full capture of the multi-state step kernel (only after the plain run exited 0)
if [ -s gpurun_out/r2_bench_multistate.json ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_multistate_launches.csv \
      python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_list.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:jmb_step_kernel -s 5 -c 1 -o gpurun_out/r2_multistate_step \
      python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_full.log 2>&1
fi
