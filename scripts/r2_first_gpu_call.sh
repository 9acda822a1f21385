#!/bin/bash
# Round 2, GPU call 1: the whole GPU suite without any xfail marker, compute-sanitizer memcheck over every step-kernel
# variant, first numbers of the multi-state workload (bench line, launch list, one full ncu capture).
#   gpurun --timeout 1500 -- 'bash scripts/r2_first_gpu_call.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "gpu suite: exit $?" | tee -a gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python scripts/sanitize_chain.py > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck: exit $?" | tee -a gpurun_out/r2_memcheck.log
tail -8 gpurun_out/r2_memcheck.log
python bench.py --workload multistate > gpurun_out/r2_bench_multistate.json 2> gpurun_out/r2_bench_multistate.err
tail -c 1200 gpurun_out/r2_bench_multistate.json
if [ -s gpurun_out/r2_bench_multistate.json ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_multistate_launches.csv \
      python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_list.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:jmb_step_kernel -s 5 -c 1 -o gpurun_out/r2_multistate_step \
      python bench.py --workload multistate --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_ms_full.log 2>&1
fi
