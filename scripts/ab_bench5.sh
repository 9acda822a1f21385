#!/bin/bash
# A/B of library builds on the survfitJM workload: scripts/ab_bench5.sh "<lib1> <lib2> ..." [steps]
for lib in $1; do
  JMB_LIB=$lib python bench.py --workload config5 --steps ${2:-5} --warmup 1 --no-cpu-baseline 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', 'ms/pass', round(d['ms_per_step'],3), 'M subject-draws/s', round(d['value']/1e6,1))"
done
