#!/bin/bash
# A/B of library builds: scripts/ab_bench.sh "<lib1> <lib2> ..." "<workload1> ..." [steps]
for lib in $1; do
  for w in $2; do
    JMB_LIB=$lib python bench.py --workload $w --steps ${3:-500} --warmup 10 --no-cpu-baseline --chains 0 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', 'ms/step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'value', round(d['value']/1e6,1))"
  done
done
