#!/bin/bash
# A/B of an environment switch on the default workload(s): scripts/ab_env.sh "VAR=a VAR=b ..." "<workloads>" [steps]
for kv in $1; do
  for w in $2; do
    env $kv python bench.py --workload $w --steps ${3:-1000} --warmup 10 --no-cpu-baseline --chains 4 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$kv', '$w', 'ms/step', round(d['ms_per_step'],5), 'kernel_ms', round(d['roofline']['kernel_ms'],5), 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), 'multi', round(d['multi_chain']['value']/1e6,1))"
  done
done
