#!/bin/bash
# Round 2: the driver's scaling command (--steps 20 --warmup 5) at N = 1, 2, 4, 8 on one box, plus a 1000-step run at N = 8.
set -u
mkdir -p gpurun_out
run() { # tag nproc extra-args...
  local tag=$1 np=$2; shift 2
  if [ "$np" -eq 1 ]; then python bench.py --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r2_scale8_$tag.json 2> gpurun_out/r2_scale8_$tag.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $np --no-cpu-baseline "$@" \
      > gpurun_out/r2_scale8_$tag.json 2> gpurun_out/r2_scale8_$tag.err; fi
  echo "$tag exit $?"
}
run N1 1
run N8 8 --steps 20 --warmup 5
run N2 2 --steps 20 --warmup 5
run N4 4 --steps 20 --warmup 5
run N8b 8 --steps 20 --warmup 5
run N8_1000 8 --steps 1000 --warmup 10
python - <<'PY'
import json, glob
base = None
for tag in ("N1", "N2", "N4", "N8", "N8b", "N8_1000"):
    f = "gpurun_out/r2_scale8_%s.json" % tag
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        if tag == "N1": base = l["value"]
        print(tag, "n_gpus", l["n_gpus"], "value=%.1fM ms_per_step=%.4f e2e=%.1fM multi=%.1fM" % (l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6, (l.get("multi_chain") or {}).get("value", 0) / 1e6),
              "eff=%.3f" % (l["value"] / l["n_gpus"] / base), l.get("sharded_parity"))
    except Exception as e:
        print(f, "FAILED", e)
PY
