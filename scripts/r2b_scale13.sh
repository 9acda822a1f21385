#!/bin/bash
# Round 2 (second session), GPU call 13 (8 GPUs): the driver's scaling command at N = 1, 2, 4, 8.
set -u
mkdir -p gpurun_out
run() { # tag, nproc
  local tag=$1 np=$2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline \
      > gpurun_out/r2b_s13_scale_$tag.json 2> gpurun_out/r2b_s13_scale_$tag.err
  echo "$tag exit $?"
}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_s13_scale_N1.json 2> gpurun_out/r2b_s13_scale_N1.err
run N8 8
run N4 4
run N2 2
python - <<'PY'
import json, glob
base = json.loads(open("gpurun_out/r2b_s13_scale_N1.json").read().strip().splitlines()[-1])["value"]
for f in sorted(glob.glob("gpurun_out/r2b_s13_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(12), "n_gpus", l["n_gpus"], "value=%.1fM ms_per_step=%.4f kernel_ms=%.4f e2e=%.1fM eff=%.4f multi=%.1fM" % (l["value"] / 1e6, l["ms_per_step"], l["roofline"]["kernel_ms"], l["e2e"]["value"] / 1e6, l["value"] / base / l["n_gpus"], (l.get("multi_chain") or {}).get("value", 0) / 1e6), (l.get("sharded_parity") or {}).get("max_rel_vs_single_rank"), (l.get("sharded_parity") or {}).get("identical_accepts"))
    except Exception as e:
        print(f, "FAILED", e)
PY
