#!/bin/bash
# Round 2: multi-GPU checks on N GPUs of one box: sharded-chain parity tests, then the driver's scaling command
# (--steps 20 --warmup 5) at 1 and N ranks, with the last-CTA publish on and off.
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/r2_multi_pytest_N$N.log 2>&1
echo "2-GPU parity tests: exit $?"; tail -3 gpurun_out/r2_multi_pytest_N$N.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_N1.json 2> gpurun_out/r2_scale_N1.err
run() { # tag, nproc, env...
  local tag=$1 np=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline \
      > gpurun_out/r2_scale_$tag.json 2> gpurun_out/r2_scale_$tag.err
  echo "$tag exit $?"
}
run N${N}_pub $N JMB_PUBLISH=1
run N${N}_nopub $N JMB_PUBLISH=0
run N${N}_pub_b $N JMB_PUBLISH=1
if [ "$N" -gt 2 ]; then run N2_pub 2 JMB_PUBLISH=1; fi
if [ "$N" -gt 4 ]; then run N4_pub 4 JMB_PUBLISH=1; fi
env JMB_PUBLISH=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 1000 --warmup 10 --no-cpu-baseline \
      > gpurun_out/r2_scale_N${N}_1000.json 2> gpurun_out/r2_scale_N${N}_1000.err
python - <<'PY'
import json, glob
base = None
for f in sorted(glob.glob("gpurun_out/r2_scale_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        if f.endswith("N1.json"): base = l["value"]
        print(f.split("/")[-1], "n_gpus", l["n_gpus"], "value=%.1fM ms_per_step=%.4f e2e=%.1fM" % (l["value"] / 1e6, l["ms_per_step"], l["e2e"]["value"] / 1e6),
              "eff=%.3f" % (l["value"] / l["n_gpus"] / base) if base else "", l.get("sharded_parity"))
    except Exception as e:
        print(f, "FAILED", e)
PY
