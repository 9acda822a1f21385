#!/bin/bash
# Round 2 (second session), GPU call 12 (N GPUs): stamped mailbox words; 2-GPU parity tests; the driver's scaling command at 1 and N ranks.
set -u
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -m gpu -p no:cacheprovider > gpurun_out/r2b_s12_multi_pytest_N$N.log 2>&1
echo "2-GPU parity tests: exit $?"; tail -3 gpurun_out/r2b_s12_multi_pytest_N$N.log
run() { # tag, nproc, env...
  local tag=$1 np=$2; shift 2
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline \
      > gpurun_out/r2b_s12_scale_$tag.json 2> gpurun_out/r2b_s12_scale_$tag.err
  echo "$tag exit $?"
}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_s12_scale_N1.json 2> gpurun_out/r2b_s12_scale_N1.err
run N${N} $N JMB_X=0
run N${N}_pub $N JMB_PUBLISH=1
run N${N}_b $N JMB_X=0
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_s12_scale_N1_b.json 2> gpurun_out/r2b_s12_scale_N1_b.err
if [ "$N" -gt 2 ]; then run N2 2 JMB_X=0; fi
if [ "$N" -gt 4 ]; then run N4 4 JMB_X=0; fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_s12_*.json")):
    try:
        l = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1][8:-5].ljust(16), "n_gpus", l["n_gpus"], l["roofline"]["kernel"], "value=%.1fM ms_per_step=%.4f kernel_ms=%.4f e2e=%.1fM" % (l["value"] / 1e6, l["ms_per_step"], l["roofline"]["kernel_ms"], l["e2e"]["value"] / 1e6), (l.get("sharded_parity") or {}).get("max_rel_vs_single_rank"), (l.get("sharded_parity") or {}).get("identical_accepts"))
    except Exception as e:
        print(f, "FAILED", e)
PY
