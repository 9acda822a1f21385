"""Two-GPU run of the sharded CUDA path (NCCL all-reduce of the per-iteration sums inside the
library): must reproduce the single-GPU chain.  Skipped on boxes with fewer than two GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, cfg, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    from jmbayes_b200 import api
    from jmbayes_b200.cohorts import make_cohort
    from jmbayes_b200.sharding import shard_cohort
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_cohort(cfg, n=n)
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=10)
    sh = shard_cohort(c, rank, world)
    with api.Fit(sh["Data"], sh["Covs"]["b"], c["interval_cens"], device=rank, subject_offset=sh["subject_offset"]) as fit:
        api.connect_ranks(fit, world, rank, dist)
        fit.set_draw(sh["Data"])
        r = fit.lap_rwm_C(sh["inits"], sh["priors"], sh["scales"], sh["Covs"], ctl)
    parts = [None] * world
    dist.all_gather_object(parts, (sh["subject_offset"], r["mcmc"]["b"], r["mcmc"]["alphas"]))
    if rank == 0:
        full = api.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], device=0)
        b = np.concatenate([p[1] for p in sorted(parts, key=lambda x: x[0])], axis=0)
        ok = (np.allclose(b, full["mcmc"]["b"], rtol=1e-9, atol=1e-12)
              and np.allclose(r["mcmc"]["alphas"], full["mcmc"]["alphas"], rtol=1e-9)
              and np.array_equal(parts[0][2], parts[1][2])          # identical decisions on both ranks
              and np.allclose(r["logWeights"], full["logWeights"], rtol=1e-11)
              and np.allclose(r["extras"]["trace_surv"], full["extras"]["trace_surv"], rtol=1e-11))
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", ["config3", "config4"])
def test_two_gpu_sharded_chain_equals_single_gpu(cuda_lib, cfg):
    if cuda_lib.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, 1000, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True
