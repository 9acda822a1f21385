"""world_size-2 run of the sharded path on CPU (gloo): host-side logic of the multi-GPU design.

Each rank runs the oracle chain on its contiguous block of subjects and all-reduces the three
per-iteration sums, exactly what the CUDA library does with ncclAllReduce; the merged result must
reproduce the single-process chain (same counter-RNG keys through ``subject_offset``).
"""
import os
import socket
import sys

import numpy as np
import pytest

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
    from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort
    from jmbayes_b200.sharding import shard_cohort
    from oracle import oracle
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_multistate_cohort(n=n) if cfg == "multistate" else make_cohort(cfg, n=n)
    ms = bool(c["multiState"])
    ctl = dict(c["control"]); ctl.update(n_burnin=15, n_iter=10, n_block=5, n_thin=5)
    sh = shard_cohort(c, rank, world)

    def allreduce(buf):
        t = torch.from_numpy(buf.copy())
        dist.all_reduce(t)
        buf[:] = t.numpy()

    oracle.set_rowsum_mode(1)
    r = oracle.lap_rwm_C(sh["inits"], sh["Data"], sh["priors"], sh["scales"], sh["Covs"], ctl, c["interval_cens"], ms,
                         subject_offset=sh["subject_offset"], allreduce=allreduce)
    # gather the shards' b draws on rank 0
    parts = [None] * world
    dist.all_gather_object(parts, (sh["subject_offset"], r["mcmc"]["b"], r["scales"]["sigma"]["b"]))
    if rank == 0:
        full = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], ms)
        b = np.concatenate([p[1] for p in sorted(parts, key=lambda x: x[0])], axis=0)
        sg = np.concatenate([p[2] for p in sorted(parts, key=lambda x: x[0])])
        ok = (np.allclose(b, full["mcmc"]["b"], rtol=1e-9, atol=1e-12)
              and np.allclose(sg, full["scales"]["sigma"]["b"], rtol=1e-10)
              and np.allclose(r["mcmc"]["alphas"], full["mcmc"]["alphas"], rtol=1e-9)
              and np.allclose(r["mcmc"]["Bs_gammas"], full["mcmc"]["Bs_gammas"], rtol=1e-9, atol=1e-12)
              and np.allclose(r["mcmc"]["tau_Bs_gammas"], full["mcmc"]["tau_Bs_gammas"], rtol=1e-9)
              and np.allclose(r["logWeights"], full["logWeights"], rtol=1e-11)
              and np.allclose(r["extras"]["trace_surv"], full["extras"]["trace_surv"], rtol=1e-11))
        ret.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("cfg", ["config3", "config4", "multistate"])
def test_two_rank_sharded_chain_equals_single_process(cfg):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cfg, 260, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True


def test_shard_bounds_cover_and_balance():
    sys.path.insert(0, ROOT)
    from jmbayes_b200.sharding import shard_bounds
    rng = np.random.default_rng(0)
    w = 1 + rng.poisson(9, 1000)
    for world in (1, 2, 3, 8):
        b = shard_bounds(1000, world, w)
        assert b[0] == 0 and b[-1] == 1000 and np.all(np.diff(b) > 0)
        loads = np.add.reduceat(w, b[:-1])
        assert loads.max() / loads.mean() < 1.05
    assert list(shard_bounds(3, 3)) == [0, 1, 2, 3]


# ---- survfitJM (config 5): subjects shard over ranks, no collective on the data path -----------------------------
def _svft_worker(rank, world, port, n, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from jmbayes_b200 import survfit as sv
    from jmbayes_b200.cohorts import make_cohort
    from jmbayes_b200.sharding import shard_bounds
    from oracle import oracle_svft
    dist.init_process_group("gloo", rank=rank, world_size=world)
    c = make_cohort("config3", n=n)
    means, draws = sv.posterior_draws(c, M=12)
    b = shard_bounds(n, world)
    sel = np.arange(int(b[rank]), int(b[rank + 1]))
    d = sv.newdata_designs(c, L=3, subjects=sel)
    r = oracle_svft.survfitJM(d, means, draws, subject_offset=int(b[rank]))
    parts = [None] * world
    dist.all_gather_object(parts, (int(b[rank]), r["summaries"], r["modes"]))      # a gather of results, nothing else
    if rank == 0:
        full = oracle_svft.survfitJM(sv.newdata_designs(c, L=3), means, draws)
        parts = sorted(parts, key=lambda x: x[0])
        S = np.concatenate([p[1] for p in parts], axis=2)
        m = np.concatenate([p[2] for p in parts], axis=0)
        ret.put(bool(np.array_equal(S, full["summaries"]) and np.array_equal(m, full["modes"])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_survfit_equals_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_svft_worker, args=(r, 2, port, 30, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) is True
