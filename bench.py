#!/usr/bin/env python
"""bench.py - subject-iterations/s of the mvJointModelBayes MCMC hot path (lap_rwm_C loop).

One "step" = one MCMC iteration over the whole cohort (every subject's b_i update, the
survival-block step, the Gibbs draws and the adaptation): n subject-iterations.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config4_shard|config3|config2|config5|summaries|multistate]
  python bench.py --impl reference ...      # restated reference CPU path on the host cores

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how each field is obtained.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "subject-iterations/sec of mvJointModelBayes MCMC"
UNIT = "subject-iterations/s"
# algorithmic bytes per subject-iteration (compact fp64 layout, one chain)
ALG_BYTES = {"config2": 2588.0, "config3": 6892.0, "config4": 12128.0}
WORKLOADS = {
    # name: (cohort config, subjects per GPU)
    "config4_shard": ("config4", 125000),   # 1M-subject config 4 split over 8 GPUs: 125k per GPU
    "config3": ("config3", 100000),
    "config2": ("config2", 100000),
    # survfitJM dynamic predictions (BASELINE config 5): M = 200 draws x L = 10 horizons per subject; one "step" =
    # the whole per-subject loop (mode search + MH over the draws + horizon integrals + summaries) for every subject
    "config5": ("config4", 25000),
    # posterior summaries of the random effects of the config-4 shard: 125 000 subjects x q = 6 series of
    # M = 500 kept draws, the ten statistics of mvJointModelBayes()$statistics per series; one "step" = all series
    "summaries": ("config4", 125000),
    # multi-state fit (multiState = TRUE): illness-death cohort, 2-3 transition rows per subject, 3 strata
    "multistate": ("multistate", 50000),
}
STATS_M, STATS_Q = 500, 6
STATS_METRIC = "series/sec of mvJointModelBayes posterior summaries (ten statistics per series of M=500 kept draws)"
STATS_UNIT = "series/s"
SVFT_M, SVFT_L = 200, 10
SVFT_METRIC = "subject-draws/sec of survfitJM dynamic predictions (M=200 draws x L=10 horizons per subject)"
SVFT_UNIT = "subject-draws/s"
CPU_SAMPLE_N = 20000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="config4_shard", choices=sorted(WORKLOADS))
    ap.add_argument("--n-per-gpu", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the sharded-chain parity check that precedes the timing")
    ap.add_argument("--chains", type=int, default=4, help="chains in flight for the extra multi_chain measurement (0: skip)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for j, nm in enumerate(names) if any(len(r) >= 6 and r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py executes oracle/): restated reference chain on host cores
# ------------------------------------------------------------------------------------------------
def _ref_kind():
    """"reference": oracle/_ref (the reference's own lap_rwm_C source compiled against the stand-in
    Armadillo/Rcpp header) when it is built or buildable; else "port": the C restatement."""
    try:
        from oracle import ref
        return "reference" if ref.build() else "port"
    except Exception:
        return "port"


def _cpu_worker(args):
    cfg, n, iters, chain_id, kind = args
    from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort
    c = make_multistate_cohort(n=n) if cfg == "multistate" else make_cohort(cfg, n=n)
    ms = bool(c.get("multiState", False))
    ctl = dict(c["control"])
    if kind == "reference":
        from oracle import ref
        nb = ctl["n_block"]
        iters = max(nb, (iters // nb) * nb)          # the reference runs whole blocks
        ctl.update(n_burnin=iters - nb, n_iter=nb, n_thin=nb)
        ref.lib()
        t0 = time.perf_counter()
        ref.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], ms, chain_id=chain_id)
        return time.perf_counter() - t0, iters
    from oracle import oracle
    oracle.set_rowsum_mode(0)        # the reference's own cumsum-difference rowsum
    t0 = time.perf_counter()
    oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], ms,
                     chain_id=chain_id, max_iters=iters)
    return time.perf_counter() - t0, iters


def cpu_reference(cfg: str, n: int, iters: int, workers: int, kind: str):
    """Independent chains, one per worker, like the reference's foreach over n_cores
    (R/mvJointModelBayes.R:860-905).  Returns (subject-iterations/s over all workers, seconds, iterations)."""
    import multiprocessing as mp
    if workers <= 1:
        dt, it = _cpu_worker((cfg, n, iters, 0, kind))
        return n * it / dt, dt, it
    ctx = mp.get_context("fork")
    # each worker times its own chain (cohort generation excluded); the slowest worker bounds the job
    with ctx.Pool(workers) as pool:
        res = pool.map(_cpu_worker, [(cfg, n, iters, w, kind) for w in range(workers)])
    dt = max(r[0] for r in res)
    it = res[0][1]
    return workers * n * it / dt, dt, it


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    workers = max(1, cores - 1)          # n_cores = detectCores() - 1 (R/mvJointModelBayes.R:14)
    n = CPU_SAMPLE_N
    from oracle import oracle
    oracle.build()
    kind = _ref_kind()
    # size the sample so that the whole run finishes within a few minutes: ~20 s of chain per worker
    t_probe, it_probe = _cpu_worker((cfg, n, 50, 0, kind))
    per_iter = t_probe / it_probe
    iters_total = max(50, int(20.0 / max(per_iter, 1e-4)))
    val, dt, iters_total = cpu_reference(cfg, n, iters_total, workers, kind)
    what = ("the reference's own lap_rwm_C source (src/fast_LAPRWM.cpp) compiled against the stand-in Armadillo/Rcpp header"
            if kind == "reference" else "restated lap_rwm_C (C port)")
    sample = f"{workers} independent chains x {iters_total} iterations x {n} subjects of {cfg}; {what}; g++ -O2"
    port = None
    if kind == "reference":   # the stand-in linear algebra is unoptimised: also time the plain-C restatement
        pv, pdt, pit = cpu_reference(cfg, n, iters_total, workers, "port")
        port = {"value": pv, "unit": UNIT, "cores": workers, "kind": "port",
                "sample": f"{workers} chains x {pit} iterations x {n} subjects; restated lap_rwm_C, gcc -O3 -march=native"}
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / iters_total, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "cohort": cfg, "subjects": n, "note": "reference CPU path on the host cores, one chain per worker (R absent from the image)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if port:
        line["cpu_baseline_port"] = port
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# CUDA arm
# ------------------------------------------------------------------------------------------------
def _sharded_parity(api, dist, world, rank, local, n=2000, iters=40):
    """Correctness of the N-rank path, checked on the device before anything is timed: a 2 000-subject config-4 cohort
    sharded over the ranks must walk the chain a single rank walks (same accept decisions on every rank) and the
    restated reference (oracle, used as the checker only)."""
    from jmbayes_b200.cohorts import make_cohort
    from jmbayes_b200.sharding import shard_cohort
    c = make_cohort("config4", n=n, seed=77)
    ctl = dict(c["control"]); ctl.update(n_burnin=iters - 10, n_iter=10, n_block=10, n_thin=5)
    sh = shard_cohort(c, rank, world) if world > 1 else dict(c, subject_offset=0)
    with api.Fit(sh["Data"], sh["Covs"]["b"], c["interval_cens"], device=local, subject_offset=sh["subject_offset"]) as f:
        if world > 1:
            api.connect_ranks(f, world, rank, dist)
        f.set_draw(sh["Data"])
        r = f.lap_rwm_C(sh["inits"], sh["priors"], sh["scales"], sh["Covs"], ctl)
    mine = (int(sh["subject_offset"]), r["mcmc"]["b"], r["extras"]["accept_b"], r["extras"]["trace_surv"], r["mcmc"]["alphas"])
    parts = [mine]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, mine)
    out = None
    if rank == 0:
        parts.sort(key=lambda x: x[0])
        b = np.concatenate([x[1] for x in parts], axis=0)
        acc = np.concatenate([x[2] for x in parts])
        same_decisions = all(np.array_equal(x[3], parts[0][3]) and np.array_equal(x[4], parts[0][4]) for x in parts)
        rel = lambda u, v: float(np.max(np.abs(u - v) / np.maximum(np.abs(v), 1e-12)))  # noqa: E731
        if world > 1:
            one = api.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], device=local)
        else:
            one = r
        from oracle import oracle
        oracle.build(); oracle.set_rowsum_mode(1)
        orc = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
        out = {"ranks": world, "subjects": n, "iterations": iters,
               "max_rel_vs_single_rank": max(rel(b, one["mcmc"]["b"]), rel(parts[0][3], one["extras"]["trace_surv"])),
               "max_rel_vs_oracle": max(rel(b, orc["mcmc"]["b"]), rel(parts[0][3], orc["extras"]["trace_surv"])),
               "identical_accepts": bool(np.array_equal(acc, one["extras"]["accept_b"]) and np.array_equal(acc, orc["extras"]["accept_b"])
                                         and same_decisions)}
        if not (out["identical_accepts"] and out["max_rel_vs_single_rank"] <= 1e-9 and out["max_rel_vs_oracle"] <= 1e-9):
            raise RuntimeError(f"sharded chain differs from the single-rank chain / the oracle: {out}")
    return out


def run_cuda(args):
    import torch
    import torch.distributed as dist
    from jmbayes_b200 import api, build
    from jmbayes_b200.cohorts import make_cohort, shared_model

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(build.LIB):
        raise RuntimeError("libjmb200.so missing: run __graft_entry__.build()")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path has no CPU fallback")

    parity = None if args.no_parity else _sharded_parity(api, dist, world, rank, local)

    cfg, n_default = WORKLOADS[args.workload]
    n = args.n_per_gpu or n_default
    n_total = n * world
    shared = shared_model(cfg, n_total)
    c = make_cohort(cfg, n=n, seed=1000 + rank, shared=shared)
    Data = c["Data"]
    fit = api.Fit(Data, c["Covs"]["b"], c["interval_cens"], device=local, subject_offset=rank * n)
    if world > 1:
        api.connect_ranks(fit, world, rank, dist)
    info = fit.layout_info()
    ctl = dict(c["control"])
    n_block = int(ctl["n_block"])
    # the timed window must hold its share of the per-block work (scale adaptation at the block end, the draw kernels):
    # with fewer steps than a block the window is centred on a block end
    preroll = max(0, n_block - args.warmup - args.steps // 2) if args.steps < n_block else 0
    total_needed = preroll + args.warmup + args.steps
    ctl.update(n_burnin=max(total_needed * 2, 100), n_iter=50, n_thin=50)   # never reaches a kept iteration
    fit.set_draw(Data)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)          # runs from before the warm-up to the end of the last timed region (e2e)
    if rank == 0:
        sampler.start()

    # ---- device-resident timing: K iterations, inputs in HBM, CUDA events on the library's stream ----
    # The last LEAD warm-up iterations are enqueued in the same call as the K timed ones: on N > 1 ranks the start event then
    # follows their all-reduce on the device, so the ranks' clocks start together (a host barrier leaves them hundreds of
    # microseconds apart, which a 20-step window would count as iteration time).  barrier + synchronize on both sides stay.
    lead = min(2, args.warmup)
    fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    if preroll:
        fit.chain_iterate(preroll)
    if args.warmup - lead > 0:
        fit.chain_iterate(args.warmup - lead)
    barrier()
    ms, _ = fit.chain_iterate(args.steps, lead=lead)
    launches = fit.launch_count_timed()
    barrier()
    fit.chain_end(fetch=False)
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_total * args.steps / (ms_max * 1e-3)

    # ---- per-launch duration of the step kernel (events around every launch) ----
    fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    fit.chain_iterate(args.warmup)
    ksteps = min(max(args.steps, 20), 100)
    _, kms = fit.chain_iterate(ksteps, time_kernels=True)
    fit.chain_end(fetch=False)
    k_ms = kms / ksteps

    # ---- several chains of the fit in flight (the reference's parallel axis); reported next to the C = 1 headline ----
    multi = None
    if args.chains > 1:
        fit.chain_slots(args.chains)
        for k in range(args.chains):
            fit.select_chain(k)
            fit.set_draw(Data)
            fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl, chain_id=k)
        fit.chains_iterate(args.chains, args.warmup)
        barrier()
        msm = fit.chains_iterate(args.chains, args.steps)
        barrier()
        for k in range(args.chains):
            fit.select_chain(k)
            fit.chain_end(fetch=False)
        fit.select_chain(0)
        tm = torch.tensor([msm], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        multi = {"chains": args.chains, "value": args.chains * n_total * args.steps / (float(tm.item()) * 1e-3), "unit": UNIT,
                 "ms_per_step_all_chains": float(tm.item()) / args.steps,
                 "what": "independent chains (draws) of the same fit advanced together on their own streams"}

    # ---- end to end through the public API with host buffers: one chain of the reference's own length
    #      (n_burnin = 1000, n_iter = 300, n_block = 50: R/mvJointModelBayes.R:8-12) per call = the per-draw step
    #      (betas -> device Xbetas_calc), chain_begin (H2D of b / scales / parameters from host memory), 1 300
    #      iterations, chain_end (D2H of the kept b, parameters, scales).  The fixed-effects designs are uploaded once
    #      per fit (like jmb_create), outside the timed region.  One untimed warm call, then best of three; the chain
    #      length does not depend on --steps. ----
    ctl_e = dict(c["control"])
    e2e_total = -(-(ctl_e["n_burnin"] + ctl_e["n_iter"]) // ctl_e["n_block"]) * ctl_e["n_block"]
    fit.set_xdesigns(Data)

    def e2e_call():
        barrier()
        t0 = time.perf_counter()
        fit.set_draw_betas(Data["betas"], Data["sigmas"], Data["invD"])
        ta_ = time.perf_counter()
        fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl_e)
        tb_ = time.perf_counter()
        fit.chain_iterate(e2e_total)
        tc_ = time.perf_counter()
        res_ = fit.chain_end()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        te = torch.tensor([t1 - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()), (ta_ - t0, tb_ - ta_, tc_ - tb_, t1 - tc_), res_

    ctl_w = dict(ctl_e); ctl_w.update(n_burnin=ctl_e["n_block"], n_iter=ctl_e["n_block"], n_thin=ctl_e["n_block"])
    fit.set_draw_betas(Data["betas"], Data["sigmas"], Data["invD"])          # warm call: allocations, first-touch of the staging buffers
    fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl_w)
    fit.chain_iterate(2 * ctl_e["n_block"])
    fit.chain_end()
    e2e_runs, res = [], None
    for _ in range(3):
        dt_, ph_, res = e2e_call()
        e2e_runs.append(dt_)
        print(f"[rank {rank}] e2e phases: set_draw_betas {ph_[0] * 1e3:.2f} ms chain_begin {ph_[1] * 1e3:.2f} ms iterate {ph_[2] * 1e3:.1f} ms "
              f"chain_end {ph_[3] * 1e3:.2f} ms", file=sys.stderr, flush=True)
    e2e_value = n_total * e2e_total / min(e2e_runs)
    clocks = sampler.stop() if rank == 0 else None
    h2d = 8 * sum(len(b) for b in Data["betas"]) + 8 * (fit.q * fit.q + fit.O) + 8 * n * (fit.q + 1)
    d2h = sum(v.nbytes for v in res["mcmc"].values()) + res["scales"]["sigma"]["b"].nbytes + 8 * n + 16 * e2e_total

    # ---- roofline of the step kernel: bytes that really cross HBM (the record layout's compulsory bytes, cross-checked
    #      against ncu's dram__bytes of the same kernel) over the CUDA-event launch duration ----
    prof, traffic, traffic_src = None, None, None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {}).get(info["kernel"])
        if prof:
            traffic = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) * (n / prof["subjects"])
            traffic_src = prof["source"]
    except Exception:
        pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    layout_bytes = info["shared_bytes_per_subject"] + info["chain_bytes_per_subject"]
    # bytes that cross HBM per launch: ncu's dram__bytes of this very kernel when a capture is committed (the kernel skips
    # the lower-interval rows of warps without interval-censored subjects, and over-fetches a little elsewhere), else the
    # record layout's bytes; never more than the layout's bytes
    real_bytes = min(traffic, layout_bytes * n) if traffic else layout_bytes * n
    achieved = real_bytes / (k_ms * 1e-3) / 1e9
    alg8d = ALG_BYTES[cfg]
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "traffic_source": traffic_src, "kernel": info["kernel"], "kernel_ms": k_ms,
            "alg_bytes_per_subject_iteration": real_bytes / n,
            "alg_source": ("ncu dram__bytes_read + dram__bytes_write of this kernel, per subject (" + str(traffic_src) + "); " if traffic else "") +
                          "layout bytes of the fused design (DESIGN.md section 2): design record + per-draw record read once, state read + "
                          "written, b-step draws written once per chunk and read once; Wlong(s) and the cached bases never exist",
            "layout_bytes_per_subject_iteration": layout_bytes, "frac_layout": layout_bytes * n / (k_ms * 1e-3) / 1e9 / peak,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
            "whole_step_gbs": real_bytes / (ms_max / args.steps * 1e-3) / 1e9,
            "survey_8d": {"alg_bytes_per_subject_iteration": alg8d, "gbs": alg8d * n / (k_ms * 1e-3) / 1e9,
                          "note": "SURVEY.md 8(d) compact layout that materialises Wlong(s) and the cached bases; the fused kernel does not "
                                  "move these bytes, so this is NOT a fraction of the roofline (it can exceed the peak)"}}
    if prof and "inst_per_subject" in prof:
        sm_hz = 1e6 * float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
        sms = 148
        roof["floors_us"] = {"dram": real_bytes / (peak * 1e9) * 1e6,
                             "issue": prof["inst_per_subject"] * n / (4 * sms * sm_hz) * 1e6,
                             # fp64 pipe: from the instruction count (two issue cycles per warp instruction) or from the capture's
                             # pipe-active share of its own duration
                             "fp64": (2.0 * prof["fp64_inst_per_subject"] * n / (4 * sms * sm_hz) * 1e6 if "fp64_inst_per_subject" in prof
                                      else prof.get("fp64_pipe_active_pct", 0.0) / 100.0 * prof.get("ncu_time_us", 0.0) * (n / prof["subjects"])),
                             "source": traffic_src}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "cohort": cfg, "subjects_per_gpu": n, "subjects_total": n_total,
                   "outcomes": fit.O, "q": fit.q, "terms": fit.T, "K": fit.K, "interval_cens": bool(c["interval_cens"]),
                   "parallelism": f"subject-shards x{world}", "chains_per_launch": 1,
                   "window": f"iterations {preroll + args.warmup}..{preroll + args.warmup + args.steps - 1} of a chain with n_block = {n_block}"
                             + (" (centred on a block end: holds the scale adaptation and its share of the draw kernels)" if preroll else ""),
                   "l2": "per-iteration working set %.2f GB per GPU >> 126 MB L2 (inputs larger than L2)" % (layout_bytes * n / 1e9)},
        "roofline": roof,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_total, "d2h_bytes_per_step": d2h / e2e_total,
                "runs_s": e2e_runs, "iterations_per_call": e2e_total,
                "what": f"set_draw_betas (device Xbetas_calc) + lap_rwm_C({e2e_total} iterations: the reference's n_burnin = 1000, n_iter = 300, "
                        "n_block = 50) through the C ABI with host buffers, wall clock, max over ranks; one warm call, best of 3"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if parity is not None:
        line["sharded_parity"] = parity
    if multi:
        line["multi_chain"] = multi
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        kind = _ref_kind()
        t_probe, it_probe = _cpu_worker((cfg, CPU_SAMPLE_N, 50, 0, kind))
        iters = max(50, int(12.0 / max(t_probe / it_probe, 1e-4)))
        val, dt, iters = cpu_reference(cfg, CPU_SAMPLE_N, iters, 1, kind)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{iters} iterations x {CPU_SAMPLE_N} subjects of {cfg}, single thread ({dt:.1f} s); "
                                          + ("the reference's own lap_rwm_C source compiled against the stand-in Armadillo/Rcpp header"
                                             if kind == "reference" else "restated lap_rwm_C (C port)")}
    fit.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# multi-state fits (multiState = TRUE): same metric, generic step kernel with one hazard-row block per
# transition row.  Every rank generates its own illness-death cohort (weak scaling; the per-iteration sums are still
# all-reduced).  The algorithmic bytes are the record layout's (jmb_layout_info): SURVEY 8(d) gives no figure for this shape.
# ------------------------------------------------------------------------------------------------
def run_multistate(args):
    import torch
    import torch.distributed as dist
    from jmbayes_b200 import api, build
    from jmbayes_b200.cohorts import make_multistate_cohort

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(build.LIB):
        raise RuntimeError("libjmb200.so missing: run __graft_entry__.build()")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path has no CPU fallback")
    n = args.n_per_gpu or WORKLOADS["multistate"][1]
    n_total = n * world
    c = make_multistate_cohort(n=n, seed=3000 + rank)
    Data = c["Data"]
    fit = api.Fit(Data, c["Covs"]["b"], False, multiState=True, device=local, subject_offset=rank * n)
    if world > 1:
        api.connect_ranks(fit, world, rank, dist)
    info = fit.layout_info()
    ctl = dict(c["control"])
    ctl.update(n_burnin=max((args.warmup + args.steps) * 2, 100), n_iter=50, n_thin=50)
    fit.set_draw(Data)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local)          # from before the warm-up to the end of the last timed region (e2e)
    if rank == 0:
        sampler.start()
    lead = min(2, args.warmup)             # see run_cuda: the clock starts after the last warm-up exchange on the device
    fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    if args.warmup - lead > 0:
        fit.chain_iterate(args.warmup - lead)
    barrier()
    ms, _ = fit.chain_iterate(args.steps, lead=lead)
    launches = fit.launch_count_timed()
    barrier()
    fit.chain_end(fetch=False)
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = n_total * args.steps / (ms_max * 1e-3)

    fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    fit.chain_iterate(args.warmup)
    ksteps = min(max(args.steps, 20), 100)
    _, kms = fit.chain_iterate(ksteps, time_kernels=True)
    fit.chain_end(fetch=False)
    k_ms = kms / ksteps

    # end to end with host buffers: per-draw vectors (jmb_set_draw), chain_begin, one chain of the reference's length
    # (n_burnin = 1000, n_iter = 300, n_block = 50), chain_end; one warm call, best of three
    ctl_e = dict(c["control"])
    e2e_total = -(-(ctl_e["n_burnin"] + ctl_e["n_iter"]) // ctl_e["n_block"]) * ctl_e["n_block"]

    def e2e_call(ctl_x, iters):
        barrier()
        t0 = time.perf_counter()
        fit.set_draw(Data)
        fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl_x)
        fit.chain_iterate(iters)
        res_ = fit.chain_end()
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item()), res_

    ctl_w = dict(ctl_e); ctl_w.update(n_burnin=ctl_e["n_block"], n_iter=ctl_e["n_block"], n_thin=ctl_e["n_block"])
    e2e_call(ctl_w, 2 * ctl_e["n_block"])
    e2e_runs, res = [], None
    for _ in range(3):
        dt_, res = e2e_call(ctl_e, e2e_total)
        e2e_runs.append(dt_)
    e2e_value = n_total * e2e_total / min(e2e_runs)
    clocks = sampler.stop() if rank == 0 else None
    h2d = (sum(np.asarray(x).nbytes for key in ("Xbetas", "XXbetas", "XXsbetas") for x in Data[key])
           + 8 * (fit.q * fit.q + fit.O) + 8 * n * (fit.q + 1))
    d2h = sum(v.nbytes for v in res["mcmc"].values()) + res["scales"]["sigma"]["b"].nbytes + 8 * n + 16 * e2e_total

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = info["shared_bytes_per_subject"] + info["chain_bytes_per_subject"]
    traffic, traffic_src = None, None
    try:      # bytes that cross HBM per launch from a committed ncu capture of this kernel, never more than the layout's
        prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("multistate", {}).get(info["kernel"] + ":multistate")
        if prof:
            traffic = (prof["dram_bytes_read"] + prof["dram_bytes_write"]) * (n / prof["subjects"])
            traffic_src = prof["source"]
    except Exception:
        pass
    achieved = (min(traffic, alg * n) if traffic else alg * n) / (k_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "multistate", "cohort": "illness-death, 2 outcomes (gaussian + binomial), q = 4, 3 strata x 7 B-spline coefficients",
                   "subjects_per_gpu": n, "subjects_total": n_total, "transition_rows_per_gpu": int(fit.nT), "outcomes": fit.O, "q": fit.q,
                   "terms": fit.T, "K": fit.K, "multiState": True, "parallelism": f"subject-shards x{world}", "chains_per_launch": 1,
                   "l2": "per-iteration working set %.2f GB per GPU (inputs larger than the 126 MB L2)" % (alg * n / 1e9)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "kernel": info["kernel"] + ":multistate", "kernel_ms": k_ms, "alg_bytes_per_subject_iteration": alg,
                     "frac_layout": alg * n / (k_ms * 1e-3) / 1e9 / peak,
                     "alg_source": "record layout (jmb_layout_info): design + per-draw records, state read/write, b-step draws",
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / e2e_total, "d2h_bytes_per_step": d2h / e2e_total,
                "runs_s": e2e_runs, "iterations_per_call": e2e_total,
                "what": f"jmb_set_draw + lap_rwm_C({e2e_total} iterations, multiState = TRUE) through the C ABI with host buffers, wall clock, one warm call, best of 3"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        kind = _ref_kind()
        t_probe, it_probe = _cpu_worker(("multistate", CPU_SAMPLE_N, 50, 0, kind))
        iters = max(50, int(12.0 / max(t_probe / it_probe, 1e-4)))
        val, dt, iters = cpu_reference("multistate", CPU_SAMPLE_N, iters, 1, kind)
        line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": 1, "kind": kind,
                                "sample": f"{iters} iterations x {CPU_SAMPLE_N} subjects of the multi-state cohort, single thread ({dt:.1f} s)"}
    fit.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# BASELINE config 5: survfitJM.mvJMbayes() loop (R/survfitJM.mvJMbayes.R:277-414)
# ------------------------------------------------------------------------------------------------
def _svft_case(n, seed, dense_basis=True):
    from jmbayes_b200 import survfit as sv
    from jmbayes_b200.cohorts import make_cohort
    c = make_cohort("config4", n=n, seed=seed)                       # new subjects of the config-4 fit (BASELINE config 5)
    d = sv.newdata_designs(c, L=SVFT_L, dense_basis=dense_basis)     # the CPU restatement reads the reference's dense bases
    means, draws = sv.posterior_draws(c, M=SVFT_M)
    return d, means, draws


def _svft_cpu(n, workers):
    """The restated R loop (oracle/jmb_oracle_svft.c) on the host cores, subjects split over worker threads."""
    from oracle import oracle, oracle_svft
    oracle.build()
    d, means, draws = _svft_case(n, 4242)
    t0 = time.perf_counter()
    oracle_svft.survfitJM(d, means, draws, threads=workers, full=False)
    dt = time.perf_counter() - t0
    return n * SVFT_M / dt, dt


def run_config5_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    workers = max(1, (os.cpu_count() or 1) - 1)
    _, dt0 = _svft_cpu(40 * workers, workers)              # probe, then ~15 s of CPU work
    n = int(min(40000, max(40 * workers, 40 * workers * 15.0 / max(dt0, 1e-3))))
    val, dt = _svft_cpu(n, workers)
    line = {"impl": "reference", "metric": SVFT_METRIC, "value": val, "unit": SVFT_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "config5", "subjects": n, "M": SVFT_M, "L": SVFT_L,
                       "note": "restated survfitJM loop (R absent from the image): optim BFGS + MH over M draws + horizon integrals"},
            "cpu_baseline": {"value": val, "unit": SVFT_UNIT, "cores": workers, "kind": "port",
                             "sample": f"{n} subjects x {SVFT_M} draws x {SVFT_L} horizons, {workers} worker threads ({dt:.1f} s)"},
            "e2e": {"value": val, "unit": SVFT_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_config5(args):
    import torch
    import torch.distributed as dist
    from jmbayes_b200 import build
    from jmbayes_b200 import survfit as sv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(build.LIB):
        raise RuntimeError("libjmb200.so missing: run __graft_entry__.build()")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path has no CPU fallback")
    n = args.n_per_gpu or WORKLOADS["config5"][1]
    steps, warmup = min(args.steps, 20), min(args.warmup, 3)
    # subjects shard over ranks, no collective on this path; the B-spline rows are evaluated by the library from
    # (knots, quadrature times), so the dense n*L*K x 17 basis never exists on the host
    d, means, draws = _svft_case(n, 5000 + rank, dense_basis=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    h = sv.SurvFit(d, device=local, subject_offset=rank * n)
    for _ in range(max(warmup, 1)):
        h.survfitJM(means, draws, full=False)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = h.info()["launches"]
    dev_ms, t0 = 0.0, time.perf_counter()
    for _ in range(steps):                                 # every step: params H2D, the kernels, summaries D2H
        h.survfitJM(means, draws, full=False, hessians=False)
        dev_ms += h.info()["last_kernel_ms"]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    info = h.info()
    t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall = float(t[0].item()), float(t[1].item())
    n_total = n * world
    value = n_total * SVFT_M * steps / (dev_ms * 1e-3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    rec_bytes = info["record_bytes_per_subject"]
    out_bytes = 8.0 * SVFT_M * SVFT_L * 2 + 8.0 * (4 * SVFT_L + 6 * 6 + 6 * 2 + 1)   # S written + re-read by the summary kernel, results
    achieved = (rec_bytes + out_bytes) * n / (dev_ms / steps * 1e-3) / 1e9
    pbytes = 8.0 * (SVFT_M + 1) * 70
    line = {
        "metric": SVFT_METRIC, "value": value, "unit": SVFT_UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "config5", "cohort": "config-4 fit (3 outcomes, q = 6, value + slope), all subjects as newdata truncated at 0.5 T", "subjects_per_gpu": n,
                   "subjects_total": n_total, "M": SVFT_M, "L": SVFT_L, "q": 6, "outcomes": 3, "K": 15,
                   "parallelism": f"subject-shards x{world} (no collective)",
                   "l2": "each subject's record (%.1f kB) is read once from HBM and re-used ~2900 times from L1/L2" % (rec_bytes / 1e3)},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "jmb_svft_static<SvSpecThree>", "kernel_ms": dev_ms / steps,
                     "alg_bytes_per_subject": rec_bytes + out_bytes,
                     "note": "compulsory bytes are ~1e-3 of what the time would allow: this kernel is bound by fp64 issue "
                             "(~2900 quadrature / visit evaluations per subject), not by HBM; see profiles/"},
        "e2e": {"value": n_total * SVFT_M * steps / wall, "unit": SVFT_UNIT, "h2d_bytes_per_step": pbytes,
                "d2h_bytes_per_step": 8.0 * n * (4 * SVFT_L + 6 + 6 + 1) + 12.0 * n,
                "what": "survfitJM() through the C ABI with host buffers: parameter H2D, kernels, modes / summaries / b.new / accept rates D2H, wall clock"},
        "gpu_launches": int(info["launches"] - l0), "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = max(1, (os.cpu_count() or 1) - 1)
        _, dt0 = _svft_cpu(40 * workers, workers)
        nc = int(min(40000, max(40 * workers, 40 * workers * 12.0 / max(dt0, 1e-3))))
        val, dt = _svft_cpu(nc, workers)
        line["cpu_baseline"] = {"value": val, "unit": SVFT_UNIT, "cores": workers, "kind": "port",
                                "sample": f"{nc} subjects x {SVFT_M} draws x {SVFT_L} horizons, restated R loop, {workers} threads ({dt:.1f} s)"}
    h.close()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _stats_cpu(S, seed=99):
    """R's summary_fun is a serial apply(): the restated R code in plain C (oracle/jmb_oracle_stats.c) on one core."""
    import numpy as np
    from oracle import oracle, oracle_stats
    oracle.build()
    rng = np.random.default_rng(seed)
    X = np.cumsum(rng.normal(size=(S, STATS_M)), axis=1) * 0.05 + rng.normal(size=(S, STATS_M))
    w = oracle_stats.stand(rng.normal(size=STATS_M))
    oracle_stats.statistics_c(X[:8], w)
    t0 = time.perf_counter()
    oracle_stats.statistics_c(X, w)
    dt = time.perf_counter() - t0
    return S / dt, dt


def run_summaries_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    _, dt0 = _stats_cpu(500)
    S = int(min(100000, max(500, 500 * 15.0 / max(dt0, 1e-3))))
    val, dt = _stats_cpu(S)
    line = {"impl": "reference", "metric": STATS_METRIC, "value": val, "unit": STATS_UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "summaries", "series": S, "M": STATS_M,
                       "note": "restated summary_fun sweeps in plain C (R absent from the image), serial as R's apply()"},
            "cpu_baseline": {"value": val, "unit": STATS_UNIT, "cores": 1, "kind": "port",
                             "sample": f"{S} series x {STATS_M} draws, ten statistics, plain-C restatement, one thread as R's apply() ({dt:.1f} s)"},
            "e2e": {"value": val, "unit": STATS_UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_summaries(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from jmbayes_b200 import build, summaries

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(build.LIB):
        raise RuntimeError("libjmb200.so missing: run __graft_entry__.build()")
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    n = args.n_per_gpu or WORKLOADS["summaries"][1]
    S, M = n * STATS_Q, STATS_M
    steps, warmup = min(args.steps, 10), max(min(args.warmup, 3), 3)
    # synthetic kept draws of b (n x q x M, series contiguous): autocorrelated chains around subject-specific centres;
    # 3 GB per GPU, larger than L2, so no flush is needed between steps
    g = torch.Generator(device=dev).manual_seed(7000 + rank)
    x = torch.empty((M, S), dtype=torch.float64, device=dev)
    centre = torch.randn(S, dtype=torch.float64, device=dev, generator=g)
    rho = torch.rand(S, dtype=torch.float64, device=dev, generator=g) * 0.8
    cur = torch.randn(S, dtype=torch.float64, device=dev, generator=g)
    for m in range(M):
        cur = rho * cur + torch.randn(S, dtype=torch.float64, device=dev, generator=g)
        x[m] = centre + 0.5 * cur
    w = summaries.stand(np.random.default_rng(7).normal(size=M))
    outs = {name: torch.empty(S * (2 if name in ("CIs", "wCIs") else 1), dtype=torch.float64, device=dev) for name in summaries.NAMES}
    ptrs = {k: v.data_ptr() for k, v in outs.items()}

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        summaries.device_statistics(x.data_ptr(), S, M, 1, S, w, ptrs, device=local)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev_ms = 0.0
    for _ in range(steps):
        dev_ms += summaries.device_statistics(x.data_ptr(), S, M, 1, S, w, ptrs, device=local)   # CUDA events on the launch stream
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    # e2e: the same series from pinned host memory through the public call, chunked copies overlapped with the kernel
    xh = torch.empty((M, S), dtype=torch.float64).pin_memory()
    xh.copy_(x)
    torch.cuda.synchronize()
    xn = xh.numpy()
    summaries.series_statistics(xn, S, M, 1, S, w, device=local)
    e2e_steps = max(1, min(steps, 3))
    barrier()
    walls = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        summaries.series_statistics(xn, S, M, 1, S, w, device=local)
        walls.append(time.perf_counter() - t0)
    wall = sum(walls) / e2e_steps
    t = torch.tensor([dev_ms, wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, wall = float(t[0].item()), float(t[1].item())
    S_total = S * world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = 8.0 * (M + 12)
    achieved = alg * S / (dev_ms / steps * 1e-3) / 1e9
    line = {
        "metric": STATS_METRIC, "value": S_total * steps / (dev_ms * 1e-3), "unit": STATS_UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": dev_ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": "summaries", "what": "mcmc$b of the config-4 shard: n x q x M kept draws, series contiguous",
                   "subjects_per_gpu": n, "q": STATS_Q, "series_per_gpu": S, "series_total": S_total, "M": M,
                   "parallelism": f"series-shards x{world} (no collective)", "l2": "inputs (3 GB per GPU) larger than L2"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "kernel": "jmb_stats_kernel", "kernel_ms": dev_ms / steps, "alg_bytes_per_series": alg,
                     "note": "each series is read once (M doubles) and 12 doubles are written; the time goes to the warp-level sort "
                             "and the binned kernel density of modes() (fp64 issue from shared memory), not to HBM; see profiles/"},
        "e2e": {"value": S_total / wall, "unit": STATS_UNIT, "h2d_bytes_per_step": 8.0 * M * S, "d2h_bytes_per_step": 96.0 * S,
                "what": "jmb_mcmc_statistics() through the C ABI on pinned host draws: chunked H2D overlapped with the kernel, "
                        "all ten statistics D2H, wall clock", "wall_ms_per_call": [round(1e3 * t, 1) for t in walls]},
        "gpu_launches": steps, "clocks": clocks,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        _, dt0 = _stats_cpu(500)
        Sc = int(min(100000, max(500, 500 * 12.0 / max(dt0, 1e-3))))
        val, dt = _stats_cpu(Sc)
        line["cpu_baseline"] = {"value": val, "unit": STATS_UNIT, "cores": 1, "kind": "port",
                                "sample": f"{Sc} series x {M} draws, ten statistics, plain-C restatement of the serial R sweeps, one thread ({dt:.1f} s)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.workload == "summaries":
        return run_summaries_reference(args) if args.impl == "reference" else run_summaries(args)
    if args.workload == "config5":
        return run_config5_reference(args) if args.impl == "reference" else run_config5(args)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "multistate":
        run_multistate(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
