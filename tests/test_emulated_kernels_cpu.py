"""The library's OWN kernels, run on CPU threads.

oracle/emu compiles jmbayes_b200/csrc/jmb200.cu (+ jmb_kernels.cuh, jmb_device.cuh, jmb_laplace.cu, jmb_accuracy.cu,
jmb_survfit.cu/.cuh, jmb_summaries.cu) - the sources a B200 runs, not a restatement - with g++ against a stand-in CUDA runtime in which a kernel launch executes its CTAs
with one OS thread per CUDA thread and __syncthreads / __syncwarp / __shfl_*_sync are barriers over exactly the threads they
name (oracle/emu/shim/cuda_runtime.h).  That is test infrastructure in the sense of oracle/: nothing in the product loads
it, it is never timed and it is no fallback.  It lets this CPU suite check, on small cohorts, the kernel arithmetic, the
block / warp / half-warp-group logic and the host orchestration of paths that have not been on a GPU yet (multi-state fits,
the accuracy sums), and it re-checks GPU-verified paths (so a disagreement would point at the emulator, not at the kernels).
The compile-time specialised kernels are emulated as they are; of the TMA-staged kernel everything but its four PTX helpers
(mbarrier.init / arrive.expect_tx / try_wait, cp.async.bulk), which oracle/emu/shim/emu_mbarrier.h models (transaction counts,
phase parity, copy published on completion), so its two-stage per-warp pipeline logic runs too.
"""
import os

import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort
from oracle import np_twin


@pytest.fixture(scope="module")
def emu(oracle):
    """jmbayes_b200.api bound to the emulated library for the duration of this module."""
    import importlib.util
    from jmbayes_b200 import api
    spec = importlib.util.spec_from_file_location("make_emu", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                            "oracle", "emu", "make_emu.py"))
    make_emu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(make_emu)
    lib = make_emu.build()
    import ctypes as C
    saved = (api._lib_cache, os.environ.get("JMB_KERNEL"))
    os.environ["JMB_KERNEL"] = "generic"               # default for this module; one test below selects the other kernels
    api._lib_cache = api._bind(C.CDLL(lib))            # test-only injection: the product never loads this library
    yield api
    api._lib_cache = saved[0]
    if saved[1] is None:
        os.environ.pop("JMB_KERNEL", None)
    else:
        os.environ["JMB_KERNEL"] = saved[1]


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


@pytest.mark.parametrize("cfg", ["config2", "config4"])
def test_emulated_generic_kernel_matches_oracle(emu, oracle, cfg):
    """GPU-verified shapes first: the emulation of the generic kernel (16 and 32 lanes per subject) reproduces the oracle."""
    c = make_cohort(cfg, n=40)
    D, ini = c["Data"], c["inits"]
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], c["interval_cens"], ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    with emu.Fit(D, c["Covs"]["b"], c["interval_cens"]) as fit:
        assert fit.layout_info()["kernel"] == "generic"
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    for k in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb"):
        assert _rel(got[k], ref[k]) <= 1e-12, k
    ctl = dict(c["control"]); ctl.update(n_burnin=10, n_iter=10, n_block=5, n_thin=5)
    r1 = emu.lap_rwm_C(ini, D, c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
    r0 = oracle.lap_rwm_C(ini, D, c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert np.allclose(r1["mcmc"]["b"], r0["mcmc"]["b"], rtol=1e-10, atol=1e-13)


def test_emulated_multistate_kernel_matches_oracle(emu, oracle):
    """jmb_step_kernel<G, true> + the per-stratum draws of the finalize kernel: pieces, chain, strata precisions."""
    c = make_multistate_cohort(n=60)
    D, ini = c["Data"], c["inits"]
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], False, ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"], multiState=True)
    with emu.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        assert fit.nT == D["event"].size and fit.nRisks == 3
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    for k in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
        assert got[k].shape == ref[k].shape and _rel(got[k], ref[k]) <= 1e-12, k
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=5)
    r1 = emu.lap_rwm_C(ini, D, c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=3)
    r0 = oracle.lap_rwm_C(ini, D, c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=3)
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert r1["mcmc"]["tau_Bs_gammas"].shape == (3, 4)
    for k in ("b", "Bs_gammas", "gammas", "alphas", "tau_Bs_gammas"):
        assert np.allclose(r1["mcmc"][k], r0["mcmc"][k], rtol=1e-10, atol=1e-13), k
    assert np.allclose(r1["logWeights"], r0["logWeights"], rtol=1e-12)
    assert np.allclose(r1["scales"]["sigma"]["b"], r0["scales"]["sigma"]["b"], rtol=1e-12)
    assert np.array_equal(r1["extras"]["accept_b"], r0["extras"]["accept_b"])


def test_emulated_multistate_kernel_reproduces_reference_golden(emu):
    """... and the golden chain of the reference's own lap_rwm_C source (shrinkage priors + adaptCov)."""
    from test_multistate_cpu import check_chain, make_golden, _gold
    name = "chain_multistate_shrink_adapt"
    c, ctl, pr, chain = make_golden.ms_inputs(name)
    r = emu.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, True, chain_id=chain)
    check_chain(r, _gold(name), rtol=1e-8, atol=1e-11)


def test_emulated_one_transition_per_subject_is_the_ordinary_fit(emu):
    c = make_cohort("config3", n=40)
    ctl = dict(c["control"]); ctl.update(n_burnin=10, n_iter=10, n_block=5, n_thin=5)
    res = [emu.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, ms) for ms in (False, True)]
    assert _rel(res[1]["extras"]["trace_surv"], res[0]["extras"]["trace_surv"]) <= 1e-13
    assert np.allclose(res[1]["mcmc"]["b"], res[0]["mcmc"]["b"], rtol=1e-12, atol=1e-14)


def test_emulated_multistate_laplace_step_and_device_xbetas(emu, oracle):
    c = make_multistate_cohort(n=50)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl = oracle.wlong(D, c["Covs"]["b"], False, b, 0, multiState=True)
    Wls = oracle.wlong(D, c["Covs"]["b"], False, b, 1, multiState=True)
    Bs, g, a = (np.asarray(ini[k], dtype=float) for k in ("Bs_gammas", "gammas", "alphas"))
    tau, temp = 3.5, 0.7
    want = np_twin.logPosterior(temp, D, Wl, Wls, Bs, g, a, pr, tau)
    W = np.hstack([np.asarray(D["W1s"]), np.asarray(D["W2s"]), Wls])
    Int = np.asarray(D["Pw"]) * np.exp(W @ np.concatenate([Bs, g, a]))
    ecs = np.asarray(D["event"]) @ np.hstack([np.asarray(D["W1"]), np.asarray(D["W2"]), Wl])
    pen = np.concatenate([tau * (pr["Tau_Bs_gammas"] @ (Bs - pr["mean_Bs_gammas"])), pr["Tau_gammas"] @ (g - pr["mean_gammas"]),
                          pr["Tau_alphas"] @ (a - pr["mean_alphas"])])
    score = temp * (ecs - W.T @ Int) - pen
    with emu.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        fit.set_wlong(Wl, Wls)
        lp = fit.logPosterior(temp, Bs, g, a, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        gr = fit.gradient_logPosterior(temp, Bs, g, a, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        assert abs(lp - want) <= 1e-12 * abs(want)
        assert np.allclose(gr[:-1], score, rtol=1e-10, atol=1e-10)
        # Xbetas_calc on the "device": same per-draw records as the host vectors
        fit.set_draw(D)
        e0 = fit.eval_logpost_re(b, Bs, g, a)
        fit.set_xdesigns(D)
        fit.set_draw_betas(D["betas"], D["sigmas"], D["invD"])
        e1 = fit.eval_logpost_re(b, Bs, g, a)
    for k in ("log_pyb", "log_h", "H", "log_ptb"):
        assert _rel(e1[k], e0[k]) <= 1e-13, k


@pytest.mark.parametrize("n,seed,na", [(2, 1, 0.0), (300, 2, 0.0), (600, 4, 0.15), (257, 5, 0.05)])
def test_emulated_accuracy_sums_match_restated_loops(emu, n, seed, na):
    from jmbayes_b200 import accuracy
    from oracle import oracle_accuracy as oa
    from test_accuracy_cpu import make_case
    case = make_case(n, seed, na)
    want = oa.aucJM_pairs(*case)
    got = accuracy.aucJM_pairs(*case)
    assert np.isclose(got["num"], want[1], rtol=1e-12) and np.isclose(got["den"], want[2], rtol=1e-12)
    assert (np.isnan(got["auc"]) and np.isnan(want[0])) or np.isclose(got["auc"], want[0], rtol=1e-12)
    rng = np.random.default_rng(seed)
    a, d, cc, w = rng.uniform(size=300), rng.uniform(size=200), rng.uniform(size=100), rng.uniform(size=100)
    cc[3] = np.nan
    for loss in ("square", "absolute"):
        assert np.isclose(accuracy.prederrJM_sum(a, d, cc, w, 700, loss), oa.prederrJM_sum(a, d, cc, w, 700, loss), rtol=1e-12)


@pytest.mark.parametrize("n,seed,na", [(1, 1, 0.0), (700, 2, 0.0), (2500, 3, 0.0), (300, 4, 0.05), (50_000, 3, 0.0)])
def test_emulated_roc_sums_match_restated_outer_product(emu, n, seed, na):
    from jmbayes_b200 import accuracy
    from oracle import oracle_accuracy as oa
    from test_accuracy_cpu import make_case
    Time, event, Th, pi, pi2, _ = make_case(n, seed, na)
    want = oa.rocJM(Time, event, Th, pi, pi2)
    got = accuracy.rocJM(Time, event, Th, pi, pi2)
    for k in ("nTP", "nFP", "nFN", "nTN", "TP", "FP"):
        assert np.allclose(got[k], want[k], rtol=1e-11, atol=1e-11, equal_nan=True), k
    # the kappa-type measures are 0/0-like where every (or no) probability is below the threshold: compare inside
    below = np.sum(np.asarray(pi)[:, None] < want["thrs"][None, :], axis=0)
    inside = (below > 0) & (below < np.sum(~np.isnan(pi)))
    for k in ("qSN", "qSP", "qOverall"):
        assert np.allclose(got[k][inside], want[k][inside], rtol=1e-9, atol=1e-9, equal_nan=True), k
    if not na:
        assert np.isclose(got["F1score"], want["F1score"], equal_nan=True) and np.isclose(got["Youden"], want["Youden"], equal_nan=True)


def test_emulated_summaries_and_survfit_kernels_match_their_oracles(emu):
    """The checks of __graft_entry__.smoke() for the posterior summaries and survfitJM, on the emulated kernels."""
    from jmbayes_b200 import summaries, survfit as sv
    from oracle import oracle_stats, oracle_svft
    rng = np.random.default_rng(3)
    X = np.cumsum(rng.normal(size=(48, 160)), axis=1) * 0.1 + rng.normal(size=(48, 160))
    w = oracle_stats.stand(rng.normal(size=160))
    gs, _ = summaries.series_statistics(np.ascontiguousarray(X), 48, 160, 160, 1, w)
    rs = oracle_stats.statistics(X, w)
    for name in ("postMeans", "postwMeans", "StDev", "wStDev", "CIs", "wCIs", "EffectiveSize", "StErr", "Pvalues"):
        assert np.allclose(gs[name], rs[name], rtol=1e-9, atol=1e-12), name
    c5 = make_cohort("config3", n=24)
    d5 = sv.newdata_designs(c5, L=3)
    means, draws = sv.posterior_draws(c5, M=6)
    o5 = oracle_svft.modes(d5, means)
    with sv.SurvFit(d5) as h:
        g5 = h.modes(means)
        s5 = h.sample(draws, o5["modes"], o5["hessians"])
    r5 = oracle_svft.sample(d5, draws, o5["modes"], o5["hessians"])
    assert np.max(np.abs(g5["modes"] - o5["modes"])) < 1e-6
    assert np.allclose(s5["summaries"], r5["summaries"], rtol=1e-9)


@pytest.mark.parametrize("cfg,mode,G,early", [("config4", None, 4, 0), ("config4", None, 8, 1), ("config4", None, 16, 1), ("config4", None, 2, 0),
                                              ("config4", None, 8, 0), ("config3", None, 4, 1), ("config3", None, 16, 1), ("config3", None, 16, 0),
                                              ("config2", None, 4, 1), ("config2", None, 8, 0),
                                              ("config4", None, 8, 2), ("config4", None, 4, 2), ("config3", None, 16, 2), ("config2", None, 4, 2),
                                              ("config4", None, 8, 3), ("config4", None, 16, 3), ("config3", None, 8, 3), ("config2", None, 8, 3),
                                              ("config2", "tma", 0, 0), ("config4", "tma", 0, 0), ("config3", "direct", 0, 0)])
def test_emulated_static_kernels_match_oracle(emu, oracle, cfg, mode, G, early):
    """The hot-path kernels against the oracle chain: jmb_step_rows (G lanes per subject, rows strided by G; the default),
    jmb_step_tma (two-stage bulk-copy + mbarrier pipeline per warp; 70 subjects on the one-SM emulated device give every warp
    three pipeline steps) and jmb_step_static."""
    saved = (os.environ.get("JMB_KERNEL"), os.environ.get("JMB_ROWS_G"), os.environ.get("JMB_ROWS_EARLY"))
    try:
        os.environ.pop("JMB_KERNEL", None); os.environ.pop("JMB_ROWS_G", None); os.environ.pop("JMB_ROWS_EARLY", None)
        if mode:
            os.environ["JMB_KERNEL"] = mode
        if G:
            os.environ["JMB_ROWS_G"] = str(G)
            os.environ["JMB_ROWS_EARLY"] = str(early)
        c = make_cohort(cfg, n=70)
        ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=10, n_block=10, n_thin=5)
        with emu.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
            want = {None: "static-rows%d%s:" % (G, ("", "e", "p", "q")[early]), "tma": "static-tma:", "direct": "static:"}[mode]
            assert fit.layout_info()["kernel"].startswith(want)
            fit.set_draw(c["Data"])
            r1 = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    finally:
        for key, val in zip(("JMB_KERNEL", "JMB_ROWS_G", "JMB_ROWS_EARLY"), saved):
            if val is None:
                os.environ.pop(key, None)
            else:
                os.environ[key] = val
    oracle.set_rowsum_mode(1)
    r0 = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert np.allclose(r1["mcmc"]["b"], r0["mcmc"]["b"], rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("cfg,chunk", [("config4", "5"), ("config2", "2")])
def test_emulated_blocks_of_several_draw_chunks_match_oracle(emu, oracle, cfg, chunk, monkeypatch):
    """Blocks of several draw chunks (JMB_RNG_CHUNK < n_block): the draws are keyed by (iteration, subject), so the chunk length
    must not change the chain.  Same counter-RNG specification as the oracle: the whole trajectory must agree."""
    monkeypatch.delenv("JMB_KERNEL", raising=False)
    monkeypatch.setenv("JMB_RNG_CHUNK", chunk)
    c = make_cohort(cfg, n=70)
    ctl = dict(c["control"]); ctl.update(n_burnin=40, n_iter=20, n_block=20, n_thin=5)
    with emu.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
        fit.set_draw(c["Data"])
        r1 = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    oracle.set_rowsum_mode(1)
    r0 = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert np.allclose(r1["mcmc"]["b"], r0["mcmc"]["b"], rtol=1e-10, atol=1e-13)
    assert np.array_equal(r1["extras"]["accept_b"], r0["extras"]["accept_b"])


def test_emulated_step_lists_do_not_change_the_chain(emu, monkeypatch):
    """The cost-balanced step lists (longest processing time first over the (CTA, warp) slots) and the CTA-contiguous assignment of
    equal counts walk the same chain: only the grouping of the per-slot partial sums differs."""
    monkeypatch.delenv("JMB_KERNEL", raising=False)
    c = make_cohort("config4", n=150)
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=10, n_block=10, n_thin=5)
    out = {}
    for lpt in ("1", "0"):
        monkeypatch.setenv("JMB_ROWS_LPT", lpt)
        with emu.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
            assert fit.layout_info()["kernel"].startswith("static-rows")
            fit.set_draw(c["Data"])
            out[lpt] = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    assert _rel(out["1"]["extras"]["trace_surv"], out["0"]["extras"]["trace_surv"]) <= 1e-12
    assert np.array_equal(out["1"]["extras"]["accept_b"], out["0"]["extras"]["accept_b"])
    assert np.allclose(out["1"]["mcmc"]["b"], out["0"]["mcmc"]["b"], rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("early", ["1", "3"])
def test_emulated_visit_segments_beyond_the_stage_are_read_in_place(emu, oracle, early, monkeypatch):
    """A stage too small for the longer visit segments (JMB_ROWS_CAP): those subjects read their visit rows from the record through
    the same loop (pointer selected per subject), the others from the stage - same chain as the oracle."""
    monkeypatch.delenv("JMB_KERNEL", raising=False)
    monkeypatch.setenv("JMB_ROWS_G", "8"); monkeypatch.setenv("JMB_ROWS_EARLY", early); monkeypatch.setenv("JMB_ROWS_CAP", "40")
    c = make_cohort("config4", n=70)
    v = np.bincount(np.asarray(c["Data"]["idL"][0]).astype(int))
    assert (6 * v[v > 0] > 40).any() and (6 * v[v > 0] <= 40).any()          # both paths are taken (6 stored doubles per visit)
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=10, n_block=10, n_thin=5)
    with emu.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
        fit.set_draw(c["Data"])
        r1 = fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)
    oracle.set_rowsum_mode(1)
    r0 = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"])
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert np.allclose(r1["mcmc"]["b"], r0["mcmc"]["b"], rtol=1e-10, atol=1e-13)
