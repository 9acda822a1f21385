"""Multi-state fits (``multiState = TRUE``) on the CUDA path, through the C ABI, against the CPU oracle, the reference's
golden vectors and a size-independent property.  Tolerances as in test_gpu_parity.py (1e-10 relative on the pieces of
the b-step target against the exact-rowsum oracle).

"""
import importlib.util
import os

import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort

pytestmark = [pytest.mark.gpu]
RTOL = 1e-10

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def test_one_transition_per_subject_is_the_ordinary_fit(cuda_lib, monkeypatch):
    """One transition row per subject, one stratum: the multi-state kernel must walk the chain of the generic kernel."""
    monkeypatch.setenv("JMB_KERNEL", "generic")
    c = make_cohort("config3", n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=10)
    ini = c["inits"]
    res = []
    for ms in (False, True):
        with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False, multiState=ms) as fit:
            assert fit.layout_info()["kernel"] == "generic"
            fit.set_draw(c["Data"])
            ev = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
            res.append((ev, fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)))
    (e0, r0), (e1, r1) = res
    for k in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
        assert _rel(e1[k], e0[k]) <= 1e-13, k
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-12
    assert np.allclose(r1["mcmc"]["b"], r0["mcmc"]["b"], rtol=1e-10, atol=1e-12)
    for k in ("Bs_gammas", "gammas", "alphas", "tau_Bs_gammas"):
        assert np.allclose(r1["mcmc"][k], r0["mcmc"][k], rtol=1e-10), k


@pytest.mark.parametrize("n", [400, 23])
def test_logpost_pieces_match_oracle_multistate(cuda_lib, oracle, n):
    c = make_multistate_cohort(n=n)
    D, ini = c["Data"], c["inits"]
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], False, ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"], multiState=True)
    with cuda_lib.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        assert fit.nT == D["event"].size and fit.nRisks == 3
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    for key in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
        assert got[key].shape == ref[key].shape, key
        assert np.isfinite(got[key]).all(), key
        assert _rel(got[key], ref[key]) <= RTOL, (key, _rel(got[key], ref[key]))
    # the b-step target: log p(y|b) + rowsum(log_ptb, idT_rsum) + log p(b)  (src/fast_LAPRWM.cpp:669)
    idT = np.asarray(D["idT2"][0])
    tot = got["log_pyb"] + np.bincount(idT, weights=got["log_ptb"]) + got["log_pb"]
    tot_ref = ref["log_pyb"] + np.bincount(idT, weights=ref["log_ptb"]) + ref["log_pb"]
    assert _rel(tot, tot_ref) <= RTOL


def test_chain_matches_oracle_multistate(cuda_lib, oracle):
    c = make_multistate_cohort(n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=60, n_iter=40, n_block=10, n_thin=10)
    oracle.set_rowsum_mode(1)
    ref = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=3)
    got = cuda_lib.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=3)
    assert _rel(got["extras"]["trace_surv"], ref["extras"]["trace_surv"]) <= 1e-9
    assert np.allclose(got["mcmc"]["b"], ref["mcmc"]["b"], rtol=1e-8, atol=1e-11)
    assert got["mcmc"]["tau_Bs_gammas"].shape == (3, 4)                       # nRisks x n_out
    for key in ("Bs_gammas", "gammas", "alphas", "tau_Bs_gammas"):
        assert np.allclose(got["mcmc"][key], ref["mcmc"][key], rtol=1e-8, atol=1e-12), key
    assert np.allclose(got["logWeights"], ref["logWeights"], rtol=1e-10)
    assert np.allclose(got["scales"]["sigma"]["b"], ref["scales"]["sigma"]["b"], rtol=1e-9)
    assert np.allclose(got["extras"]["accept_b"], ref["extras"]["accept_b"])
    assert 0.05 < got["extras"]["accept_b"].mean() < 0.8


def test_multistate_specialised_and_generic_kernels_agree(cuda_lib, monkeypatch):
    """The illness-death shape has a compile-time-specialised row-strided kernel (hazard-row blocks of the subject's
    transition rows walked in one flattened loop); every group width, with and without the early stage, must walk the
    chain of the runtime-descriptor kernel."""
    c = make_multistate_cohort(n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=10)

    def run(label):
        with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False, multiState=True) as fit:
            k = fit.layout_info()["kernel"]
            assert k.startswith(label) and (label == "generic" or k.endswith(":multistate_illness_death")), k
            fit.set_draw(c["Data"])
            return fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)

    monkeypatch.setenv("JMB_KERNEL", "generic")
    g = run("generic")
    monkeypatch.delenv("JMB_KERNEL")
    for G, early in [(None, None)] + [(w, e) for w in ("4", "8", "16") for e in ("0", "1", "2", "3")]:
        if G:
            monkeypatch.setenv("JMB_ROWS_G", G); monkeypatch.setenv("JMB_ROWS_EARLY", early)
        r = run("static-rows" + (G + {"0": ":", "1": "e:", "2": "p:", "3": "q:"}[early] if G else ""))
        assert _rel(r["extras"]["trace_surv"], g["extras"]["trace_surv"]) <= 1e-10, (G, early)
        assert np.allclose(r["mcmc"]["b"], g["mcmc"]["b"], rtol=1e-9, atol=1e-12), (G, early)
        assert np.array_equal(r["extras"]["accept_b"], g["extras"]["accept_b"]), (G, early)
        assert np.allclose(r["mcmc"]["tau_Bs_gammas"], g["mcmc"]["tau_Bs_gammas"], rtol=1e-9), (G, early)


@pytest.mark.parametrize("name", sorted(make_golden.MS_CASES))
def test_cuda_reproduces_reference_multistate_golden(cuda_lib, name):
    from test_multistate_cpu import check_chain
    c, ctl, pr, chain = make_golden.ms_inputs(name)
    r = cuda_lib.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, True, chain_id=chain)
    check_chain(r, dict(np.load(os.path.join(HERE, "golden", name + ".npz"))), rtol=1e-7, atol=1e-10)


def test_multistate_limits_are_loud(cuda_lib):
    c = make_multistate_cohort(n=40)
    with pytest.raises(cuda_lib.JmbError):            # not combinable with interval censoring
        cuda_lib.Fit(c["Data"], c["Covs"]["b"], True, multiState=True)
    D = dict(c["Data"]); D["idT2"] = [np.asarray(D["idT2"][0])[::-1].copy()]
    with pytest.raises(cuda_lib.JmbError):            # transition rows must be grouped by subject
        cuda_lib.Fit(D, c["Covs"]["b"], False, multiState=True)
    with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False, multiState=True) as fit:
        with pytest.raises(cuda_lib.JmbError):        # Wlong has one row per transition
            fit.set_wlong(np.zeros((fit.n, fit.A)), np.zeros((fit.n * fit.K, fit.A)))


def test_device_xbetas_calc_multistate(cuda_lib):
    """jmb_set_xdesigns + jmb_set_draw_betas (Xbetas_calc on the device, R/supportFuns_mvJointModelBayes.R:43-55) fill the
    same per-draw records as the host vectors of jmb_set_draw, one block of term rows per transition row."""
    c = make_multistate_cohort(n=300)
    D, ini = c["Data"], c["inits"]
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=20, n_block=10, n_thin=10)
    with cuda_lib.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        fit.set_draw(D)
        e0 = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        r0 = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
    with cuda_lib.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        fit.set_xdesigns(D)
        fit.set_draw_betas(D["betas"], D["sigmas"], D["invD"])
        e1 = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        r1 = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
    for k in ("log_pyb", "log_h", "H", "log_ptb"):
        assert _rel(e1[k], e0[k]) <= 1e-12, k
    assert _rel(r1["extras"]["trace_surv"], r0["extras"]["trace_surv"]) <= 1e-9


def test_laplace_step_multistate(cuda_lib, oracle):
    """logPosterior / gradient_logPosterior (src/fast_logPost.cpp:8-26, src/fast_score.cpp:8-50) on the nT transition
    rows of a multi-state fit, as marglogLik2 calls them (R/margLogLik2.R:42-65), against NumPy."""
    from oracle import np_twin
    c = make_multistate_cohort(n=300)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl = oracle.wlong(D, c["Covs"]["b"], False, b, 0, multiState=True)
    Wls = oracle.wlong(D, c["Covs"]["b"], False, b, 1, multiState=True)
    Bs, g, a = (np.asarray(ini[k], dtype=float) for k in ("Bs_gammas", "gammas", "alphas"))
    tau, temp = 3.5, 0.7
    want = np_twin.logPosterior(temp, D, Wl, Wls, Bs, g, a, pr, tau)
    ev = np.asarray(D["event"])
    W = np.hstack([np.asarray(D["W1s"]), np.asarray(D["W2s"]), Wls])
    Int = np.asarray(D["Pw"]) * np.exp(W @ np.concatenate([Bs, g, a]))
    ecs = ev @ np.hstack([np.asarray(D["W1"]), np.asarray(D["W2"]), Wl])
    pen = np.concatenate([tau * (pr["Tau_Bs_gammas"] @ (Bs - pr["mean_Bs_gammas"])), pr["Tau_gammas"] @ (g - pr["mean_gammas"]),
                          pr["Tau_alphas"] @ (a - pr["mean_alphas"])])
    score = temp * (ecs - W.T @ Int) - pen
    zb = Bs - pr["mean_Bs_gammas"]
    last = tau * (-0.5 * zb @ pr["Tau_Bs_gammas"] @ zb + (pr["A_tau_Bs_gammas"] - 1.0) / tau - pr["B_tau_Bs_gammas"])
    with cuda_lib.Fit(D, c["Covs"]["b"], False, multiState=True) as fit:
        fit.set_wlong(Wl, Wls)
        lp = fit.logPosterior(temp, Bs, g, a, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        gr = fit.gradient_logPosterior(temp, Bs, g, a, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
    assert abs(lp - want) <= 1e-10 * abs(want)
    assert np.allclose(gr[:-1], score, rtol=1e-9, atol=1e-9) and np.isclose(gr[-1], last, rtol=1e-10)
