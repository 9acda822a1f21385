"""Golden vectors produced by the reference's OWN sources (tests/golden/make_golden.py:
/root/reference/src/*.cpp compiled against the stand-in Armadillo/Rcpp header, fed the DESIGN.md
random streams).  CPU: the restated oracle must reproduce them; GPU: the CUDA path must.
"""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from jmbayes_b200 import _abi
from jmbayes_b200.cohorts import make_cohort
from oracle import np_twin

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)
CASES = make_golden.CASES


def _load(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def _inputs(name):
    cfg, n, ctl_over, pr_over, chain = CASES[name]
    c = make_cohort(cfg, n=n)
    ctl = dict(c["control"]); ctl.update(ctl_over)
    pr = dict(c["priors"]); pr.update(pr_over)
    return c, ctl, pr, chain


def _check_chain(r, gold, rtol, atol):
    for k, v in r["mcmc"].items():
        g = gold["mcmc_" + k]
        assert v.shape == g.shape, k
        assert np.allclose(v, g, rtol=rtol, atol=atol), (k, float(np.max(np.abs(v - g))))
    assert np.allclose(r["logWeights"], gold["logWeights"], rtol=rtol)
    assert np.allclose(r["scales"]["sigma"]["b"], gold["sigma_b"], rtol=rtol)
    sg = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    assert np.allclose(sg, gold["sigma_surv"], rtol=rtol)
    for k in ("Bs_gammas", "gammas", "alphas"):
        assert np.allclose(r["scales"]["Sigma"][k], gold["Sigma_" + k], rtol=max(rtol, 1e-8), atol=1e-14), k


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_reference_golden(oracle, name):
    c, ctl, pr, chain = _inputs(name)
    oracle.set_rowsum_mode(0)          # the reference's own cumsum-difference rowsum
    r = oracle.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=chain)
    _check_chain(r, _load(name), rtol=1e-11, atol=1e-13)


def test_golden_keeps_the_reference_thinning_quirk():
    g = _load("chain_config2_thin1")       # 0-based iter vs 1-based keep: last column never written
    assert np.all(g["mcmc_alphas"][:, -1] == 0.0) and np.all(g["mcmc_alphas"][:, :-1] != 0.0)


@pytest.mark.parametrize("name", ["chain_config3", "chain_config4"])
def test_live_reference_source_matches_oracle(oracle, name):
    """Where the reference tree is mounted: rebuild oracle/_ref from it and compare directly."""
    from oracle import ref
    if not os.path.exists(os.path.join(ref.REF, "src", "fast_LAPRWM.cpp")):
        pytest.skip("reference tree not present")
    c, ctl, pr, chain = _inputs(name)
    ctl = dict(ctl); ctl.update(n_burnin=25, n_iter=15, n_block=5, n_thin=5)
    oracle.set_rowsum_mode(0)
    a = ref.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=9)
    b = oracle.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=9)
    for k in a["mcmc"]:
        assert np.allclose(a["mcmc"][k], b["mcmc"][k], rtol=1e-12, atol=1e-14), k
    assert np.allclose(a["logWeights"], b["logWeights"], rtol=1e-12)


def test_oracle_surv_logposterior_and_score_match_reference_golden(oracle):
    g = _load("surv_config3")
    c = make_cohort("config3", n=200)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl, Wls = np.asfortranarray(np_twin.wlong(D, b, 0)), np.asfortranarray(np_twin.wlong(D, b, 1))
    L = oracle.lib()
    Bs, ga, al = np.array(ini["Bs_gammas"]), np.array(ini["gammas"]), np.array(ini["alphas"])
    tau, temp = float(g["tau"][0]), float(g["temp"][0])
    P = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))  # noqa: E731
    keep = [P(x) for x in (D["event"], D["W1"], D["W1s"], Bs, D["W2"], D["W2s"], ga, Wl, Wls, al, D["Pw"], pr["mean_Bs_gammas"],
                           pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    L.orc_logPosterior.restype = C.c_double
    L.orc_logPosterior.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 17 + [C.c_double] * 3
    lp = L.orc_logPosterior(temp, Wl.shape[0], Wls.shape[0], len(Bs), len(ga), len(al),
                            *[k.ctypes.data_as(_abi.c_double_p) for k in keep], tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
    assert abs(lp - g["logPosterior"][0]) <= 1e-12 * abs(g["logPosterior"][0])
    ev = np.asarray(D["event"])
    ecs = [np.ascontiguousarray(ev @ np.asarray(M)) for M in (D["W1"], D["W2"], Wl)]
    keep2 = [P(x) for x in (ecs[0], D["W1s"], Bs, ecs[1], D["W2s"], ga, ecs[2], Wls, al, D["Pw"], pr["mean_Bs_gammas"],
                            pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    sc = np.zeros(len(Bs) + len(ga) + len(al) + 1)
    L.orc_gradient_logPosterior.restype = None
    L.orc_gradient_logPosterior.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 16 + [C.c_double] * 3 + [_abi.c_double_p]
    L.orc_gradient_logPosterior(temp, Wls.shape[0], len(Bs), len(ga), len(al), *[k.ctypes.data_as(_abi.c_double_p) for k in keep2],
                                tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"], sc.ctypes.data_as(_abi.c_double_p))
    assert np.allclose(sc, g["score"], rtol=1e-11, atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_reproduces_reference_golden(cuda_lib, name):
    """CUDA path, through the C ABI, against the reference's own output (same random streams).
    Tolerance 1e-9: the reference's cumsum rowsum and Armadillo-order sums differ from the device's
    per-segment butterflies in the last bits; accept decisions are identical."""
    c, ctl, pr, chain = _inputs(name)
    r = cuda_lib.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=chain)
    _check_chain(r, _load(name), rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
def test_cuda_logposterior_and_score_match_reference_golden(cuda_lib):
    """jmb_logPosterior / jmb_gradient_logPosterior against the reference's own fast_logPost.cpp /
    fast_score.cpp output (tolerance 1e-10 relative, BASELINE.json north_star)."""
    g = _load("surv_config3")
    c = make_cohort("config3", n=200)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl, Wls = np_twin.wlong(D, b, 0), np_twin.wlong(D, b, 1)
    tau, temp = float(g["tau"][0]), float(g["temp"][0])
    with cuda_lib.Fit(D, c["Covs"]["b"], False) as fit:
        fit.set_wlong(Wl, Wls)
        lp = fit.logPosterior(temp, ini["Bs_gammas"], ini["gammas"], ini["alphas"], pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        sc = fit.gradient_logPosterior(temp, ini["Bs_gammas"], ini["gammas"], ini["alphas"], pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        # a second point (cache must not be reused) and finite differences of the device value
        Bs2 = np.asarray(ini["Bs_gammas"]) + 0.01
        lp2 = fit.logPosterior(temp, Bs2, ini["gammas"], ini["alphas"], pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        sc2 = fit.gradient_logPosterior(temp, Bs2, ini["gammas"], ini["alphas"], pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
        h = 1e-5
        a_p = np.asarray(ini["alphas"]).copy(); a_p[0] += h
        a_m = np.asarray(ini["alphas"]).copy(); a_m[0] -= h
        fd = (fit.logPosterior(temp, Bs2, ini["gammas"], a_p, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
              - fit.logPosterior(temp, Bs2, ini["gammas"], a_m, pr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])) / (2 * h)
    assert abs(lp - g["logPosterior"][0]) <= 1e-10 * abs(g["logPosterior"][0])
    assert np.allclose(sc, g["score"], rtol=1e-10, atol=1e-9)
    assert lp2 != lp and not np.allclose(sc2, sc)
    B, p = len(ini["Bs_gammas"]), len(ini["gammas"])
    assert np.isclose(sc2[B + p], fd, rtol=1e-5)


# ---- lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (control$update_RE = FALSE) ---------------------------------
def _check_wore(r, gold, rtol, nogam):
    for k, v in r["mcmc"].items():
        if nogam and k in ("gammas", "phi_gammas", "tau_gammas"):
            continue
        g = gold["mcmc_" + k]
        assert np.allclose(np.asarray(v).reshape(g.shape), g, rtol=rtol, atol=1e-12), (k, float(np.max(np.abs(np.asarray(v).reshape(g.shape) - g))))
    assert np.allclose(r["logPost"], gold["logPost"], rtol=rtol)
    sg = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    keep = [0, 2] if nogam else [0, 1, 2]
    assert np.allclose(sg[keep], gold["sigma_surv"][keep], rtol=rtol)
    for k in ("Bs_gammas", "alphas"):
        assert np.allclose(r["scales"]["Sigma"][k], gold["Sigma_" + k], rtol=max(rtol, 1e-8), atol=1e-14), k


@pytest.mark.parametrize("name", sorted(make_golden.WORE_CASES))
def test_oracle_wore_reproduces_reference_golden(oracle, name):
    c, D, ini, pr, cv, ctl, chain = make_golden.wore_inputs(name)
    r = oracle.lap_rwm_C_woRE(ini, D, pr, c["scales"], cv, ctl, c["Covs"]["b"], chain_id=chain)
    _check_wore(r, _load(name), 1e-11, make_golden.WORE_CASES[name][4])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(make_golden.WORE_CASES))
def test_cuda_wore_reproduces_reference_golden(cuda_lib, name):
    c, D, ini, pr, cv, ctl, chain = make_golden.wore_inputs(name)
    nogam = make_golden.WORE_CASES[name][4]
    Dfit = dict(c["Data"]) if not nogam else D        # the fit's designs (W2 dropped for the _nogammas flavour)
    with cuda_lib.Fit(Dfit if not nogam else D, c["Covs"]["b"], False) as fit:
        fit.set_wlong(D["Wlong"], D["Wlongs"])
        r = fit.lap_rwm_C_woRE(ini, pr, c["scales"], cv, ctl, chain_id=chain)
    _check_wore(r, _load(name), 1e-9, nogam)


# ---- survfitJM evaluators: log_post_RE_svft / survPred_svft_2 ------------------------------------------------
def test_oracle_svft_pieces_match_reference_golden(oracle):
    """log_post_RE_svft = log p(y|b) + log p(b) - H and survPred_svft_2 = -H (src/survfitJM_mvJMbayes.cpp:140-187)."""
    g = _load("svft_config3")
    c, _ = make_golden.svft_inputs()
    ini = c["inits"]
    oracle.set_rowsum_mode(1)
    r = oracle.eval_logpost_re(c["Data"], c["Covs"]["b"], False, ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    assert np.allclose(r["log_pyb"] + r["log_pb"] - r["H"], g["log_post"], rtol=1e-11)
    assert np.allclose(-r["H"], g["surv_pred"], rtol=1e-11)


@pytest.mark.gpu
def test_cuda_svft_matches_reference_golden(cuda_lib, oracle):
    g = _load("svft_config3")
    c, Dmiss = make_golden.svft_inputs()
    ini = c["inits"]
    with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False) as fit:
        fit.set_draw(c["Data"])
        lp = fit.log_post_RE_svft(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        sp = fit.survPred_svft_2(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    assert np.max(np.abs(lp - g["log_post"]) / np.abs(g["log_post"])) <= 1e-10
    assert np.max(np.abs(sp - g["surv_pred"]) / np.abs(g["surv_pred"])) <= 1e-10
    # new subjects without measurements of one outcome (log_longF_svft skips it, :109-110)
    with cuda_lib.Fit(Dmiss, c["Covs"]["b"], False) as fit:
        fit.set_draw(Dmiss)
        lp2 = fit.log_post_RE_svft(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        assert np.array_equal(np.diff(fit.row_ptr(1))[[3, 17, 159]], [0, 0, 0])
    full = np.ones(160, dtype=bool); full[[3, 17, 159]] = False
    assert np.max(np.abs(lp2[full] - g["log_post"][full]) / np.abs(g["log_post"][full])) <= 1e-10
    oracle.set_rowsum_mode(1)
    r = oracle.eval_logpost_re(Dmiss, c["Covs"]["b"], False, ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    assert np.allclose(lp2, r["log_pyb"] + r["log_pb"] - r["H"], rtol=1e-10)
    assert np.all(lp2[~full] != g["log_post"][~full])
