"""marglogLik2 (R/margLogLik2.R): the restated Laplace step (oracle/jmb_oracle_mll.c) against SciPy / NumPy,
and the native driver of the CUDA library (jmb_marglogLik2) against the restatement."""
import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort


def _case(oracle, n=400, cfg="config3"):
    c = make_cohort(cfg, n=n)
    D = dict(c["Data"]); ini = c["inits"]
    D["Wlong"] = oracle.wlong(c["Data"], c["Covs"]["b"], False, ini["b"], 0)
    D["Wlongs"] = oracle.wlong(c["Data"], c["Covs"]["b"], False, ini["b"], 1)
    th = dict(Bs_gammas=ini["Bs_gammas"], gammas=ini["gammas"], alphas=ini["alphas"], tau_Bs_gammas=200.0)
    return c, D, th


def test_nearpd_and_eigen(oracle):
    rng = np.random.default_rng(4)
    A = rng.normal(size=(25, 25)); S = A @ A.T / 25 + 0.05 * np.eye(25)
    assert np.allclose(oracle.nearPD(S), S, rtol=1e-9, atol=1e-12)                  # PD input: unchanged
    ind = S.copy(); w, V = np.linalg.eigh(ind); w[:3] = [-0.5, -0.1, -1e-3]; ind = V @ np.diag(w) @ V.T
    ind = 0.5 * (ind + ind.T)
    X = oracle.nearPD(ind)
    ev = np.linalg.eigvalsh(X)
    assert np.allclose(X, X.T) and ev.min() > 0 and ev.min() >= 0.9e-8 * ev.max()     # posd.tol = 1e-08
    wpos = np.where(w > 1e-6 * w.max(), w, 0.0)                                      # one projection step dominates here
    assert np.linalg.norm(X - ind) <= 2.0 * np.linalg.norm(V @ np.diag(wpos) @ V.T - ind) + 1e-9
    assert np.allclose(oracle.nearPD(np.array([[-3.0]])), [[3.0]])


def test_oracle_marglogLik2_against_scipy(oracle):
    from scipy.optimize import minimize
    c, D, th = _case(oracle)
    pr = c["priors"]
    r = oracle.marglogLik2(th, D, pr, fixed_tau_Bs_gammas=True)
    assert r["convergence"] == 0 and r["counts"][1] < 100
    B, p, A = 17, 2, 6
    W1, W1s, W2, W2s = (np.asarray(D[k]) for k in ("W1", "W1s", "W2", "W2s"))
    Wl, Wls, Pw, ev = np.asarray(D["Wlong"]), np.asarray(D["Wlongs"]), np.asarray(D["Pw"]), np.asarray(D["event"])

    def fn(x):      # -logPosterior (src/fast_logPost.cpp:8-26), NumPy
        Bs, g, a = x[:B], x[B:B + p], x[B + p:]
        lh = W1 @ Bs + W2 @ g + Wl @ a
        H = Pw * np.exp(W1s @ Bs + W2s @ g + Wls @ a)
        val = (ev * lh).sum() - H.sum() - 0.5 * 200.0 * Bs @ pr["Tau_Bs_gammas"] @ Bs - 0.5 * g @ pr["Tau_gammas"] @ g \
            - 0.5 * a @ pr["Tau_alphas"] @ a + (pr["A_tau_Bs_gammas"] - 1) * np.log(200.0) - pr["B_tau_Bs_gammas"] * 200.0
        return -val
    x0 = np.concatenate([th["Bs_gammas"], th["gammas"], th["alphas"]])
    res = minimize(fn, x0, method="BFGS", options=dict(gtol=1e-8, maxiter=2000))
    assert abs(r["opt_value"] - fn(r["par"])) <= 1e-9 * abs(r["opt_value"])
    assert r["opt_value"] <= res.fun + 1e-5 * max(1.0, abs(res.fun))                   # optim stops on reltol of the value
    sign, logdet = np.linalg.slogdet(r["hessian"])
    assert sign > 0 and abs(r["value"] - (0.5 * (25 * np.log(2 * np.pi) - logdet) - r["opt_value"])) <= 1e-9 * abs(r["value"])
    # Covs: blocks of the (nearPD'ed) inverse of a finite-difference Hessian of fn at the mode
    h = 1e-4; Hn = np.zeros((25, 25)); x = r["par"]
    g0 = lambda y: np.array([(fn(y + h * e) - fn(y - h * e)) / (2 * h) for e in np.eye(25)])
    for i in range(25):
        e = np.zeros(25); e[i] = h
        Hn[:, i] = (g0(x + e) - g0(x - e)) / (2 * h)
    iH = np.linalg.inv(0.5 * (Hn + Hn.T))
    assert np.allclose(r["Covs"]["alphas"], iH[19:, 19:], rtol=2e-2, atol=1e-6)
    assert np.allclose(r["Covs"]["Bs_gammas"], iH[:17, :17], rtol=5e-2, atol=1e-5)
    assert np.all(np.linalg.eigvalsh(r["Covs"]["Bs_gammas"]) > 0)
    # free tau: the last coordinate is log tau_Bs_gammas
    th2 = dict(th); th2["tau_Bs_gammas"] = np.log(200.0)
    r2 = oracle.marglogLik2(th2, D, pr, fixed_tau_Bs_gammas=False)
    assert r2["par"].shape == (26,) and np.isfinite(r2["value"]) and "Covs" not in r2


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["config2", "config3"])
def test_cuda_marglogLik2_matches_the_restatement(cuda_lib, oracle, cfg):
    c, D, th = _case(oracle, n=600, cfg=cfg)
    pr = c["priors"]
    with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False) as fit:
        fit.set_wlong(D["Wlong"], D["Wlongs"])
        for fixed in (True, False):
            t = dict(th)
            if not fixed:
                t["tau_Bs_gammas"] = np.log(200.0)
            g = fit.marglogLik2(t, pr, fixed_tau_Bs_gammas=fixed)
            o = oracle.marglogLik2(t, D, pr, fixed_tau_Bs_gammas=fixed)
            assert g["convergence"] == o["convergence"] == 0
            assert np.allclose(g["par"], o["par"], rtol=0, atol=1e-6)
            assert abs(g["opt_value"] - o["opt_value"]) <= 1e-10 * abs(o["opt_value"])
            assert abs(g["value"] - o["value"]) <= 1e-7 * abs(o["value"])
            assert np.allclose(g["hessian"], o["hessian"], rtol=1e-5, atol=1e-6 * np.abs(o["hessian"]).max())
            if fixed:
                for k in ("Bs_gammas", "gammas", "alphas"):
                    assert np.allclose(g["Covs"][k], o["Covs"][k], rtol=1e-4, atol=1e-6 * np.abs(o["Covs"][k]).max()), k
        with pytest.raises(cuda_lib.JmbError, match="failed to find initial values"):
            bad = dict(th); bad["alphas"] = np.full_like(th["alphas"], 400.0)             # exp overflow: non-finite start,
            fit.marglogLik2(bad, pr, fixed_tau_Bs_gammas=True)                            # and at the jittered restart too
        if cfg == "config2":
            # a start exactly on the edge of the finite region: exp() overflows at the starting values, the jittered restart
            # of R/margLogLik2.R:95-100 lands inside and BFGS reaches the same mode
            edge = dict(th)
            lo, hi = 0.0, 400.0
            for _ in range(60):                                                        # bisect the overflow threshold in alpha
                mid = 0.5 * (lo + hi)
                t = dict(th); t["alphas"] = np.full_like(th["alphas"], mid)
                v = fit.logPosterior(1.0, t["Bs_gammas"], t["gammas"], t["alphas"], pr, 200.0, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
                lo, hi = (mid, hi) if np.isfinite(v) else (lo, mid)
            found = None
            for a0 in np.linspace(hi, hi + 0.5, 26):                                   # just above: restart jitter (sd 0.1) may rescue
                edge["alphas"] = np.full_like(th["alphas"], a0)
                try:
                    r = fit.marglogLik2(edge, pr, fixed_tau_Bs_gammas=True)
                except cuda_lib.JmbError:
                    continue
                if r["restarts"] == 1:
                    found = r
                    break
            ref = fit.marglogLik2(th, pr, fixed_tau_Bs_gammas=True)
            assert ref["restarts"] == 0
            if found is not None:        # the fixed jitter stream moved alpha below the threshold for one of the starts
                assert found["convergence"] in (0, 1) and np.isfinite(found["value"])
