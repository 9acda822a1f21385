"""CPU tests of the posterior-summaries restatement (oracle/oracle_stats.py) and of the C ABI of
include/jmb200_summaries.h (symbols, struct layout, loud failure without a GPU).  R is absent: the restatement is checked
against independent NumPy / SciPy implementations and analytic cases (parity unpinned, see the module header)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import oracle_stats as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ar1(rng, S, M, rho, loc=0.0):
    e = rng.normal(size=(S, M))
    X = np.empty((S, M))
    X[:, 0] = e[:, 0] / np.sqrt(1 - rho * rho)
    for m in range(1, M):
        X[:, m] = rho * X[:, m - 1] + e[:, m]
    return X + loc


def test_moments_and_quantiles_against_numpy():
    rng = np.random.default_rng(0)
    X = rng.gamma(2.0, 1.5, size=(40, 333)) - 2.0
    w = st.stand(rng.normal(size=333) * 2)
    assert abs(w.sum() - 1) < 1e-14
    assert np.allclose(st.post_means(X), X.mean(1), rtol=1e-14)
    assert np.allclose(st.st_dev(X), X.std(1, ddof=1), rtol=1e-13)
    assert np.allclose(st.post_wmeans(X, w), X @ w, rtol=1e-13)
    assert np.allclose(st.w_st_dev(X, w) ** 2, [np.cov(x, aweights=w) for x in X], rtol=1e-12)
    assert np.allclose(st.cis(X), np.quantile(X, [0.025, 0.975], axis=1).T, rtol=1e-13)
    assert np.allclose(st.p_values(X), 2 * np.minimum((X >= 0).mean(1), (X < 0).mean(1)))


def test_levinson_solves_the_yule_walker_equations():
    from scipy.linalg import solve_toeplitz
    rng = np.random.default_rng(1)
    x = _ar1(rng, 1, 700, 0.7)[0]
    M = x.size
    xc = x - x.mean()
    om = st.order_max(M)
    assert om == 28
    r = np.array([(xc[: M - l] * xc[l:]).sum() / M for l in range(om + 1)])
    coefs, v = st.levinson(r)
    for l in (1, 2, 7, om):
        phi = solve_toeplitz(r[:l], r[1 : l + 1])
        assert np.allclose(coefs[l, 1 : l + 1], phi, atol=1e-12)
        assert np.isclose(v[l], r[0] - phi @ r[1 : l + 1], rtol=1e-11)


def test_effective_size_of_known_processes():
    rng = np.random.default_rng(2)
    M = 600
    iid = rng.normal(size=(300, M))
    assert abs(np.median(st.effective_size(iid)) / M - 1) < 0.08
    rho = 0.5
    ar = _ar1(rng, 300, M, rho, loc=3.0)
    ess = st.effective_size(ar)
    assert abs(np.median(ess) / (M * (1 - rho) / (1 + rho)) - 1) < 0.15
    assert np.allclose(st.std_err(ar), st.st_dev(ar) / np.sqrt(ess))
    # a constant and an exactly linear series have a zero spectrum -> effective size 0, standard error Inf / NaN
    deg = np.stack([np.full(M, 2.5), 1.0 + 0.25 * np.arange(M)])
    assert np.array_equal(st.effective_size(deg), [0.0, 0.0])


def test_weighted_quantiles_follow_hmisc_i_over_n_plus_1():
    # five points, weights 1..5 (not normalised): n = 15, ecdf = cumsum / 16 = (1, 3, 6, 10, 15) / 16, (0, x1) prepended
    x = np.array([[3.0, 1.0, 2.0, 5.0, 4.0]])
    w = np.array([3.0, 1.0, 2.0, 5.0, 4.0])
    q = st.w_cis(x, w, probs=(0.25, 0.5))[0]
    assert np.isclose(q[0], 2.0 + (0.25 - 3 / 16) / (3 / 16))        # between (3/16, 2) and (6/16, 3)
    assert np.isclose(q[1], 3.0 + (0.5 - 6 / 16) / (4 / 16))         # between (6/16, 3) and (10/16, 4)
    # beyond the last ecdf value: rule = 2 -> the largest draw; below the first: the smallest
    assert st.w_cis(x, w, probs=(0.01, 0.99))[0].tolist() == [1.0, 5.0]
    # ties are one point carrying the summed weight; zero weights are dropped
    x2 = np.array([[1.0, 2.0, 2.0, 3.0, 9.0]])
    w2 = np.array([1.0, 1.0, 1.0, 1.0, 0.0])
    q2 = st.w_cis(x2, w2, probs=(0.5, 0.9))[0]                       # knots (1/5, 1), (3/5, 2), (4/5, 3)
    assert np.isclose(q2[0], 1.0 + (0.5 - 0.2) / 0.4)
    assert q2[1] == 3.0
    # weights that sum to one (the reference's use): the ecdf ends at 1/2, the upper limit is the maximum
    rng = np.random.default_rng(3)
    X = rng.normal(size=(6, 200))
    wn = st.stand(rng.normal(size=200))
    q3 = st.w_cis(X, wn)
    assert np.array_equal(q3[:, 1], X.max(1))
    assert np.all((q3[:, 0] > X.min(1)) & (q3[:, 0] < np.median(X, 1)))


def test_binned_density_and_mode():
    rng = np.random.default_rng(4)
    x = rng.normal(1.5, 2.0, size=400)
    xs = np.sort(x)[None, :]
    bw = st.ADJUST * st.bw_nrd(xs, st.st_dev(x[None, :]))[0]
    xg, yg = st.density_grid(x, bw)
    assert xg.size == 1000 and np.isclose(xg[0], x.min() - 3 * bw) and np.isclose(xg[-1], x.max() + 3 * bw)
    assert abs(np.trapezoid(yg, xg) - 1.0) < 2e-3                    # cut = 3 bandwidths holds all but ~0.1 % of the mass
    # against the exact Gaussian kernel density estimate: binning error is O(cell^2), and density.default spaces the kernel
    # ordinates 2 (up - lo) / (2n - 1) apart but the cells (up - lo) / (n - 1), a 0.05 % narrower kernel
    exact = np.exp(-0.5 * ((xg[:, None] - x[None, :]) / bw) ** 2).sum(1) / (x.size * bw * np.sqrt(2 * np.pi))
    assert np.max(np.abs(yg - exact)) < 1e-3 * exact.max()
    mode = st.post_modes(x[None, :])[0]
    assert abs(mode - xg[np.argmax(exact)]) <= (xg[1] - xg[0]) * 1.01
    # a symmetric sample has its mode at the centre; a constant series has bw = 0 -> NA
    sym = np.concatenate([-np.linspace(0.1, 3, 150) ** 1.5, np.linspace(0.1, 3, 150) ** 1.5])[None, :] + 7.0
    assert abs(st.post_modes(sym)[0] - 7.0) < 0.02
    assert np.isnan(st.post_modes(np.full((1, 50), 3.0))[0])


def test_numpy_and_c_restatements_agree(oracle):
    """Two independently written restatements of the R code (oracle_stats.py, jmb_oracle_stats.c): <= 1e-11 relative, modes in
    the same cell of the 1000-point grid (NumPy's FFT against the C radix-2 FFT)."""
    rng = np.random.default_rng(6)
    M, S = 500, 240
    X = _ar1(rng, S, M, 0.6)
    X[:40] = np.exp(0.6 * X[:40])
    X[40:80] = rng.standard_t(3, size=(40, M))
    X[80:120] = 250.0 + 1e-3 * X[80:120]
    X[120] = 3.0
    X[121] = 0.5 * np.arange(M)
    X[122] = np.round(X[122], 1)
    w = st.stand(rng.normal(size=M) * 1.5)
    w[::9] = 0.0
    a, c = st.statistics(X, w), st.statistics_c(X, w)
    scale = np.maximum(np.abs(a["postMeans"]), a["StDev"])
    for name in ("postMeans", "postwMeans", "StDev", "wStDev"):
        assert np.max(np.abs(a[name] - c[name]) / np.maximum(scale, 1e-300)) < 1e-11, name
    for name in ("CIs", "wCIs"):
        assert np.max(np.abs(a[name] - c[name]) / scale[:, None]) < 1e-12, name
    assert np.array_equal(a["Pvalues"], c["Pvalues"])
    ok = a["EffectiveSize"] > 0
    assert np.array_equal(ok, c["EffectiveSize"] > 0) and not ok[120] and not ok[121]
    assert np.max(np.abs(a["EffectiveSize"][ok] / c["EffectiveSize"][ok] - 1)) < 1e-10
    assert np.isnan(a["postModes"][120]) and np.isnan(c["postModes"][120])
    fin = np.isfinite(a["postModes"])
    cell = (X.max(1) - X.min(1) + 18 * st.bw_nrd(np.sort(X, 1), a["StDev"])) / 999
    assert np.all(np.abs(a["postModes"][fin] - c["postModes"][fin]) <= 1e-6 * cell[fin] + 1e-12 * scale[fin])


def test_cabi_exports_the_summaries_symbols_and_struct_layout():
    from jmbayes_b200 import build, summaries
    lib = build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "jmb200_summaries.h")).read()
    names = sorted(set(re.findall(r"\b(jmb_[a-z_0-9A-Z]+)\s*\(", hdr)))
    assert {"jmb_mcmc_statistics", "jmb_stand_weights"} <= set(names)
    L = C.CDLL(lib)
    for nm in names:
        assert hasattr(L, nm), nm
    src = '#include <stdio.h>\n#include "jmb200_summaries.h"\nint main(){printf("%zu\\n", sizeof(jmb_stats_out));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")], check=True)
        size = int(subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout)
    assert size == C.sizeof(summaries.StatsOut)


def test_stand_weights_host_function_and_loud_failure_without_a_gpu():
    import torch
    from jmbayes_b200 import api, summaries
    rng = np.random.default_rng(5)
    ll = rng.normal(size=37) * 30 - 4000.0
    assert np.allclose(summaries.stand(ll), st.stand(ll), rtol=1e-14, atol=0)
    X = rng.normal(size=(3, 64))
    with pytest.raises(api.JmbError):
        summaries.series_statistics(X, 3, 1025, 1025, 1, np.ones(1025))          # M beyond JMB_STATS_MAX_M
    with pytest.raises(api.JmbError):
        summaries.series_statistics(X, 3, 64, 7, 3, np.ones(64))                 # neither layout
    if not torch.cuda.is_available():
        with pytest.raises(api.JmbError, match="no CUDA device"):
            summaries.series_statistics(X, 3, 64, 64, 1, np.ones(64))


def test_statistics_margins_and_layouts_with_a_stand_in_backend(monkeypatch):
    """Host logic of summaries.statistics(): apply margins, strides and output shapes, checked on CPU by serving
    series_statistics from the restatement (the GPU test of the same name runs the real kernel)."""
    from jmbayes_b200 import summaries

    def fake(x, S, M, ss, ds, weights, probs=(0.025, 0.975), device=0, want=summaries.NAMES):
        flat = np.asarray(x).reshape(-1)
        X = np.stack([flat[s * ss : s * ss + (M - 1) * ds + 1 : ds] for s in range(S)])
        r = st.statistics_c(X, weights, probs)
        return {k: r[k] for k in want}, 0.0

    monkeypatch.setattr(summaries, "series_statistics", fake)
    monkeypatch.setattr(summaries, "stand", st.stand)
    rng = np.random.default_rng(12)
    M, n, q = 60, 7, 3
    mcmc = {"alphas": rng.normal(size=(M, 2)), "D": rng.normal(size=(M, q, q)), "b": rng.normal(size=(n, q, M)),
            "tau": rng.gamma(2.0, size=M), "none": None}
    res = summaries.statistics(mcmc, LogLiks=rng.normal(size=M))["statistics"]
    assert "none" not in res["postMeans"]
    assert np.allclose(res["postMeans"]["alphas"], mcmc["alphas"].mean(0))
    assert np.allclose(res["postMeans"]["D"], mcmc["D"].mean(0))
    assert np.allclose(res["postMeans"]["b"], mcmc["b"].mean(2))
    assert np.isclose(res["postMeans"]["tau"], mcmc["tau"].mean())
    assert res["CIs"]["b"].shape == (2, n, q) and res["CIs"]["D"].shape == (2, q, q) and res["CIs"]["tau"].shape == (2,)
    assert np.allclose(res["CIs"]["b"][0], np.quantile(mcmc["b"], 0.025, axis=2))
    assert np.allclose(res["CIs"]["D"][1], np.quantile(mcmc["D"], 0.975, axis=0))
    assert np.allclose(res["StDev"]["alphas"], mcmc["alphas"].std(0, ddof=1))
