"""GPU parity of jmb_mcmc_statistics (include/jmb200_summaries.h) against the restated R code
(oracle/oracle_stats.py), through the C ABI.  Tolerances (fp64, different summation orders): moments and quantiles 1e-12
relative, effective sizes 1e-9 relative (Levinson recursion of up to 30 orders), modes: the same cell of density()'s
1000-point grid for >= 99 % of the series and never more than one cell apart (the argmax of two nearly tied cells may
flip between the FFT convolution of the restatement and the direct sum of the kernel)."""
import numpy as np
import pytest

from jmbayes_b200 import summaries
from oracle import oracle_stats as st

pytestmark = pytest.mark.gpu


def _series(rng, S, M):
    """A mix of the shapes posterior draws take: autocorrelated, skewed, heavy-tailed, shifted far from zero, tiny scale."""
    e = rng.normal(size=(S, M))
    rho = rng.uniform(-0.3, 0.9, size=S)
    X = np.empty((S, M))
    X[:, 0] = e[:, 0]
    for m in range(1, M):
        X[:, m] = rho * X[:, m - 1] + e[:, m]
    k = S // 5
    X[k : 2 * k] = np.exp(0.6 * X[k : 2 * k])
    X[2 * k : 3 * k] = rng.standard_t(3, size=(k, M))
    X[3 * k : 4 * k] = 250.0 + 1e-3 * X[3 * k : 4 * k]
    X[4 * k :] *= rng.uniform(1e-4, 1e3, size=(S - 4 * k, 1))
    return X


def _compare(got, X, w, probs=(0.025, 0.975)):
    ref = st.statistics(X, w, probs)
    sd = ref["StDev"]
    scale = np.maximum(np.abs(ref["postMeans"]), sd)
    for name in ("postMeans", "postwMeans"):
        assert np.max(np.abs(got[name] - ref[name]) / scale) < 1e-12, name
    for name in ("StDev", "wStDev"):
        assert np.max(np.abs(got[name] / ref[name] - 1)) < 1e-11, name
    for name in ("CIs", "wCIs"):
        assert np.max(np.abs(got[name] - ref[name]) / scale[:, None]) < 1e-12, name
    assert np.array_equal(got["Pvalues"], ref["Pvalues"])
    assert np.max(np.abs(got["EffectiveSize"] / ref["EffectiveSize"] - 1)) < 1e-9
    assert np.max(np.abs(got["StErr"] / ref["StErr"] - 1)) < 1e-9
    cell = (X.max(1) - X.min(1) + 6 * 3 * st.bw_nrd(np.sort(X, 1), sd)) / 999
    off = np.abs(got["postModes"] - ref["postModes"]) / cell
    assert np.max(off) < 1.0 + 1e-6
    assert np.mean(off < 1e-6) >= 0.99


@pytest.mark.parametrize("M", [500, 200, 37, 1000])
def test_statistics_match_the_restated_r_code_in_both_layouts(cuda_lib, M):
    rng = np.random.default_rng(100 + M)
    S = 203
    X = _series(rng, S, M)
    w = st.stand(rng.normal(size=M) * 1.5)
    got, ms = summaries.series_statistics(np.ascontiguousarray(X), S, M, M, 1, w)       # draws contiguous (M x p matrices)
    assert ms > 0
    _compare(got, X, w)
    got2, _ = summaries.series_statistics(np.ascontiguousarray(X.T), S, M, 1, S, w)     # series contiguous (n x q x M)
    for name in summaries.NAMES:
        assert np.array_equal(got[name], got2[name], equal_nan=True), name


def test_degenerate_series_ties_and_zero_weights(cuda_lib):
    rng = np.random.default_rng(7)
    M = 120
    X = rng.normal(size=(6, M))
    X[0] = 4.0                                   # constant: sd 0, ESS 0, mode NA
    X[1] = 0.5 * np.arange(M) - 3                # exactly linear: ESS 0
    X[2] = np.round(X[2], 1)                     # ties
    X[3, ::2] = X[3, 1::2]                       # pairs of equal values
    w = st.stand(rng.normal(size=M))
    w[::7] = 0.0                                 # zero weights are dropped by wtd.quantile
    got, _ = summaries.series_statistics(np.ascontiguousarray(X), 6, M, M, 1, w, probs=(0.1, 0.3))
    ref = st.statistics(X, w, probs=(0.1, 0.3))
    assert got["EffectiveSize"][0] == 0 and got["EffectiveSize"][1] == 0
    assert np.isnan(got["postModes"][0]) and np.isnan(ref["postModes"][0])
    assert got["StDev"][0] == 0
    for name in ("CIs", "wCIs", "postMeans", "postwMeans"):
        assert np.allclose(got[name], ref[name], rtol=1e-12, atol=1e-13), name
    assert np.allclose(got["EffectiveSize"][2:], ref["EffectiveSize"][2:], rtol=1e-9)
    assert np.allclose(got["postModes"][1:], ref["postModes"][1:], rtol=0, atol=0.02)


def test_mcmc_list_with_the_reference_margins(cuda_lib):
    """statistics(mcmc): columns of M x p matrices, cells of M x q x q arrays, cells (i, j) of the n x q x M array b."""
    rng = np.random.default_rng(11)
    M, n, q = 300, 57, 4
    mcmc = {"alphas": rng.normal(0.4, 0.1, size=(M, 3)), "D": rng.normal(size=(M, q, q)) + 2.0,
            "b": rng.normal(size=(n, q, M)) * 0.7, "tau_Bs_gammas": rng.gamma(3.0, size=M)}
    ll = rng.normal(size=M)
    res = summaries.statistics(mcmc, LogLiks=ll)
    w = st.stand(ll)
    assert np.allclose(res["weights"], w, rtol=1e-14)
    s = res["statistics"]
    assert s["postMeans"]["alphas"].shape == (3,) and s["CIs"]["alphas"].shape == (2, 3)
    assert s["postMeans"]["D"].shape == (q, q) and s["CIs"]["D"].shape == (2, q, q)
    assert s["postMeans"]["b"].shape == (n, q) and s["wCIs"]["b"].shape == (2, n, q)
    assert np.allclose(s["postMeans"]["b"], mcmc["b"].mean(2), rtol=1e-12, atol=1e-15)
    assert np.allclose(s["StDev"]["D"], mcmc["D"].std(0, ddof=1), rtol=1e-12)
    assert np.allclose(s["CIs"]["b"][1], np.quantile(mcmc["b"], 0.975, axis=2), rtol=1e-12)
    assert np.isclose(s["postwMeans"]["tau_Bs_gammas"], mcmc["tau_Bs_gammas"] @ w, rtol=1e-12)
    ref = st.statistics(mcmc["b"].reshape(n * q, M, order="F"), w)
    assert np.allclose(s["EffectiveSize"]["b"].reshape(-1, order="F"), ref["EffectiveSize"], rtol=1e-9)
    assert np.allclose(s["wCIs"]["b"][0].reshape(-1, order="F"), ref["wCIs"][:, 0], rtol=1e-12)


def test_streamed_host_input_equals_one_launch(cuda_lib):
    """More series than one staging chunk holds (256 MB): the chunked host path equals per-series results."""
    rng = np.random.default_rng(13)
    M, S = 64, 700_000                                            # 358 MB of draws -> two chunks
    X = rng.normal(size=(M, S))                                   # series contiguous
    w = np.full(M, 1.0 / M)
    got, _ = summaries.series_statistics(X, S, M, 1, S, w, want=("postMeans", "StDev", "CIs"))
    assert np.allclose(got["postMeans"], X.mean(0), rtol=0, atol=1e-14)
    assert np.allclose(got["StDev"], X.std(0, ddof=1), rtol=1e-12)
    pick = np.r_[0:50, 524280:524300, S - 50 : S]
    assert np.allclose(got["CIs"][pick], np.quantile(X[:, pick], [0.025, 0.975], axis=0).T, rtol=1e-12)
