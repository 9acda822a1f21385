"""CPU suite: the oracle against known-answer vectors, the independent NumPy twin and analytic
self-checks; the C-ABI library loads and exports what include/jmb200.h declares.

The reference ships no golden vectors for this path: the only published
known-answer data that applies are the Random123 Philox4x32-10 vectors for the counter RNG.
"""
import ctypes as C
import os
import re

import numpy as np
import pytest

from jmbayes_b200 import _abi
from jmbayes_b200.cohorts import gauss_kronrod, make_cohort, spline_design
from oracle import np_twin

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


# ---- random streams -------------------------------------------------------------------------
def test_philox_known_answers(oracle):
    """Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, out in kat:
        assert tuple(oracle.philox(key[0], key[1], *ctr)) == out


def test_normal_and_gamma_draw_moments(oracle):
    L = oracle.lib()
    z = np.zeros(2)
    zs = []
    for i in range(20000):
        L.orc_rnorm_pair(7, 3, i, 11, 0, 0, z.ctypes.data_as(_abi.c_double_p))
        zs.extend(z.tolist())
    zs = np.array(zs)
    assert abs(zs.mean()) < 0.02 and abs(zs.var() - 1.0) < 0.03
    assert abs(np.mean(zs ** 4) - 3.0) < 0.15
    for shape, scale in [(8.5, 0.01), (1.0, 2.0), (0.6, 1.0)]:
        g = np.array([L.orc_rgamma(1, 2, i, 5, shape, scale) for i in range(20000)])
        assert np.all(g > 0)
        assert abs(g.mean() / (shape * scale) - 1.0) < 0.03
        assert abs(g.var() / (shape * scale ** 2) - 1.0) < 0.08


# ---- pieces against the independent NumPy twin ---------------------------------------------------
@pytest.mark.parametrize("cfg", ["config2", "config3", "config4"])
def test_oracle_matches_numpy_twin(oracle, cfg):
    c = make_cohort(cfg, n=700)
    D, ini = c["Data"], c["inits"]
    oracle.set_rowsum_mode(1)
    r1 = oracle.eval_logpost_re(D, c["Covs"]["b"], c["interval_cens"], ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    r2 = np_twin.eval_logpost_re(D, c["interval_cens"], ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    for k in r1:
        assert _rel(r1[k], r2[k]) < 1e-11, (k, _rel(r1[k], r2[k]))
    for which in (0, 1) + ((2,) if c["interval_cens"] else ()):
        W1 = oracle.wlong(D, c["Covs"]["b"], c["interval_cens"], ini["b"], which)
        W2 = np_twin.wlong(D, np.asarray(ini["b"]), which)
        assert np.allclose(W1, W2, rtol=1e-14, atol=1e-15)


def test_links_and_transforms_match_twin(oracle):
    c = make_cohort("config3", n=300)
    D, ini = c["Data"], c["inits"]
    D["fams"] = ["binomial", "binomial", "poisson"]
    D["links"] = ["probit", "cloglog", "log"]
    D["y"][0] = (D["y"][0] > 1.0).astype(float)
    b = 0.3 * np.asarray(ini["b"])
    oracle.set_rowsum_mode(1)
    for trans, shift in ((["expit", "exp", "identity", "identity", "exp", "identity"], 0.0),
                         (["sqrt", "log", "log2", "log10", "identity", "sqrt"], 6.0)):
        D["trans_Funs"] = trans
        for key in ("XXbetas", "XXsbetas"):
            D[key] = [np.asarray(x) + shift for x in c["Data"][key]]
        r1 = oracle.eval_logpost_re(D, c["Covs"]["b"], False, b, ini["Bs_gammas"], ini["gammas"], 0.1 * ini["alphas"])
        r2 = np_twin.eval_logpost_re(D, False, b, ini["Bs_gammas"], ini["gammas"], 0.1 * ini["alphas"])
        for k in r1:
            assert np.isfinite(r2[k]).all(), (trans, k)
            assert _rel(r1[k], r2[k]) < 1e-11, (trans, k)


def test_rowsum_literal_vs_exact(oracle):
    """The reference's cumsum-difference rowsum (src/fast_LAPRWM.cpp:18-25) against exact sums."""
    rng = np.random.default_rng(1)
    v = 1 + rng.poisson(9, 2000)
    last = np.cumsum(v) - 1
    x = rng.normal(size=int(v.sum())) * 10
    exact = np.add.reduceat(x, np.concatenate([[0], last[:-1] + 1]))
    oracle.set_rowsum_mode(0)
    lit = oracle.rowsum(x, last)
    oracle.set_rowsum_mode(1)
    ex = oracle.rowsum(x, last)
    assert np.allclose(ex, exact, rtol=1e-14, atol=1e-13)          # long double vs pairwise double
    assert np.allclose(lit, exact, rtol=0, atol=1e-9)        # accurate only to cumsum cancellation
    assert np.max(np.abs(lit - exact)) > 0                    # and measurably not exact


# ---- analytic self-checks -------------------------------------------------------------------------
def test_gauss_kronrod_integrates_polynomials_exactly():
    sk, wk = gauss_kronrod(15)
    assert abs(wk.sum() - 2.0) < 1e-15
    for deg in range(0, 23):
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert abs(np.sum(wk * sk ** deg) - exact) < 5e-15, deg
    sk7, wk7 = gauss_kronrod(7)
    for deg in range(0, 14):
        exact = 0.0 if deg % 2 else 2.0 / (deg + 1)
        assert abs(np.sum(wk7 * sk7 ** deg) - exact) < 5e-15, deg


def test_spline_design_properties():
    kn = np.sort(np.concatenate([np.repeat([0.5, 9.0], 4), np.linspace(1, 8, 13)]))
    x = np.concatenate([[0.5, 9.0, 0.1, 9.5], np.random.default_rng(0).uniform(0.5, 9.0, 500)])
    W = spline_design(kn, x)
    assert W.shape == (504, 17)
    assert np.allclose(W[[0, 1] + list(range(4, 504))].sum(axis=1), 1.0, atol=1e-14)   # partition of unity
    assert np.all(W[2:4] == 0.0)                                  # outer.ok = TRUE -> zero rows
    assert W[1, -1] == 1.0 and W[0, 0] == 1.0                     # closed at both boundary knots
    assert np.all((W != 0).sum(axis=1) <= 4) and np.all(W >= 0)


def _surv_pieces(c):
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl, Wls = np_twin.wlong(D, b, 0), np_twin.wlong(D, b, 1)
    return D, ini, pr, np.asfortranarray(Wl), np.asfortranarray(Wls)


def _orc_logpost(L, temp, D, Wl, Wls, Bs, g, a, pr, tau):
    n, ns = Wl.shape[0], Wls.shape[0]
    P = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64)).ctypes.data_as(_abi.c_double_p)  # noqa: E731
    keep = [np.asfortranarray(np.asarray(x, dtype=np.float64)) for x in
            (D["event"], D["W1"], D["W1s"], Bs, D["W2"], D["W2s"], g, Wl, Wls, a, D["Pw"], pr["mean_Bs_gammas"],
             pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    ptr = [k.ctypes.data_as(_abi.c_double_p) for k in keep]
    L.orc_logPosterior.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 17 + [C.c_double] * 3
    return L.orc_logPosterior(temp, n, ns, len(Bs), len(g), len(a), *ptr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])


def test_logposterior_and_score(oracle):
    """logPosterior (src/fast_logPost.cpp:8-26) against the twin; gradient_logPosterior
    (src/fast_score.cpp:8-50) against central differences of logPosterior."""
    c = make_cohort("config3", n=400)
    D, ini, pr, Wl, Wls = _surv_pieces(c)
    L = oracle.lib()
    Bs, g, a = np.array(ini["Bs_gammas"]), np.array(ini["gammas"]), np.array(ini["alphas"])
    tau, temp = 3.5, 1.0
    v = _orc_logpost(L, temp, D, Wl, Wls, Bs, g, a, pr, tau)
    v2 = np_twin.logPosterior(temp, D, Wl, Wls, Bs, g, a, pr, tau)
    assert abs(v - v2) <= 1e-11 * abs(v2)
    # score
    ev = np.asarray(D["event"])
    ecs = [np.ascontiguousarray(ev @ np.asarray(M)) for M in (D["W1"], D["W2"], Wl)]
    B, p, A = len(Bs), len(g), len(a)
    out = np.zeros(B + p + A + 1)
    keep = [np.asfortranarray(np.asarray(x, dtype=np.float64)) for x in
            (ecs[0], D["W1s"], Bs, ecs[1], D["W2s"], g, ecs[2], Wls, a, D["Pw"], pr["mean_Bs_gammas"], pr["Tau_Bs_gammas"],
             pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    ptr = [k.ctypes.data_as(_abi.c_double_p) for k in keep]
    L.orc_gradient_logPosterior.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 16 + [C.c_double] * 3 + [_abi.c_double_p]
    L.orc_gradient_logPosterior.restype = None
    L.orc_gradient_logPosterior(temp, Wls.shape[0], B, p, A, *ptr, tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"],
                                out.ctypes.data_as(_abi.c_double_p))
    theta = np.concatenate([Bs, g, a, [np.log(tau)]])

    def f(th):
        return _orc_logpost(L, temp, D, Wl, Wls, th[:B], th[B:B + p], th[B + p:B + p + A], pr, float(np.exp(th[-1])))
    num = np.zeros_like(theta)
    for j in range(theta.size):
        h = 1e-5 * max(1.0, abs(theta[j]))
        e = np.zeros_like(theta); e[j] = h
        num[j] = (f(theta + e) - f(theta - e)) / (2 * h)
    # the reference's score of log(tauBs) omits the +1 of the log-Jacobian-free parametrisation: compare as coded
    assert np.allclose(out[:-1], num[:-1], rtol=2e-6, atol=1e-5)
    assert np.isclose(out[-1], num[-1], rtol=2e-6, atol=1e-5)


# ---- chain-level behaviour of the oracle --------------------------------------------------------------
def test_oracle_chain_is_deterministic_and_keeps_reference_quirks(oracle):
    c = make_cohort("config2", n=150)
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=10, n_block=10, n_thin=1)
    a = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, chain_id=5)
    b = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, chain_id=5)
    assert np.array_equal(a["mcmc"]["b"], b["mcmc"]["b"])
    d = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, chain_id=6)
    assert not np.array_equal(a["mcmc"]["alphas"], d["mcmc"]["alphas"])
    # quirk 3: 0-based iter against keep = n_burnin+1.. -> with n_thin = 1 the last
    # output column is never written
    assert a["mcmc"]["alphas"].shape == (1, 10)
    assert np.all(a["mcmc"]["alphas"][:, -1] == 0.0) and np.all(a["mcmc"]["alphas"][:, :-1] != 0.0)
    assert np.all(a["mcmc"]["b"][:, :, -1] == 0.0)
    rates = a["extras"]["accept_b"]
    assert 0.05 < rates.mean() < 0.9


# ---- the C ABI ------------------------------------------------------------------------------------------
def test_cabi_library_exports_every_declared_symbol():
    from jmbayes_b200 import build
    lib = build.build_cuda()
    hdr = open(os.path.join(ROOT, "include", "jmb200.h")).read()
    names = sorted(set(re.findall(r"\b(jmb_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 15
    L = C.CDLL(lib)
    for nm in names:
        assert hasattr(L, nm), nm


def test_struct_layouts_match_the_header():
    """sizeof of the ctypes mirrors against the C compiler's view of include/jmb200.h."""
    import subprocess
    import tempfile
    src = '#include <stdio.h>\n#include "jmb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(jmb_data), sizeof(jmb_draw), sizeof(jmb_priors), sizeof(jmb_control), sizeof(jmb_chain_in), sizeof(jmb_chain_out), sizeof(jmb_eval_out), sizeof(jmb_xdesigns), sizeof(jmb_mll_out));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(td, "s"), os.path.join(td, "s.c")], check=True)
        out = subprocess.run([os.path.join(td, "s")], capture_output=True, text=True, check=True).stdout.split()
    got = [C.sizeof(x) for x in (_abi.JmbData, _abi.JmbDraw, _abi.JmbPriors, _abi.JmbControl, _abi.JmbChainIn, _abi.JmbChainOut, _abi.JmbEvalOut,
                                 _abi.JmbXDesigns, _abi.JmbMllOut)]
    assert got == [int(x) for x in out]


def test_no_gpu_means_a_loud_error():
    from jmbayes_b200 import api, build
    build.build_cuda()
    try:
        ndev = api.device_count()
    except api.JmbError:
        ndev = 0
    if ndev > 0:
        pytest.skip("a GPU is present")
    c = make_cohort("config2", n=20)
    with pytest.raises(api.JmbError):
        api.Fit(c["Data"], c["Covs"]["b"], False)
