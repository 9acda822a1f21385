"""GPU parity of the fused survfitJM kernel (BASELINE config 5) against the restated R loop, through the C ABI
of include/jmb200_survfit.h.  Tolerances: log-posterior driven quantities 1e-9 relative when the inputs of a
stage are identical (sampler from given modes / Hessians); the BFGS mode search is an iterative fp64 process whose
two implementations sum in different orders, so modes are compared at 1e-6 absolute and Hessians at 1e-5 relative."""
import os

import numpy as np
import pytest

from jmbayes_b200 import survfit as sv
from jmbayes_b200.cohorts import make_cohort

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def osv(oracle):
    from oracle import oracle_svft
    oracle_svft.lib()
    return oracle_svft


def _case(config, n, L, M, seed=None):
    c = make_cohort(config, n=n, seed=seed)
    d = sv.newdata_designs(c, L=L)
    means, draws = sv.posterior_draws(c, M=M)
    return c, d, means, draws


def _cmp_modes(g, o, n):
    assert np.array_equal(g["convergence"], o["convergence"])
    assert np.max(np.abs(g["modes"] - o["modes"])) < 1e-6
    rel = np.abs(g["hessians"] - o["hessians"]) / np.maximum(np.abs(o["hessians"]), 1e-2)
    assert np.max(rel) < 1e-5
    # the two optimisers walk the same path: identical evaluation counts for (nearly) every subject
    assert np.mean(np.all(g["counts"] == o["counts"], axis=0)) > 0.8


@pytest.mark.parametrize("config,n,L,M", [("config3", 96, 10, 50), ("config2", 64, 4, 30), ("config4", 40, 10, 24)])
def test_modes_and_sampler_match_the_oracle(cuda_lib, osv, config, n, L, M):
    c, d, means, draws = _case(config, n, L, M)
    q = d["q"]
    with sv.SurvFit(d) as h:
        g = h.modes(means)
        o = osv.modes(d, means)
        _cmp_modes(g, o, n)
        # sampler from IDENTICAL modes / Hessians: draw-for-draw agreement
        gs = h.sample(draws, o["modes"], o["hessians"])
        os_ = osv.sample(d, draws, o["modes"], o["hessians"])
        assert np.array_equal(gs["accept_rate"], os_["accept_rate"])
        assert np.allclose(gs["b_last"], os_["b_last"], rtol=1e-9, atol=1e-11)
        assert np.allclose(gs["full_results"], os_["full_results"], rtol=1e-9, atol=0)
        assert np.allclose(gs["summaries"], os_["summaries"], rtol=1e-9, atol=0)
        assert gs["full_results"].shape == (L, M, n) and gs["summaries"].shape == (L, 4, n) and gs["modes"].shape == (n, q)
        # the whole loop in one call: same as the two stages chained on the device
        ga = h.survfitJM(means, draws)
        assert np.array_equal(ga["modes"], g["modes"]) and np.array_equal(ga["hessians"], g["hessians"])
        gb = h.sample(draws, g["modes"], g["hessians"])
        assert np.array_equal(ga["full_results"], gb["full_results"]) and np.array_equal(ga["summaries"], gb["summaries"])
        # end to end against the oracle: a decision can only flip where u sits within ~1e-7 of the ratio
        oa = osv.survfitJM(d, means, draws)
        diff = np.max(np.abs(ga["summaries"] - oa["summaries"]), axis=(0, 1))
        assert np.mean(diff > 1e-5) <= 0.05
        info = h.info()
        assert info["launches"] >= 2 and info["last_kernel_ms"] > 0


def test_generic_kernel_matches_the_specialised_one(cuda_lib, osv, monkeypatch):
    c, d, means, draws = _case("config3", 64, 5, 20)
    with sv.SurvFit(d) as h:
        a = h.survfitJM(means, draws)
    monkeypatch.setenv("JMB_SVFT_KERNEL", "generic")
    with sv.SurvFit(d) as h:
        b = h.survfitJM(means, draws)
    monkeypatch.delenv("JMB_SVFT_KERNEL")
    assert np.allclose(a["modes"], b["modes"], rtol=0, atol=1e-7)
    diff = np.max(np.abs(a["summaries"] - b["summaries"]), axis=(0, 1))
    assert np.mean(diff > 1e-6) <= 0.05


def test_ragged_horizons_missing_outcomes_log_scale(cuda_lib, osv):
    """Subjects with fewer prediction intervals (survTimes > last.time filters them, :91), a subject without
    measurements of one outcome (log_longF_svft skips it), dense (non-banded) bases, log = TRUE, other CI levels."""
    c, d, means, draws = _case("config3", 50, 6, 16)
    d = dict(d)
    nh = np.full(50, 6, dtype=np.int32); nh[[1, 7]] = 2; nh[11] = 0
    d["n_horizons"] = nh
    keep = ~np.isin(d["idL"][1], [3, 20])
    d["y_long"] = list(d["y_long"]); d["X_long"] = list(d["X_long"]); d["Z_long"] = list(d["Z_long"]); d["idL"] = list(d["idL"])
    d["y_long"][1] = d["y_long"][1][keep]; d["idL"][1] = d["idL"][1][keep]
    d["X_long"][1] = np.asfortranarray(np.asarray(d["X_long"][1])[keep]); d["Z_long"][1] = np.asfortranarray(np.asarray(d["Z_long"][1])[keep])
    # a non-trivial interaction design and a transformed term -> the generic kernel, stored U columns
    d["Us"] = list(d["Us"]); d["Us_h"] = list(d["Us_h"]); d["trans_Funs"] = list(d["trans_Funs"])
    rng = np.random.default_rng(5)
    d["Us"][2] = np.asfortranarray(rng.uniform(0.5, 1.5, size=(50 * 15, 1)))
    d["Us_h"][2] = np.asfortranarray(rng.uniform(0.5, 1.5, size=(50 * 6 * 15, 1)))
    d["trans_Funs"][3] = "expit"
    ctl = dict(log=True, CI_levels=(0.1, 0.9), seed=11)
    with sv.SurvFit(d) as h:
        g = h.modes(means, ctl)
        o = osv.modes(d, means, ctl)
        _cmp_modes(g, o, 50)
        gs = h.sample(draws, o["modes"], o["hessians"], ctl)
    os_ = osv.sample(d, draws, o["modes"], o["hessians"], ctl)
    assert np.allclose(gs["full_results"], os_["full_results"], rtol=1e-9, atol=0, equal_nan=True)
    assert np.allclose(gs["summaries"], os_["summaries"], rtol=1e-9, atol=0, equal_nan=True)
    assert np.all(np.isnan(gs["summaries"][2:, :, 1])) and np.all(np.isfinite(gs["summaries"][:2, :, 1]))
    assert np.all(np.isnan(gs["summaries"][:, :, 11]))
    assert np.all(gs["full_results"][:, :, 0] < 0)            # log scale: per-interval log survival


def test_bad_hessian_and_sharded_keys(cuda_lib, osv):
    c, d, means, draws = _case("config3", 40, 3, 12)
    o = osv.modes(d, means)
    Hbad = o["hessians"].copy(); Hbad[:, :, 6] = -np.eye(6)
    with sv.SurvFit(d) as h:
        g = h.sample(draws, o["modes"], Hbad)
        whole = h.sample(draws, o["modes"], o["hessians"])
    assert g["convergence"][6] == 20 and np.all(np.isnan(g["summaries"][:, :, 6])) and np.all(np.isfinite(g["summaries"][:, :, 7]))
    # subjects sharded over ranks reproduce the single-GPU result (draws are keyed by the global subject index)
    sel = np.arange(25, 40)
    c2 = make_cohort("config3", n=40)
    d2 = sv.newdata_designs(c2, L=3, subjects=sel)
    with sv.SurvFit(d2, subject_offset=25) as h2:
        part = h2.sample(draws, o["modes"][sel], o["hessians"][:, :, sel])
    assert np.array_equal(part["full_results"], whole["full_results"][:, :, sel])


def test_basis_from_knots_equals_dense_basis(cuda_lib, osv):
    """with W1s / W1s_h = NULL the library evaluates splineDesign(knots, x, ord = 4, outer.ok = TRUE)
    itself (band form, de Boor) from the quadrature times; results must equal the dense-basis path, including the
    all-zero rows left of the first boundary knot."""
    c = make_cohort("config3", n=80)
    dense = sv.newdata_designs(c, L=4)
    lean = sv.newdata_designs(c, L=4, dense_basis=False)
    assert lean["W1s"] is None and lean["W1s_h"] is None
    assert np.any(np.asarray(dense["W1s"]).sum(axis=1) == 0.0)        # the edge case is present
    means, draws = sv.posterior_draws(c, M=20)
    with sv.SurvFit(dense) as h:
        a = h.survfitJM(means, draws)
    with sv.SurvFit(lean) as h:
        b = h.survfitJM(means, draws)
    assert np.allclose(a["modes"], b["modes"], rtol=0, atol=1e-12)
    assert np.allclose(a["summaries"], b["summaries"], rtol=1e-12, atol=0)
    bad = dict(lean); bad["knots"] = lean["knots"][::-1].copy()
    with pytest.raises(cuda_lib.JmbError):
        sv.SurvFit(bad)
