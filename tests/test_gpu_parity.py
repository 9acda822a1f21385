"""CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance: 1e-10 relative in fp64 on every per-subject piece of the b-step target
(BASELINE.json north_star), measured against the exact-rowsum oracle (the
reference's own cumsum-difference rowsum is less accurate than that at scale); segment boundaries
bit-exact.
"""
import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def _rel(a, b, floor=1e-12):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def _variant(c, links=None, trans=None):
    D = c["Data"]
    if links:
        D["links"] = list(links)
    if trans:
        D["trans_Funs"] = list(trans)
    return c


@pytest.mark.parametrize("cfg,n", [("config2", 3000), ("config3", 3000), ("config4", 3000), ("config3", 37)])
def test_logpost_pieces_match_oracle(cuda_lib, oracle, cfg, n):
    c = make_cohort(cfg, n=n)
    D, ini = c["Data"], c["inits"]
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], c["interval_cens"], ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    with cuda_lib.Fit(D, c["Covs"]["b"], c["interval_cens"]) as fit:
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        for k in range(len(D["y"])):   # segment boundaries bit-exact
            rp = fit.row_ptr(k)
            assert np.array_equal(rp[1:] - 1, np.asarray(D["idL2"][k], dtype=np.int64))
            assert rp[0] == 0
    for key in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb"):
        assert np.isfinite(got[key]).all(), key
        assert _rel(got[key], ref[key]) <= RTOL, (key, _rel(got[key], ref[key]))
    tot_ref = ref["log_pyb"] + ref["log_ptb"] + ref["log_pb"]
    tot = got["log_pyb"] + got["log_ptb"] + got["log_pb"]
    assert _rel(tot, tot_ref) <= RTOL


def test_links_and_transforms(cuda_lib, oracle):
    c = make_cohort("config3", n=500)
    # probit + cloglog links, expit/exp/sqrt/log-type transforms of positive terms
    c["Data"]["fams"] = ["binomial", "binomial", "poisson"]
    c["Data"]["links"] = ["probit", "cloglog", "log"]
    c["Data"]["y"][0] = (c["Data"]["y"][0] > 1.0).astype(float)
    c["Data"]["trans_Funs"] = ["expit", "exp", "identity", "identity", "exp", "identity"]
    D, ini = c["Data"], c["inits"]
    b = 0.3 * np.asarray(ini["b"])
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], False, b, ini["Bs_gammas"], ini["gammas"], 0.1 * ini["alphas"])
    with cuda_lib.Fit(D, c["Covs"]["b"], False) as fit:
        fit.set_draw(D)
        got = fit.eval_logpost_re(b, ini["Bs_gammas"], ini["gammas"], 0.1 * ini["alphas"])
    for key in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
        assert np.isfinite(ref[key]).all(), key
        assert _rel(got[key], ref[key]) <= RTOL, (key, _rel(got[key], ref[key]))


@pytest.mark.parametrize("cfg", ["config2", "config3", "config4"])
def test_chain_matches_oracle_draw_for_draw(cuda_lib, oracle, cfg):
    """Same counter-RNG specification on both sides: the whole trajectory must agree."""
    c = make_cohort(cfg, n=400)
    ctl = dict(c["control"]); ctl.update(n_burnin=60, n_iter=40, n_block=20, n_thin=10)
    oracle.set_rowsum_mode(1)
    ref = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=3)
    got = cuda_lib.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=3)
    tr_ref, tr = ref["extras"]["trace_surv"], got["extras"]["trace_surv"]
    assert tr.shape == tr_ref.shape
    assert _rel(tr, tr_ref) <= 1e-9, _rel(tr, tr_ref)
    for key in ("Bs_gammas", "gammas", "alphas", "tau_Bs_gammas", "b"):
        assert ref["mcmc"][key].shape == got["mcmc"][key].shape
        assert np.allclose(got["mcmc"][key], ref["mcmc"][key], rtol=1e-8, atol=1e-10), key
    assert np.allclose(got["logWeights"], ref["logWeights"], rtol=1e-10)
    assert np.allclose(got["scales"]["sigma"]["b"], ref["scales"]["sigma"]["b"], rtol=1e-9)
    for key in ("Bs_gammas", "gammas", "alphas"):
        assert np.isclose(got["scales"]["sigma"][key], ref["scales"]["sigma"][key], rtol=1e-9)
    assert np.allclose(got["extras"]["accept_b"], ref["extras"]["accept_b"])
    assert np.allclose(got["extras"]["accept_surv"], ref["extras"]["accept_surv"])
    assert got["mcmc"]["b"].shape[2] == 4 and np.abs(got["mcmc"]["b"][:, :, -1]).sum() > 0


@pytest.mark.parametrize("cfg", ["config2", "config3", "config4"])
def test_static_and_generic_kernels_agree(cuda_lib, cfg, monkeypatch):
    """The BASELINE shapes get compile-time-specialised step kernels (row-strided groups by default, every group width;
    the TMA-staged and the direct one-lane-per-row kernels on request); the generic kernel must produce the same chain."""
    c = make_cohort(cfg, n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=10)

    def run(label):
        with cuda_lib.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
            assert fit.layout_info()["kernel"].startswith(label), fit.layout_info()["kernel"]
            fit.set_draw(c["Data"])
            return fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl)

    monkeypatch.delenv("JMB_KERNEL", raising=False)
    monkeypatch.delenv("JMB_ROWS_G", raising=False)
    monkeypatch.delenv("JMB_ROWS_EARLY", raising=False)
    r = run("static-rows")
    monkeypatch.setenv("JMB_KERNEL", "tma")
    a = run("static-tma:")
    monkeypatch.setenv("JMB_KERNEL", "direct")
    d = run("static:")
    # same per-subject arithmetic, different partial-sum grouping only
    assert _rel(a["extras"]["trace_surv"], d["extras"]["trace_surv"]) <= 1e-12
    assert np.array_equal(a["mcmc"]["b"], d["mcmc"]["b"])
    monkeypatch.setenv("JMB_KERNEL", "generic")
    b = run("generic")
    assert _rel(a["extras"]["trace_surv"], b["extras"]["trace_surv"]) <= 1e-10
    assert np.allclose(a["mcmc"]["b"], b["mcmc"]["b"], rtol=1e-9, atol=1e-12)
    assert np.allclose(a["scales"]["sigma"]["b"], b["scales"]["sigma"]["b"], rtol=1e-10)
    # the row-strided kernel sums a subject's rows in a different order (lane-strided partial sums): 1e-10
    # (with and without the bulk-copied early stage of a step)
    monkeypatch.delenv("JMB_KERNEL")
    for G, early in [(None, None)] + [(g, e) for g in ("2", "4", "8", "16") for e in ("0", "1", "2", "3")]:
        if G:
            monkeypatch.setenv("JMB_ROWS_G", G)
            monkeypatch.setenv("JMB_ROWS_EARLY", early)
            r = run("static-rows" + G + {"0": ":", "1": "e:", "2": "p:", "3": "q:"}[early])
        assert _rel(r["extras"]["trace_surv"], b["extras"]["trace_surv"]) <= 1e-10, (G, early)
        assert np.allclose(r["mcmc"]["b"], b["mcmc"]["b"], rtol=1e-9, atol=1e-12), (G, early)
        assert np.allclose(r["scales"]["sigma"]["b"], b["scales"]["sigma"]["b"], rtol=1e-10), (G, early)
        assert np.array_equal(r["extras"]["accept_b"], b["extras"]["accept_b"]), (G, early)


def test_uncompacted_designs_use_generic_layout(cuda_lib, oracle):
    """Non-constant U, time-varying W2s and no intercept columns: nothing is compacted."""
    c = make_cohort("config3", n=300)
    D = c["Data"]
    rng = np.random.default_rng(5)
    n, K = 300, 15
    D["U"] = [np.asfortranarray(1.0 + 0.1 * rng.normal(size=u.shape)) for u in D["U"]]
    D["Us"] = [np.asfortranarray(1.0 + 0.1 * rng.normal(size=u.shape)) for u in D["Us"]]
    D["W2s"] = np.asfortranarray(np.asarray(D["W2s"]) + 0.05 * rng.normal(size=D["W2s"].shape))
    D["ZZs"] = [np.asfortranarray(np.asarray(z) * (1.0 + 0.01 * rng.normal(size=(z.shape[0], 1)))) for z in D["ZZs"]]
    D["Z"] = [np.asfortranarray(np.asarray(z) * 1.5) for z in D["Z"]]
    ini = c["inits"]
    oracle.set_rowsum_mode(1)
    ref = oracle.eval_logpost_re(D, c["Covs"]["b"], False, ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    with cuda_lib.Fit(D, c["Covs"]["b"], False) as fit:
        assert fit.layout_info()["kernel"] == "generic"
        fit.set_draw(D)
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    for key in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
        assert _rel(got[key], ref[key]) <= RTOL, (key, _rel(got[key], ref[key]))
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=20, n_block=10, n_thin=10)
    r1 = oracle.lap_rwm_C(c["inits"], D, c["priors"], c["scales"], c["Covs"], ctl, False)
    r2 = cuda_lib.lap_rwm_C(c["inits"], D, c["priors"], c["scales"], c["Covs"], ctl, False)
    assert _rel(r2["extras"]["trace_surv"], r1["extras"]["trace_surv"]) <= 1e-9


def test_adapt_cov_and_shrinkage_chain(cuda_lib, oracle):
    c = make_cohort("config3", n=200)
    ctl = dict(c["control"]); ctl.update(n_burnin=40, n_iter=40, n_block=10, n_thin=8, adaptCov=True)
    pr = dict(c["priors"]); pr.update(shrink_gammas=True, shrink_alphas=True, double_gamma_alphas=True)
    oracle.set_rowsum_mode(1)
    ref = oracle.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, chain_id=1)
    got = cuda_lib.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, chain_id=1)
    assert _rel(got["extras"]["trace_surv"], ref["extras"]["trace_surv"]) <= 1e-8
    for key in ("phi_gammas", "phi_alphas", "tau_gammas", "tau_alphas", "alphas"):
        assert np.allclose(got["mcmc"][key], ref["mcmc"][key], rtol=1e-7, atol=1e-10), key
    for key in ("Bs_gammas", "gammas", "alphas"):
        assert np.allclose(got["scales"]["Sigma"][key], ref["scales"]["Sigma"][key], rtol=1e-6, atol=1e-12), key


def test_errors_are_loud(cuda_lib):
    c = make_cohort("config2", n=50)
    D = dict(c["Data"])
    bad = [np.array(x) for x in D["idL2"]]
    bad[0] = bad[0].copy(); bad[0][3] += 1
    D["idL2"] = bad
    with pytest.raises(cuda_lib.JmbError):
        cuda_lib.Fit(D, c["Covs"]["b"], False)
    with cuda_lib.Fit(c["Data"], c["Covs"]["b"], False) as fit:
        with pytest.raises(cuda_lib.JmbError):   # no draw set yet
            fit.eval_logpost_re(c["inits"]["b"], c["inits"]["Bs_gammas"], c["inits"]["gammas"], c["inits"]["alphas"])
    with pytest.raises(cuda_lib.JmbError):       # multiState without the transition-row fields
        D2 = dict(c["Data"]); D2["nRisks"] = 0
        cuda_lib.Fit(D2, c["Covs"]["b"], False, multiState=True)
    # a control for which the reference indexes past its output arrays (src/fast_LAPRWM.cpp:836-850)
    ctl = dict(c["control"]); ctl.update(n_burnin=40, n_iter=40, n_block=10, n_thin=7)
    with pytest.raises(cuda_lib.JmbError):
        cuda_lib.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False)


def test_chains_in_flight_equal_sequential_chains(cuda_lib):
    """Three chain slots advanced together on their own streams give the chains one gets by running
    them one after the other (the reference's foreach over draws, R/mvJointModelBayes.R:871-905)."""
    c = make_cohort("config4", n=300)
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=20, n_block=10, n_thin=10)
    total = 50
    with cuda_lib.Fit(c["Data"], c["Covs"]["b"], c["interval_cens"]) as fit:
        fit.set_draw(c["Data"])
        seq = [fit.lap_rwm_C(c["inits"], c["priors"], c["scales"], c["Covs"], ctl, chain_id=k) for k in range(3)]
        fit.chain_slots(3)
        for k in range(3):
            fit.select_chain(k)
            fit.set_draw(c["Data"])
            fit.chain_begin(c["inits"], c["priors"], c["scales"], c["Covs"], ctl, chain_id=k)
        ms = fit.chains_iterate(3, total)
        assert ms > 0
        par = []
        for k in range(3):
            fit.select_chain(k)
            par.append(fit.chain_end())
    for a, b in zip(seq, par):
        assert np.array_equal(a["mcmc"]["b"], b["mcmc"]["b"])
        assert np.array_equal(a["mcmc"]["alphas"], b["mcmc"]["alphas"])
        assert np.array_equal(a["extras"]["trace_surv"], b["extras"]["trace_surv"])
        assert np.array_equal(a["scales"]["sigma"]["b"], b["scales"]["sigma"]["b"])
    assert not np.array_equal(par[0]["mcmc"]["alphas"], par[1]["mcmc"]["alphas"])


# ---- Xbetas_calc on the device: set_xdesigns + set_draw_betas == host Xbetas_calc + set_draw -----------
@pytest.mark.parametrize("cfg", ["config2", "config4"])
def test_device_xbetas_equals_host_xbetas(cuda_lib, cfg):
    """R/supportFuns_mvJointModelBayes.R:43-55 evaluated by jmb_xbetas_kernel from Data$X / XX / XXs[_int] and the draw's
    betas: per-subject log-posterior pieces within 1e-12 relative of the host-Xbetas path (only the rounding of the
    two-term dot products can differ), chains identical to 1e-9, also on a second chain slot."""
    c = make_cohort(cfg, n=700)
    D, ini = c["Data"], c["inits"]
    ctl = dict(c["control"]); ctl.update(n_burnin=30, n_iter=10, n_block=10, n_thin=10)
    with cuda_lib.Fit(D, c["Covs"]["b"], c["interval_cens"]) as fit:
        fit.set_draw(D)
        ref = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        ch_ref = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
        fit.set_xdesigns(D)
        fit.set_draw_betas(D["betas"], D["sigmas"], D["invD"])
        got = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        for k in ("log_pyb", "log_pb", "log_h", "H", "log_ptb"):
            assert np.allclose(got[k], ref[k], rtol=1e-12, atol=1e-13), k
        ch = fit.lap_rwm_C(ini, c["priors"], c["scales"], c["Covs"], ctl)
        assert np.allclose(ch["extras"]["trace_surv"], ch_ref["extras"]["trace_surv"], rtol=1e-9)
        assert np.allclose(ch["mcmc"]["b"], ch_ref["mcmc"]["b"], rtol=1e-8, atol=1e-10)
        # another draw on another chain slot: different betas must give the host result for those betas
        fit.chain_slots(2)
        fit.select_chain(1)
        b2 = [b + 0.05 for b in D["betas"]]
        D2 = dict(D)
        D2["Xbetas"] = [np.asarray(D["X"][k]) @ b2[k] for k in range(len(b2))]
        for name in ("XXbetas", "XXsbetas") + (("XXsbetas_int",) if c["interval_cens"] else ()):
            src = {"XXbetas": "XX", "XXsbetas": "XXs", "XXsbetas_int": "XXs_int"}[name]
            D2[name] = [np.asarray(D[src][t]) @ b2[D["outcome"][t]][D["indFixed"][t]] for t in range(len(D["XX"]))]
        fit.set_draw_betas(b2, D["sigmas"], D["invD"])
        got2 = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        fit.set_draw(D2)
        ref2 = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
        for k in ("log_pyb", "log_h", "H", "log_ptb"):
            assert np.allclose(got2[k], ref2[k], rtol=1e-12, atol=1e-13), k
        assert not np.allclose(got2["log_pyb"], ref["log_pyb"], rtol=1e-6)
    with cuda_lib.Fit(D, c["Covs"]["b"], c["interval_cens"]) as fit:
        with pytest.raises(cuda_lib.JmbError):
            fit.set_draw_betas(D["betas"], D["sigmas"], D["invD"])       # designs not uploaded


@pytest.mark.parametrize("cfg", ["config3", "config4"])
def test_basis_from_knots_equals_dense_basis(cuda_lib, oracle, cfg):
    """jmb_create with W1 = NULL evaluates splineDesign(knots, x, ord = 4, outer.ok = TRUE) itself from
    Time / st / st_int; everything downstream must equal the dense-basis handle (band values are the same numbers)."""
    c = make_cohort(cfg, n=500)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    Wl = oracle.wlong(D, c["Covs"]["b"], c["interval_cens"], ini["b"], 0)
    Wls = oracle.wlong(D, c["Covs"]["b"], c["interval_cens"], ini["b"], 1)
    res = []
    for lean in (False, True):
        with cuda_lib.Fit(D, c["Covs"]["b"], c["interval_cens"], basis_from_knots=lean) as fit:
            fit.set_draw(D)
            ev = fit.eval_logpost_re(ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
            fit.set_wlong(Wl, Wls)
            gr = fit.gradient_logPosterior(1.0, ini["Bs_gammas"], ini["gammas"], ini["alphas"], pr, 200.0, 1.0, 0.01)
            res.append((ev, gr, fit.layout_info()["w1_band"]))
    (e0, g0, w0), (e1, g1, w1) = res
    assert w0 == w1 == 4
    for k in ("log_h", "H", "log_ptb"):
        assert np.allclose(e0[k], e1[k], rtol=1e-14, atol=0), k
    assert np.allclose(g0, g1, rtol=1e-12, atol=1e-12)
    bad = dict(D); bad["knots"] = np.asarray(D["knots"])[:-1]
    with pytest.raises(cuda_lib.JmbError):
        cuda_lib.Fit(bad, c["Covs"]["b"], c["interval_cens"], basis_from_knots=True)
