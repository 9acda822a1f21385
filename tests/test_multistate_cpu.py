"""Multi-state fits (``multiState = TRUE``; src/fast_LAPRWM.cpp:519-529,669,703-725,777-782) on the CPU: the restated
oracle against golden vectors produced by the reference's own sources (tests/golden/make_golden.py, MS_CASES), against
the live reference tree where it is mounted, against an independent NumPy computation of the per-transition pieces, and
through a size-independent property (a fit whose subjects have one transition each is the ordinary fit).
"""
import importlib.util
import os

import numpy as np
import pytest

from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort
from oracle import np_twin

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def _gold(name):
    return dict(np.load(os.path.join(HERE, "golden", name + ".npz")))


def check_chain(r, gold, rtol, atol):
    for k, v in r["mcmc"].items():
        g = gold["mcmc_" + k]
        assert v.shape == g.shape, (k, v.shape, g.shape)
        assert np.allclose(v, g, rtol=rtol, atol=atol), (k, float(np.max(np.abs(v - g))))
    assert np.allclose(r["logWeights"], gold["logWeights"], rtol=rtol)
    assert np.allclose(r["scales"]["sigma"]["b"], gold["sigma_b"], rtol=rtol)
    sg = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    assert np.allclose(sg, gold["sigma_surv"], rtol=rtol)
    for k in ("Bs_gammas", "gammas", "alphas"):
        assert np.allclose(r["scales"]["Sigma"][k], gold["Sigma_" + k], rtol=max(rtol, 1e-8), atol=1e-14), k


def test_cohort_follows_the_multistate_data_contract():
    c = make_multistate_cohort(n=80)
    D = c["Data"]
    idT = np.asarray(D["idT2"][0])
    nT, K = idT.size, c["K"]
    assert np.all(np.diff(idT) >= 0) and set(idT) == set(range(80))           # grouped by subject, everyone at risk
    assert np.array_equal(D["idT_rsum"], np.flatnonzero(np.r_[idT[1:] != idT[:-1], True]))   # R/mvJointModelBayes.R:567-568
    assert D["event"].shape == (nT,) and D["W1"].shape[0] == nT and D["W1s"].shape[0] == nT * K
    assert np.array_equal(D["idGK_fast"], (np.arange(1, nT + 1) * K - 1))
    # block-structured W1: a row is non-zero only in the columns of its own stratum (:351-359)
    for h in range(D["nRisks"]):
        own = np.zeros(D["W1"].shape[1], dtype=bool); own[D["kn_strat_first"][h]:D["kn_strat_last"][h] + 1] = True
        assert np.all(np.asarray(D["W1"])[D["strata"] == h][:, ~own] == 0.0)
        assert np.allclose(np.asarray(D["W1"])[D["strata"] == h][:, own].sum(axis=1), 1.0)
    assert c["inits"]["tau_Bs_gammas"].shape == (D["nRisks"],)                 # rep(tau, nRisks), :723-724
    assert 2 * 80 < nT <= 3 * 80 and 0 < D["event"].sum() < nT


@pytest.mark.parametrize("name", sorted(make_golden.MS_CASES))
def test_oracle_reproduces_reference_multistate_golden(oracle, name):
    c, ctl, pr, chain = make_golden.ms_inputs(name)
    oracle.set_rowsum_mode(0)          # the reference's own cumsum-difference rowsum
    r = oracle.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, True, chain_id=chain)
    g = _gold(name)
    assert g["mcmc_tau_Bs_gammas"].shape[0] == c["Data"]["nRisks"]             # nRisks x n_out, :607
    check_chain(r, g, rtol=1e-11, atol=1e-13)


def test_live_reference_source_matches_oracle_multistate(oracle):
    from oracle import ref
    if not os.path.exists(os.path.join(ref.REF, "src", "fast_LAPRWM.cpp")):
        pytest.skip("reference tree not present")
    c = make_multistate_cohort(n=50, seed=7)
    ctl = dict(c["control"]); ctl.update(n_burnin=25, n_iter=15, n_block=5, n_thin=5)
    oracle.set_rowsum_mode(0)
    a = ref.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=9)
    b = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, True, chain_id=9)
    for k in a["mcmc"]:
        assert np.allclose(a["mcmc"][k], b["mcmc"][k], rtol=1e-12, atol=1e-14), k
    assert np.allclose(a["logWeights"], b["logWeights"], rtol=1e-12)


def numpy_pieces(D, b, Bs, gam, alp):
    """log_h, H, log_ptb per transition row and their per-subject sums, straight from the Data dict
    (src/fast_LAPRWM.cpp:622-650 with idT2 / idT2s gathers)."""
    idT = np.asarray(D["idT2"][0]); nT = idT.size
    K = len(D["Pw"]) // nT
    A = sum(len(c) for c in D["col_inds"])

    def wl(xb, ZZ, U, ids):
        out = np.zeros((len(ids), A))
        for t in range(len(xb)):
            zb = np.sum(np.asarray(ZZ[t]) * b[ids][:, D["RE_inds2"][t]], axis=1)
            v = np_twin._TF[D["trans_Funs"][t]](np.asarray(xb[t]) + zb)
            out[:, D["col_inds"][t]] = np.asarray(U[t]) * v[:, None]
        return out
    log_h = np.asarray(D["W1"]) @ Bs + np.asarray(D["W2"]) @ gam + wl(D["XXbetas"], D["ZZ"], D["U"], idT) @ alp
    lin = np.asarray(D["W1s"]) @ Bs + np.asarray(D["W2s"]) @ gam + wl(D["XXsbetas"], D["ZZs"], D["Us"], np.repeat(idT, K)) @ alp
    H = (np.asarray(D["Pw"]) * np.exp(lin)).reshape(nT, K).sum(axis=1)
    ptb = np.asarray(D["event"]) * log_h - H
    return log_h, H, ptb, np.bincount(idT, weights=ptb)


def test_oracle_pieces_match_numpy_multistate(oracle):
    c = make_multistate_cohort(n=120)
    D, ini = c["Data"], c["inits"]
    b = np.asarray(ini["b"])
    oracle.set_rowsum_mode(1)
    r = oracle.eval_logpost_re(D, c["Covs"]["b"], False, b, ini["Bs_gammas"], ini["gammas"], ini["alphas"], multiState=True)
    log_h, H, ptb, _ = numpy_pieces(D, b, np.asarray(ini["Bs_gammas"]), np.asarray(ini["gammas"]), np.asarray(ini["alphas"]))
    assert r["log_h"].shape == log_h.shape
    for got, want in ((r["log_h"], log_h), (r["H"], H), (r["log_ptb"], ptb)):
        assert np.allclose(got, want, rtol=1e-12, atol=1e-13)
    assert np.allclose(r["log_pyb"], np_twin.log_longF(D, b), rtol=1e-12)


def test_one_transition_per_subject_is_the_ordinary_fit(oracle):
    """Size-independent property: with one transition row per subject and one stratum, multiState = TRUE walks
    exactly the chain of the ordinary fit."""
    c = make_cohort("config3", n=80)
    ctl = dict(c["control"]); ctl.update(n_burnin=20, n_iter=20, n_block=10, n_thin=5)
    oracle.set_rowsum_mode(1)
    a = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, False)
    b = oracle.lap_rwm_C(c["inits"], c["Data"], c["priors"], c["scales"], c["Covs"], ctl, False, True)
    for k in a["mcmc"]:
        assert np.array_equal(a["mcmc"][k], b["mcmc"][k]), k
    assert np.array_equal(a["extras"]["trace_surv"], b["extras"]["trace_surv"])
