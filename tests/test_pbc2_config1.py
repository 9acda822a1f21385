"""BASELINE config 1: pbc2 (312 subjects, 1945 visits), log(serBilir) ~ year + (year | id) with
Surv(years, status2) ~ drug + age (man/mvJointModelBayes.Rd:204-227): the reference's own CPU-runnable case.
CPU: marglogLik2 on the oracle's logPosterior / score, then the chain on the reference's own source vs the
oracle; GPU: the same flow through the CUDA library (Laplace step and chain) against the CPU results."""
import ctypes as C

import numpy as np
import pytest

from jmbayes_b200 import _abi
from jmbayes_b200.pbc2 import finish_inits, make_pbc2


def _oracle_post(oracle, pb):
    """logPosterior / gradient_logPosterior of the oracle as callables of (Bs, gammas, alphas)."""
    D, pr = pb["Data"], pb["priors"]
    L = oracle.lib()
    Wl, Wls = pb["Wlong"], pb["Wlongs"]
    ev = np.asarray(D["event"])
    ecs = [np.ascontiguousarray(ev @ np.asarray(M)) for M in (D["W1"], D["W2"], Wl)]
    P = _abi.c_double_p
    L.orc_logPosterior.restype = C.c_double
    L.orc_logPosterior.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [P] * 17 + [C.c_double] * 3
    L.orc_gradient_logPosterior.restype = None
    L.orc_gradient_logPosterior.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [P] * 16 + [C.c_double] * 3 + [P]
    F = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))  # noqa: E731
    n, ns, B = Wl.shape[0], Wls.shape[0], pb["B"]

    def lp(Bs, g, a):
        keep = [F(x) for x in (D["event"], D["W1"], D["W1s"], Bs, D["W2"], D["W2s"], g, Wl, Wls, a, D["Pw"], pr["mean_Bs_gammas"],
                               pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
        return L.orc_logPosterior(1.0, n, ns, B, 2, 1, *[k.ctypes.data_as(P) for k in keep], 200.0, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])

    def gr(Bs, g, a):
        keep = [F(x) for x in (ecs[0], D["W1s"], Bs, ecs[1], D["W2s"], g, ecs[2], Wls, a, D["Pw"], pr["mean_Bs_gammas"],
                               pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
        out = np.zeros(B + 4)
        L.orc_gradient_logPosterior(1.0, ns, B, 2, 1, *[k.ctypes.data_as(P) for k in keep], 200.0, pr["A_tau_Bs_gammas"],
                                    pr["B_tau_Bs_gammas"], out.ctypes.data_as(P))
        return out
    return lp, gr


@pytest.fixture(scope="module")
def pbc2_cpu(oracle):
    pb = make_pbc2()
    lp, gr = _oracle_post(oracle, pb)
    return finish_inits(pb, lp, gr)


def test_pbc2_design_and_stage1(pbc2_cpu):
    pb = pbc2_cpu
    D = pb["Data"]
    assert pb["n"] == 312 and len(D["y"][0]) == 1945 and int(D["event"].sum()) == 140
    assert D["W1"].shape == (312, 17) and D["W1s"].shape == (312 * 15, 17)
    assert np.allclose(D["W1s"].sum(axis=1), 1.0) and np.all((D["W1s"] != 0).sum(axis=1) <= 4)
    s1 = pb["stage1"]                       # literature values for this model: intercept ~0.49, slope ~0.18, sigma ~0.35
    assert 0.4 < s1["betas"][0] < 0.6 and 0.12 < s1["betas"][1] < 0.25 and 0.3 < s1["sigma"] < 0.4
    ini = pb["inits"]                       # Laplace step: value association of log serBilir ~1.2-1.4, older age raises the hazard
    assert 1.0 < ini["alphas"][0] < 1.6 and ini["gammas"][1] > 0
    assert np.all(np.linalg.eigvalsh(pb["Covs"]["Bs_gammas"]) > 0)


def test_pbc2_chain_reference_source_vs_oracle(oracle, pbc2_cpu):
    from oracle import ref
    pb = pbc2_cpu
    ctl = dict(pb["control"]); ctl.update(n_burnin=100, n_iter=100, n_block=50, n_thin=20)
    args = (pb["inits"], pb["Data"], pb["priors"], pb["scales"], pb["Covs"], ctl, False)
    oracle.set_rowsum_mode(0)
    o = oracle.lap_rwm_C(*args, chain_id=1)
    if ref.available() and ref.build():
        r = ref.lap_rwm_C(*args, chain_id=1)
        for k in ("b", "Bs_gammas", "gammas", "alphas", "tau_Bs_gammas"):
            assert np.allclose(r["mcmc"][k], o["mcmc"][k], rtol=1e-11, atol=1e-13), k
        assert np.allclose(r["logWeights"], o["logWeights"], rtol=1e-12)
    assert o["mcmc"]["alphas"].shape == (1, 5) and np.all(o["mcmc"]["alphas"] > 0.8)
    acc = o["extras"]["accept_b"].mean()
    assert 0.15 < acc < 0.6


@pytest.mark.gpu
def test_pbc2_on_cuda_matches_cpu(cuda_lib, oracle, pbc2_cpu):
    pb = pbc2_cpu
    D, pr = pb["Data"], pb["priors"]
    with cuda_lib.Fit(D, pb["Covs_b"], False) as fit:
        assert fit.layout_info()["kernel"].startswith("static-rows") and fit.layout_info()["kernel"].endswith(":value_gaussian")
        # marglogLik2 driven by the device functions
        fit.set_wlong(pb["Wlong"], pb["Wlongs"])
        dev = finish_inits(make_pbc2(),
                           lambda Bs, g, a: fit.logPosterior(1.0, Bs, g, a, pr, 200.0, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"]),
                           lambda Bs, g, a: fit.gradient_logPosterior(1.0, Bs, g, a, pr, 200.0, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"]))
        assert np.allclose(dev["inits"]["alphas"], pb["inits"]["alphas"], rtol=1e-5)
        assert np.allclose(dev["inits"]["Bs_gammas"], pb["inits"]["Bs_gammas"], rtol=1e-4, atol=1e-5)
        assert abs(dev["laplace_value"] - pb["laplace_value"]) < 1e-6 * abs(pb["laplace_value"])
        # the native driver (jmb_marglogLik2: R's vmmin / optimhess / cd.vec / nearPD restated) against its CPU restatement,
        # and against the SciPy-driven stand-in above (same optimum)
        B = pb["B"]
        th0 = dict(Bs_gammas=np.zeros(B), gammas=np.zeros(2), alphas=np.zeros(1), tau_Bs_gammas=200.0)
        nat = fit.marglogLik2(th0, pr, fixed_tau_Bs_gammas=True)
        Dm = dict(D); Dm["Wlong"] = pb["Wlong"]; Dm["Wlongs"] = pb["Wlongs"]
        orc = oracle.marglogLik2(th0, Dm, pr, fixed_tau_Bs_gammas=True)
        assert nat["convergence"] == orc["convergence"] == 0
        assert np.allclose(nat["par"], orc["par"], rtol=0, atol=1e-6)
        assert np.allclose(nat["Covs"]["alphas"], orc["Covs"]["alphas"], rtol=1e-4)
        assert np.allclose(nat["inits"]["alphas"], pb["inits"]["alphas"], rtol=2e-3)
        assert abs(nat["opt_value"] - pb["laplace_value"]) < 1e-5 * abs(pb["laplace_value"])
        # the chain, from the CPU-side inits so that both sides start from identical numbers
        ctl = dict(pb["control"]); ctl.update(n_burnin=100, n_iter=100, n_block=50, n_thin=20)
        fit.set_draw(D)
        got = fit.lap_rwm_C(pb["inits"], pr, pb["scales"], pb["Covs"], ctl, chain_id=1)
    oracle.set_rowsum_mode(1)
    ref_ = oracle.lap_rwm_C(pb["inits"], D, pr, pb["scales"], pb["Covs"], ctl, False, chain_id=1)
    tr, tr_ref = got["extras"]["trace_surv"], ref_["extras"]["trace_surv"]
    assert np.max(np.abs(tr - tr_ref) / np.abs(tr_ref)) <= 1e-9
    for k in ("b", "Bs_gammas", "gammas", "alphas"):
        assert np.allclose(got["mcmc"][k], ref_["mcmc"][k], rtol=1e-8, atol=1e-10), k
