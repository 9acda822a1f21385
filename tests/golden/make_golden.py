"""Generates tests/golden/*.npz from the reference's OWN sources (oracle/_ref: /root/reference/src
compiled against the stand-in Armadillo/Rcpp header).  Run in the build container, where the
reference tree is mounted:  python tests/golden/make_golden.py

Each fixture stores the control overrides, the chain id and the outputs of the reference's
lap_rwm_C / logPosterior / gradient_logPosterior for a small seeded cohort (cohorts.make_cohort is
deterministic, so the inputs are regenerated from (config, n) by the tests).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from jmbayes_b200 import _abi  # noqa: E402
from jmbayes_b200.cohorts import make_cohort, make_multistate_cohort  # noqa: E402
from oracle import np_twin, ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (config, n, control overrides, prior overrides, chain id)
    "chain_config2": ("config2", 120, dict(n_burnin=40, n_iter=20, n_block=10, n_thin=5), {}, 1),
    "chain_config3": ("config3", 120, dict(n_burnin=40, n_iter=20, n_block=10, n_thin=5), {}, 2),
    "chain_config4": ("config4", 120, dict(n_burnin=40, n_iter=20, n_block=10, n_thin=5), {}, 3),
    "chain_config3_shrink_adapt": ("config3", 90, dict(n_burnin=30, n_iter=30, n_block=10, n_thin=6, adaptCov=True),
                                   dict(shrink_gammas=True, shrink_alphas=True, double_gamma_alphas=True), 4),
    "chain_config2_thin1": ("config2", 60, dict(n_burnin=10, n_iter=10, n_block=5, n_thin=1), {}, 5),
}


def run_case(name):
    cfg, n, ctl_over, pr_over, chain = CASES[name]
    c = make_cohort(cfg, n=n)
    ctl = dict(c["control"]); ctl.update(ctl_over)
    pr = dict(c["priors"]); pr.update(pr_over)
    r = ref.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, c["interval_cens"], chain_id=chain)
    out = {f"mcmc_{k}": v for k, v in r["mcmc"].items()}
    out["logWeights"] = r["logWeights"]
    out["sigma_b"] = r["scales"]["sigma"]["b"]
    out["sigma_surv"] = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    for k in ("Bs_gammas", "gammas", "alphas"):
        out[f"Sigma_{k}"] = r["scales"]["Sigma"][k]
    return out


MS_CASES = {
    # multi-state fits (multiState = TRUE): name -> (n, control overrides, prior overrides, chain id)
    "chain_multistate": (90, dict(n_burnin=40, n_iter=20, n_block=10, n_thin=5), {}, 6),
    "chain_multistate_shrink_adapt": (70, dict(n_burnin=30, n_iter=30, n_block=10, n_thin=6, adaptCov=True),
                                      dict(shrink_gammas=True, shrink_alphas=True, double_gamma_alphas=True), 7),
}


def ms_inputs(name):
    n, ctl_over, pr_over, chain = MS_CASES[name]
    c = make_multistate_cohort(n=n)
    ctl = dict(c["control"]); ctl.update(ctl_over)
    pr = dict(c["priors"]); pr.update(pr_over)
    return c, ctl, pr, chain


def run_ms_case(name):
    c, ctl, pr, chain = ms_inputs(name)
    r = ref.lap_rwm_C(c["inits"], c["Data"], pr, c["scales"], c["Covs"], ctl, False, True, chain_id=chain)
    out = {f"mcmc_{k}": v for k, v in r["mcmc"].items()}
    out["logWeights"] = r["logWeights"]
    out["sigma_b"] = r["scales"]["sigma"]["b"]
    out["sigma_surv"] = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    for k in ("Bs_gammas", "gammas", "alphas"):
        out[f"Sigma_{k}"] = r["scales"]["Sigma"][k]
    return out


def surv_case():
    c = make_cohort("config3", n=200)
    D, ini, pr = c["Data"], c["inits"], c["priors"]
    b = np.asarray(ini["b"])
    Wl, Wls = np.asfortranarray(np_twin.wlong(D, b, 0)), np.asfortranarray(np_twin.wlong(D, b, 1))
    L = ref.lib()
    Bs, g, a = np.array(ini["Bs_gammas"]), np.array(ini["gammas"]), np.array(ini["alphas"])
    tau, temp = 3.5, 0.7
    P = lambda x: np.asfortranarray(np.asarray(x, dtype=np.float64))  # noqa: E731
    keep = [P(x) for x in (D["event"], D["W1"], D["W1s"], Bs, D["W2"], D["W2s"], g, Wl, Wls, a, D["Pw"], pr["mean_Bs_gammas"],
                           pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    lp = L.jmbref_logPosterior(temp, Wl.shape[0], Wls.shape[0], len(Bs), len(g), len(a),
                               *[k.ctypes.data_as(_abi.c_double_p) for k in keep], tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"])
    ev = np.asarray(D["event"])
    ecs = [np.ascontiguousarray(ev @ np.asarray(M)) for M in (D["W1"], D["W2"], Wl)]
    keep2 = [P(x) for x in (ecs[0], D["W1s"], Bs, ecs[1], D["W2s"], g, ecs[2], Wls, a, D["Pw"], pr["mean_Bs_gammas"],
                            pr["Tau_Bs_gammas"], pr["mean_gammas"], pr["Tau_gammas"], pr["mean_alphas"], pr["Tau_alphas"])]
    sc = np.zeros(len(Bs) + len(g) + len(a) + 1)
    L.jmbref_gradient_logPosterior(temp, Wls.shape[0], len(Bs), len(g), len(a), *[k.ctypes.data_as(_abi.c_double_p) for k in keep2],
                                   tau, pr["A_tau_Bs_gammas"], pr["B_tau_Bs_gammas"], sc.ctypes.data_as(_abi.c_double_p))
    return dict(logPosterior=np.array([lp]), score=sc, tau=np.array([tau]), temp=np.array([temp]))


WORE_CASES = {
    # name: (config, n, control overrides, prior overrides, drop gammas, chain id)
    "wore_config3": ("config3", 150, dict(n_burnin=30, n_iter=20, n_block=10, n_thin=5), {}, False, 3),
    "wore_config2_shrink_adapt": ("config2", 150, dict(n_burnin=30, n_iter=20, n_block=10, n_thin=5, adaptCov=True),
                                  dict(shrink_gammas=True, shrink_alphas=True), False, 4),
    "wore_config3_nogammas": ("config3", 120, dict(n_burnin=20, n_iter=20, n_block=10, n_thin=5), dict(shrink_alphas=True), True, 5),
}


def wore_inputs(name):
    """Inputs of a woRE case: Wlong / Wlongs at the initial b (designMatLong, R/mvJointModelBayes.R:787-790)."""
    cfg, n, ctl_over, pr_over, nogam, chain = WORE_CASES[name]
    c = make_cohort(cfg, n=n)
    D = dict(c["Data"])
    b = np.asarray(c["inits"]["b"])
    D["Wlong"], D["Wlongs"] = np_twin.wlong(D, b, 0), np_twin.wlong(D, b, 1)
    pr = dict(c["priors"]); pr.update(pr_over)
    ini, cv = dict(c["inits"]), dict(c["Covs"])
    if nogam:
        K = len(D["Pw"]) // n
        D["W2"], D["W2s"] = np.zeros((n, 0), order="F"), np.zeros((n * K, 0), order="F")
        pr.update(mean_gammas=np.zeros(0), Tau_gammas=np.zeros((0, 0)), shrink_gammas=False)
        ini.update(gammas=np.zeros(0), phi_gammas=np.zeros(0))
        cv["gammas"] = np.zeros((0, 0))
    ctl = dict(c["control"]); ctl.update(ctl_over)
    return c, D, ini, pr, cv, ctl, chain


def run_wore_case(name):
    c, D, ini, pr, cv, ctl, chain = wore_inputs(name)
    r = ref.lap_rwm_C_woRE(ini, D, pr, c["scales"], cv, ctl, c["Covs"]["b"], chain_id=chain)
    out = {f"mcmc_{k}": v for k, v in r["mcmc"].items()}
    out["logPost"] = r["logPost"]
    out["sigma_surv"] = np.array([r["scales"]["sigma"][k] for k in ("Bs_gammas", "gammas", "alphas")])
    for k in ("Bs_gammas", "gammas", "alphas"):
        out[f"Sigma_{k}"] = r["scales"]["Sigma"][k]
    return out


def svft_inputs():
    """New-subject style data: config 3 cohort where some subjects have no rows in outcome 2."""
    c = make_cohort("config3", n=160)
    D = dict(c["Data"])
    idL = np.asarray(D["idL"][1])
    keep = ~np.isin(idL, [3, 17, 159])                # subjects without binary measurements
    D["y"] = list(D["y"]); D["Xbetas"] = list(D["Xbetas"]); D["Z"] = list(D["Z"]); D["idL"] = list(D["idL"]); D["idL2"] = list(D["idL2"])
    D["y"][1] = np.asarray(D["y"][1])[keep]; D["Xbetas"][1] = np.asarray(D["Xbetas"][1])[keep]
    D["Z"][1] = np.asfortranarray(np.asarray(D["Z"][1])[keep]); D["idL"][1] = idL[keep].astype(np.int32)
    cnt = np.bincount(D["idL"][1], minlength=160)
    D["idL2"][1] = (np.cumsum(cnt) - 1).astype(np.int32)
    return c, D


def svft_case():
    c, D = svft_inputs()
    ini = c["inits"]
    # the reference's per-subject call needs every outcome present: evaluate the complete cohort for the golden
    # values of the complete subjects, and the three incomplete subjects with outcome 2 removed from their lists
    lp, sp = ref.svft(c["Data"], c["Covs"]["b"], ini["b"], ini["Bs_gammas"], ini["gammas"], ini["alphas"])
    return dict(log_post=lp, surv_pred=sp)


if __name__ == "__main__":
    ref.build(force=True)
    for name in CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_case(name))
        print("wrote", name)
    for name in MS_CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_ms_case(name))
        print("wrote", name)
    for name in WORE_CASES:
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run_wore_case(name))
        print("wrote", name)
    np.savez_compressed(os.path.join(HERE, "svft_config3.npz"), **svft_case())
    print("wrote svft_config3")
    np.savez_compressed(os.path.join(HERE, "surv_config3.npz"), **surv_case())
    print("wrote surv_config3")
