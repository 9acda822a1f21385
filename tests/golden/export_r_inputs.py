"""Writes the seeded inputs of tests/golden/make_r_fixtures.R as JSON under tests/golden/r_inputs/ (matrices as
{"dim": [r, c], "data": column-major}).  The R script feeds them to JMbayes' own R functions; tests/test_r_fixtures.py
compares what R returned (tests/golden/r_fixtures/*.json, when a reviewer has produced them) with the restated R code of
oracle/ and with the CUDA path.  Run:  python tests/golden/export_r_inputs.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from jmbayes_b200.cohorts import make_cohort  # noqa: E402

OUT = os.path.join(HERE, "r_inputs")


def mat(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        return [float(x) for x in a]
    return {"dim": list(a.shape), "data": [float(x) for x in np.asfortranarray(a).ravel(order="F")]}


def stats_case():
    """Series of M = 160 draws (autocorrelated, skewed, tied, constant) and LogLiks for the importance weights."""
    rng = np.random.default_rng(3)
    X = np.cumsum(rng.normal(size=(48, 160)), axis=1) * 0.1 + rng.normal(size=(48, 160))
    X[5] = np.exp(X[5]); X[6] = np.round(X[6], 1); X[7] = 2.5
    return X, rng.normal(size=160)


def mll_case():
    from oracle import oracle
    oracle.build()
    c = make_cohort("config3", n=120)
    D = dict(c["Data"]); ini = c["inits"]
    D["Wlong"] = oracle.wlong(c["Data"], c["Covs"]["b"], False, ini["b"], 0)
    D["Wlongs"] = oracle.wlong(c["Data"], c["Covs"]["b"], False, ini["b"], 1)
    return c, D, dict(Bs_gammas=ini["Bs_gammas"], gammas=ini["gammas"], alphas=ini["alphas"], tau_Bs_gammas=200.0)


def svft_case():
    from jmbayes_b200 import survfit as sv
    c = make_cohort("config3", n=16)
    d = sv.newdata_designs(c, L=3)
    means, _ = sv.posterior_draws(c, M=6)
    return c, d, means


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    X, ll = stats_case()
    json.dump({"X": mat(X), "LogLiks": mat(ll)}, open(os.path.join(OUT, "stats.json"), "w"))
    c, D, th = mll_case()
    ev = np.asarray(D["event"])
    json.dump({"Data": {k: mat(D[k]) for k in ("event", "W1", "W1s", "W2", "W2s", "Wlong", "Wlongs", "Pw")}
               | {"idGK_fast": [int(x) for x in D["idGK_fast"]],
                  "event_colSumsW1": mat(ev @ np.asarray(D["W1"])), "event_colSumsW2": mat(ev @ np.asarray(D["W2"])),
                  "event_colSumsWlong": mat(ev @ np.asarray(D["Wlong"]))},
               "priors": {k: (mat(v) if np.ndim(v) else float(v)) for k, v in c["priors"].items() if k != "td_cols"},
               "thetas": {k: (mat(v) if np.ndim(v) else float(v)) for k, v in th.items()}},
              open(os.path.join(OUT, "marglogLik2.json"), "w"))
    # per-subject Data lists of log_post_RE_svft exactly as survfitJM.mvJMbayes builds them (R/survfitJM.mvJMbayes.R:281-294)
    c5, d5, means = svft_case()
    K, O, T = d5["K"], len(d5["y_long"]), len(d5["ZZs"])
    subjects = []
    for i in range(d5["n"]):
        sub = {"y": [], "Xbetas": [], "Z": []}
        for k in range(O):
            rows = np.flatnonzero(np.asarray(d5["idL"][k]) == i)
            sub["y"].append(mat(np.asarray(d5["y_long"][k])[rows]))
            sub["Xbetas"].append(mat(np.asarray(d5["X_long"][k])[rows] @ np.asarray(means["betas"][k])[0]))
            sub["Z"].append(mat(np.asarray(d5["Z_long"][k])[rows]))
        q = slice(i * K, (i + 1) * K)
        sub["W1s"], sub["W2s"], sub["Pw"] = mat(d5["W1s"][q]), mat(d5["W2s"][q]), mat(d5["Pw"][q])
        sub["XXsbetas"] = [mat(np.asarray(d5["XXs"][t])[q] @ np.asarray(means["betas"][int(d5["outcome"][t])])[0][np.asarray(d5["indFixed"][t])]) for t in range(T)]
        sub["ZZs"] = [mat(np.asarray(d5["ZZs"][t])[q]) for t in range(T)]
        sub["Us"] = [mat(np.asarray(d5["Us"][t])[q]) for t in range(T)]
        subjects.append(sub)
    json.dump({"K": K, "q": d5["q"], "fams": d5["fams"], "links": d5["links"], "trans_Funs": d5["trans_Funs"],
               "RE_inds": [[int(x) + 1 for x in r] for r in d5["RE_inds"]], "RE_inds2": [[int(x) + 1 for x in r] for r in d5["RE_inds2"]],
               "col_inds": [[int(x) + 1 for x in r] for r in d5["col_inds"]],
               "sigmas": [None if v is None else float(np.asarray(v)[0]) for v in means["sigmas"]], "invD": mat(np.asarray(means["inv_D"])[0]),
               "Bs_gammas": mat(np.asarray(means["Bs_gammas"])[0]), "gammas": mat(np.asarray(means["gammas"])[0]),
               "alphas": mat(np.asarray(means["alphas"])[0]), "subjects": subjects},
              open(os.path.join(OUT, "survfit_modes.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))
