"""Independent NumPy restatement of the per-subject target of the b-step and of the survival-block
log-posterior / score.  TEST INFRASTRUCTURE ONLY (see jmb_oracle.h for the parity pin).

Written from the reference's formulas (src/fast_LAPRWM.cpp:67-109,168-202,619-650;
src/fast_logPost.cpp:8-26; src/fast_score.cpp:8-50) directly on the ``Data`` dicts, without the C
marshaller, so that it cross-checks both the C oracle and the marshalling.  Segment sums are exact
per-segment sums (``np.add.reduceat``), not the reference's cumsum difference.
"""
from __future__ import annotations

import numpy as np
from scipy.special import ndtr

_TF = {
    "identity": lambda x: x, "expit": lambda x: np.exp(x) / (1 + np.exp(x)), "exp": np.exp,
    "log": np.log, "log2": np.log2, "log10": np.log10, "sqrt": np.sqrt,
}


def _starts(last):
    last = np.asarray(last, dtype=np.int64)
    return np.concatenate([[0], last[:-1] + 1])


def log_longF(Data, b):
    """lin_predF + log_longF: src/fast_LAPRWM.cpp:67-78,168-202."""
    n = b.shape[0]
    out = np.zeros(n)
    for k, y in enumerate(Data["y"]):
        Z = np.asarray(Data["Z"][k])
        idL = np.asarray(Data["idL"][k])
        eta = np.asarray(Data["Xbetas"][k]) + np.sum(Z * b[idL][:, Data["RE_inds"][k]], axis=1)
        fam, link = Data["fams"][k], Data["links"][k]
        with np.errstate(all="ignore"):
            if fam == "gaussian":
                ld = -0.5 * ((y - eta) / Data["sigmas"][k]) ** 2
            elif fam == "binomial":
                if link == "logit":
                    pr = np.exp(eta) / (1 + np.exp(eta))
                elif link == "probit":
                    pr = ndtr(eta)
                else:
                    pr = -np.exp(-np.exp(eta)) + 1
                ld = y * np.log(pr) + (1 - y) * np.log(1 - pr)
            else:
                mu = np.exp(eta)
                ld = y * np.log(mu) - mu
        out += np.add.reduceat(ld, _starts(Data["idL2"][k]))
    return out


def wlong(Data, b, which):
    """lin_pred_matF: src/fast_LAPRWM.cpp:80-109.  which: 0 event rows, 1 quadrature, 2 interval."""
    xb = [Data["XXbetas"], Data["XXsbetas"], Data.get("XXsbetas_int")][which]
    ZZ = [Data["ZZ"], Data["ZZs"], Data.get("ZZs_int")][which]
    U = [Data["U"], Data["Us"], Data.get("Us_int")][which]
    n = b.shape[0]
    rows = len(xb[0])
    K = rows // n
    ids = np.arange(n) if which == 0 else np.repeat(np.arange(n), K)
    A = sum(len(c) for c in Data["col_inds"])
    out = np.zeros((rows, A))
    for t in range(len(xb)):
        zb = np.sum(np.asarray(ZZ[t]) * b[ids][:, Data["RE_inds2"][t]], axis=1)
        with np.errstate(all="ignore"):
            v = _TF[Data["trans_Funs"][t]](np.asarray(xb[t]) + zb)
        out[:, Data["col_inds"][t]] = np.asarray(U[t]) * v[:, None]
    return out


def eval_logpost_re(Data, interval_cens, b, Bs, gam, alp):
    """The state initialised at src/fast_LAPRWM.cpp:619-650 (== log_postREF :219-248)."""
    b = np.asarray(b)
    n = b.shape[0]
    K = len(Data["Pw"]) // n
    res = {}
    res["log_pyb"] = log_longF(Data, b)
    res["log_pb"] = -0.5 * np.sum((b @ np.asarray(Data["invD"])) * b, axis=1)
    gam = np.atleast_1d(gam)
    W2g = np.asarray(Data["W2"]) @ gam if gam.size else 0.0
    W2sg = np.asarray(Data["W2s"]) @ gam if gam.size else 0.0
    res["log_h"] = np.asarray(Data["W1"]) @ Bs + W2g + wlong(Data, b, 0) @ alp
    with np.errstate(all="ignore"):
        lin = np.asarray(Data["W1s"]) @ Bs + W2sg + wlong(Data, b, 1) @ alp
        res["H"] = (np.asarray(Data["Pw"]) * np.exp(lin)).reshape(n, K).sum(axis=1)
        if interval_cens:
            W2ig = np.asarray(Data["W2s_int"]) @ gam if gam.size else 0.0
            lin = np.asarray(Data["W1s_int"]) @ Bs + W2ig + wlong(Data, b, 2) @ alp
            res["HL"] = (np.asarray(Data["Pw_int"]) * np.exp(lin)).reshape(n, K).sum(axis=1)
            ev = np.asarray(Data["event"])
            r = np.zeros(n)
            r += np.where(ev == 1, res["log_h"], 0.0)
            r -= np.where((ev == 1) | (ev == 0), res["H"], 0.0)
            r += np.where(ev == 2, np.log(1 - np.exp(-res["H"])), 0.0)
            r += np.where(ev == 3, np.log(np.exp(-res["HL"]) - np.exp(-res["H"])), 0.0)
            res["log_ptb"] = r
        else:
            res["HL"] = res["H"].copy()
            res["log_ptb"] = np.asarray(Data["event"]) * res["log_h"] - res["H"]
    return res


def logPosterior(temp, Data, Wlong, Wlongs, Bs, gam, alp, priors, tauBs):
    """src/fast_logPost.cpp:8-26."""
    ev = np.asarray(Data["event"])
    gam = np.atleast_1d(gam)
    log_h = np.asarray(Data["W1"]) @ Bs + (np.asarray(Data["W2"]) @ gam if gam.size else 0) + Wlong @ alp
    H = np.asarray(Data["Pw"]) * np.exp(np.asarray(Data["W1s"]) @ Bs + (np.asarray(Data["W2s"]) @ gam if gam.size else 0) + Wlongs @ alp)
    zb = Bs - priors["mean_Bs_gammas"]
    za = alp - priors["mean_alphas"]
    out = temp * (np.sum(ev * log_h) - np.sum(H)) - 0.5 * tauBs * zb @ priors["Tau_Bs_gammas"] @ zb \
        - 0.5 * za @ priors["Tau_alphas"] @ za \
        + (priors["A_tau_Bs_gammas"] - 1) * np.log(tauBs) - priors["B_tau_Bs_gammas"] * tauBs
    if gam.size:
        zg = gam - priors["mean_gammas"]
        out -= 0.5 * zg @ priors["Tau_gammas"] @ zg
    return float(out)
