"""BASELINE config 1: the pbc2 joint model of man/mvJointModelBayes.Rd:204-227,

    mvglmer(log(serBilir) ~ year + (year | id))  +  coxph(Surv(years, status2) ~ drug + age)
    -> mvJointModelBayes(..., timeVar = "year")

built without R from the columns in tests/golden/pbc2_columns.npz (extracted from the reference's
data/pbc2.rda by tests/golden/make_pbc2.py).  Host-side data preparation only - the hot path runs
through jmbayes_b200.api.

* Stage 1 (mvglmer, JAGS/Stan in the reference; out of scope and absent from the image) is replaced by
  a documented surrogate: a maximum-likelihood linear mixed model fitted by EM, whose estimates play
  the role of one posterior draw (betas, sigma, D) and whose conditional means / covariances of b_i
  play the role of mvglmer's b draws (inits$b and Covs$b = chol(var(b_i)), R/mvJointModelBayes.R:726-731).
* Designs follow R/mvJointModelBayes.R: Time < 1e-4 -> 1e-4 (:45); P = Time/2, st = outer(P, sk + 1)
  (:117-138); 13 interior knots at the inner points of seq(q05, q95, length = 15), boundary knots
  range(Time, st) repeated 4 times (:209-218); W1 / W1s = cubic B-spline designs (:382-383,389-392);
  W2 = model.matrix(~ drug + age)[, -1] (:385-386); value term XX = ZZ = [1, t], U = 1 (:407-504);
  P-spline penalty and the default priors (:588-650).
* inits / Covs: marglogLik2 (R/margLogLik2.R) restated: BFGS on -logPosterior with the analytic
  gradient (tau_Bs_gammas fixed at 200, :716-722), Hessian by central differences of the gradient
  (cd.vec, :76-92), Covs = diagonal blocks of its inverse (:114-131).  The log-posterior and gradient
  are passed in as callables, so the same routine runs on the oracle (CPU tests) or on the CUDA library.
"""
from __future__ import annotations

import os

import numpy as np

from .cohorts import CONTROL_DEFAULTS, gauss_kronrod, spline_design

HERE = os.path.dirname(os.path.abspath(__file__))
COLUMNS = os.path.join(os.path.dirname(HERE), "tests", "golden", "pbc2_columns.npz")


def _F(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def fit_lmm(y, X, Z, ids, n, iters=300):
    """ML fit of y = X beta + Z b_i + eps by EM; returns beta, sigma, D, E[b_i|y] (n x q), Var[b_i|y] (n x q x q)."""
    q = Z.shape[1]
    beta = np.linalg.lstsq(X, y, rcond=None)[0]
    sigma2 = float(np.var(y - X @ beta))
    D = np.eye(q)
    ZtZ = np.zeros((n, q, q))
    for a in range(q):
        for c in range(q):
            ZtZ[:, a, c] = np.bincount(ids, Z[:, a] * Z[:, c], minlength=n)
    for _ in range(iters):
        r = y - X @ beta
        Ztr = np.stack([np.bincount(ids, Z[:, a] * r, minlength=n) for a in range(q)], axis=1)
        V = np.linalg.inv(ZtZ / sigma2 + np.linalg.inv(D)[None])
        b = np.einsum("nij,nj->ni", V, Ztr) / sigma2
        zb = np.sum(Z * b[ids], axis=1)
        beta = np.linalg.lstsq(X, y - zb, rcond=None)[0]
        res = y - X @ beta - zb
        sigma2 = float((res @ res + np.einsum("nij,nij->", ZtZ, V)) / y.size)
        D = (np.einsum("ni,nj->ij", b, b) + V.sum(axis=0)) / n
    return beta, np.sqrt(sigma2), D, b, V


def marglogLik2(log_post, grad, B, p, A, tau_Bs=200.0):
    """R/margLogLik2.R with fixed_tau_Bs_gammas = TRUE: returns (inits dict, Covs dict, opt value).

    ``log_post(Bs, gammas, alphas)`` and ``grad(Bs, gammas, alphas)`` evaluate logPosterior /
    gradient_logPosterior (all B + p + A + 1 entries) at temp = 1, tauBs = ``tau_Bs``."""
    from scipy.optimize import minimize
    d = B + p + A

    def split(th):
        return th[:B], th[B:B + p], th[B + p:]

    def fn(th):
        return -log_post(*split(th))

    def gr(th):
        return -np.asarray(grad(*split(th)))[:d]       # head(out, -1), :73-75

    opt = minimize(fn, np.zeros(d), jac=gr, method="BFGS", options=dict(gtol=1e-6, maxiter=500))
    x = opt.x
    H = np.zeros((d, d))                                # cd.vec, :79-92
    ex = np.maximum(np.abs(x), 1.0)
    for i in range(d):
        x1, x2 = x.copy(), x.copy()
        x1[i] += 1e-3 * ex[i]
        x2[i] -= 1e-3 * ex[i]
        H[:, i] = (gr(x1) - gr(x2)) / (x1[i] - x2[i])
    H = 0.5 * (H + H.T)
    invH = np.linalg.inv(H)
    invH = 0.5 * (invH + invH.T)
    w, Q = np.linalg.eigh(invH)                         # nearPD stand-in: clip the spectrum
    invH = (Q * np.maximum(w, 1e-10 * w.max())) @ Q.T
    Bs, g, a = split(x)
    inits = dict(Bs_gammas=Bs.copy(), gammas=g.copy(), alphas=a.copy(), tau_Bs_gammas=np.array([tau_Bs]))
    Covs = dict(Bs_gammas=_F(invH[:B, :B]), gammas=_F(invH[B:B + p, B:B + p]), alphas=_F(invH[B + p:, B + p:]))
    return inits, Covs, float(opt.fun)


def make_pbc2(columns: str = COLUMNS, K: int = 15) -> dict:
    """``Data`` / ``priors`` / ``scales`` / ``Covs$b`` / ``inits$b`` of config 1 (survival inits come from
    :func:`marglogLik2`, see :func:`finish_inits`)."""
    c = np.load(columns)
    ids = c["id"].astype(np.int64)
    n = int(c["years"].size)
    t = c["year"].astype(np.float64)
    y = np.log(c["serBilir"].astype(np.float64))
    N = y.size
    X = np.column_stack([np.ones(N), t])
    beta, sigma, D, b, V = fit_lmm(y, X, X, ids, n)
    Time = np.maximum(c["years"].astype(np.float64), 1e-4)
    event = c["status2"].astype(np.float64)
    sk, wk = gauss_kronrod(K)
    P = Time / 2.0
    st = np.outer(P, sk + 1.0).reshape(-1)              # c(t(st)): K consecutive rows per subject
    Pw = np.repeat(P, K) * np.tile(wk, n)
    pp = np.quantile(Time, [0.05, 0.95])
    kn = np.linspace(pp[0], pp[1], 15)[1:-1]
    kn = kn[kn < Time.max()]
    rr = np.array([min(Time.min(), st.min()), max(Time.max(), st.max())])
    knots = np.sort(np.concatenate([np.repeat(rr, 4), kn]))
    W1, W1s = spline_design(knots, Time), spline_design(knots, st)
    B = W1.shape[1]
    W2 = np.column_stack([c["drug"].astype(np.float64), c["age"].astype(np.float64)])
    counts = np.bincount(ids, minlength=n)
    row_ptr = np.concatenate([[0], np.cumsum(counts)])
    Data = dict(
        y=[y], Xbetas=[X @ beta], X=[_F(X)], Z=[_F(X.copy())],
        RE_inds=[np.array([0, 1], dtype=np.int32)], RE_inds2=[np.array([0, 1], dtype=np.int32)],
        idL=[ids.astype(np.int32)], idL2=[(row_ptr[1:] - 1).astype(np.int32)],
        sigmas=[float(sigma)], invD=_F(np.linalg.inv(D)), fams=["gaussian"], links=["identity"],
        trans_Funs=["identity"], Time=Time, event=event,
        idGK_fast=(np.arange(1, n + 1) * K - 1).astype(np.int32), idT_rsum=np.arange(n, dtype=np.int32),
        W1=W1, W1s=W1s, W2=_F(W2), W2s=_F(np.repeat(W2, K, axis=0)),
        U=[_F(np.ones((n, 1)))], Us=[_F(np.ones((n * K, 1)))], col_inds=[np.array([0], dtype=np.int32)],
        XXbetas=[beta[0] + beta[1] * Time], XXsbetas=[beta[0] + beta[1] * st],
        ZZ=[_F(np.column_stack([np.ones(n), Time]))], ZZs=[_F(np.column_stack([np.ones(n * K), st]))],
        Pw=Pw, st=st, knots=knots,
        Levent1=np.zeros(0, dtype=bool), Levent01=np.zeros(0, dtype=bool), Levent2=np.zeros(0, dtype=bool),
        Levent3=np.zeros(0, dtype=bool), W1s_int=np.zeros((0, 0), order="F"), W2s_int=np.zeros((0, 0), order="F"),
        Us_int=[np.zeros((0, 0), order="F")], XXsbetas_int=[np.zeros(0)], ZZs_int=[np.zeros((0, 0), order="F")],
        Pw_int=np.zeros(0))
    DD = np.diff(np.eye(B), n=2, axis=0)
    Tau_Bs = DD.T @ DD + 1e-6 * np.eye(B)
    priors = dict(
        mean_Bs_gammas=np.zeros(B), Tau_Bs_gammas=_F(Tau_Bs), mean_gammas=np.zeros(2), Tau_gammas=_F(0.01 * np.eye(2)),
        mean_alphas=np.zeros(1), Tau_alphas=_F(0.01 * np.eye(1)), td_cols=[], rank_Tau_td_alphas=0,
        A_tau_Bs_gammas=1.0, B_tau_Bs_gammas=0.01, rank_Tau_Bs_gammas=float(np.linalg.matrix_rank(Tau_Bs)),
        shrink_gammas=False, A_tau_gammas=0.1, B_tau_gammas=0.1, rank_Tau_gammas=2.0, A_phi_gammas=0.5, B_phi_gammas=0.01,
        shrink_alphas=False, double_gamma_alphas=False, A_tau_alphas=0.5, B_tau_alphas=0.1, rank_Tau_alphas=1.0,
        A_phi_alphas=0.5, B_phi_alphas=0.01, A_nu_alphas=0.5, B_nu_alphas=1.0, A_xi_alphas=0.5, B_xi_alphas=1.0)
    Rb = np.transpose(np.linalg.cholesky(V), (0, 2, 1))          # upper factors
    Wlong = _F((beta[0] + b[:, 0] + (beta[1] + b[:, 1]) * Time)[:, None])
    bs = np.repeat(b, K, axis=0)
    Wlongs = _F((beta[0] + bs[:, 0] + (beta[1] + bs[:, 1]) * st)[:, None])
    return dict(Data=Data, priors=priors, Covs_b=np.asfortranarray(np.transpose(Rb, (1, 2, 0))),
                b=_F(b), scales=dict(b=np.full(n, 5.76 / 2), Bs_gammas=5.76 / B, gammas=5.76 / 2, alphas=5.76 / 1),
                control=dict(CONTROL_DEFAULTS), Wlong=Wlong, Wlongs=Wlongs, n=n, K=K, B=B,
                stage1=dict(betas=beta, sigma=sigma, D=D), interval_cens=False, multiState=False)


def finish_inits(pb: dict, log_post, grad) -> dict:
    """Complete a :func:`make_pbc2` result with marglogLik2's initial values and proposal covariances."""
    inits, Covs, val = marglogLik2(log_post, grad, pb["B"], 2, 1)
    inits.update(b=pb["b"], tau_gammas=1.0, tau_alphas=1.0, phi_gammas=np.ones(2), phi_alphas=np.ones(1),
                 tau_td_alphas=np.zeros(0))
    Covs["b"] = pb["Covs_b"]
    out = dict(pb)
    out.update(inits=inits, Covs=Covs, laplace_value=val)
    return out
