"""Seeded synthetic cohorts in the reference's ``Data`` contract.

Builds, without R, the lists that ``mvJointModelBayes()`` hands to ``lap_rwm_C``
(reference: R/mvJointModelBayes.R:651-733 builds ``Data``/``priors``/``inits``/``scales``/``Covs``;
src/fast_LAPRWM.cpp:434-601 reads them).  Field names are the reference's.  Matrices are
column-major (Fortran order) float64 like R matrices; every index vector is **0-based** here (the
Rcpp shim subtracts 1 where R is 1-based).

Shapes follow SURVEY.md section 8(d):

* ``config2``: 1 gaussian outcome, random intercept+slope (q=2), current-value association.
* ``config3``: gaussian + binomial(logit) + poisson(log), q=6, value+slope terms (T=6, A=6).
* ``config4``: config3 plus left truncation and interval censoring (status in {0,1,2,3}).

The quadrature grid is the reference's Gauss-Kronrod rule (R/gaussKronrod.R:3-10), the baseline
hazard design a cubic B-spline basis with the reference's knot rule (R/mvJointModelBayes.R:209-218),
the P-spline penalty ``crossprod(diff(I, differences=2)) + 1e-6 I`` (:600-601).
"""
from __future__ import annotations

import numpy as np

# R/gaussKronrod.R:3-10 (7 Gauss nodes first, then the 8 Kronrod nodes; not sorted)
GK_SK = np.array([
    -0.949107912342758524526189684047851, -0.741531185599394439863864773280788,
    -0.405845151377397166906606412076961, 0.0,
    0.405845151377397166906606412076961, 0.741531185599394439863864773280788,
    0.949107912342758524526189684047851, -0.991455371120812639206854697526329,
    -0.864864423359769072789712788640926, -0.586087235467691130294144838258730,
    -0.207784955007898467600689403773245, 0.207784955007898467600689403773245,
    0.586087235467691130294144838258730, 0.864864423359769072789712788640926,
    0.991455371120812639206854697526329])
GK_WK15 = np.array([
    0.063092092629978553290700663189204, 0.140653259715525918745189590510238,
    0.190350578064785409913256402421014, 0.209482141084727828012999174891714,
    0.190350578064785409913256402421014, 0.140653259715525918745189590510238,
    0.063092092629978553290700663189204, 0.022935322010529224963732008058970,
    0.104790010322250183839876322541518, 0.169004726639267902826583426598550,
    0.204432940075298892414161999234649, 0.204432940075298892414161999234649,
    0.169004726639267902826583426598550, 0.104790010322250183839876322541518,
    0.022935322010529224963732008058970])
GK_WK7 = np.array([
    0.129484966168869693270611432679082, 0.279705391489276667901467771423780,
    0.381830050505118944950369775488975, 0.417959183673469387755102040816327,
    0.381830050505118944950369775488975, 0.279705391489276667901467771423780,
    0.129484966168869693270611432679082])

H0_SCALE = 0.02   # Weibull-type baseline h0(t) = H0_SCALE * 1.3 * t^0.3

CONTROL_DEFAULTS = dict(  # R/mvJointModelBayes.R:8-22
    n_iter=300, n_burnin=1000, n_block=50, n_thin=300, target_acc=0.234,
    c0=1.0, c1=0.8, eps1=1e-6, eps2=1e-5, eps3=1e4, adaptCov=False, seed=1)


def gauss_kronrod(k: int = 15):
    """R/gaussKronrod.R:1-19."""
    if k == 7:
        return GK_SK[:7].copy(), GK_WK7.copy()
    return GK_SK.copy(), GK_WK15.copy()


def spline_design(knots: np.ndarray, x: np.ndarray, order: int = 4) -> np.ndarray:
    """``splines::splineDesign(knots, x, ord, outer.ok=TRUE)`` (R/mvJointModelBayes.R:382-383).

    Cox-de Boor recursion; rows outside ``[knots[ord-1], knots[-ord]]`` are zero; the basis is
    right-closed at the last knot (row at the right boundary is ``(0,...,0,1)``).
    """
    t = np.asarray(knots, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64).ravel()
    nb = len(t) - order
    lo, hi = t[order - 1], t[nb]
    inside = (x >= lo) & (x <= hi)
    # interval index j with t[j] <= x < t[j+1]; right boundary goes to the last non-empty span
    j = np.searchsorted(t, x, side="right") - 1
    j = np.clip(j, order - 1, nb - 1)
    out = np.zeros((x.size, nb), order="F")
    # de Boor's local basis evaluation, vectorised over x
    Bl = np.zeros((x.size, order))
    Bl[:, 0] = 1.0
    dl = np.zeros((x.size, order))
    dr = np.zeros((x.size, order))
    for d in range(1, order):
        dr[:, d - 1] = t[j + d] - x
        dl[:, d - 1] = x - t[j + 1 - d]
        saved = np.zeros(x.size)
        for r in range(d):
            den = dr[:, r] + dl[:, d - 1 - r]
            term = np.where(den != 0.0, Bl[:, r] / np.where(den != 0.0, den, 1.0), 0.0)
            Bl[:, r] = saved + dr[:, r] * term
            saved = dl[:, d - 1 - r] * term
        Bl[:, d] = saved
    rows = np.arange(x.size)
    for r in range(order):
        out[rows, j - (order - 1) + r] = Bl[:, r]
    out[~inside, :] = 0.0
    return out


def _F(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def make_cohort(config: str = "config3", n: int = 1000, seed: int | None = None,
                K: int = 15, mean_visits: float = 9.0, draw_jitter: float = 0.02,
                shared: dict | None = None) -> dict:
    """Return ``dict(Data=, priors=, inits=, scales=, Covs=, control=, truth=, interval_cens=)``.

    ``seed`` defaults to the SURVEY section 8(d) seed of the named config.  ``shared`` (from
    :func:`shared_model`) fixes everything that must be identical on all shards of one fit - the
    spline knots, the per-draw fixed effects and the initial survival parameters - so that each
    rank can generate its own block of subjects independently.
    """
    cfgs = {"config2": 20250, "config3": 20251, "config4": 20252}
    if config not in cfgs:
        raise ValueError(f"unknown config {config!r}")
    rng = np.random.default_rng(cfgs[config] if seed is None else seed)
    multi = config in ("config3", "config4")
    interval = config == "config4"
    sk, wk = gauss_kronrod(K)

    # ---- truth -------------------------------------------------------------------------
    if multi:
        fams = ["gaussian", "binomial", "poisson"]
        links = ["identity", "logit", "log"]
        betas = [np.array([1.0, 0.3]), np.array([-0.5, 0.25]), np.array([0.5, 0.1])]
        sd = np.array([0.77, 0.32, 0.6, 0.2, 0.4, 0.1])
        Rho = np.full((6, 6), 0.2)
        np.fill_diagonal(Rho, 1.0)
        D = Rho * np.outer(sd, sd)
        alphas_true = np.array([0.4, 0.3, 0.15, 0.2, 0.2, 0.3])
    else:
        fams, links = ["gaussian"], ["identity"]
        betas = [np.array([1.0, 0.3])]
        D = np.array([[0.6, 0.05], [0.05, 0.1]])
        alphas_true = np.array([0.4])
    sigma_true = 0.4
    gammas_true = np.array([-0.3, 0.2])
    O = len(fams)
    q = 2 * O
    b = rng.multivariate_normal(np.zeros(q), D, size=n)           # n x q
    W2 = np.column_stack([rng.binomial(1, 0.5, n).astype(float), rng.normal(size=n)])

    # ---- event times by inversion of the cumulative hazard on a grid --------------------
    t_max = 10.0

    def log_haz(t, rows):
        """log h_i(t) at truth for subjects ``rows`` (t broadcastable to rows x G)."""
        lh = np.log(H0_SCALE * 1.3) + 0.3 * np.log(t) + (W2[rows] @ gammas_true)[:, None]
        a = 0
        for k in range(O):
            b0 = b[rows, 2 * k][:, None]
            b1 = b[rows, 2 * k + 1][:, None]
            value = betas[k][0] + b0 + (betas[k][1] + b1) * t
            lh = lh + alphas_true[a] * value
            a += 1
            if multi:
                lh = lh + alphas_true[a] * (betas[k][1] + b1)
                a += 1
        return lh

    grid = np.linspace(0.0, t_max, 401)[1:]
    mid = grid - 0.5 * (grid[1] - grid[0])
    T_true = np.empty(n)
    E = rng.exponential(size=n)
    for s in range(0, n, 50000):
        rows = np.arange(s, min(n, s + 50000))
        cum = np.cumsum(np.exp(np.clip(log_haz(mid[None, :], rows), -50, 50)) * (grid[1] - grid[0]), axis=1)
        idx = (cum < E[rows, None]).sum(axis=1)
        T_true[rows] = np.where(idx >= grid.size, np.inf, grid[np.minimum(idx, grid.size - 1)])
    T_true = T_true * (1.0 + 1e-3 * rng.uniform(-1, 1, n))     # break grid ties
    Time = np.minimum(T_true, t_max)
    status = (T_true <= t_max).astype(float)                 # 1 event, 0 right-censored
    TimeL0 = np.zeros(n)                                      # left-truncation (entry) time
    TimeI = np.zeros(n)                                       # lower limit for interval-censored
    if interval:
        TimeL0 = np.minimum(rng.uniform(0, 2, n), 0.8 * Time)  # entry before the event time
        u = rng.uniform(size=n)
        ev = status == 1
        left = ev & (u < 0.15)
        intv = ev & (u >= 0.15) & (u < 0.30)
        status[left] = 2.0                                    # left-censored at Time
        # interval-censored: event known to lie in (Time, Time + U(0,4)]
        TimeI[intv] = Time[intv]
        Time = np.where(intv, Time + rng.uniform(0, 4, n), Time)
        status[intv] = 3.0
    Time = np.maximum(Time, 1e-4)                             # R/mvJointModelBayes.R:45

    # ---- visits -------------------------------------------------------------------------
    obs_end = np.where(status == 3.0, TimeI, Time) if interval else Time
    v = 1 + rng.poisson(mean_visits, n)
    N = int(v.sum())
    row_ptr = np.concatenate([[0], np.cumsum(v)])
    idL = np.repeat(np.arange(n), v)
    tt = rng.uniform(size=N) * obs_end[idL]
    tt[row_ptr[:-1]] = 0.0
    # sort within subject
    order = np.lexsort((tt, idL))
    tt = tt[order]
    idL2 = (row_ptr[1:] - 1).astype(np.int64)

    y, X, Z, Xbetas_truth = [], [], [], []
    for k in range(O):
        Xk = np.column_stack([np.ones(N), tt])
        eta = Xk @ betas[k] + b[idL, 2 * k] + b[idL, 2 * k + 1] * tt
        if fams[k] == "gaussian":
            yk = eta + sigma_true * rng.normal(size=N)
        elif fams[k] == "binomial":
            yk = rng.binomial(1, 1.0 / (1.0 + np.exp(-eta))).astype(float)
        else:
            yk = rng.poisson(np.exp(np.clip(eta, -30, 8))).astype(float)
        y.append(yk)
        X.append(_F(Xk))
        Z.append(_F(Xk.copy()))

    # ---- per-draw quantities (truth + jitter), R/mvJointModelBayes.R:760-770 -------------
    betas_draw = [bk + draw_jitter * rng.normal(size=2) for bk in betas]
    sigma_draw = sigma_true * np.exp(draw_jitter * rng.normal())
    if shared is not None:
        betas_draw, sigma_draw = shared["betas_draw"], shared["sigma_draw"]
    invD = np.linalg.inv(D)
    invD = 0.5 * (invD + invD.T)

    # ---- quadrature grid, R/mvJointModelBayes.R:117-152 -----------------------------------
    P = (Time - TimeL0) / 2.0
    st = np.outer(P, sk) + ((Time + TimeL0) / 2.0)[:, None]   # n x K
    st_vec = st.reshape(-1)                                   # c(t(st)): K consecutive rows per subject
    Pw = np.repeat(P, K) * np.tile(wk, n)
    if interval:
        upper_int = np.where(status == 3.0, TimeI, TimeL0)    # zero-length interval unless status 3
        P_int = (upper_int - TimeL0) / 2.0
        st_int = np.outer(P_int, sk) + ((upper_int + TimeL0) / 2.0)[:, None]
        st_int_vec = st_int.reshape(-1)
        Pw_int = np.repeat(P_int, K) * np.tile(wk, n)

    # ---- baseline hazard design, R/mvJointModelBayes.R:209-218,382-394 --------------------
    pp = np.quantile(Time, [0.05, 0.95])
    kn = np.linspace(pp[0], pp[1], 15)[1:-1]
    kn = kn[kn < Time.max()]
    rr = np.array([min(Time.min(), st_vec.min()), max(Time.max(), st_vec.max())])
    if interval:   # keep the lower-interval nodes inside the basis range
        rr = np.array([min(rr[0], st_int_vec.min()), max(rr[1], st_int_vec.max())])
    knots = np.sort(np.concatenate([np.repeat(rr, 4), kn]))
    if shared is not None:
        knots = shared["knots"]
    W1 = spline_design(knots, Time)
    W1s = spline_design(knots, st_vec)
    B = W1.shape[1]
    W2s = _F(np.repeat(W2, K, axis=0))

    # ---- association terms, R/mvJointModelBayes.R:407-504 ---------------------------------
    XXbetas, XXsbetas, XXsbetas_int = [], [], []
    XX, XXs, XXs_int, indFixed = [], [], [], []     # designs behind XXbetas = Xbetas_calc(XX, betas, indFixed, outcome)
    ZZ, ZZs, ZZs_int, U, Us, Us_int = [], [], [], [], [], []
    RE_inds = [np.array([2 * k, 2 * k + 1], dtype=np.int32) for k in range(O)]
    RE_inds2, col_inds, trans_Funs, outcome_of_term = [], [], [], []

    def add_term(k, kind):
        col = len(RE_inds2)
        if kind == "value":
            f = lambda t: betas_draw[k][0] + betas_draw[k][1] * t
            z = lambda t: np.column_stack([np.ones_like(t), t])
            RE_inds2.append(RE_inds[k].copy())
            indFixed.append(np.array([0, 1], dtype=np.int32))
        else:  # slope: fixed = ~1 with indFixed = 2, random = ~1 with indRandom = 2
            f = lambda t: np.full_like(t, betas_draw[k][1])
            z = lambda t: np.ones((t.size, 1))
            RE_inds2.append(RE_inds[k][1:].copy())
            indFixed.append(np.array([1], dtype=np.int32))
        XX.append(_F(z(Time))); XXs.append(_F(z(st_vec)))
        if interval:
            XXs_int.append(_F(z(st_int_vec)))
        XXbetas.append(f(Time)); XXsbetas.append(f(st_vec))
        ZZ.append(_F(z(Time))); ZZs.append(_F(z(st_vec)))
        U.append(_F(np.ones((n, 1)))); Us.append(_F(np.ones((n * K, 1))))
        if interval:
            XXsbetas_int.append(f(st_int_vec)); ZZs_int.append(_F(z(st_int_vec)))
            Us_int.append(_F(np.ones((n * K, 1))))
        col_inds.append(np.array([col], dtype=np.int32))
        trans_Funs.append("identity")
        outcome_of_term.append(k)

    for k in range(O):
        add_term(k, "value")
        if multi:
            add_term(k, "slope")
    A = len(RE_inds2)

    Data = dict(
        y=y, Xbetas=[X[k] @ betas_draw[k] for k in range(O)], X=X, Z=Z,
        RE_inds=RE_inds, RE_inds2=RE_inds2,
        idL=[idL.astype(np.int32) for _ in range(O)], idL2=[idL2.astype(np.int32) for _ in range(O)],
        sigmas=[sigma_draw if f == "gaussian" else float("nan") for f in fams],
        invD=_F(invD), fams=fams, links=links, trans_Funs=trans_Funs,
        Time=Time, event=status.copy(),
        idGK_fast=(np.arange(1, n + 1) * K - 1).astype(np.int32),
        idT_rsum=np.arange(n, dtype=np.int32),
        W1=W1, W1s=W1s, W2=_F(W2), W2s=W2s,
        U=U, Us=Us, col_inds=col_inds,
        row_inds_U=np.arange(n, dtype=np.int32), row_inds_Us=np.arange(n * K, dtype=np.int32),
        XXbetas=XXbetas, XXsbetas=XXsbetas, ZZ=ZZ, ZZs=ZZs,
        idT=[np.arange(n, dtype=np.int32) for _ in range(A)],
        idTs=[np.repeat(np.arange(n, dtype=np.int32), K) for _ in range(A)],
        Pw=Pw, nRisks=1, kn_strat_first=np.array([0]), kn_strat_last=np.array([B - 1]),
        outcome=np.array(outcome_of_term, dtype=np.int32), st=st_vec, knots=knots,
        XX=XX, XXs=XXs, XXs_int=XXs_int, indFixed=indFixed, betas=[np.asarray(bk, dtype=np.float64) for bk in betas_draw],
    )
    Data["idT2"] = Data["idT"]
    Data["idT2s"] = Data["idTs"]
    if interval:
        ev = status
        Data.update(
            Levent1=(ev == 1), Levent01=(ev == 1) | (ev == 0), Levent2=(ev == 2), Levent3=(ev == 3),
            W1s_int=spline_design(knots, st_int_vec), W2s_int=W2s.copy(order="F"),
            Us_int=Us_int, XXsbetas_int=XXsbetas_int, ZZs_int=ZZs_int, Pw_int=Pw_int, st_int=st_int_vec)
    else:
        e0 = np.zeros(0, dtype=bool)
        Data.update(Levent1=e0, Levent01=e0, Levent2=e0, Levent3=e0,
                    W1s_int=np.zeros((0, 0), order="F"), W2s_int=np.zeros((0, 0), order="F"),
                    Us_int=[np.zeros((0, 0), order="F")], XXsbetas_int=[np.zeros(0)],
                    ZZs_int=[np.zeros((0, 0), order="F")], Pw_int=np.zeros(0))

    # ---- priors, R/mvJointModelBayes.R:588-650 ---------------------------------------------
    DD = np.diff(np.eye(B), n=2, axis=0)
    Tau_Bs = DD.T @ DD + 1e-6 * np.eye(B)
    priors = dict(
        mean_Bs_gammas=np.zeros(B), Tau_Bs_gammas=_F(Tau_Bs),
        mean_gammas=np.zeros(2), Tau_gammas=_F(0.01 * np.eye(2)),
        mean_alphas=np.zeros(A), Tau_alphas=_F(0.01 * np.eye(A)),
        td_cols=[], rank_Tau_td_alphas=0,
        A_tau_Bs_gammas=1.0, B_tau_Bs_gammas=0.01,
        rank_Tau_Bs_gammas=float(np.linalg.matrix_rank(Tau_Bs)),
        shrink_gammas=False, A_tau_gammas=0.1, B_tau_gammas=0.1, rank_Tau_gammas=2.0,
        A_phi_gammas=0.5, B_phi_gammas=0.01,
        shrink_alphas=False, double_gamma_alphas=False,
        A_tau_alphas=0.5, B_tau_alphas=0.1, rank_Tau_alphas=float(A),
        A_phi_alphas=0.5, B_phi_alphas=0.01, A_nu_alphas=0.5, B_nu_alphas=1.0,
        A_xi_alphas=0.5, B_xi_alphas=1.0)

    # ---- inits / scales / Covs, R/mvJointModelBayes.R:716-733 ------------------------------
    # Bs_gammas: least-squares fit of the true log baseline hazard on the basis at the nodes
    sub = rng.choice(n * K, size=min(n * K, 20000), replace=False)
    tgt = np.log(H0_SCALE * 1.3) + 0.3 * np.log(st_vec[sub])
    Wsub = np.asarray(W1s[sub])
    Bs0 = np.linalg.solve(Wsub.T @ Wsub + 1e-3 * Tau_Bs + 1e-8 * np.eye(B), Wsub.T @ tgt)
    n_ev = max(1.0, float((status == 1).sum()))
    g0 = gammas_true + 0.01 * rng.normal(size=2)
    a0 = alphas_true + 0.01 * rng.normal(size=A)
    if shared is not None:
        Bs0, g0, a0, n_ev = shared["Bs0"], shared["g0"], shared["a0"], shared["n_ev"]
    inits = dict(
        b=_F(b + 0.05 * rng.normal(size=b.shape)),
        Bs_gammas=Bs0, gammas=g0, alphas=a0,
        tau_Bs_gammas=np.array([200.0]), tau_gammas=1.0, tau_alphas=1.0,
        phi_gammas=np.ones(2), phi_alphas=np.ones(A), tau_td_alphas=np.zeros(0))
    # Covs$b[,,i]: upper Cholesky factor of the Laplace covariance of b_i (GLMM part) at truth
    Prec = np.broadcast_to(invD, (n, q, q)).copy()
    for k in range(O):
        eta = X[k] @ betas[k] + b[idL, 2 * k] + b[idL, 2 * k + 1] * tt
        if fams[k] == "gaussian":
            w = np.full(N, 1.0 / sigma_true ** 2)
        elif fams[k] == "binomial":
            pr = 1.0 / (1.0 + np.exp(-eta)); w = pr * (1 - pr)
        else:
            w = np.exp(np.clip(eta, -30, 8))
        s0 = np.add.reduceat(w, row_ptr[:-1]); s1 = np.add.reduceat(w * tt, row_ptr[:-1])
        s2 = np.add.reduceat(w * tt * tt, row_ptr[:-1])
        Prec[:, 2 * k, 2 * k] += s0; Prec[:, 2 * k, 2 * k + 1] += s1
        Prec[:, 2 * k + 1, 2 * k] += s1; Prec[:, 2 * k + 1, 2 * k + 1] += s2
    Vb = np.linalg.inv(Prec)
    Rb = np.transpose(np.linalg.cholesky(Vb), (0, 2, 1))     # upper factors, n x q x q
    Covs = dict(
        b=np.asfortranarray(np.transpose(Rb, (1, 2, 0))),    # q x q x n cube like R's array
        Bs_gammas=_F(np.eye(B) * (4.0 / n_ev)), gammas=_F(np.eye(2) * (2.0 / n_ev)),
        alphas=_F(np.eye(A) * (2.0 / n_ev)))
    scales = dict(b=np.full(n, 5.76 / q), Bs_gammas=5.76 / B, gammas=5.76 / 2, alphas=5.76 / A)
    truth = dict(b=b, betas=betas, D=D, sigma=sigma_true, gammas=gammas_true, alphas=alphas_true,
                 row_ptr=row_ptr, times=tt, TimeL=TimeL0, TimeI=TimeI)
    return dict(Data=Data, priors=priors, inits=inits, scales=scales, Covs=Covs,
                control=dict(CONTROL_DEFAULTS), truth=truth, interval_cens=interval,
                multiState=False, config=config, n=n, K=K,
                shared=dict(knots=knots, betas_draw=betas_draw, sigma_draw=sigma_draw, Bs0=Bs0,
                            g0=g0, a0=a0, n_ev=n_ev))


def make_multistate_cohort(n: int = 300, seed: int = 20253, K: int = 15, mean_visits: float = 6.0,
                           draw_jitter: float = 0.02) -> dict:
    """Illness-death cohort in the reference's multi-state ``Data`` contract (``multiState = TRUE``,
    R/mvJointModelBayes.R:76-101,139-150,351-378,559-570,588-600,651-670; src/fast_LAPRWM.cpp:519-529).

    Every subject is at risk for transitions 1 (healthy -> ill) and 2 (healthy -> dead) on ``(0, stop]``; subjects who
    fall ill add a row for transition 3 (ill -> dead) on ``(T_ill, stop]``.  The survival part therefore has
    ``nT = sum(2 or 3)`` rows grouped by subject (``idT2``), the baseline hazard is stratified by transition
    (``nRisks = 3``; block-structured ``W1`` / ``Tau_Bs_gammas``; ``kn_strat_first/last``), ``W2`` holds
    transition-specific covariate effects, term 0 (value of the gaussian outcome) interacts with the transition
    (``U`` = stratum indicators, 3 columns of ``Wlong``) and term 1 (value of the binary outcome's linear predictor,
    ``expit``-transformed) has the default ``U`` of ones.
    """
    rng = np.random.default_rng(seed)
    sk, wk = gauss_kronrod(K)
    fams, links = ["gaussian", "binomial"], ["identity", "logit"]
    betas = [np.array([1.0, 0.3]), np.array([-0.5, 0.25])]
    sd = np.array([0.77, 0.32, 0.6, 0.2])
    Rho = np.full((4, 4), 0.2); np.fill_diagonal(Rho, 1.0)
    D = Rho * np.outer(sd, sd)
    O, q, nR = 2, 4, 3
    sigma_true = 0.4
    gammas_true = np.array([-0.3, 0.2, 0.1])                  # effect of x on each transition
    alphas_true = np.array([0.3, 0.2, 0.4, 0.5])              # value1 x transition (3), expit(value2)
    base = np.log(np.array([0.04, 0.02, 0.08]) * 1.3)
    b = rng.multivariate_normal(np.zeros(q), D, size=n)
    x = rng.normal(size=n)
    t_max = 10.0

    def log_haz(h, t, rows):
        v1 = betas[0][0] + b[rows, 0][:, None] + (betas[0][1] + b[rows, 1][:, None]) * t
        v2 = betas[1][0] + b[rows, 2][:, None] + (betas[1][1] + b[rows, 3][:, None]) * t
        return (base[h] + 0.3 * np.log(t) + gammas_true[h] * x[rows][:, None] + alphas_true[h] * v1
                + alphas_true[3] / (1.0 + np.exp(-v2)))

    grid = np.linspace(0.0, t_max, 401)[1:]
    dg = grid[1] - grid[0]
    mid = grid - 0.5 * dg
    rows_all = np.arange(n)

    def draw_time(h, start):
        haz = np.exp(np.clip(log_haz(h, mid[None, :], rows_all), -50, 50)) * dg
        haz = np.where(mid[None, :] > start[:, None], haz, 0.0)
        cum = np.cumsum(haz, axis=1)
        E = rng.exponential(size=n)
        idx = (cum < E[:, None]).sum(axis=1)
        t = np.where(idx >= grid.size, np.inf, grid[np.minimum(idx, grid.size - 1)])
        return t * (1.0 + 1e-3 * rng.uniform(-1, 1, n))

    T01, T02 = draw_time(0, np.zeros(n)), draw_time(1, np.zeros(n))
    stop1 = np.maximum(np.minimum(np.minimum(T01, T02), t_max), 1e-4)
    ill = (T01 < T02) & (T01 < t_max)
    dead_direct = (T02 <= T01) & (T02 < t_max)
    T12 = draw_time(2, np.where(ill, stop1, t_max))
    stop2 = np.minimum(np.maximum(T12, stop1 + 1e-3), t_max)
    # transition rows, grouped by subject
    idT, trans, TimeL, TimeR, event = [], [], [], [], []
    for i in range(n):
        idT += [i, i]; trans += [0, 1]; TimeL += [0.0, 0.0]; TimeR += [stop1[i], stop1[i]]
        event += [float(ill[i]), float(dead_direct[i])]
        if ill[i]:
            idT.append(i); trans.append(2); TimeL.append(stop1[i]); TimeR.append(stop2[i]); event.append(float(T12[i] < t_max))
    idT = np.array(idT, dtype=np.int32); trans = np.array(trans); TimeL = np.array(TimeL); TimeR = np.array(TimeR)
    event = np.array(event)
    nT = idT.size
    last_time = np.where(ill, stop2, stop1)

    # ---- visits ---------------------------------------------------------------------------------
    v = 1 + rng.poisson(mean_visits, n)
    N = int(v.sum())
    row_ptr = np.concatenate([[0], np.cumsum(v)])
    idL = np.repeat(np.arange(n), v)
    tt = rng.uniform(size=N) * last_time[idL]
    tt[row_ptr[:-1]] = 0.0
    tt = tt[np.lexsort((tt, idL))]
    idL2 = (row_ptr[1:] - 1).astype(np.int32)
    y, X, Z = [], [], []
    for k in range(O):
        Xk = np.column_stack([np.ones(N), tt])
        eta = Xk @ betas[k] + b[idL, 2 * k] + b[idL, 2 * k + 1] * tt
        y.append(eta + sigma_true * rng.normal(size=N) if k == 0 else rng.binomial(1, 1.0 / (1.0 + np.exp(-eta))).astype(float))
        X.append(_F(Xk)); Z.append(_F(Xk.copy()))
    betas_draw = [bk + draw_jitter * rng.normal(size=2) for bk in betas]
    sigma_draw = sigma_true * np.exp(draw_jitter * rng.normal())
    invD = np.linalg.inv(D); invD = 0.5 * (invD + invD.T)

    # ---- quadrature grid (R/mvJointModelBayes.R:121-135) -------------------------------------------
    P = (TimeR - TimeL) / 2.0
    st_vec = (np.outer(P, sk) + ((TimeR + TimeL) / 2.0)[:, None]).reshape(-1)
    Pw = np.repeat(P, K) * np.tile(wk, nT)
    trans_s = np.repeat(trans, K)

    # ---- stratified baseline hazard (R/mvJointModelBayes.R:154-208,351-373) -----------------------
    W1_blocks, W1s_blocks, knots_strat, Tau_blocks = [], [], [], []
    for h in range(nR):
        th, sh = TimeR[trans == h], st_vec[trans_s == h]
        pp = np.quantile(th, [0.05, 0.95])
        kn = np.linspace(pp[0], pp[1], 5)[1:-1]                 # lng.in.kn.multiState = 5
        rr = np.array([min(th.min(), sh.min()), max(th.max(), sh.max())])
        knots = np.sort(np.concatenate([np.repeat(rr, 4), kn]))
        w1 = np.zeros((nT, len(knots) - 4), order="F"); w1[trans == h] = spline_design(knots, th)
        w1s = np.zeros((nT * K, len(knots) - 4), order="F"); w1s[trans_s == h] = spline_design(knots, sh)
        W1_blocks.append(w1); W1s_blocks.append(w1s); knots_strat.append(w1.shape[1])
        DD = np.diff(np.eye(w1.shape[1]), n=2, axis=0)
        Tau_blocks.append(DD.T @ DD + 1e-6 * np.eye(w1.shape[1]))
    W1, W1s = _F(np.hstack(W1_blocks)), _F(np.hstack(W1s_blocks))
    knots_strat = np.array(knots_strat)
    B = int(knots_strat.sum())
    kn_last = (np.cumsum(knots_strat) - 1).astype(np.int32)
    kn_first = (np.cumsum(knots_strat) - knots_strat).astype(np.int32)
    Tau_Bs = np.zeros((B, B))
    for h in range(nR):
        Tau_Bs[kn_first[h]:kn_last[h] + 1, kn_first[h]:kn_last[h] + 1] = Tau_blocks[h]
    W2 = np.zeros((nT, nR)); W2[np.arange(nT), trans] = x[idT]
    W2s = _F(np.repeat(W2, K, axis=0))

    # ---- association terms ---------------------------------------------------------------------------
    ind = np.zeros((nT, nR)); ind[np.arange(nT), trans] = 1.0
    zfun = lambda t: np.column_stack([np.ones_like(t), t])
    XXbetas = [zfun(TimeR) @ betas_draw[k] for k in range(O)]
    XXsbetas = [zfun(st_vec) @ betas_draw[k] for k in range(O)]
    ZZ = [_F(zfun(TimeR)) for _ in range(O)]; ZZs = [_F(zfun(st_vec)) for _ in range(O)]
    U = [_F(ind), _F(np.ones((nT, 1)))]
    Us = [_F(np.repeat(ind, K, axis=0)), _F(np.ones((nT * K, 1)))]
    RE_inds = [np.array([0, 1], dtype=np.int32), np.array([2, 3], dtype=np.int32)]
    col_inds = [np.array([0, 1, 2], dtype=np.int32), np.array([3], dtype=np.int32)]
    A = 4
    idTs = np.repeat(idT, K)
    last_row = np.flatnonzero(np.r_[idT[1:] != idT[:-1], True]).astype(np.int32)
    e0 = np.zeros(0, dtype=bool)
    Data = dict(
        y=y, Xbetas=[X[k] @ betas_draw[k] for k in range(O)], X=X, Z=Z, RE_inds=RE_inds, RE_inds2=[r.copy() for r in RE_inds],
        idL=[idL.astype(np.int32) for _ in range(O)], idL2=[idL2 for _ in range(O)],
        sigmas=[sigma_draw, float("nan")], invD=_F(invD), fams=fams, links=links, trans_Funs=["identity", "expit"],
        Time=TimeR, event=event, idGK_fast=(np.arange(1, nT + 1) * K - 1).astype(np.int32), idT_rsum=last_row,
        W1=W1, W1s=W1s, W2=_F(W2), W2s=W2s, U=U, Us=Us, col_inds=col_inds,
        row_inds_U=np.arange(nT, dtype=np.int32), row_inds_Us=np.arange(nT * K, dtype=np.int32),
        XXbetas=XXbetas, XXsbetas=XXsbetas, ZZ=ZZ, ZZs=ZZs,
        idT=[idT for _ in range(O)], idTs=[idTs for _ in range(O)], idT2=[idT for _ in range(O)], idT2s=[idTs for _ in range(O)],
        Pw=Pw, nRisks=nR, kn_strat_first=kn_first, kn_strat_last=kn_last, outcome=np.array([0, 1], dtype=np.int32), st=st_vec,
        strata=trans.astype(np.int32), TimeL=TimeL,
        # designs behind XXbetas / XXsbetas = Xbetas_calc(XX, betas, indFixed, outcome) (R/supportFuns_mvJointModelBayes.R:43-55)
        XX=[_F(zfun(TimeR)) for _ in range(O)], XXs=[_F(zfun(st_vec)) for _ in range(O)], XXs_int=[],
        indFixed=[np.array([0, 1], dtype=np.int32) for _ in range(O)], betas=[np.asarray(bk, dtype=np.float64) for bk in betas_draw],
        Levent1=e0, Levent01=e0, Levent2=e0, Levent3=e0, W1s_int=np.zeros((0, 0), order="F"), W2s_int=np.zeros((0, 0), order="F"),
        Us_int=[np.zeros((0, 0), order="F")], XXsbetas_int=[np.zeros(0)], ZZs_int=[np.zeros((0, 0), order="F")], Pw_int=np.zeros(0))
    priors = dict(
        mean_Bs_gammas=np.zeros(B), Tau_Bs_gammas=_F(Tau_Bs), mean_gammas=np.zeros(nR), Tau_gammas=_F(0.01 * np.eye(nR)),
        mean_alphas=np.zeros(A), Tau_alphas=_F(0.01 * np.eye(A)), td_cols=[], rank_Tau_td_alphas=0,
        A_tau_Bs_gammas=1.0, B_tau_Bs_gammas=0.01, rank_Tau_Bs_gammas=float(np.linalg.matrix_rank(Tau_Bs)),
        shrink_gammas=False, A_tau_gammas=0.1, B_tau_gammas=0.1, rank_Tau_gammas=float(nR), A_phi_gammas=0.5, B_phi_gammas=0.01,
        shrink_alphas=False, double_gamma_alphas=False, A_tau_alphas=0.5, B_tau_alphas=0.1, rank_Tau_alphas=float(A),
        A_phi_alphas=0.5, B_phi_alphas=0.01, A_nu_alphas=0.5, B_nu_alphas=1.0, A_xi_alphas=0.5, B_xi_alphas=1.0)
    tgt = base[trans_s] + 0.3 * np.log(st_vec)
    Ws = np.asarray(W1s)
    Bs0 = np.linalg.solve(Ws.T @ Ws + 1e-3 * Tau_Bs + 1e-8 * np.eye(B), Ws.T @ tgt)
    n_ev = max(1.0, float(event.sum()))
    inits = dict(
        b=_F(b + 0.05 * rng.normal(size=b.shape)), Bs_gammas=Bs0, gammas=gammas_true + 0.01 * rng.normal(size=nR),
        alphas=alphas_true + 0.01 * rng.normal(size=A), tau_Bs_gammas=np.full(nR, 200.0),     # rep(., nRisks), :723-724
        tau_gammas=1.0, tau_alphas=1.0, phi_gammas=np.ones(nR), phi_alphas=np.ones(A), tau_td_alphas=np.zeros(0))
    Prec = np.broadcast_to(invD, (n, q, q)).copy()
    for k in range(O):
        eta = X[k] @ betas[k] + b[idL, 2 * k] + b[idL, 2 * k + 1] * tt
        pr = 1.0 / (1.0 + np.exp(-eta))
        w = np.full(N, 1.0 / sigma_true ** 2) if k == 0 else pr * (1 - pr)
        s0 = np.add.reduceat(w, row_ptr[:-1]); s1 = np.add.reduceat(w * tt, row_ptr[:-1]); s2 = np.add.reduceat(w * tt * tt, row_ptr[:-1])
        Prec[:, 2 * k, 2 * k] += s0; Prec[:, 2 * k, 2 * k + 1] += s1; Prec[:, 2 * k + 1, 2 * k] += s1; Prec[:, 2 * k + 1, 2 * k + 1] += s2
    Rb = np.transpose(np.linalg.cholesky(np.linalg.inv(Prec)), (0, 2, 1))
    Covs = dict(b=np.asfortranarray(np.transpose(Rb, (1, 2, 0))), Bs_gammas=_F(np.eye(B) * (4.0 / n_ev)),
                gammas=_F(np.eye(nR) * (2.0 / n_ev)), alphas=_F(np.eye(A) * (2.0 / n_ev)))
    scales = dict(b=np.full(n, 5.76 / q), Bs_gammas=5.76 / B, gammas=5.76 / nR, alphas=5.76 / A)
    truth = dict(b=b, betas=betas, D=D, sigma=sigma_true, gammas=gammas_true, alphas=alphas_true, row_ptr=row_ptr, times=tt)
    return dict(Data=Data, priors=priors, inits=inits, scales=scales, Covs=Covs, control=dict(CONTROL_DEFAULTS), truth=truth,
                interval_cens=False, multiState=True, config="multistate", n=n, K=K)


def shared_model(config: str, n_total: int, pilot_n: int = 20000, K: int = 15) -> dict:
    """Fit-wide quantities every shard of a multi-GPU run must agree on, derived from a small
    seeded pilot cohort: spline knots (boundary knots widened to the whole possible time range),
    the per-draw fixed effects, and the initial survival parameters / proposal scales."""
    pilot = make_cohort(config, n=pilot_n, K=K)
    sh = dict(pilot["shared"])
    kn = sh["knots"][4:-4]
    hi = 14.0 if config == "config4" else 10.0
    sh["knots"] = np.sort(np.concatenate([np.repeat([0.0, hi], 4), kn]))
    # refit the initial spline coefficients on the widened basis
    st = pilot["Data"]["st"]
    sub = np.random.default_rng(11).choice(st.size, size=min(st.size, 40000), replace=False)
    Wsub = np.asarray(spline_design(sh["knots"], st[sub]))
    B = Wsub.shape[1]
    DD = np.diff(np.eye(B), n=2, axis=0)
    tgt = np.log(H0_SCALE * 1.3) + 0.3 * np.log(st[sub])
    sh["Bs0"] = np.linalg.solve(Wsub.T @ Wsub + 1e-3 * (DD.T @ DD) + 1e-8 * np.eye(B), Wsub.T @ tgt)
    sh["n_ev"] = sh["n_ev"] * n_total / pilot_n
    return sh
