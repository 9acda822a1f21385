"""Host-side mirror of ``survfitJM.mvJMbayes()``'s compute loop (BASELINE config 5) over the C ABI of
``include/jmb200_survfit.h``.

R is not available in this image, so this module plays the role of the R code around the ``.Call``s:
``newdata_designs`` builds, for synthetic new subjects, the per-subject lists ``survMats.last[[i]]`` and
``survMats[[i]][[l]]`` (R/survfitJM.mvJMbayes.R:85-116,149-276) flattened over subjects, ``posterior_draws``
stands in for ``object$mcmc`` / ``object$statistics$postMeans``, and :func:`survfitJM` runs the mode search,
the Metropolis-Hastings loop over the M posterior draws and the horizon integrals (:277-414) on the GPU.
There is no CPU fallback: without ``libjmb200.so`` or a CUDA device every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._abi import c_double_p, c_double_pp, c_int32_p, c_int32_pp, c_int64_p
from .cohorts import _F, gauss_kronrod, spline_design


class SvftData(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_outcomes", C.c_int32), ("q", C.c_int32), ("n_terms", C.c_int32),
        ("n_alphas", C.c_int32), ("n_Bs", C.c_int32), ("n_gammas", C.c_int32), ("K", C.c_int32),
        ("L", C.c_int32), ("subject_offset", C.c_int64),
        ("N", c_int64_p), ("fams", c_int32_p), ("links", c_int32_p), ("q_k", c_int32_p),
        ("RE_inds", c_int32_pp), ("p_k", c_int32_p), ("y", c_double_pp), ("X", c_double_pp),
        ("Z", c_double_pp), ("idL", c_int32_pp),
        ("r_t", c_int32_p), ("RE_inds2", c_int32_pp), ("u_t", c_int32_p), ("col_inds", c_int32_pp),
        ("trans_Funs", c_int32_p), ("outcome", c_int32_p), ("f_t", c_int32_p), ("indFixed", c_int32_pp),
        ("W1s", c_double_p), ("W2s", c_double_p), ("Pw", c_double_p),
        ("XXs", c_double_pp), ("ZZs", c_double_pp), ("Us", c_double_pp),
        ("n_horizons", c_int32_p), ("W1s_h", c_double_p), ("W2s_h", c_double_p), ("Pw_h", c_double_p),
        ("XXs_h", c_double_pp), ("ZZs_h", c_double_pp), ("Us_h", c_double_pp),
        ("knots", c_double_p), ("n_knots", C.c_int32), ("ord", C.c_int32), ("st", c_double_p), ("st_h", c_double_p),
    ]


class SvftParams(C.Structure):
    _fields_ = [("M", C.c_int32), ("betas", c_double_pp), ("sigmas", c_double_pp), ("invD", c_double_p),
                ("Bs_gammas", c_double_p), ("gammas", c_double_p), ("alphas", c_double_p)]


class SvftControl(C.Structure):
    _fields_ = [("scale", C.c_double), ("df", C.c_double), ("log", C.c_int32), ("CI_levels", C.c_double * 2),
                ("seed", C.c_uint32), ("maxit", C.c_int32), ("parscale", C.c_double), ("reltol", C.c_double),
                ("ndeps", C.c_double), ("cd_eps", C.c_double)]


class SvftOut(C.Structure):
    _fields_ = [("modes", c_double_p), ("hessians", c_double_p), ("convergence", c_int32_p), ("counts", c_int32_p),
                ("summaries", c_double_p), ("full", c_double_p), ("accept_rate", c_double_p), ("b_last", c_double_p)]


# ------------------------------------------------------------------------------------------------
def marshal_data(sv: dict, subject_offset: int = 0):
    k = _abi.Keep()
    d = SvftData()
    d.n, d.K, d.L, d.q = int(sv["n"]), int(sv["K"]), int(sv["L"]), int(sv["q"])
    d.n_outcomes, d.n_terms = len(sv["y_long"]), len(sv["XXs"])
    d.n_alphas = int(sum(len(c) for c in sv["col_inds"]))
    from_knots = sv.get("W1s") is None            # bases evaluated by the library from (knots, times)
    d.n_Bs = int(len(sv["knots"]) - sv.get("ord", 4)) if from_knots else int(np.asarray(sv["W1s"]).shape[1])
    W2s = np.asarray(sv["W2s"])
    d.n_gammas = int(W2s.shape[1]) if W2s.ndim == 2 else 0
    d.subject_offset = int(subject_offset)
    d.N = k.i64([len(y) for y in sv["y_long"]])
    d.fams = k.i32([_abi.FAMS[f] for f in sv["fams"]])
    d.links = k.i32([_abi.LINKS[f] for f in sv["links"]])
    d.q_k = k.i32([len(r) for r in sv["RE_inds"]])
    d.RE_inds = k.i32_list(sv["RE_inds"])
    d.p_k = k.i32([np.asarray(x).shape[1] for x in sv["X_long"]])
    d.y, d.X, d.Z = k.f64_list(sv["y_long"]), k.f64_list(sv["X_long"]), k.f64_list(sv["Z_long"])
    d.idL = k.i32_list(sv["idL"])
    d.r_t = k.i32([len(r) for r in sv["RE_inds2"]])
    d.RE_inds2 = k.i32_list(sv["RE_inds2"])
    d.u_t = k.i32([len(c) for c in sv["col_inds"]])
    d.col_inds = k.i32_list(sv["col_inds"])
    d.trans_Funs = k.i32([_abi.TRANS[f] for f in sv["trans_Funs"]])
    d.outcome = k.i32(sv["outcome"])
    d.f_t = k.i32([len(f) for f in sv["indFixed"]])
    d.indFixed = k.i32_list(sv["indFixed"])
    d.W2s, d.Pw = k.f64(sv["W2s"]), k.f64(sv["Pw"])
    if from_knots:
        d.knots, d.n_knots, d.ord = k.f64(sv["knots"]), int(len(sv["knots"])), int(sv.get("ord", 4))
        d.st, d.st_h = k.f64(sv["st"]), k.f64(sv["st_h"])
    else:
        d.W1s = k.f64(sv["W1s"])
    d.XXs, d.ZZs, d.Us = k.f64_list(sv["XXs"]), k.f64_list(sv["ZZs"]), k.f64_list(sv["Us"])
    d.n_horizons = k.i32(sv["n_horizons"])
    d.W2s_h, d.Pw_h = k.f64(sv["W2s_h"]), k.f64(sv["Pw_h"])
    if not from_knots:
        d.W1s_h = k.f64(sv["W1s_h"])
    d.XXs_h, d.ZZs_h, d.Us_h = k.f64_list(sv["XXs_h"]), k.f64_list(sv["ZZs_h"]), k.f64_list(sv["Us_h"])
    return d, k


def marshal_params(par: dict):
    """``par``: betas (list of M x p_k), sigmas (list of [M] or None), inv_D (M x q x q), Bs_gammas (M x B),
    gammas (M x p), alphas (M x A) - the mcmc matrices of the reference, draws in rows."""
    k = _abi.Keep()
    p = SvftParams()
    betas = [np.atleast_2d(np.asarray(b, dtype=np.float64)) for b in par["betas"]]
    p.M = int(betas[0].shape[0])
    p.betas = k.f64_list(betas)
    sig = []
    for s in par["sigmas"]:
        sig.append(k.f64(np.atleast_1d(np.asarray(s, dtype=np.float64))) if s is not None else C.cast(None, c_double_p))
    arr = (c_double_p * max(len(sig), 1))(*sig)
    k.refs.append(arr)
    p.sigmas = C.cast(arr, c_double_pp)
    invD = np.asarray(par["inv_D"], dtype=np.float64)
    if invD.ndim == 2:
        invD = invD[None, :, :]
    p.invD = k.f64(invD)                       # Fortran order: element (m, j, l) at m + j M + l M q
    p.Bs_gammas = k.f64(np.atleast_2d(np.asarray(par["Bs_gammas"], dtype=np.float64)))
    g = np.atleast_2d(np.asarray(par["gammas"], dtype=np.float64))
    p.gammas = k.f64(g)
    p.alphas = k.f64(np.atleast_2d(np.asarray(par["alphas"], dtype=np.float64)))
    return p, k


CONTROL_DEFAULTS = dict(scale=1.6, df=4.0, log=False, CI_levels=(0.025, 0.975), seed=1, maxit=200, parscale=0.1,
                        reltol=float(np.sqrt(np.finfo(np.float64).eps)), ndeps=1e-3, cd_eps=1e-3)


def marshal_control(control: dict | None = None):
    c = dict(CONTROL_DEFAULTS)
    c.update(control or {})
    s = SvftControl()
    s.scale, s.df, s.log = float(c["scale"]), float(c["df"]), int(bool(c["log"]))
    s.CI_levels[0], s.CI_levels[1] = float(c["CI_levels"][0]), float(c["CI_levels"][1])
    s.seed = int(c["seed"]) & 0xFFFFFFFF
    s.maxit, s.parscale, s.reltol = int(c["maxit"]), float(c["parscale"]), float(c["reltol"])
    s.ndeps, s.cd_eps = float(c["ndeps"]), float(c["cd_eps"])
    return s


def alloc_out(n, q, L, M, full=True, modes=None, hessians=None, want_hessians=True):
    bufs = dict(
        modes=np.zeros((n, q), order="F") if modes is None else np.asfortranarray(np.array(modes, dtype=np.float64)),
        hessians=np.zeros((q, q, n), order="F") if hessians is None else np.asfortranarray(np.array(hessians, dtype=np.float64)),
        convergence=np.zeros(n, dtype=np.int32), counts=np.zeros((2, n), dtype=np.int32, order="F"),
        summaries=np.full((L, 4, n), np.nan, order="F"), accept_rate=np.zeros(n), b_last=np.zeros((n, q), order="F"))
    if full:
        bufs["full"] = np.full((L, M, n), np.nan, order="F")
    o = SvftOut()
    for name, a in bufs.items():
        if name == "hessians" and not want_hessians:
            continue
        setattr(o, name, a.ctypes.data_as(c_int32_p if a.dtype == np.int32 else c_double_p))
    return o, bufs


# ------------------------------------------------------------------------------------------------
def _bind(L):
    vp = C.c_void_p
    L.jmb_svft_create.argtypes = [C.POINTER(SvftData), C.c_int, C.POINTER(vp)]
    L.jmb_svft_destroy.argtypes = [vp]
    L.jmb_svft_modes.argtypes = [vp, C.POINTER(SvftParams), C.POINTER(SvftControl), C.POINTER(SvftOut)]
    L.jmb_svft_sample.argtypes = [vp, C.POINTER(SvftParams), C.POINTER(SvftControl), C.POINTER(SvftOut)]
    L.jmb_survfitJM.argtypes = [vp, C.POINTER(SvftParams), C.POINTER(SvftParams), C.POINTER(SvftControl), C.POINTER(SvftOut)]
    L.jmb_svft_info.argtypes = [vp, C.POINTER(C.c_float), c_int64_p, c_double_p]
    return L


def _lib():
    from . import api
    L = api.lib()
    if not getattr(L, "_svft_bound", False):
        _bind(L)
        L._svft_bound = True
    return L


def _check(rc):
    if rc != 0:
        from . import api
        raise api.JmbError(_lib().jmb_last_error().decode())


class SurvFit:
    """The designs of a set of new subjects, resident on one GPU (``survMats.last`` + ``survMats``)."""

    def __init__(self, sv: dict, device: int = 0, subject_offset: int = 0):
        self._h = C.c_void_p()
        d, keep = marshal_data(sv, subject_offset)
        self.n, self.q, self.L, self.K = d.n, d.q, d.L, d.K
        _check(_lib().jmb_svft_create(C.byref(d), device, C.byref(self._h)))
        del keep

    def close(self):
        if self._h:
            _lib().jmb_svft_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @staticmethod
    def _result(bufs):
        return dict(modes=bufs["modes"], hessians=bufs["hessians"], convergence=bufs["convergence"], counts=bufs["counts"],
                    summaries=bufs["summaries"], full_results=bufs.get("full"), accept_rate=bufs["accept_rate"],
                    b_last=bufs["b_last"])

    def modes(self, post_means: dict, control: dict | None = None):
        """modes.b and opt$hessian of every subject (R/survfitJM.mvJMbayes.R:277-305)."""
        p, k = marshal_params(post_means)
        out, bufs = alloc_out(self.n, self.q, self.L, 1, full=False)
        ct = marshal_control(control)
        _check(_lib().jmb_svft_modes(self._h, C.byref(p), C.byref(ct), C.byref(out)))
        del k
        return self._result(bufs)

    def sample(self, draws: dict, modes, hessians, control: dict | None = None, full: bool = True):
        """The MH loop over the M draws and the horizon integrals from given modes / Hessians (:306-414)."""
        p, k = marshal_params(draws)
        out, bufs = alloc_out(self.n, self.q, self.L, p.M, full=full, modes=modes, hessians=hessians)
        ct = marshal_control(control)
        _check(_lib().jmb_svft_sample(self._h, C.byref(p), C.byref(ct), C.byref(out)))
        del k
        return self._result(bufs)

    def survfitJM(self, post_means: dict, draws: dict, control: dict | None = None, full: bool = True, hessians: bool = True):
        """The whole per-subject loop of survfitJM.mvJMbayes (:277-414).  ``hessians=False``: opt$hessian stays on the
        device (survfitJM itself only uses it for Vars.b / invVars.b)."""
        pm, k1 = marshal_params(post_means)
        pd, k2 = marshal_params(draws)
        out, bufs = alloc_out(self.n, self.q, self.L, pd.M, full=full, want_hessians=hessians)
        ct = marshal_control(control)
        _check(_lib().jmb_survfitJM(self._h, C.byref(pm), C.byref(pd), C.byref(ct), C.byref(out)))
        del k1, k2
        return self._result(bufs)

    def info(self):
        ms, ln, rb = C.c_float(0), C.c_int64(0), C.c_double(0)
        _check(_lib().jmb_svft_info(self._h, C.byref(ms), C.byref(ln), C.byref(rb)))
        return dict(last_kernel_ms=ms.value, launches=ln.value, record_bytes_per_subject=rb.value)


def survfitJM(sv: dict, post_means: dict, draws: dict, control: dict | None = None, device: int = 0, full: bool = True):
    """``survfitJM(object, newdata, survTimes, last.time, M, scale, log, CI.levels, seed)``'s compute, on the GPU."""
    with SurvFit(sv, device) as h:
        return h.survfitJM(post_means, draws, control, full)


# ------------------------------------------------------------------------------------------------
# synthetic new subjects (config 5): designs as R/survfitJM.mvJMbayes.R:85-116,149-276 builds them
# ------------------------------------------------------------------------------------------------
def newdata_designs(cohort: dict, L: int = 10, horizon_step: float = 0.5, last_frac: float = 0.5,
                    subjects: np.ndarray | None = None, dense_basis: bool = True) -> dict:
    """All (or ``subjects``) of a cohort as ``newdata`` truncated at ``last.time = last_frac * Time``, with
    ``L`` prediction times ``last.time + (1..L) * horizon_step`` each."""
    D, truth, K = cohort["Data"], cohort["truth"], cohort["K"]
    n_all = cohort["n"]
    sel = np.arange(n_all) if subjects is None else np.asarray(subjects)
    n = sel.size
    O = len(D["y"])
    sk, wk = gauss_kronrod(K)
    Time = np.asarray(D["Time"])[sel]
    last_time = np.maximum(last_frac * Time, 1e-3)
    row_ptr, tt = truth["row_ptr"], truth["times"]

    # longitudinal rows with time <= last.time (the first visit is at 0)
    keep_rows, idL = [], []
    for j, i in enumerate(sel):
        r = np.arange(row_ptr[i], row_ptr[i + 1])
        r = r[tt[r] <= last_time[j]]
        keep_rows.append(r)
        idL.append(np.full(r.size, j, dtype=np.int32))
    rows = np.concatenate(keep_rows)
    idL = np.concatenate(idL)
    y_long = [np.asarray(D["y"][k])[rows] for k in range(O)]
    X_long = [_F(np.asarray(D["X"][k])[rows]) for k in range(O)]
    Z_long = [_F(np.asarray(D["Z"][k])[rows]) for k in range(O)]

    # GK_points_postRE = outer(last.time / 2, sk + 1); Pw = P * wk (:97-101,206)
    P = last_time / 2.0
    st = np.outer(P, sk + 1.0).reshape(-1)
    Pw = np.repeat(P, K) * np.tile(wk, n)
    # prediction intervals (:85-96,99-102)
    upper = last_time[:, None] + horizon_step * np.arange(1, L + 1)[None, :]
    lower = np.concatenate([last_time[:, None], upper[:, :-1]], axis=1)
    P1, P2 = (upper + lower) / 2.0, (upper - lower) / 2.0
    st_h = (P2[:, :, None] * sk[None, None, :] + P1[:, :, None]).reshape(-1)      # row (i L + l) K + k
    Pw_h = (P2[:, :, None] * wk[None, None, :]).reshape(-1)

    knots = D["knots"]
    W2 = np.asarray(D["W2"])[sel]
    outcome = np.asarray(D["outcome"], dtype=np.int32)
    RE_inds2 = D["RE_inds2"]

    def term_designs(t_vec):
        XXs, ZZs, Us = [], [], []
        for t, k in enumerate(outcome):
            if len(RE_inds2[t]) == 2:       # value: fixed = ~ 1 + time, random = ~ 1 + time
                XXs.append(_F(np.column_stack([np.ones_like(t_vec), t_vec])))
                ZZs.append(_F(np.column_stack([np.ones_like(t_vec), t_vec])))
            else:                           # slope: fixed = ~ 1 (indFixed = 2), random = ~ 1 (indRandom = 2)
                XXs.append(_F(np.ones((t_vec.size, 1))))
                ZZs.append(_F(np.ones((t_vec.size, 1))))
            Us.append(_F(np.ones((t_vec.size, 1))))
        return XXs, ZZs, Us

    XXs, ZZs, Us = term_designs(st)
    XXs_h, ZZs_h, Us_h = term_designs(st_h)
    indFixed = [np.array([0, 1], dtype=np.int32) if len(RE_inds2[t]) == 2 else np.array([1], dtype=np.int32)
                for t in range(len(outcome))]
    return dict(
        n=n, K=K, L=L, q=int(np.asarray(cohort["Covs"]["b"]).shape[0]),
        fams=D["fams"], links=D["links"], RE_inds=D["RE_inds"], RE_inds2=RE_inds2, col_inds=D["col_inds"],
        trans_Funs=D["trans_Funs"], outcome=outcome, indFixed=indFixed,
        y_long=y_long, X_long=X_long, Z_long=Z_long, idL=[idL.copy() for _ in range(O)],
        W1s=spline_design(knots, st) if dense_basis else None, W2s=_F(np.repeat(W2, K, axis=0)), Pw=Pw, XXs=XXs, ZZs=ZZs, Us=Us,
        n_horizons=np.full(n, L, dtype=np.int32),
        W1s_h=spline_design(knots, st_h) if dense_basis else None, W2s_h=_F(np.repeat(W2, K * L, axis=0)), Pw_h=Pw_h,
        knots=np.asarray(knots, dtype=np.float64), ord=4, st=st, st_h=st_h,
        XXs_h=XXs_h, ZZs_h=ZZs_h, Us_h=Us_h,
        last_time=last_time, times_to_pred=upper, subjects=sel)


def posterior_draws(cohort: dict, M: int = 200, seed: int = 7, jitter: float = 0.03):
    """Stand-ins for ``object$statistics$postMeans`` and the ``M`` sub-sampled rows of ``object$mcmc``
    (R/survfitJM.mvJMbayes.R:117-128,306-329): truth-centred means and Gaussian / Wishart-type jitter."""
    rng = np.random.default_rng(seed)
    truth, ini, D = cohort["truth"], cohort["inits"], cohort["Data"]
    O = len(D["y"])
    q = truth["D"].shape[0]
    means = dict(
        betas=[np.asarray(b, dtype=np.float64)[None, :] for b in truth["betas"]],
        sigmas=[np.array([truth["sigma"]]) if f == "gaussian" else None for f in D["fams"]],
        inv_D=np.linalg.inv(truth["D"])[None, :, :],
        Bs_gammas=np.asarray(ini["Bs_gammas"])[None, :], gammas=np.asarray(truth["gammas"])[None, :],
        alphas=np.asarray(truth["alphas"])[None, :])
    invD = np.empty((M, q, q))
    for m in range(M):
        G = rng.normal(size=(40 * q, q)) @ np.linalg.cholesky(truth["D"]).T
        Dm = G.T @ G / (40 * q)
        invD[m] = np.linalg.inv(0.5 * (Dm + Dm.T))
    draws = dict(
        betas=[means["betas"][k] + jitter * rng.normal(size=(M, means["betas"][k].shape[1])) for k in range(O)],
        sigmas=[truth["sigma"] * np.exp(jitter * rng.normal(size=M)) if f == "gaussian" else None for f in D["fams"]],
        inv_D=invD,
        Bs_gammas=means["Bs_gammas"] + 2 * jitter * rng.normal(size=(M, means["Bs_gammas"].shape[1])),
        gammas=means["gammas"] + jitter * rng.normal(size=(M, means["gammas"].shape[1])),
        alphas=means["alphas"] + jitter * rng.normal(size=(M, means["alphas"].shape[1])))
    return means, draws
