"""ctypes mirror of ``include/jmb200.h`` and marshalling of the reference's R lists (as Python
dicts with the reference's field names, see ``cohorts.py``) into the flat C structs.

The oracle's structs (``oracle/jmb_oracle.h``) have the same field order on purpose, so the test
suite fills both with this one marshaller; an independently written NumPy twin
(``oracle/np_twin.py``) consumes the dicts directly and guards against marshalling mistakes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_double_pp = C.POINTER(c_double_p)
c_int32_pp = C.POINTER(c_int32_p)

FAMS = {"gaussian": 0, "binomial": 1, "poisson": 2}
LINKS = {"identity": 0, "logit": 1, "probit": 2, "cloglog": 3, "log": 4}
TRANS = {"identity": 0, "expit": 1, "exp": 2, "log": 3, "log2": 4, "log10": 5, "sqrt": 6}


class JmbData(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("n_outcomes", C.c_int32), ("q", C.c_int32), ("n_terms", C.c_int32),
        ("n_alphas", C.c_int32), ("n_Bs", C.c_int32), ("n_gammas", C.c_int32), ("K", C.c_int32),
        ("interval_cens", C.c_int32), ("multi_state", C.c_int32), ("subject_offset", C.c_int64),
        ("N", c_int64_p), ("fams", c_int32_p), ("links", c_int32_p), ("q_k", c_int32_p),
        ("RE_inds", c_int32_pp), ("y", c_double_pp), ("Z", c_double_pp), ("idL", c_int32_pp),
        ("idL2", c_int32_pp),
        ("event", c_double_p), ("W1", c_double_p), ("W1s", c_double_p), ("W1s_int", c_double_p),
        ("W2", c_double_p), ("W2s", c_double_p), ("W2s_int", c_double_p), ("Pw", c_double_p),
        ("Pw_int", c_double_p), ("idGK_fast", c_int32_p),
        ("r_t", c_int32_p), ("RE_inds2", c_int32_pp), ("u_t", c_int32_p), ("col_inds", c_int32_pp),
        ("trans_Funs", c_int32_p),
        ("ZZ", c_double_pp), ("ZZs", c_double_pp), ("ZZs_int", c_double_pp),
        ("U", c_double_pp), ("Us", c_double_pp), ("Us_int", c_double_pp),
        ("Covs_b", c_double_p),
        ("knots", c_double_p), ("n_knots", C.c_int32), ("ord", C.c_int32), ("Time", c_double_p), ("st", c_double_p),
        ("st_int", c_double_p),
        ("nT", C.c_int32), ("nRisks", C.c_int32), ("idT", c_int32_p), ("kn_strat_first", c_int32_p),
        ("kn_strat_last", c_int32_p),
    ]


class JmbDraw(C.Structure):
    _fields_ = [("Xbetas", c_double_pp), ("XXbetas", c_double_pp), ("XXsbetas", c_double_pp),
                ("XXsbetas_int", c_double_pp), ("sigmas", c_double_p), ("invD", c_double_p)]


class JmbXDesigns(C.Structure):
    _fields_ = [("p_k", c_int32_p), ("X", c_double_pp), ("outcome", c_int32_p), ("f_t", c_int32_p),
                ("indFixed", c_int32_pp), ("XX", c_double_pp), ("XXs", c_double_pp), ("XXs_int", c_double_pp)]


class JmbPriors(C.Structure):
    _fields_ = [
        ("mean_Bs_gammas", c_double_p), ("Tau_Bs_gammas", c_double_p),
        ("mean_gammas", c_double_p), ("Tau_gammas", c_double_p),
        ("mean_alphas", c_double_p), ("Tau_alphas", c_double_p),
        ("A_tau_Bs_gammas", C.c_double), ("B_tau_Bs_gammas", C.c_double), ("rank_Tau_Bs_gammas", C.c_double),
        ("shrink_gammas", C.c_int32),
        ("A_tau_gammas", C.c_double), ("B_tau_gammas", C.c_double), ("rank_Tau_gammas", C.c_double),
        ("A_phi_gammas", C.c_double), ("B_phi_gammas", C.c_double),
        ("shrink_alphas", C.c_int32), ("double_gamma_alphas", C.c_int32),
        ("A_tau_alphas", C.c_double), ("B_tau_alphas", C.c_double), ("rank_Tau_alphas", C.c_double),
        ("A_phi_alphas", C.c_double), ("B_phi_alphas", C.c_double),
        ("A_xi_alphas", C.c_double), ("B_xi_alphas", C.c_double),
        ("A_nu_alphas", C.c_double), ("B_nu_alphas", C.c_double),
        ("n_td", C.c_int32), ("td_len", c_int32_p), ("td_cols", c_int32_pp),
        ("rank_Tau_td_alphas", C.c_double),
    ]


class JmbControl(C.Structure):
    _fields_ = [("n_iter", C.c_int32), ("n_burnin", C.c_int32), ("n_block", C.c_int32),
                ("n_thin", C.c_int32), ("target_acc", C.c_double), ("eps1", C.c_double),
                ("eps2", C.c_double), ("eps3", C.c_double), ("c0", C.c_double), ("c1", C.c_double),
                ("adaptCov", C.c_int32), ("seed", C.c_uint32), ("chain_id", C.c_uint32)]


class JmbChainIn(C.Structure):
    _fields_ = [("b", c_double_p), ("Bs_gammas", c_double_p), ("gammas", c_double_p),
                ("alphas", c_double_p), ("tau_Bs_gammas", c_double_p), ("tau_gammas", C.c_double),
                ("tau_alphas", C.c_double), ("phi_gammas", c_double_p), ("phi_alphas", c_double_p),
                ("tau_td_alphas", c_double_p), ("sigma_b", c_double_p),
                ("sigma_Bs_gammas", C.c_double), ("sigma_gammas", C.c_double),
                ("sigma_alphas", C.c_double), ("Sigma_Bs_gammas", c_double_p),
                ("Sigma_gammas", c_double_p), ("Sigma_alphas", c_double_p)]


_OUT_FIELDS = ["b", "Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas", "tau_Bs_gammas",
               "tau_gammas", "tau_alphas", "tau_td_alphas", "logWeights", "sigma_b", "sigma_surv",
               "Sigma_Bs_gammas", "Sigma_gammas", "Sigma_alphas", "accept_b", "accept_surv",
               "trace_surv"]


class JmbChainOut(C.Structure):
    _fields_ = [(f, c_double_p) for f in _OUT_FIELDS]


class JmbMllOut(C.Structure):
    _fields_ = [("value", C.c_double), ("opt_value", C.c_double), ("convergence", C.c_int32), ("counts", C.c_int32 * 2)] + \
               [(f, c_double_p) for f in ("par", "hessian", "Cov_Bs_gammas", "Cov_gammas", "Cov_alphas")] + [("restarts", C.c_int32)]


class JmbEvalOut(C.Structure):
    _fields_ = [(f, c_double_p) for f in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb")]


# ------------------------------------------------------------------------------------------------
class Keep:
    """Keeps the numpy buffers and pointer arrays alive for as long as the C structs are used."""

    def __init__(self):
        self.refs = []

    def f64(self, a, fortran=True):
        a = np.asarray(a, dtype=np.float64)
        a = np.asfortranarray(a) if (fortran and a.ndim > 1) else np.ascontiguousarray(a)
        if a.size == 0:
            a = np.zeros(1)
        self.refs.append(a)
        return a.ctypes.data_as(c_double_p)

    def i32(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
        if a.size == 0:
            a = np.zeros(1, dtype=np.int32)
        self.refs.append(a)
        return a.ctypes.data_as(c_int32_p)

    def i64(self, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int64))
        self.refs.append(a)
        return a.ctypes.data_as(c_int64_p)

    def f64_list(self, lst):
        arr = (c_double_p * max(len(lst), 1))(*[self.f64(x) for x in lst])
        self.refs.append(arr)
        return C.cast(arr, c_double_pp)

    def i32_list(self, lst):
        arr = (c_int32_p * max(len(lst), 1))(*[self.i32(x) for x in lst])
        self.refs.append(arr)
        return C.cast(arr, c_int32_pp)


def marshal_data(Data: dict, Covs_b, interval_cens: bool, multi_state: bool = False,
                 subject_offset: int = 0, K: int | None = None, basis_from_knots: bool = False):
    """``Data`` + ``Covs$b`` -> (JmbData, Keep).  ``basis_from_knots``: pass knots / Time / st (/ st_int) instead of the
    dense W1 / W1s (/ W1s_int); the library evaluates splineDesign itself."""
    k = Keep()
    d = JmbData()
    n = int(np.asarray(Covs_b).shape[2]) if multi_state else int(np.asarray(Data["event"]).shape[0])
    O = len(Data["y"])
    T = len(Data["XXbetas"])
    ns = int(np.asarray(Data["Pw"]).shape[0])
    d.n, d.n_outcomes, d.n_terms = n, O, T
    d.q = int(np.asarray(Covs_b).shape[0])
    W1 = np.asarray(Data["W1"])
    W2 = np.asarray(Data["W2"])
    d.n_Bs = int(W1.shape[1])
    d.n_gammas = int(W2.shape[1]) if W2.ndim == 2 else 0
    d.n_alphas = int(sum(len(c) for c in Data["col_inds"]))
    nT = int(np.asarray(Data["event"]).shape[0])
    d.K = int(K if K is not None else ns // nT)
    if multi_state:     # Data$idT2 (subject of each transition row), nRisks, kn_strat_* (R/mvJointModelBayes.R:651-670)
        d.nT, d.nRisks = nT, int(Data["nRisks"])
        d.idT = k.i32(np.asarray(Data["idT2"][0]))
        d.kn_strat_first, d.kn_strat_last = k.i32(Data["kn_strat_first"]), k.i32(Data["kn_strat_last"])
    d.interval_cens = int(bool(interval_cens))
    d.multi_state = int(bool(multi_state))
    d.subject_offset = int(subject_offset)
    d.N = k.i64([len(y) for y in Data["y"]])
    d.fams = k.i32([FAMS[f] for f in Data["fams"]])
    d.links = k.i32([LINKS[f] for f in Data["links"]])
    d.q_k = k.i32([len(r) for r in Data["RE_inds"]])
    d.RE_inds = k.i32_list(Data["RE_inds"])
    d.y = k.f64_list(Data["y"])
    d.Z = k.f64_list(Data["Z"])
    d.idL = k.i32_list(Data["idL"])
    d.idL2 = k.i32_list(Data["idL2"])
    d.event = k.f64(Data["event"])
    if basis_from_knots:
        d.knots, d.n_knots, d.ord = k.f64(Data["knots"]), int(len(Data["knots"])), int(Data.get("ordSpline", 4))
        d.Time, d.st = k.f64(Data["Time"]), k.f64(Data["st"])
        if interval_cens:
            d.st_int = k.f64(Data["st_int"])
    else:
        d.W1, d.W1s = k.f64(W1), k.f64(Data["W1s"])
    d.W2, d.W2s = k.f64(W2), k.f64(Data["W2s"])
    d.Pw = k.f64(Data["Pw"])
    d.idGK_fast = k.i32(Data["idGK_fast"])
    d.r_t = k.i32([len(r) for r in Data["RE_inds2"]])
    d.RE_inds2 = k.i32_list(Data["RE_inds2"])
    d.u_t = k.i32([len(c) for c in Data["col_inds"]])
    d.col_inds = k.i32_list(Data["col_inds"])
    d.trans_Funs = k.i32([TRANS[f] for f in Data["trans_Funs"]])
    d.ZZ, d.ZZs = k.f64_list(Data["ZZ"]), k.f64_list(Data["ZZs"])
    d.U, d.Us = k.f64_list(Data["U"]), k.f64_list(Data["Us"])
    if interval_cens:
        if not basis_from_knots:
            d.W1s_int = k.f64(Data["W1s_int"])
        d.W2s_int, d.Pw_int = k.f64(Data["W2s_int"]), k.f64(Data["Pw_int"])
        d.ZZs_int, d.Us_int = k.f64_list(Data["ZZs_int"]), k.f64_list(Data["Us_int"])
    d.Covs_b = k.f64(Covs_b)
    return d, k


def marshal_draw(Data: dict, interval_cens: bool):
    k = Keep()
    w = JmbDraw()
    w.Xbetas = k.f64_list(Data["Xbetas"])
    w.XXbetas = k.f64_list(Data["XXbetas"])
    w.XXsbetas = k.f64_list(Data["XXsbetas"])
    if interval_cens:
        w.XXsbetas_int = k.f64_list(Data["XXsbetas_int"])
    sig = [1.0 if (s is None or (isinstance(s, float) and np.isnan(s))) else float(s) for s in Data["sigmas"]]
    w.sigmas = k.f64(sig)
    w.invD = k.f64(Data["invD"])
    return w, k


def marshal_xdesigns(Data: dict, interval_cens: bool):
    """Data$X / XX / XXs / XXs_int + indFixed + outcome -> JmbXDesigns."""
    k = Keep()
    x = JmbXDesigns()
    x.p_k = k.i32([np.asarray(m).shape[1] for m in Data["X"]])
    x.X = k.f64_list(Data["X"])
    x.outcome = k.i32(Data["outcome"])
    x.f_t = k.i32([len(f) for f in Data["indFixed"]])
    x.indFixed = k.i32_list(Data["indFixed"])
    x.XX, x.XXs = k.f64_list(Data["XX"]), k.f64_list(Data["XXs"])
    if interval_cens:
        x.XXs_int = k.f64_list(Data["XXs_int"])
    return x, k


def marshal_priors(priors: dict):
    k = Keep()
    p = JmbPriors()
    for name in ("mean_Bs_gammas", "Tau_Bs_gammas", "mean_gammas", "Tau_gammas", "mean_alphas", "Tau_alphas"):
        setattr(p, name, k.f64(priors[name]))
    for name in ("A_tau_Bs_gammas", "B_tau_Bs_gammas", "rank_Tau_Bs_gammas", "A_tau_gammas",
                 "B_tau_gammas", "rank_Tau_gammas", "A_phi_gammas", "B_phi_gammas", "A_tau_alphas",
                 "B_tau_alphas", "rank_Tau_alphas", "A_phi_alphas", "B_phi_alphas", "A_xi_alphas",
                 "B_xi_alphas", "A_nu_alphas", "B_nu_alphas", "rank_Tau_td_alphas"):
        setattr(p, name, float(priors[name]))
    for name in ("shrink_gammas", "shrink_alphas", "double_gamma_alphas"):
        setattr(p, name, int(bool(priors[name])))
    td = priors.get("td_cols", []) or []
    p.n_td = len(td)
    p.td_len = k.i32([len(t) for t in td])
    p.td_cols = k.i32_list(td)
    return p, k


def marshal_control(control: dict, chain_id: int = 0):
    c = JmbControl()
    for name in ("n_iter", "n_burnin", "n_block", "n_thin"):
        setattr(c, name, int(control[name]))
    for name in ("target_acc", "eps1", "eps2", "eps3", "c0", "c1"):
        setattr(c, name, float(control[name]))
    c.adaptCov = int(bool(control.get("adaptCov", False)))
    c.seed = int(control.get("seed", 1)) & 0xFFFFFFFF
    c.chain_id = int(chain_id) & 0xFFFFFFFF
    return c


def marshal_chain_in(inits: dict, scales: dict, Covs: dict):
    k = Keep()
    s = JmbChainIn()
    s.b = k.f64(inits["b"])
    for name in ("Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas", "tau_td_alphas"):
        setattr(s, name, k.f64(np.atleast_1d(inits.get(name, np.zeros(0)))))
    s.tau_Bs_gammas = k.f64(np.atleast_1d(inits["tau_Bs_gammas"]))
    s.tau_gammas, s.tau_alphas = float(inits["tau_gammas"]), float(inits["tau_alphas"])
    s.sigma_b = k.f64(scales["b"])
    s.sigma_Bs_gammas = float(scales["Bs_gammas"])
    s.sigma_gammas = float(scales["gammas"])
    s.sigma_alphas = float(scales["alphas"])
    s.Sigma_Bs_gammas = k.f64(Covs["Bs_gammas"])
    s.Sigma_gammas = k.f64(Covs["gammas"])
    s.Sigma_alphas = k.f64(Covs["alphas"])
    return s, k


def alloc_chain_out(n, q, B, p, A, n_td, n_out, total_iters, extras=True, n_risks=1):
    """Caller-allocated outputs shaped like the list lap_rwm_C returns."""
    no = max(n_out, 0)
    bufs = dict(
        b=np.zeros((n, q, no), order="F"), Bs_gammas=np.zeros((B, no), order="F"),
        gammas=np.zeros((p, no), order="F"), alphas=np.zeros((A, no), order="F"),
        phi_gammas=np.zeros((p, no), order="F"), phi_alphas=np.zeros((A, no), order="F"),
        tau_Bs_gammas=np.zeros((n_risks, no), order="F"), tau_gammas=np.zeros(no), tau_alphas=np.zeros(no),
        tau_td_alphas=np.zeros((n_td, no), order="F"), logWeights=np.zeros(no),
        sigma_b=np.zeros(n), sigma_surv=np.zeros(3), Sigma_Bs_gammas=np.zeros((B, B), order="F"),
        Sigma_gammas=np.zeros((p, p), order="F"), Sigma_alphas=np.zeros((A, A), order="F"))
    if extras:
        bufs.update(accept_b=np.zeros(n), accept_surv=np.zeros(1),
                    trace_surv=np.zeros(2 * max(total_iters, 1)))
    out = JmbChainOut()
    for name, a in bufs.items():
        if a.size:
            setattr(out, name, a.ctypes.data_as(c_double_p))
    return out, bufs


def chain_dims(control: dict):
    """its * n_block, n_out: src/fast_LAPRWM.cpp:584-586."""
    total_it = int(control["n_iter"]) + int(control["n_burnin"])
    its = -(-total_it // int(control["n_block"]))
    return its * int(control["n_block"]), int(control["n_iter"]) // int(control["n_thin"])
