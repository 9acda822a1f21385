"""ctypes front-end of the CPU oracle (oracle/jmb_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs, never by the product package.  Pinned against the reference's own sources compiled in
oracle/_ref (see jmb_oracle.h for what that pin does and does not cover).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from jmbayes_b200 import _abi   # struct layout shared by specification (same field order)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libjmb_oracle.so")
_lib = None


def _cpu_stamp() -> str:
    """-march=native objects must not travel between hosts with different CPUs."""
    try:
        txt = open("/proc/cpuinfo").read()
        flags = next((ln for ln in txt.splitlines() if ln.startswith("flags")), "")
        model = next((ln for ln in txt.splitlines() if ln.startswith("model name")), "")
        return hashlib.sha1((model + flags).encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force: bool = False) -> str:
    src = [os.path.join(HERE, f) for f in ("jmb_oracle.c", "jmb_oracle.h", "jmb_oracle_svft.c", "jmb_oracle_svft.h", "jmb_oracle_mll.c", "jmb_oracle_mll.h", "jmb_oracle_stats.c")]
    stamp_file = os.path.join(HERE, "_build", "cpu.stamp")
    stamp = _cpu_stamp()
    old = open(stamp_file).read() if os.path.exists(stamp_file) else ""
    stale = (force or not os.path.exists(LIB) or old != stamp
             or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in src))
    if stale:
        subprocess.run(["make", "-C", HERE, "-s", "-B", "all"], check=True)
        open(stamp_file, "w").write(stamp)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB)
        L.orc_u01.restype = C.c_double
        L.orc_u01.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_rgamma.restype = C.c_double
        L.orc_rgamma.argtypes = [C.c_uint32] * 4 + [C.c_double, C.c_double]
        L.orc_philox4x32.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
        L.orc_rnorm_pair.argtypes = [C.c_uint32] * 6 + [_abi.c_double_p]
        L.orc_rowsum.argtypes = [_abi.c_double_p, C.c_int64, _abi.c_int32_p, C.c_int32, _abi.c_double_p]
        L.orc_eval_logpost_re.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw)] + [_abi.c_double_p] * 4 + [C.POINTER(_abi.JmbEvalOut)]
        L.orc_lap_rwm.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw), C.POINTER(_abi.JmbPriors),
                                  C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn),
                                  C.POINTER(_abi.JmbChainOut), C.c_int]
        L.orc_wlong.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw), _abi.c_double_p, C.c_int, _abi.c_double_p]
        L.orc_logPosterior.restype = C.c_double
        L.orc_set_allreduce.argtypes = [C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(_abi.c_double_p)


def set_rowsum_mode(mode: int):
    """0 = literal cumsum difference (reference), 1 = exact per-segment sums."""
    lib().orc_set_rowsum_mode(int(mode))


def philox(seed, chain, c0, c1, c2, c3):
    out = (C.c_uint32 * 4)()
    lib().orc_philox4x32(seed, chain, c0, c1, c2, c3, out)
    return [int(x) for x in out]


def rowsum(x, last):
    x = np.ascontiguousarray(x, dtype=np.float64)
    last = np.ascontiguousarray(last, dtype=np.int32)
    out = np.zeros(last.size)
    lib().orc_rowsum(_p(x), x.size, last.ctypes.data_as(_abi.c_int32_p), last.size, _p(out))
    return out


def eval_logpost_re(Data, Covs_b, interval_cens, b, Bs_gammas, gammas, alphas, multiState=False):
    d, k1 = _abi.marshal_data(Data, Covs_b, interval_cens, multiState)
    w, k2 = _abi.marshal_draw(Data, interval_cens)
    n = d.n
    nT = d.nT if multiState else n        # log_h / H / log_ptb: one value per transition row
    res = {k: np.zeros(n if k in ("log_pyb", "log_pb") else nT) for k in ("log_pyb", "log_pb", "log_h", "H", "HL", "log_ptb")}
    out = _abi.JmbEvalOut()
    for k, a in res.items():
        setattr(out, k, _p(a))
    b = np.asfortranarray(np.asarray(b, dtype=np.float64))
    g = np.ascontiguousarray(np.atleast_1d(np.asarray(gammas, dtype=np.float64)))
    if g.size == 0:
        g = np.zeros(1)
    Bs = np.ascontiguousarray(np.asarray(Bs_gammas, dtype=np.float64))
    al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
    rc = lib().orc_eval_logpost_re(C.byref(d), C.byref(w), _p(b), _p(Bs), _p(g), _p(al), C.byref(out))
    assert rc == 0
    return res


def wlong(Data, Covs_b, interval_cens, b, which, multiState=False):
    d, k1 = _abi.marshal_data(Data, Covs_b, interval_cens, multiState)
    w, k2 = _abi.marshal_draw(Data, interval_cens)
    nT = d.nT if multiState else d.n
    rows = nT if which == 0 else nT * d.K
    out = np.zeros((rows, d.n_alphas), order="F")
    b = np.asfortranarray(np.asarray(b, dtype=np.float64))
    lib().orc_wlong(C.byref(d), C.byref(w), _p(b), which, _p(out))
    return out


def lap_rwm_C(initials, Data, priors, scales, Covs, control, interval_cens=False, multiState=False,
              chain_id: int = 0, max_iters: int = 0, subject_offset: int = 0, allreduce=None):
    """Restated reference chain (src/fast_LAPRWM.cpp:431-896) with the DESIGN.md random streams.

    ``allreduce(buf: np.ndarray)`` (in place) is the hook for sharded runs."""
    d, k1 = _abi.marshal_data(Data, Covs["b"], interval_cens, multiState, subject_offset)
    w, k2 = _abi.marshal_draw(Data, interval_cens)
    pr, k3 = _abi.marshal_priors(priors)
    ct = _abi.marshal_control(control, chain_id)
    ci, k4 = _abi.marshal_chain_in(initials, scales, Covs)
    total, n_out = _abi.chain_dims(control)
    out, bufs = _abi.alloc_chain_out(d.n, d.q, d.n_Bs, d.n_gammas, d.n_alphas, pr.n_td, n_out, total,
                                      n_risks=d.nRisks if multiState else 1)
    L = lib()
    cb = None
    if allreduce is not None:
        CB = C.CFUNCTYPE(None, _abi.c_double_p, C.c_int, C.c_void_p)

        def _hook(ptr, ln, _ctx):
            arr = np.ctypeslib.as_array(ptr, shape=(ln,))
            allreduce(arr)
        cb = CB(_hook)
        L.orc_set_allreduce(C.cast(cb, C.c_void_p), None)
    try:
        rc = L.orc_lap_rwm(C.byref(d), C.byref(w), C.byref(pr), C.byref(ct), C.byref(ci), C.byref(out), int(max_iters))
    finally:
        if allreduce is not None:
            L.orc_set_allreduce(None, None)
    if rc != 0:
        raise RuntimeError(f"oracle lap_rwm failed with code {rc}")
    mcmc = {k: bufs[k] for k in ("b", "Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas",
                                 "tau_Bs_gammas", "tau_gammas", "tau_alphas", "tau_td_alphas")}
    return dict(mcmc=mcmc, logWeights=bufs["logWeights"],
                scales=dict(sigma=dict(b=bufs["sigma_b"], Bs_gammas=float(bufs["sigma_surv"][0]),
                                       gammas=float(bufs["sigma_surv"][1]), alphas=float(bufs["sigma_surv"][2])),
                            Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"],
                                       alphas=bufs["Sigma_alphas"])),
                extras=dict(accept_b=bufs["accept_b"], accept_surv=bufs["accept_surv"], trace_surv=bufs["trace_surv"]))


def lap_rwm_C_woRE(initials, Data, priors, scales, Covs, control, Covs_b, chain_id: int = 0):
    """lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (src/fast_LAPRWM.cpp:934-1297) on Data$Wlong / Data$Wlongs."""
    d, k1 = _abi.marshal_data(Data, Covs_b, False)
    pr, k3 = _abi.marshal_priors(priors)
    ct = _abi.marshal_control(control, chain_id)
    ini = dict(initials)
    ini.setdefault("b", np.zeros((d.n, d.q)))
    sc = dict(scales); sc.setdefault("b", np.zeros(d.n))
    ci, k4 = _abi.marshal_chain_in(ini, sc, Covs)
    total, n_out = _abi.chain_dims(control)
    out, bufs = _abi.alloc_chain_out(1, 1, d.n_Bs, d.n_gammas, d.n_alphas, 0, n_out, total)
    Wl = np.asfortranarray(np.asarray(Data["Wlong"], dtype=np.float64))
    Wls = np.asfortranarray(np.asarray(Data["Wlongs"], dtype=np.float64))
    logPost = np.zeros(max(n_out, 1))
    L = lib()
    L.orc_lap_rwm_woRE.argtypes = [C.POINTER(_abi.JmbData), _abi.c_double_p, _abi.c_double_p, C.POINTER(_abi.JmbPriors),
                     C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn), C.POINTER(_abi.JmbChainOut), _abi.c_double_p]
    rc = L.orc_lap_rwm_woRE(C.byref(d), Wl.ctypes.data_as(_abi.c_double_p), Wls.ctypes.data_as(_abi.c_double_p), C.byref(pr),
              C.byref(ct), C.byref(ci), C.byref(out), logPost.ctypes.data_as(_abi.c_double_p))
    if rc != 0:
        raise RuntimeError(f"orc_lap_rwm_woRE failed with code {rc}")
    mcmc = {k: bufs[k] for k in ("Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas", "tau_Bs_gammas",
                                 "tau_gammas", "tau_alphas")}
    return dict(mcmc=mcmc, logPost=logPost[:n_out],
                scales=dict(sigma=dict(Bs_gammas=float(bufs["sigma_surv"][0]), gammas=float(bufs["sigma_surv"][1]),
                                       alphas=float(bufs["sigma_surv"][2])),
                            Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"], alphas=bufs["Sigma_alphas"])),
                extras=dict(trace_surv=bufs.get("trace_surv")))


# ---- marglogLik2 (R/margLogLik2.R) ------------------------------------------------------------------------------------
class _MllData(C.Structure):
    _fields_ = [("n", C.c_int32), ("ns", C.c_int64), ("B", C.c_int32), ("p", C.c_int32), ("A", C.c_int32)] + \
               [(f, _abi.c_double_p) for f in ("event", "W1", "W1s", "W2", "W2s", "Wlong", "Wlongs", "Pw")]


class _MllOut(C.Structure):
    _fields_ = [("value", C.c_double), ("opt_value", C.c_double), ("convergence", C.c_int32), ("counts", C.c_int32 * 2)] + \
               [(f, _abi.c_double_p) for f in ("par", "hessian", "Cov_Bs_gammas", "Cov_gammas", "Cov_alphas")]


def marglogLik2(thetas, Data, priors, temp=1.0, fixed_tau_Bs_gammas=False):
    """marglogLik2(thetas, Data, priors, temp, fixed_tau_Bs_gammas) restated (oracle/jmb_oracle_mll.c).  ``thetas``:
    dict(Bs_gammas, gammas, alphas, tau_Bs_gammas); tau is on the log scale when it is not fixed (:39-40)."""
    k = _abi.Keep()
    W1, W2 = np.asarray(Data["W1"]), np.asarray(Data["W2"])
    d = _MllData()
    d.n, d.ns = int(W1.shape[0]), int(np.asarray(Data["Pw"]).shape[0])
    d.B, d.p, d.A = int(W1.shape[1]), int(W2.shape[1]) if W2.ndim == 2 else 0, int(np.asarray(Data["Wlong"]).shape[1])
    for f in ("event", "W1", "W1s", "W2", "W2s", "Wlong", "Wlongs", "Pw"):
        setattr(d, f, k.f64(Data[f]))
    pr, kp = _abi.marshal_priors(priors)
    P = d.B + d.p + d.A
    dim = P if fixed_tau_Bs_gammas else P + 1
    th = np.concatenate([np.asarray(thetas["Bs_gammas"], dtype=np.float64).ravel(), np.asarray(thetas.get("gammas", []), dtype=np.float64).ravel(),
                         np.asarray(thetas["alphas"], dtype=np.float64).ravel()])
    tau = float(np.asarray(thetas["tau_Bs_gammas"]).ravel()[0])
    if not fixed_tau_Bs_gammas:
        th = np.concatenate([th, [tau]])
    bufs = dict(par=np.zeros(dim), hessian=np.zeros((dim, dim), order="F"), Cov_Bs_gammas=np.zeros((d.B, d.B), order="F"),
                Cov_gammas=np.zeros((max(d.p, 1), max(d.p, 1)), order="F"), Cov_alphas=np.zeros((d.A, d.A), order="F"))
    out = _MllOut()
    for name, a in bufs.items():
        setattr(out, name, _p(a))
    L = lib()
    L.orc_marglogLik2.argtypes = [C.POINTER(_MllData), C.POINTER(_abi.JmbPriors), _abi.c_double_p, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(_MllOut)]
    rc = L.orc_marglogLik2(C.byref(d), C.byref(pr), _p(np.ascontiguousarray(th)), tau, float(temp), int(bool(fixed_tau_Bs_gammas)), C.byref(out))
    if rc:
        raise RuntimeError(f"orc_marglogLik2 failed with code {rc}")
    res = dict(value=out.value, opt_value=out.opt_value, convergence=out.convergence, counts=(out.counts[0], out.counts[1]),
               par=bufs["par"], hessian=bufs["hessian"])
    if fixed_tau_Bs_gammas:
        res["Covs"] = dict(Bs_gammas=bufs["Cov_Bs_gammas"], gammas=bufs["Cov_gammas"][:d.p, :d.p], alphas=bufs["Cov_alphas"])
        res["inits"] = dict(Bs_gammas=bufs["par"][:d.B], gammas=bufs["par"][d.B:d.B + d.p], alphas=bufs["par"][d.B + d.p:P],
                            tau_Bs_gammas=tau)
    return res


def nearPD(M):
    M = np.asfortranarray(M, dtype=np.float64)
    out = np.zeros_like(M, order="F")
    L = lib()
    L.orc_nearPD.argtypes = [_abi.c_double_p, C.c_int, _abi.c_double_p]
    L.orc_nearPD(_p(M), M.shape[0], _p(out))
    return out
