"""ctypes front-end of oracle/_ref/libjmb_ref.so: the reference's OWN C++ sources
(/root/reference/src/fast_LAPRWM.cpp, fast_logPost.cpp, fast_score.cpp), compiled where they lie
against the stand-in header oracle/ref_shim/RcppArmadillo.h.

TEST INFRASTRUCTURE ONLY.  Available only where the reference tree is present (this container);
the GPU box uses the prebuilt .so when it travelled with the snapshot.  Used to check the restated
oracle (jmb_oracle.c) against the reference's actual code, and to generate tests/golden/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from jmbayes_b200 import _abi

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libjmb_ref.so")
REF = os.environ.get("JMB_REFERENCE", "/root/reference")
_lib = None


def available() -> bool:
    return os.path.exists(LIB) or os.path.exists(os.path.join(REF, "src", "fast_LAPRWM.cpp"))


def build(force: bool = False) -> str | None:
    have_src = os.path.exists(os.path.join(REF, "src", "fast_LAPRWM.cpp"))
    if not have_src:
        return LIB if os.path.exists(LIB) else None
    deps = [os.path.join(HERE, f) for f in ("ref_driver.cpp", "ref_shim/RcppArmadillo.h", "jmb_oracle.c", "jmb_oracle.h")]
    stale = force or not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in deps)
    if stale:
        subprocess.run(["make", "-C", HERE, "-s", "ref", f"REF={REF}"] + (["-B"] if force else []), check=True)
    return LIB


SHIM = os.path.join(HERE, "_ref", "libjmb_shim.so")
_shim = None
_use_shim = False


def build_shim(force: bool = False) -> str:
    """oracle/_ref/libjmb_shim.so: the SAME driver (ref_driver.cpp: R lists in, flat results out) linked with the Rcpp glue of
    integration/fast_LAPRWM_cuda.cpp instead of the reference's sources - the exported functions with their bodies replaced
    by calls into libjmb200.so.  Needs the CUDA library (and a GPU to run)."""
    from jmbayes_b200 import build as jb
    jb.build_cuda()
    subprocess.run(["make", "-C", HERE, "-s", "shim"] + (["-B"] if force else []), check=True)
    return SHIM


class use_shim:
    """Context manager: the calls of this module go to the Rcpp glue + CUDA library instead of the reference's own bodies."""

    def __enter__(self):
        global _use_shim
        build_shim()
        _use_shim = True
        return self

    def __exit__(self, *exc):
        global _use_shim
        _use_shim = False


def _bind(path):
    L = C.CDLL(path)
    L.jmbref_last_error.restype = C.c_char_p
    L.jmbref_lap_rwm.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw), C.POINTER(_abi.JmbPriors),
                                 C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn), C.POINTER(_abi.JmbChainOut)]
    L.jmbref_logPosterior.restype = C.c_double
    L.jmbref_logPosterior.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 17 + [C.c_double] * 3
    L.jmbref_gradient_logPosterior.restype = None
    L.jmbref_gradient_logPosterior.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 16 + [C.c_double] * 3 + [_abi.c_double_p]
    return L


def lib():
    global _lib, _shim
    if _use_shim:
        if _shim is None:
            _shim = _bind(build_shim())
        return _shim
    if _lib is None:
        if build() is None:
            raise RuntimeError("oracle/_ref is not available (no reference tree and no prebuilt library)")
        L = C.CDLL(LIB)
        L.jmbref_last_error.restype = C.c_char_p
        L.jmbref_lap_rwm.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw), C.POINTER(_abi.JmbPriors),
                                     C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn), C.POINTER(_abi.JmbChainOut)]
        L.jmbref_logPosterior.restype = C.c_double
        L.jmbref_logPosterior.argtypes = [C.c_double, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 17 + [C.c_double] * 3
        L.jmbref_gradient_logPosterior.restype = None
        L.jmbref_gradient_logPosterior.argtypes = [C.c_double, C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [_abi.c_double_p] * 16 + [C.c_double] * 3 + [_abi.c_double_p]
        _lib = L
    return _lib


def lap_rwm_C(initials, Data, priors, scales, Covs, control, interval_cens=False, multiState=False,
              chain_id: int = 0, subject_offset: int = 0):
    """The reference's lap_rwm_C (src/fast_LAPRWM.cpp:431-896), fed the DESIGN.md random streams."""
    d, k1 = _abi.marshal_data(Data, Covs["b"], interval_cens, multiState, subject_offset)
    w, k2 = _abi.marshal_draw(Data, interval_cens)
    pr, k3 = _abi.marshal_priors(priors)
    ct = _abi.marshal_control(control, chain_id)
    ci, k4 = _abi.marshal_chain_in(initials, scales, Covs)
    total, n_out = _abi.chain_dims(control)
    out, bufs = _abi.alloc_chain_out(d.n, d.q, d.n_Bs, d.n_gammas, d.n_alphas, pr.n_td, n_out, total, extras=False,
                                      n_risks=d.nRisks if multiState else 1)
    rc = lib().jmbref_lap_rwm(C.byref(d), C.byref(w), C.byref(pr), C.byref(ct), C.byref(ci), C.byref(out))
    if rc != 0:
        raise RuntimeError("reference lap_rwm_C: " + lib().jmbref_last_error().decode())
    mcmc = {k: bufs[k] for k in ("b", "Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas",
                                 "tau_Bs_gammas", "tau_gammas", "tau_alphas", "tau_td_alphas")}
    return dict(mcmc=mcmc, logWeights=bufs["logWeights"],
                scales=dict(sigma=dict(b=bufs["sigma_b"], Bs_gammas=float(bufs["sigma_surv"][0]),
                                       gammas=float(bufs["sigma_surv"][1]), alphas=float(bufs["sigma_surv"][2])),
                            Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"],
                                       alphas=bufs["Sigma_alphas"])))


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
    L.jmbref_lap_rwm_woRE.argtypes = [C.POINTER(_abi.JmbData), _abi.c_double_p, _abi.c_double_p, C.POINTER(_abi.JmbPriors),
                     C.POINTER(_abi.JmbControl), C.POINTER(_abi.JmbChainIn), C.POINTER(_abi.JmbChainOut), _abi.c_double_p]
    rc = L.jmbref_lap_rwm_woRE(C.byref(d), Wl.ctypes.data_as(_abi.c_double_p), Wls.ctypes.data_as(_abi.c_double_p), C.byref(pr),
              C.byref(ct), C.byref(ci), C.byref(out), logPost.ctypes.data_as(_abi.c_double_p))
    if rc != 0:
        raise RuntimeError(f"jmbref_lap_rwm_woRE failed with code {rc}")
    mcmc = {k: bufs[k] for k in ("Bs_gammas", "gammas", "alphas", "phi_gammas", "phi_alphas", "tau_Bs_gammas",
                                 "tau_gammas", "tau_alphas")}
    return dict(mcmc=mcmc, logPost=logPost[:n_out],
                scales=dict(sigma=dict(Bs_gammas=float(bufs["sigma_surv"][0]), gammas=float(bufs["sigma_surv"][1]),
                                       alphas=float(bufs["sigma_surv"][2])),
                            Sigma=dict(Bs_gammas=bufs["Sigma_Bs_gammas"], gammas=bufs["Sigma_gammas"], alphas=bufs["Sigma_alphas"])),
                extras=dict(trace_surv=bufs.get("trace_surv")))


def svft(Data, Covs_b, b, Bs_gammas, gammas, alphas):
    """The reference's log_post_RE_svft and survPred_svft_2 (src/survfitJM_mvJMbayes.cpp:190-271) evaluated
    one subject at a time on the per-subject slices of ``Data``; returns (log_post[n], surv_pred[n])."""
    d, k1 = _abi.marshal_data(Data, Covs_b, False)
    w, k2 = _abi.marshal_draw(Data, False)
    n = d.n
    b = np.asfortranarray(np.asarray(b, dtype=np.float64))
    g = np.ascontiguousarray(np.atleast_1d(np.asarray(gammas, dtype=np.float64)))
    if g.size == 0:
        g = np.zeros(1)
    Bs = np.ascontiguousarray(np.asarray(Bs_gammas, dtype=np.float64))
    al = np.ascontiguousarray(np.asarray(alphas, dtype=np.float64))
    lp, sp = np.zeros(n), np.zeros(n)
    P = _abi.c_double_p
    L = lib()
    L.jmbref_svft.argtypes = [C.POINTER(_abi.JmbData), C.POINTER(_abi.JmbDraw)] + [P] * 6
    rc = L.jmbref_svft(C.byref(d), C.byref(w), b.ctypes.data_as(P), Bs.ctypes.data_as(P), g.ctypes.data_as(P),
                       al.ctypes.data_as(P), lp.ctypes.data_as(P), sp.ctypes.data_as(P))
    if rc != 0:
        raise RuntimeError("reference svft: " + L.jmbref_last_error().decode())
    return lp, sp
