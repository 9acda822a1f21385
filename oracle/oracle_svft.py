"""ctypes front-end of the survfitJM restatement (oracle/jmb_oracle_svft.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.
"""
from __future__ import annotations

import ctypes as C
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from jmbayes_b200 import survfit as sv
from jmbayes_b200._abi import c_double_p

from . import oracle

_bound = False


def lib():
    global _bound
    L = oracle.lib()
    if not _bound:
        D, P, Ct, O = C.POINTER(sv.SvftData), C.POINTER(sv.SvftParams), C.POINTER(sv.SvftControl), C.POINTER(sv.SvftOut)
        L.orc_svft_modes.argtypes = [D, P, Ct, O]
        L.orc_svft_sample.argtypes = [D, P, Ct, O]
        L.orc_survfitJM.argtypes = [D, P, P, Ct, O]
        L.orc_survfitJM_range.argtypes = [D, P, P, Ct, O, C.c_int, C.c_int]
        L.orc_svft_log_post.restype = C.c_double
        L.orc_svft_log_post.argtypes = [D, P, C.c_int, C.c_int, c_double_p]
        L.orc_svft_surv_pred.restype = C.c_double
        L.orc_svft_surv_pred.argtypes = [D, P, C.c_int, C.c_int, C.c_int, c_double_p]
        L.orc_sv_inverse.argtypes = [c_double_p, C.c_int, c_double_p]
        L.orc_sv_eigen_sym.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p]
        L.orc_sv_summary.argtypes = [c_double_p, C.c_int, c_double_p, c_double_p]
        _bound = True
    return L


def _result(bufs):
    return sv.SurvFit._result(bufs)


def modes(data: dict, post_means: dict, control: dict | None = None, subject_offset: int = 0):
    d, k1 = sv.marshal_data(data, subject_offset)
    p, k2 = sv.marshal_params(post_means)
    out, bufs = sv.alloc_out(d.n, d.q, d.L, 1, full=False)
    ct = sv.marshal_control(control)
    rc = lib().orc_svft_modes(C.byref(d), C.byref(p), C.byref(ct), C.byref(out))
    if rc:
        raise RuntimeError(f"orc_svft_modes failed with code {rc}")
    return _result(bufs)


def sample(data: dict, draws: dict, modes, hessians, control: dict | None = None, subject_offset: int = 0):
    d, k1 = sv.marshal_data(data, subject_offset)
    p, k2 = sv.marshal_params(draws)
    out, bufs = sv.alloc_out(d.n, d.q, d.L, p.M, full=True, modes=modes, hessians=hessians)
    ct = sv.marshal_control(control)
    rc = lib().orc_svft_sample(C.byref(d), C.byref(p), C.byref(ct), C.byref(out))
    if rc:
        raise RuntimeError(f"orc_svft_sample failed with code {rc}")
    return _result(bufs)


def survfitJM(data: dict, post_means: dict, draws: dict, control: dict | None = None, subject_offset: int = 0,
              threads: int = 1, full: bool = True):
    """R/survfitJM.mvJMbayes.R:277-414 restated; ``threads`` > 1 splits the subjects over worker threads
    (subjects are independent; used by the CPU baseline only)."""
    d, k1 = sv.marshal_data(data, subject_offset)
    pm, k2 = sv.marshal_params(post_means)
    pd, k3 = sv.marshal_params(draws)
    out, bufs = sv.alloc_out(d.n, d.q, d.L, pd.M, full=full)
    ct = sv.marshal_control(control)
    L = lib()
    if threads <= 1:
        rc = L.orc_survfitJM(C.byref(d), C.byref(pm), C.byref(pd), C.byref(ct), C.byref(out))
        if rc:
            raise RuntimeError(f"orc_survfitJM failed with code {rc}")
    else:
        edges = np.linspace(0, d.n, threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            rcs = list(ex.map(lambda t: L.orc_survfitJM_range(C.byref(d), C.byref(pm), C.byref(pd), C.byref(ct), C.byref(out),
                                                              int(edges[t]), int(edges[t + 1])), range(threads)))
        if any(rcs):
            raise RuntimeError(f"orc_survfitJM_range failed: {rcs}")
    return _result(bufs)


def log_post(data: dict, params: dict, m: int, i: int, b):
    d, k1 = sv.marshal_data(data)
    p, k2 = sv.marshal_params(params)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().orc_svft_log_post(C.byref(d), C.byref(p), int(m), int(i), b.ctypes.data_as(c_double_p))


def surv_pred(data: dict, params: dict, m: int, i: int, l: int, b):
    d, k1 = sv.marshal_data(data)
    p, k2 = sv.marshal_params(params)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return lib().orc_svft_surv_pred(C.byref(d), C.byref(p), int(m), int(i), int(l), b.ctypes.data_as(c_double_p))


def inverse(A):
    A = np.asfortranarray(A, dtype=np.float64)
    out = np.zeros_like(A, order="F")
    rc = lib().orc_sv_inverse(A.ctypes.data_as(c_double_p), A.shape[0], out.ctypes.data_as(c_double_p))
    return out, rc


def eigen_sym(S):
    S = np.asfortranarray(S, dtype=np.float64)
    val = np.zeros(S.shape[0])
    vec = np.zeros_like(S, order="F")
    lib().orc_sv_eigen_sym(S.ctypes.data_as(c_double_p), S.shape[0], val.ctypes.data_as(c_double_p), vec.ctypes.data_as(c_double_p))
    return val, vec


def summary(x, probs=(0.025, 0.975)):
    x = np.ascontiguousarray(x, dtype=np.float64)
    pr = np.ascontiguousarray(probs, dtype=np.float64)
    out = np.zeros(4)
    lib().orc_sv_summary(x.ctypes.data_as(c_double_p), x.size, pr.ctypes.data_as(c_double_p), out.ctypes.data_as(c_double_p))
    return out
