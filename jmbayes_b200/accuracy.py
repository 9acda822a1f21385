"""Predictive-accuracy sums on the device: host side of ``include/jmb200_accuracy.h``.

``aucJM.mvJMbayes`` (``R/aucJM.mvJMbayes.R:66-141``) and ``prederrJM.mvJMbayes`` (``R/prederrJM.mvJMbayes.R:104-108``) obtain
conditional survival probabilities from ``survfitJM`` (``jmbayes_b200.survfit``) and reduce them over all pairs of subjects /
over three groups of subjects; those reductions run here.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import c_double_p


def _lib():
    from . import api
    L = api.lib()
    if not getattr(L, "_acc_bound", False):
        L.jmb_aucJM_pairs.argtypes = [C.c_int, C.c_int32, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, c_double_p,
                                      c_double_p, C.POINTER(C.c_float)]
        L.jmb_prederrJM_sum.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, c_double_p, C.c_int32, c_double_p, C.c_int32,
                                        c_double_p, c_double_p, c_double_p]
        L.jmb_rocJM_sums.argtypes = [C.c_int, C.c_int32, c_double_p, c_double_p, C.c_double, c_double_p, c_double_p, C.c_int32,
                                     c_double_p, c_double_p, c_double_p, c_double_p, c_double_p]
        L._acc_bound = True
    return L


def _check(rc):
    if rc != 0:
        from . import api
        raise api.JmbError(_lib().jmb_last_error().decode())


def _vec(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())


def aucJM_pairs(Time, event, Thoriz, pi_u_t, pi2, pi3, device: int = 0):
    """The pair sums of ``aucJM`` for the subjects of ``newdata2`` in their (Time-sorted) order.

    ``pi_u_t``: ``S_i(Thoriz | Tstart)``; ``pi2``: ``1 - S_i(Thoriz | T_i)`` (read for subjects censored before ``Thoriz``);
    ``pi3``: ``S_j(Thoriz | T_j)`` (likewise).  Returns ``dict(auc, num, den, kernel_ms)``."""
    T, d, p, p2, p3 = (_vec(x) for x in (Time, event, pi_u_t, pi2, pi3))
    n = T.size
    if not (d.size == p.size == p2.size == p3.size == n):
        raise ValueError("aucJM_pairs: vectors of different lengths")
    out = np.zeros(3)
    ms = C.c_float(0)
    P = lambda a: a.ctypes.data_as(c_double_p)  # noqa: E731
    _check(_lib().jmb_aucJM_pairs(int(device), n, P(T), P(d), float(Thoriz), P(p), P(p2), P(p3), P(out), C.byref(ms)))
    return dict(auc=float(out[0]), num=float(out[1]), den=float(out[2]), kernel_ms=ms.value)


def prederrJM_sum(S_alive, S_dead, S_cens, w_cens, nr: int, lossFun: str = "square", device: int = 0) -> float:
    """``(1/nr) * sum(L(1 - S_alive), L(0 - S_dead), w L(1 - S_cens) + (1 - w) L(0 - S_cens), na.rm = TRUE)``."""
    loss = {"absolute": 0, "square": 1}[lossFun]
    a, dd, cc, w = (_vec(x) for x in (S_alive, S_dead, S_cens, w_cens))
    if cc.size != w.size:
        raise ValueError("prederrJM_sum: S_cens and w_cens differ in length")
    out = np.zeros(1)
    P = lambda v: v.ctypes.data_as(c_double_p) if v.size else None  # noqa: E731
    _check(_lib().jmb_prederrJM_sum(int(device), loss, int(nr), a.size, P(a), dd.size, P(dd), cc.size, P(cc), P(w), P(out)))
    return float(out[0])


def rocJM(Time, event, Thoriz, pi_u_t, pi2, thrs=None, device: int = 0) -> dict:
    """The list ``rocJM.mvJMbayes`` returns (``R/rocJM.mvJMbayes.R:88-108``) from the per-subject vectors: the sums over the
    subjects (``nTP``, ``nFP``, ``Q``, ``sum(ind)``) on the device, the arithmetic on the 101 thresholds here."""
    T, d, p, p2 = (_vec(x) for x in (Time, event, pi_u_t, pi2))
    thrs = np.linspace(0.0, 1.0, 101) if thrs is None else _vec(thrs)
    n, m = T.size, thrs.size
    nTP, nFP, Q, s = np.zeros(m), np.zeros(m), np.zeros(m), np.zeros(2)
    P = lambda a: a.ctypes.data_as(c_double_p)  # noqa: E731
    _check(_lib().jmb_rocJM_sums(int(device), n, P(T), P(d), float(Thoriz), P(p), P(p2), m, P(thrs), P(nTP), P(nFP), P(Q), P(s)))
    return _roc_tail(nTP, nFP, Q, float(s[0]), float(s[1]), n, thrs)


def _roc_tail(nTP, nFP, Q, sum_ind, sum_not_ind, n, thrs):
    """R/rocJM.mvJMbayes.R:90-105."""
    with np.errstate(all="ignore"):
        nFN = sum_ind - nTP
        TP = nTP / sum_ind
        nTN = sum_not_ind - nFP                 # sum(1 - ind) - nFP, :93
        FP = nFP / sum_not_ind
        Qd = 1 - Q
        k10 = (TP - Q) / Qd
        k00 = (1 - FP - Qd) / Q
        Pm = sum_ind / n
        k05 = (Pm * Qd * k10 + (1 - Pm) * Q * k00) / (Pm * Qd + (1 - Pm) * Q)
        f1 = 2 * nTP / (2 * nTP + nFN + nFP)
        youden = TP - FP
    return dict(TP=TP, FP=FP, nTP=nTP, nFN=nFN, nFP=nFP, nTN=nTN, qSN=k10, qSP=k00, qOverall=k05, thrs=thrs,
                F1score=_median_at_max(thrs, f1), Youden=_median_at_max(thrs, youden))


def _median_at_max(thrs, x):
    """median(thrs[x == max(x)]) (R/rocJM.mvJMbayes.R:103,105); max() without na.rm is NA as soon as one element is."""
    m = np.max(x) if not np.any(np.isnan(x)) else float("nan")
    return float("nan") if np.isnan(m) else float(np.median(thrs[x == m]))
