"""NumPy restatement of the pair loop of aucJM.mvJMbayes and of the sum of prederrJM.mvJMbayes.  TEST INFRASTRUCTURE ONLY
(see oracle/jmb_oracle.h): only tests/ may import it.

PARITY: unpinned - R is absent from the image and the reference ships no fixtures for these functions; the restatement
follows R/aucJM.mvJMbayes.R:66-141 and R/prederrJM.mvJMbayes.R:104-108 line by line (combn pairs, the four indicator
vectors, the in-place weighting, sum(, na.rm = TRUE)) and is cross-checked against an independent O(n) / sort-based
evaluation of the same sums (tests/test_accuracy_cpu.py).
"""
from __future__ import annotations

import numpy as np


def aucJM_pairs(Time, event, Thoriz, pi_u_t, pi2, pi3):
    """R/aucJM.mvJMbayes.R:66-141 on per-subject vectors (subjects in newdata2 order).  pi2[i] = 1 - S_i(Thoriz | T_i) (:98,
    :138), pi3[j] = S_j(Thoriz | T_j) (:114,:139); returns (auc, numerator, denominator)."""
    Time, event, pi, pi2, pi3 = (np.asarray(x, dtype=np.float64) for x in (Time, event, pi_u_t, pi2, pi3))
    n = Time.size
    if n < 2:
        return float("nan"), 0.0, 0.0                              # :142-144
    i, j = np.triu_indices(n, k=1)                                 # combn(ids, 2): first member before second (:67)
    Ti, Tj, di, dj = Time[i], Time[j], event[i], event[j]
    ev_i, ce_i, ce_j = (di == 1) | (di == 3), (di == 0) | (di == 2), (dj == 0) | (dj == 2)
    ind1 = (Ti <= Thoriz) & ev_i & (Tj > Thoriz)                   # :74
    ind2 = (Ti <= Thoriz) & ce_i & (Tj > Thoriz)                   # :75
    ind3 = (Ti <= Thoriz) & ev_i & (Tj <= Thoriz) & ce_j           # :76
    ind4 = (Ti <= Thoriz) & ce_i & (Tj <= Thoriz) & ce_j           # :77
    ind = (ind1 | ind2 | ind3 | ind4).astype(np.float64)           # :79
    ind[ind2] = ind[ind2] * pi2[i[ind2]]                           # :98-99
    ind[ind3] = ind[ind3] * pi3[j[ind3]]                           # :114-115
    ind[ind4] = ind[ind4] * pi2[i[ind4]] * pi3[j[ind4]]            # :138-140
    with np.errstate(invalid="ignore"):
        less = np.where(np.isnan(pi[i]) | np.isnan(pi[j]), np.nan, (pi[i] < pi[j]).astype(np.float64))
        num = np.nansum(less * ind)                                # :141, sum(, na.rm = TRUE)
    den = np.nansum(ind)
    return (num / den if den != 0 else float("nan")), float(num), float(den)


def prederrJM_sum(S_alive, S_dead, S_cens, w_cens, nr, lossFun="square"):
    """R/prederrJM.mvJMbayes.R:12-16,104-108."""
    L = np.abs if lossFun == "absolute" else np.square
    a, d, c, w = (np.asarray(x, dtype=np.float64).ravel() for x in (S_alive, S_dead, S_cens, w_cens))
    terms = np.concatenate([L(1 - a), L(0 - d), w * L(1 - c) + (1 - w) * L(0 - c)])
    return float(np.nansum(terms) / nr)


def _r_sum(x, axis=None):
    """R's sum() / colSums() / mean(): one sequential pass in subject order with an extended-precision (long double)
    accumulator, rounded to double at the end (R src/main/summary.c rsum(), src/main/array.c do_colsum(); R's own code,
    not under /root/reference).  NumPy's pairwise np.sum() differs from it in the last bits, which matters where the
    reference subtracts two sums of the same terms (nFN, nTN are exactly 0 in R where every probability is below the
    threshold)."""
    x = np.asarray(x, dtype=np.float64)
    if axis is None:
        return np.float64(np.cumsum(x.ravel(), dtype=np.longdouble)[-1]) if x.size else np.float64(0.0)
    return np.cumsum(x, axis=axis, dtype=np.longdouble).take(-1, axis=axis).astype(np.float64)


def rocJM(Time, event, Thoriz, pi_u_t, pi2):
    """R/rocJM.mvJMbayes.R:68-105 on per-subject vectors, with the n x 101 outer() matrix as there."""
    Time, event, pi, pi2 = (np.asarray(x, dtype=np.float64) for x in (Time, event, pi_u_t, pi2))
    ind1 = (Time < Thoriz) & ((event == 1) | (event == 3))            # :69
    ind2 = (Time < Thoriz) & ((event == 0) | (event == 2))            # :71
    ind = (ind1 | ind2).astype(np.float64)                            # :72
    ind[ind2] = ind[ind2] * pi2[ind2]                                 # :84-85
    thrs = np.linspace(0.0, 1.0, 101)                                 # :88
    with np.errstate(all="ignore"):
        lt = np.where(np.isnan(pi)[:, None], np.nan, (pi[:, None] < thrs[None, :]).astype(np.float64))   # outer(pi.u.t, thrs, "<")
        nTP = _r_sum(lt * ind[:, None], axis=0)                       # :89
        nFN = _r_sum(ind) - nTP
        TP = nTP / _r_sum(ind)
        nFP = _r_sum(lt * (1 - ind)[:, None], axis=0)                 # :92
        nTN = _r_sum(1 - ind) - nFP
        FP = nFP / _r_sum(1 - ind)
        Q = _r_sum(lt, axis=0) / lt.shape[0]                          # :95
        Qd = 1 - Q
        k10 = (TP - Q) / Qd
        k00 = (1 - FP - Qd) / Q
        P = _r_sum(ind) / ind.size
        k05 = (P * Qd * k10 + (1 - P) * Q * k00) / (P * Qd + (1 - P) * Q)
        f1 = 2 * nTP / (2 * nTP + nFN + nFP)
        youden = TP - FP
    return dict(TP=TP, FP=FP, nTP=nTP, nFN=nFN, nFP=nFP, nTN=nTN, qSN=k10, qSP=k00, qOverall=k05, thrs=thrs,
                F1score=_median_at_max(thrs, f1), Youden=_median_at_max(thrs, youden))


def _median_at_max(thrs, x):
    """median(thrs[x == max(x)]) (R/rocJM.mvJMbayes.R:103,105); max() without na.rm is NA as soon as one element is."""
    m = np.max(x) if not np.any(np.isnan(x)) else float("nan")
    return float("nan") if np.isnan(m) else float(np.median(thrs[x == m]))
