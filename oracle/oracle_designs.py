"""NumPy / SciPy restatement of right_rows, right_rows_mstate (R/supportFuns_mvJointModelBayes.R:16-40) and of
splines::splineDesign(knots, x, ord, outer.ok = TRUE).  TEST INFRASTRUCTURE ONLY (see oracle/jmb_oracle.h).

PARITY: right_rows follows the cited lines literally (split by subject, findInterval per subject with R's default
arguments = number of sorted values <= x, ind[ind < 1] <- 1, row names of the subject's block); it is integer work and is
additionally checked against a brute-force scan.  splineDesign is R's `splines` package (not under /root/reference): restated
through scipy.interpolate.BSpline.design_matrix (same Cox - de Boor basis), with outer.ok's zero rows and splineDesign's
right-closed last interval."""
from __future__ import annotations

import numpy as np


def find_interval(x, vec):
    """R's findInterval(x, vec) with default arguments: the number of (sorted) vec elements <= x."""
    return np.searchsorted(np.asarray(vec), np.asarray(x), side="right")


def right_rows(times, ids, Q_points, idT=None):
    times, ids = np.asarray(times, dtype=np.float64), np.asarray(ids)
    Q = np.atleast_2d(np.asarray(Q_points, dtype=np.float64))
    uniq = ids[np.sort(np.unique(ids, return_index=True)[1])]                 # factor(ids, levels = unique(ids))
    blocks = [np.flatnonzero(ids == u) for u in uniq]                          # split(row.names(data), fids)
    rows_of_Q = range(Q.shape[0]) if idT is None else list(np.asarray(idT))
    out = []
    for r, i in enumerate(rows_of_Q):
        ind = find_interval(Q[r], times[blocks[i]])                            # mapply(findInterval, Q_points, split(times, fids))
        ind = np.where(ind < 1, 1, ind)                                        # ind[ind < 1] <- 1
        out.append(blocks[i][ind - 1])                                         # mapply(`[`, rownams_id, split(ind, col(ind)))
    return np.concatenate(out).astype(np.int64)                                # data[c(ind), ]


def spline_design(knots, x, ord=4):
    from scipy.interpolate import BSpline
    knots, x = np.asarray(knots, dtype=np.float64), np.ravel(np.asarray(x, dtype=np.float64))
    nb = knots.size - ord
    D = np.zeros((x.size, nb))
    inside = (x >= knots[ord - 1]) & (x <= knots[nb])
    if inside.any():
        D[inside] = BSpline.design_matrix(x[inside], knots, ord - 1).toarray()
    return D


def data_rows(X, rows, time_col=None, new_time=None):
    """data[c(ind), ] (R/supportFuns_mvJointModelBayes.R:24) followed by dataL.id2[[timeVar]] <- c(t(st)) (R/mvJointModelBayes.R:257,264,268)."""
    out = np.array(np.asarray(X, dtype=np.float64)[np.asarray(rows, dtype=np.int64)], order="F")
    if time_col is not None:
        out[:, time_col] = np.asarray(new_time, dtype=np.float64)
    return out
