"""Host-side mirror of include/jmb200_designs.h: the design construction that precedes the MCMC path -
``right_rows`` / ``right_rows_mstate`` (R/supportFuns_mvJointModelBayes.R:16-40) and ``splines::splineDesign`` in band
form (R/mvJointModelBayes.R:382-383,389-392) - on the device.  No CPU fallback: without the library or a device the calls raise."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import api

_bound = False


def _lib():
    global _bound
    L = api.lib()
    if not _bound:
        i64p, dp, i32p = C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int32)
        L.jmb_right_rows.argtypes = [C.c_int, C.c_int32, C.c_int32, i64p, dp, dp, i64p]
        L.jmb_right_rows_mstate.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_int32, i64p, dp, i32p, dp, i64p]
        L.jmb_spline_design.argtypes = [C.c_int, dp, C.c_int32, C.c_int32, dp, C.c_int64, i32p, dp, dp]
        _bound = True
    return L


def _check(rc):
    if rc:
        raise api.JmbError(api.lib().jmb_last_error().decode())


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def right_rows(times, ids, Q_points, idT=None, device: int = 0) -> np.ndarray:
    """0-based rows of the long data that ``right_rows(data, times, ids, Q_points)`` selects (``data[c(ind), ]``): for every
    subject (in order of first appearance in ``ids``) and every column of ``Q_points`` the last visit at or before the point,
    the first visit when there is none.  ``idT`` (0-based subject of each row of ``Q_points``) gives ``right_rows_mstate``."""
    times = np.ascontiguousarray(times, dtype=np.float64)
    ids = np.asarray(ids)
    Q = np.ascontiguousarray(np.atleast_2d(np.asarray(Q_points, dtype=np.float64)))
    change = np.flatnonzero(np.r_[True, ids[1:] != ids[:-1]])
    if np.unique(ids).size != change.size:
        raise ValueError("right_rows: the long data must be grouped by subject")
    row_ptr = np.ascontiguousarray(np.r_[change, ids.size], dtype=np.int64)
    n, K = change.size, Q.shape[1]
    out = np.zeros(Q.shape[0] * K, dtype=np.int64)
    if idT is None:
        if Q.shape[0] != n:
            raise ValueError("right_rows: one row of Q_points per subject")
        _check(_lib().jmb_right_rows(int(device), n, K, _p(row_ptr, C.c_int64), _p(times, C.c_double), _p(Q, C.c_double), _p(out, C.c_int64)))
    else:
        idT = np.ascontiguousarray(idT, dtype=np.int32)
        _check(_lib().jmb_right_rows_mstate(int(device), n, Q.shape[0], K, _p(row_ptr, C.c_int64), _p(times, C.c_double),
                                            _p(idT, C.c_int32), _p(Q, C.c_double), _p(out, C.c_int64)))
    return out


def spline_design(knots, x, ord: int = 4, dense: bool = False, device: int = 0):
    """``splineDesign(knots, x, ord, outer.ok = TRUE)``: (offsets [n], band values [n, ord]) and, with ``dense``, the
    n x (len(knots) - ord) matrix R returns."""
    knots = np.ascontiguousarray(knots, dtype=np.float64)
    x = np.ascontiguousarray(np.ravel(x), dtype=np.float64)
    off = np.zeros(x.size, dtype=np.int32)
    vals = np.zeros((x.size, ord))
    D = np.zeros((x.size, knots.size - ord), order="F") if dense else None
    _check(_lib().jmb_spline_design(int(device), _p(knots, C.c_double), knots.size, int(ord), _p(x, C.c_double), x.size,
                                    _p(off, C.c_int32), _p(vals, C.c_double), _p(D, C.c_double) if dense else None))
    return (off, vals, D) if dense else (off, vals)
