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
        L.jmb_data_rows.argtypes = [C.c_int, dp, C.c_int64, C.c_int32, i64p, C.c_int64, C.c_int32, dp, dp]
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


def data_rows(X, rows, time_col=None, new_time=None, device: int = 0) -> np.ndarray:
    """``data[c(ind), ]`` with ``data[[timeVar]] <- c(t(st))`` (R/supportFuns_mvJointModelBayes.R:24, R/mvJointModelBayes.R:256-268) on
    the numeric columns ``X`` [N, p] of the long data: the rows ``right_rows`` selected, the time column replaced by the quadrature
    times.  Returns [M, p] (Fortran order, as R holds it)."""
    X = np.asfortranarray(np.asarray(X, dtype=np.float64))
    if X.ndim != 2:
        raise ValueError("data_rows: X must be a matrix")
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    tc = -1 if time_col is None else int(time_col)
    nt = None if time_col is None else np.ascontiguousarray(new_time, dtype=np.float64)
    if nt is not None and nt.size != rows.size:
        raise ValueError("data_rows: one new time per selected row")
    out = np.zeros((rows.size, X.shape[1]), order="F")
    _check(_lib().jmb_data_rows(int(device), _p(X, C.c_double), X.shape[0], X.shape[1], _p(rows, C.c_int64), rows.size, tc,
                                _p(nt, C.c_double) if nt is not None else None, _p(out, C.c_double)))
    return out
