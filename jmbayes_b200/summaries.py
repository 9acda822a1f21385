"""Posterior summaries of the kept draws: host side of ``include/jmb200_summaries.h``.

Mirrors the tail of ``mvJointModelBayes()`` (``R/mvJointModelBayes.R:925-971``): ``weights <- stand(LogLiks)`` and the ten
``summary_fun(FUN)`` sweeps over the ``mcmc`` list, with the reference's names (``postMeans``, ``postwMeans``, ``postModes``,
``EffectiveSize``, ``StDev``, ``wStDev``, ``StErr``, ``CIs``, ``wCIs``, ``Pvalues``) and its ``apply`` margins.  Every number
is computed by ``jmb_mcmc_statistics`` on the GPU; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._abi import c_double_p

NAMES = ("postMeans", "postwMeans", "postModes", "EffectiveSize", "StDev", "wStDev", "StErr", "CIs", "wCIs", "Pvalues")
_WIDTH = {"CIs": 2, "wCIs": 2}


class StatsOut(C.Structure):
    _fields_ = [(name, c_double_p) for name in NAMES]


def _lib():
    from . import api
    L = api.lib()
    if not getattr(L, "_stats_bound", False):
        L.jmb_stand_weights.argtypes = [c_double_p, C.c_int32, c_double_p]
        L.jmb_mcmc_statistics.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int64, c_double_p,
                                          c_double_p, C.c_int, C.POINTER(StatsOut), C.POINTER(C.c_float)]
        L._stats_bound = True
    return L


def _check(rc):
    if rc != 0:
        from . import api
        raise api.JmbError(_lib().jmb_last_error().decode())


def release():
    """Free the device buffers ``jmb_mcmc_statistics`` keeps between calls."""
    _check(_lib().jmb_stats_release())


def stand(LogLiks):
    """``stand()`` (``R/mvJointModelBayes.R:940-945``)."""
    x = np.ascontiguousarray(LogLiks, dtype=np.float64)
    w = np.empty_like(x)
    _check(_lib().jmb_stand_weights(x.ctypes.data_as(c_double_p), x.size, w.ctypes.data_as(c_double_p)))
    return w


def series_statistics(x, S, M, series_stride, draw_stride, weights, probs=(0.025, 0.975), device=0, want=NAMES):
    """Statistics of S series of M draws laid out with the given strides (in doubles) in the host array ``x``.

    Returns ``(dict name -> array [S] or [S, 2], kernel_ms)``.
    """
    x = np.asarray(x)
    if x.dtype != np.float64:
        raise TypeError("x must be float64")
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pr = np.ascontiguousarray(probs, dtype=np.float64)
    bufs = {name: np.empty((S, _WIDTH[name]) if name in _WIDTH else S) for name in want}
    out = StatsOut()
    for name, b in bufs.items():
        setattr(out, name, b.ctypes.data_as(c_double_p))
    ms = C.c_float(0.0)
    _check(_lib().jmb_mcmc_statistics(device, C.c_void_p(x.ctypes.data), S, M, series_stride, draw_stride,
                                      w.ctypes.data_as(c_double_p), pr.ctypes.data_as(c_double_p), 0, C.byref(out), C.byref(ms)))
    return bufs, float(ms.value)


def device_statistics(x_ptr, S, M, series_stride, draw_stride, weights, out_ptrs: dict, probs=(0.025, 0.975), device=0):
    """The same on buffers that already live on the device (raw device addresses); returns the kernel time in ms."""
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pr = np.ascontiguousarray(probs, dtype=np.float64)
    out = StatsOut()
    for name, p in out_ptrs.items():
        setattr(out, name, C.cast(C.c_void_p(p), c_double_p))
    ms = C.c_float(0.0)
    _check(_lib().jmb_mcmc_statistics(device, C.c_void_p(x_ptr), S, M, series_stride, draw_stride,
                                      w.ctypes.data_as(c_double_p), pr.ctypes.data_as(c_double_p), 1, C.byref(out), C.byref(ms)))
    return float(ms.value)


def _margins(x):
    """summary_fun's ``dd`` (``:929-931``): 2 for a matrix, c(2, 3) for an M x q x q array, c(1, 2) otherwise (``b``)."""
    if x.ndim == 2:
        return (1,), 0
    if x.ndim == 3 and x.shape[1] == x.shape[2]:
        return (1, 2), 0
    if x.ndim == 3:
        return (0, 1), 2
    raise ValueError("mcmc entries must be vectors, matrices or 3-d arrays")


def statistics(mcmc: dict, LogLiks=None, weights=None, probs=(0.025, 0.975), device=0):
    """``statistics`` of ``mvJointModelBayes()``: ``{statistic: {parameter: array}}`` plus the weights.

    ``mcmc`` holds the reference's arrays in R's layout seen from NumPy in Fortran order: M x p matrices, M x q x q arrays,
    and ``b`` as n x q x M.  ``CIs`` / ``wCIs`` come back as R returns them from ``apply``: the two limits first.
    """
    if weights is None:
        weights = stand(LogLiks)
    res = {name: {} for name in NAMES}
    for key, val in mcmc.items():
        if val is None:
            continue
        x = np.asarray(val, dtype=np.float64)
        if x.ndim == 1:
            xf, S, M, ss, ds, shape = np.ascontiguousarray(x), 1, x.size, x.size, 1, ()
        else:
            keep, draw_axis = _margins(x)
            xf = np.asfortranarray(x)
            shape = tuple(x.shape[a] for a in keep)
            S, M = int(np.prod(shape)), x.shape[draw_axis]
            ss, ds = (M, 1) if draw_axis == 0 else (1, S)
        out, _ = series_statistics(xf.reshape(-1, order="F") if x.ndim > 1 else xf, S, M, ss, ds, weights, probs, device)
        for name in NAMES:
            a = out[name]
            if name in _WIDTH:
                res[name][key] = a.T.reshape((2,) + shape, order="F") if shape else a.reshape(2)
            else:
                res[name][key] = a.reshape(shape, order="F") if shape else float(a[0])
    return {"weights": np.asarray(weights), "statistics": res}
