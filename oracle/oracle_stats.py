"""NumPy restatement of the `statistics` block of mvJointModelBayes().

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.

PARITY UNPINNED: R is absent from this image and the reference holds no fixture for these summaries.  Every function
restates the R code it cites; where the arithmetic lives outside /root/reference the published algorithm is restated:
  * base R `stats` (R >= 3.x; DESCRIPTION pins no version): mean, var / sd, quantile (type 7), cov.wt, bw.nrd,
    density.default (linear binning BinDist + FFT convolution + approx), ar.yw.default (acf with divisor n, Levinson
    recursion `eureka`, AIC order selection), all.equal.numeric;
  * density.default is restated as it stood up to R 4.3 (the releases JMbayes 0.8 was developed against): the kernel
    ordinates are spaced 2 (up - lo) / (2n - 1) apart while the cells are (up - lo) / (n - 1) wide.  R 4.4.0 changed the
    grid ("old.coords = FALSE", densities differ by a factor of about 0.999); density(old.coords = TRUE) is this algorithm.
  * Hmisc (Imports of DESCRIPTION:13, no version pinned): wtd.quantile(type = "i/(n+1)") through wtd.Ecdf / wtd.table.
Checked in tests/test_stats_cpu.py against numpy.quantile, numpy.cov(aweights), scipy.linalg.solve_toeplitz and
analytic cases.

A "series" is the vector of the M kept draws of one scalar of the mcmc list (R/mvJointModelBayes.R:925-938 applies FUN over
every series: columns of the M x p matrices, cells of M x q x q arrays, cells (i, j) of the n x q x M array `b`).
All functions here take X of shape (S, M): one series per row.
"""
from __future__ import annotations

import numpy as np

N_USER = 1000          # modes(): density(y, bw = "nrd", adjust = 3, n = 1000)   R/modes.R:3
ADJUST = 3.0
CUT = 3.0


def stand(loglik):
    """stand() R/mvJointModelBayes.R:940-945: importance weights from the log-likelihood contributions."""
    x = np.asarray(loglik, float)
    upp = np.nanmax(x) + np.log(x.size)
    w = np.exp(x - upp)
    return w / w.sum()


def post_means(X):
    """mean(x, na.rm = TRUE): R's two-pass mean (summary.c: s / n, then the mean of the residuals added)."""
    X = np.asarray(X, float)
    m = X.sum(1) / X.shape[1]
    return m + (X - m[:, None]).sum(1) / X.shape[1]


def post_wmeans(X, w):
    """wmean() R/mvJointModelBayes.R:948: sum(x * weights)."""
    return (np.asarray(X, float) * np.asarray(w, float)[None, :]).sum(1)


def st_dev(X):
    """sd(x): sqrt(sum((x - mean)^2) / (n - 1)) (cov.c two-pass)."""
    X = np.asarray(X, float)
    m = post_means(X)
    return np.sqrt(((X - m[:, None]) ** 2).sum(1) / (X.shape[1] - 1))


def w_st_dev(X, w):
    """wsd() R/mvJointModelBayes.R:949: sqrt(cov.wt(cbind(x), wt = weights)$cov), method "unbiased"."""
    X = np.asarray(X, float)
    wt = np.asarray(w, float)
    wt = wt / wt.sum()
    center = (X * wt[None, :]).sum(1)
    xc = np.sqrt(wt)[None, :] * (X - center[:, None])
    return np.sqrt((xc * xc).sum(1) / (1.0 - (wt * wt).sum()))


def quantile7(Xs, p):
    """quantile(x, probs, type = 7) on rows sorted ascending (stats:::quantile.default)."""
    M = Xs.shape[1]
    index = (M - 1) * p
    lo = int(np.floor(index))
    hi = int(np.ceil(index))
    h = index - lo
    qs = Xs[:, lo].copy()
    upd = (index > lo) & (Xs[:, hi] != qs)
    qs[upd] = (1.0 - h) * qs[upd] + h * Xs[upd, hi]
    return qs


def cis(X, probs=(0.025, 0.975)):
    """CIs = quantile(x, probs = c(0.025, 0.975)) R/mvJointModelBayes.R:966; returns (S, 2)."""
    Xs = np.sort(np.asarray(X, float), axis=1)
    return np.stack([quantile7(Xs, p) for p in probs], axis=1)


def w_cis(X, w, probs=(0.025, 0.975)):
    """wCIs = Hmisc::wtd.quantile(x, weights, probs, type = "i/(n+1)", na.rm = TRUE) R/mvJointModelBayes.R:967-969.

    Hmisc: observations with zero weight are dropped; wtd.table sorts x and sums the weights of tied values;
    wtd.Ecdf: n = sum(weights) (normwt = FALSE), ecdf = cumsum(weights) / (n + 1), the point (0, x[1]) is prepended when
    ecdf[1] > 0; the quantiles are approx(ecdf, x, xout = probs, rule = 2)$y.  (With weights that sum to one the ecdf
    stops at 1/2, so the upper limit is the largest draw: that is what the reference computes.)
    """
    X = np.asarray(X, float)
    w = np.asarray(w, float)
    out = np.empty((X.shape[0], len(probs)))
    for s in range(X.shape[0]):
        keep = w != 0
        x, ww = X[s, keep], w[keep]
        order = np.argsort(x, kind="stable")
        x, ww = x[order], ww[order]
        ux, start = np.unique(x, return_index=True)
        sw = np.add.reduceat(ww, start)
        n = sw.sum()
        cdf = np.cumsum(sw) / (n + 1.0)
        if cdf[0] > 0:
            ux = np.concatenate([[ux[0]], ux])
            cdf = np.concatenate([[0.0], cdf])
        for j, p in enumerate(probs):
            out[s, j] = _approx1(cdf, ux, p)
    return out


def _approx1(x, y, v):
    """approx(x, y, xout = v, rule = 2) for strictly increasing x (R_approxfun, linear)."""
    if v < x[0]:
        return y[0]
    if v > x[-1]:
        return y[-1]
    i, j = 0, x.size - 1
    while i < j - 1:                     # the bisection of approx1() in R's approx.c
        ij = (i + j) // 2
        if v < x[ij]:
            j = ij
        else:
            i = ij
    if v == x[j]:
        return y[j]
    if v == x[i]:
        return y[i]
    return y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]))


def p_values(X):
    """computeP() R/computeP.R:1-6."""
    X = np.asarray(X, float)
    above = (X >= 0).mean(1)
    below = (X < 0).mean(1)
    return 2.0 * np.minimum(above, below)


def order_max(M):
    """ar.yw.default: order.max = min(n.obs - 1, floor(10 * log10(n.obs)))."""
    return int(min(M - 1, np.floor(10.0 * np.log10(M))))


def levinson(r):
    """AR coefficients and innovation variances of orders 1..len(r)-1 from autocovariances r[0..]: what eureka
    (R src/library/stats/src/eureka.f, called by ar.yw.default) returns as coefs[l, 1..l] and vars[l]."""
    om = r.size - 1
    coefs = np.zeros((om + 1, om + 1))
    v = np.empty(om + 1)
    v[0] = r[0]
    for l in range(1, om + 1):
        acc = r[l]
        for j in range(1, l):
            acc -= coefs[l - 1, j] * r[l - j]
        k = acc / v[l - 1]
        coefs[l, l] = k
        for j in range(1, l):
            coefs[l, j] = coefs[l - 1, j] - k * coefs[l - 1, l - j]
        v[l] = v[l - 1] * (1.0 - k * k)
    return coefs, v


def spectrum0_ar(x):
    """spectrum0.ar() of one series, R/effectiveSize.R:4-19 (copied there from coda)."""
    M = x.size
    t = np.arange(1, M + 1, dtype=float)
    tb, xb = t.mean(), x.mean()
    beta = ((t - tb) * (x - xb)).sum() / ((t - tb) ** 2).sum()
    res = (x - xb) - beta * (t - tb)                   # lm.fit(cbind(1, seq_len(nrx)), x)$residuals
    sd_res = np.sqrt(((res - res.mean()) ** 2).sum() / (M - 1))
    # identical(all.equal(sd(res), 0), TRUE): relative comparison unless the target is below the tolerance 1.5e-8,
    # then absolute: "equal" exactly when sd(res) <= 1.5e-8
    if not (sd_res > 1.5e-8):
        return 0.0
    om = order_max(M)
    xc = x - xb
    r = np.array([(xc[: M - l] * xc[l:]).sum() / M for l in range(om + 1)])     # acf(type = "covariance"), divisor n
    coefs, v = levinson(r)
    aic = M * np.log(v) + 2.0 * np.arange(om + 1) + 2.0
    order = int(np.argmin(aic))                         # (0:order.max)[xaic == 0]: the first minimum
    var_pred = v[order] * M / (M - (order + 1))
    return var_pred / (1.0 - coefs[order, 1 : order + 1].sum()) ** 2


def effective_size(X):
    """effectiveSize() R/effectiveSize.R:1-24: nrow(x) * var(x) / spectrum0.ar(x), 0 where the spectrum is 0."""
    X = np.asarray(X, float)
    M = X.shape[1]
    spec = np.array([spectrum0_ar(x) for x in X])
    var = st_dev(X) ** 2
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(spec == 0, 0.0, M * var / spec)


def std_err(X):
    """stdErr() R/stdErr.R:1-7: sqrt(var / effectiveSize)."""
    X = np.asarray(X, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sqrt(st_dev(X) ** 2 / effective_size(X))


def bw_nrd(Xs, sd):
    """bw.nrd(x): 1.06 * min(sqrt(var(x)), IQR / 1.34) * n^-0.2 (rows of Xs sorted)."""
    h = (quantile7(Xs, 0.75) - quantile7(Xs, 0.25)) / 1.34
    return 1.06 * np.minimum(sd, h) * Xs.shape[1] ** (-0.2)


def density_grid(x, bw):
    """density.default(x, bw, n = 1000) of one series: returns (x grid [1000], y [1000]).

    n = max(n, 512) rounded up to a power of two (1024); lo = from - 4 bw, up = to + 4 bw with from / to = range -/+ cut bw;
    y = BinDist(x, 1 / N, lo, up, n) (linear binning, 2n cells); kords = dnorm(seq(0, 2 (up - lo), length = 2n) folded, sd = bw);
    kords = pmax(0, Re(fft(fft(y) * Conj(fft(kords)), inverse = TRUE))[1:n] / 2n); approx(xords, kords, seq(from, to, n.user)).
    """
    N = x.size
    n = 1 << int(np.ceil(np.log2(max(N_USER, 512))))
    frm, to = x.min() - CUT * bw, x.max() + CUT * bw
    lo, up = frm - 4.0 * bw, to + 4.0 * bw
    y = np.zeros(2 * n)
    xdelta = (up - lo) / (n - 1)
    xpos = (x - lo) / xdelta
    ix = np.floor(xpos).astype(np.int64)
    fx = xpos - ix
    wi = 1.0 / N
    for i in range(N):                                  # BinDist (R src/library/stats/src/massdist.c)
        if 0 <= ix[i] <= n - 2:
            y[ix[i]] += wi * (1.0 - fx[i])
            y[ix[i] + 1] += wi * fx[i]
        elif ix[i] == -1:
            y[0] += wi * fx[i]
        elif ix[i] == n - 1:
            y[ix[i]] += wi * (1.0 - fx[i])
    kords = np.linspace(0.0, 2.0 * (up - lo), 2 * n)
    kords[n + 1 :] = -kords[n - 1 : 0 : -1]
    kords = np.exp(-0.5 * (kords / bw) ** 2) / (bw * np.sqrt(2.0 * np.pi))
    conv = np.fft.ifft(np.fft.fft(y) * np.conj(np.fft.fft(kords))).real[:n]     # numpy's ifft already divides by 2n
    dens = np.maximum(0.0, conv)
    xords = np.linspace(lo, up, n)
    xg = np.linspace(frm, to, N_USER)
    return xg, np.interp(xg, xords, dens)


def post_modes(X):
    """modes() R/modes.R:1-5: d <- density(y, bw = "nrd", adjust = 3, n = 1000); d$x[which.max(d$y)]; NA on error."""
    X = np.asarray(X, float)
    Xs = np.sort(X, axis=1)
    bw = ADJUST * bw_nrd(Xs, st_dev(X))
    out = np.full(X.shape[0], np.nan)
    for s in range(X.shape[0]):
        if not np.isfinite(bw[s]) or bw[s] <= 0:        # density.default stops: 'bw' is not positive
            continue
        xg, yg = density_grid(X[s], bw[s])
        out[s] = xg[int(np.argmax(yg))]
    return out


def statistics(X, weights, probs=(0.025, 0.975)):
    """All ten entries of `statistics` (R/mvJointModelBayes.R:959-971) for the series in the rows of X."""
    X = np.asarray(X, float)
    return {
        "postMeans": post_means(X), "postwMeans": post_wmeans(X, weights), "postModes": post_modes(X),
        "EffectiveSize": effective_size(X), "StDev": st_dev(X), "wStDev": w_st_dev(X, weights), "StErr": std_err(X),
        "CIs": cis(X, probs), "wCIs": w_cis(X, weights, probs), "Pvalues": p_values(X),
    }


C_ROWS = ("postMeans", "postwMeans", "postModes", "EffectiveSize", "StDev", "wStDev", "StErr", "CIs0", "CIs1", "wCIs0", "wCIs1", "Pvalues")


def statistics_c(X, weights, probs=(0.025, 0.975)):
    """The same ten entries from the independently written plain-C restatement (oracle/jmb_oracle_stats.c)."""
    import ctypes as C

    from . import oracle
    L = oracle.lib()
    dp = C.POINTER(C.c_double)
    L.orc_stats.argtypes = [dp, C.c_long, C.c_int, dp, dp, dp]
    X = np.ascontiguousarray(X, dtype=np.float64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    pr = np.ascontiguousarray(probs, dtype=np.float64)
    S, M = X.shape
    out = np.empty((12, S))
    rc = L.orc_stats(X.ctypes.data_as(dp), S, M, w.ctypes.data_as(dp), pr.ctypes.data_as(dp), out.ctypes.data_as(dp))
    if rc:
        raise RuntimeError(f"orc_stats failed with code {rc}")
    res = {name: out[k] for k, name in enumerate(C_ROWS) if not name[-1].isdigit()}
    res["CIs"] = np.stack([out[7], out[8]], axis=1)
    res["wCIs"] = np.stack([out[9], out[10]], axis=1)
    return res
