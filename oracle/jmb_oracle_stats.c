/*
 * jmb_oracle_stats.c - plain-C restatement of the `statistics` block of mvJointModelBayes()
 * (R/mvJointModelBayes.R:925-971; R/modes.R, R/effectiveSize.R, R/stdErr.R, R/computeP.R).
 *
 * TEST INFRASTRUCTURE ONLY: used by tests/ and by bench.py's CPU-baseline legs, never by the product.
 * Written independently of oracle/oracle_stats.py (the NumPy restatement); tests/test_stats_cpu.py requires the two to
 * agree.  PARITY UNPINNED: R and Hmisc are absent from this image and the reference holds no fixture for these numbers; the
 * functions follow the published algorithms of base R's stats package (mean, var, quantile type 7, cov.wt, bw.nrd,
 * density.default = BinDist + FFT convolution + approx, ar.yw.default = acf + eureka + AIC) and of Hmisc's
 * wtd.quantile(type = "i/(n+1)") / wtd.Ecdf / wtd.table.
 * density.default is the algorithm of R <= 4.3 (= density(old.coords = TRUE) from R 4.4.0 on), see oracle_stats.py.
 *
 * Series are the rows of X (S x M, row-major).  Output: out[k * S + s], k = 0 postMeans, 1 postwMeans, 2 postModes,
 * 3 EffectiveSize, 4 StDev, 5 wStDev, 6 StErr, 7 / 8 CIs lower / upper, 9 / 10 wCIs lower / upper, 11 Pvalues.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define N_CELLS 1024 /* density.default: n = 1000 -> 2^ceiling(log2(1000)) */
#define N_USER 1000

static int cmp_double(const void* a, const void* b) {
    const double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

typedef struct {
    double x, w;
} xw_t;
static int cmp_xw(const void* a, const void* b) {
    const double x = ((const xw_t*)a)->x, y = ((const xw_t*)b)->x;
    return (x > y) - (x < y);
}

/* quantile(x, p, type = 7) of sorted xs */
static double q7(const double* xs, int M, double p) {
    const double index = (double)(M - 1) * p;
    const int lo = (int)floor(index), hi = (int)ceil(index);
    double qs = xs[lo];
    if (index > (double)lo && xs[hi] != qs) {
        const double h = index - (double)lo;
        qs = (1.0 - h) * qs + h * xs[hi];
    }
    return qs;
}

/* iterative radix-2 complex FFT of length 2 * N_CELLS, sign = -1 forward, +1 inverse (unnormalised, as R's fft) */
static void fft(double* re, double* im, int n, int sign) {
    static _Thread_local double twr[N_CELLS], twi[N_CELLS];
    static _Thread_local int have = 0;
    if (!have) {
        for (int k = 0; k < N_CELLS; ++k) { twr[k] = cos(2.0 * M_PI * k / (2.0 * N_CELLS)); twi[k] = sin(2.0 * M_PI * k / (2.0 * N_CELLS)); }
        have = 1;
    }
    for (int i = 1, j = 0; i < n; ++i) {
        int bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) {
            double t = re[i]; re[i] = re[j]; re[j] = t;
            t = im[i]; im[i] = im[j]; im[j] = t;
        }
    }
    for (int len = 2; len <= n; len <<= 1) {
        const int stride = n / len;
        for (int i = 0; i < n; i += len) {
            for (int k = 0; k < len / 2; ++k) {
                const double wr = twr[k * stride], wi = sign * twi[k * stride];
                const int a = i + k, b = i + k + len / 2;
                const double xr = re[b] * wr - im[b] * wi, xi = re[b] * wi + im[b] * wr;
                re[b] = re[a] - xr; im[b] = im[a] - xi;
                re[a] += xr; im[a] += xi;
            }
        }
    }
}

/* approx(x, y, xout = v, rule = 2) for increasing x (R's approx1: bisection, then linear interpolation) */
static double approx1(const double* x, const double* y, int n, double v) {
    if (v < x[0]) return y[0];
    if (v > x[n - 1]) return y[n - 1];
    int i = 0, j = n - 1;
    while (i < j - 1) {
        const int ij = (i + j) / 2;
        if (v < x[ij]) j = ij; else i = ij;
    }
    if (v == x[j]) return y[j];
    if (v == x[i]) return y[i];
    return y[i] + (y[j] - y[i]) * ((v - x[i]) / (x[j] - x[i]));
}

/* modes(): density(y, bw = "nrd", adjust = 3, n = 1000); d$x[which.max(d$y)] */
static double post_mode(const double* x, const double* xs, int M, double sd) {
    const double h = (q7(xs, M, 0.75) - q7(xs, M, 0.25)) / 1.34;
    const double bw = 3.0 * (1.06 * fmin(sd, h) * pow((double)M, -0.2));
    if (!(bw > 0.0) || !isfinite(bw)) return NAN;               /* density.default: 'bw' is not positive */
    const int n = N_CELLS, n2 = 2 * N_CELLS;
    const double from = xs[0] - 3.0 * bw, to = xs[M - 1] + 3.0 * bw, lo = from - 4.0 * bw, up = to + 4.0 * bw;
    static _Thread_local double yr[2 * N_CELLS], yi[2 * N_CELLS], kr[2 * N_CELLS], ki[2 * N_CELLS], xords[N_CELLS], dens[N_CELLS];
    memset(yr, 0, sizeof yr); memset(yi, 0, sizeof yi); memset(ki, 0, sizeof ki);
    const double xdelta = (up - lo) / (double)(n - 1), wi = 1.0 / (double)M;
    for (int i = 0; i < M; ++i) {                                /* BinDist (massdist.c) */
        const double xpos = (x[i] - lo) / xdelta;
        const int ix = (int)floor(xpos);
        const double fx = xpos - ix;
        if (0 <= ix && ix <= n - 2) { yr[ix] += wi * (1.0 - fx); yr[ix + 1] += wi * fx; }
        else if (ix == -1) yr[0] += wi * fx;
        else if (ix == n - 1) yr[ix] += wi * (1.0 - fx);
    }
    /* kords <- seq.int(0, 2 * (up - lo), length.out = 2n); kords[(n + 2):(2n)] <- -kords[n:2]; dnorm(kords, sd = bw) */
    const double step = 2.0 * (up - lo) / (double)(n2 - 1);
    for (int m = 0; m < n2; ++m) {
        const double k = (m <= n) ? step * m : step * (n2 - m);
        kr[m] = exp(-0.5 * (k / bw) * (k / bw)) / (bw * sqrt(2.0 * M_PI));
    }
    fft(yr, yi, n2, -1);
    fft(kr, ki, n2, -1);
    for (int m = 0; m < n2; ++m) {                               /* fft(y) * Conj(fft(kords)) */
        const double a = yr[m], b = yi[m], c = kr[m], d = -ki[m];
        yr[m] = a * c - b * d; yi[m] = a * d + b * c;
    }
    fft(yr, yi, n2, +1);
    for (int j = 0; j < n; ++j) { dens[j] = fmax(0.0, yr[j] / (double)n2); xords[j] = lo + j * (up - lo) / (double)(n - 1); }
    double best = -1.0, arg = NAN;
    for (int t = 0; t < N_USER; ++t) {
        const double xt = (t == N_USER - 1) ? to : from + t * (to - from) / (double)(N_USER - 1);
        const double v = approx1(xords, dens, n, xt);
        if (v > best) { best = v; arg = xt; }
    }
    return arg;
}

/* spectrum0.ar of one series (R/effectiveSize.R:4-19): 0 for a (numerically) exactly linear series, else var.pred /
 * (1 - sum(ar))^2 of ar(x, aic = TRUE) = ar.yw: order.max = min(n - 1, floor(10 log10 n)), acf with divisor n, Levinson
 * recursion, AIC n log(var.pred) + 2 order + 2, first minimum, var.pred scaled by n / (n - (order + 1)) */
static double spectrum0_ar(const double* x, int M, double mean) {
    const double tb = 0.5 * (M + 1);
    double sty = 0.0, stt = 0.0;
    for (int m = 0; m < M; ++m) { const double t = (m + 1) - tb; sty += t * (x[m] - mean); stt += t * t; }
    const double beta = sty / stt;
    double rss = 0.0, rsum = 0.0;
    for (int m = 0; m < M; ++m) { const double r = (x[m] - mean) - beta * ((m + 1) - tb); rss += r * r; rsum += r; }
    const double rmean = rsum / M;
    const double sd_res = sqrt((rss - M * rmean * rmean) / (M - 1));
    if (!(sd_res > 1.5e-8)) return 0.0;                          /* identical(all.equal(sd(res), 0), TRUE) */
    int om = (int)floor(10.0 * log10((double)M));
    if (om > M - 1) om = M - 1;
    double r[64] = {0.0}, phi[64] = {0.0}, prev[64];
    for (int l = 0; l <= om; ++l) {
        double s = 0.0;
        for (int m = 0; m + l < M; ++m) s += (x[m] - mean) * (x[m + l] - mean);
        r[l] = s / M;
    }
    double v = r[0], best = M * log(v) + 2.0, vbest = v, sumphi = 0.0;
    int order = 0;
    for (int l = 1; l <= om; ++l) {
        double acc = r[l];
        for (int j = 1; j < l; ++j) acc -= phi[j] * r[l - j];
        const double k = acc / v;
        memcpy(prev, phi, sizeof phi);
        for (int j = 1; j < l; ++j) phi[j] = prev[j] - k * prev[l - j];
        phi[l] = k;
        v *= 1.0 - k * k;
        const double aic = M * log(v) + 2.0 * l + 2.0;
        if (aic < best) {
            best = aic; order = l; vbest = v; sumphi = 0.0;
            for (int j = 1; j <= l; ++j) sumphi += phi[j];
        }
    }
    const double var_pred = vbest * M / (double)(M - (order + 1));
    return var_pred / ((1.0 - sumphi) * (1.0 - sumphi));
}

/* Hmisc::wtd.quantile(x, weights, probs, type = "i/(n+1)") */
static void wtd_quantiles(const double* x, const double* w, int M, const double probs[2], double out[2]) {
    xw_t* a = (xw_t*)malloc(sizeof(xw_t) * (size_t)M);
    double* ux = (double*)malloc(sizeof(double) * (size_t)(M + 1));
    double* cdf = (double*)malloc(sizeof(double) * (size_t)(M + 1));
    int n = 0;
    for (int m = 0; m < M; ++m) if (w[m] != 0.0) { a[n].x = x[m]; a[n].w = w[m]; ++n; }      /* zero weights dropped */
    if (n == 0) { out[0] = out[1] = NAN; free(a); free(ux); free(cdf); return; }
    qsort(a, (size_t)n, sizeof(xw_t), cmp_xw);
    int u = 0;
    double tot = 0.0;
    for (int i = 0; i < n;) {                                    /* wtd.table: ties summed */
        int j = i;
        double s = 0.0;
        while (j < n && a[j].x == a[i].x) { s += a[j].w; ++j; }
        tot += s;
        ux[u + 1] = a[i].x; cdf[u + 1] = tot;
        ++u; i = j;
    }
    for (int k = 1; k <= u; ++k) cdf[k] /= tot + 1.0;            /* ecdf = cumsum / (n + 1) */
    const double* px = ux + 1; const double* pc = cdf + 1; int len = u;
    if (cdf[1] > 0.0) { ux[0] = ux[1]; cdf[0] = 0.0; px = ux; pc = cdf; len = u + 1; }
    for (int e = 0; e < 2; ++e) out[e] = approx1(pc, px, len, probs[e]);
    free(a); free(ux); free(cdf);
}

int orc_stats(const double* X, long S, int M, const double* w, const double probs[2], double* out) {
    if (M < 4 || M > 100000) return 1;
    double sumw = 0.0, sumwt2 = 0.0;
    for (int m = 0; m < M; ++m) sumw += w[m];
    for (int m = 0; m < M; ++m) sumwt2 += (w[m] / sumw) * (w[m] / sumw);
    double* xs = (double*)malloc(sizeof(double) * (size_t)M);
    for (long s = 0; s < S; ++s) {
        const double* x = X + s * (long)M;
        long double acc = 0.0L;                                  /* R's mean: long double sum, then one refinement */
        for (int m = 0; m < M; ++m) acc += x[m];
        long double mu = acc / M, t = 0.0L;
        for (int m = 0; m < M; ++m) t += x[m] - mu;
        mu += t / M;
        const double mean = (double)mu;
        double ss = 0.0, wm = 0.0;
        int ge = 0;
        for (int m = 0; m < M; ++m) { ss += (x[m] - mean) * (x[m] - mean); wm += x[m] * w[m]; ge += x[m] >= 0.0; }
        const double var = ss / (M - 1), center = wm / sumw;
        double ws = 0.0;
        for (int m = 0; m < M; ++m) ws += (w[m] / sumw) * (x[m] - center) * (x[m] - center);     /* cov.wt, unbiased */
        memcpy(xs, x, sizeof(double) * (size_t)M);
        qsort(xs, (size_t)M, sizeof(double), cmp_double);
        const double spec = spectrum0_ar(x, M, mean);
        const double ess = spec == 0.0 ? 0.0 : M * var / spec;
        double wq[2];
        wtd_quantiles(x, w, M, probs, wq);
        out[0 * S + s] = mean;
        out[1 * S + s] = wm;
        out[2 * S + s] = post_mode(x, xs, M, sqrt(var));
        out[3 * S + s] = ess;
        out[4 * S + s] = sqrt(var);
        out[5 * S + s] = sqrt(ws / (1.0 - sumwt2));
        out[6 * S + s] = sqrt(var / ess);
        out[7 * S + s] = q7(xs, M, probs[0]);
        out[8 * S + s] = q7(xs, M, probs[1]);
        out[9 * S + s] = wq[0];
        out[10 * S + s] = wq[1];
        const double above = (double)ge / M, below = (double)(M - ge) / M;
        out[11 * S + s] = 2.0 * fmin(above, below);
    }
    free(xs);
    return 0;
}
