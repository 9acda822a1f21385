/*
 * jmb_oracle_svft.c - CPU restatement of survfitJM.mvJMbayes()'s per-subject loops (BASELINE config 5).
 *
 * TEST INFRASTRUCTURE ONLY (see jmb_oracle.h): the checker of the CUDA path in jmbayes_b200/csrc/
 * jmb_survfit.cu and the CPU baseline of bench.py's config-5 workload.  Nothing in the product calls it.
 *
 * What is restated, with the reference lines each function follows (JMbayes source tree):
 *   log_post_RE_svft / survPred_svft_2     src/survfitJM_mvJMbayes.cpp:57-187 (dense designs, cumsum rowsum)
 *   Xbetas_calc                            R/supportFuns_mvJointModelBayes.R:43-55
 *   cd()                                   R/cd.R:1-15
 *   rmvt() / dmvt(prop = TRUE)             R/rmvt.R, R/dmvt.R
 *   the mode search, MH loop and summaries R/survfitJM.mvJMbayes.R:277-414
 * Third-party arithmetic that is NOT under /root/reference and is restated from its published algorithm:
 *   optim(method = "BFGS", hessian = TRUE)  R's own C code (R >= 3.0, src/appl/optim.c: vmmin(), and
 *                                           src/library/stats/src/optim.c: fminfn / fmingr parscale handling
 *                                           and optimhess()); JMbayes' DESCRIPTION pins no R version.
 *   solve(), eigen(symmetric = TRUE)        LAPACK through R; restated as Gauss-Jordan with partial pivoting
 *                                           and a cyclic Jacobi iteration (eigenvector signs are arbitrary in
 *                                           both; only the distribution of rmvt() depends on them)
 *   quantile(type = 7), median, rowMeans    R's stats / base
 *   rnorm / rchisq / runif                  R's RNG -> the Philox streams of DESIGN.md section 4
 * PARITY PIN: log_post_RE_svft / survPred_svft_2 are pinned to the reference's own source through
 * tests/golden/svft_config3.npz (tests/test_reference_golden.py); the R-level loop has no reference
 * fixture and no R here, so for it parity is "unpinned": checked against analytic properties
 * (tests/test_survfit_cpu.py) and an independent NumPy/SciPy evaluation only.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "jmb_oracle.h"
#include "jmb_oracle_svft.h"

#define SV_KEY 0x53564654u   /* second half of the Philox key of every survfit stream */
enum { SV_ST_NORM = 0, SV_ST_UNIF = 1 };

static double sv_trans(int tf, double x) {     /* src/survfitJM_mvJMbayes.cpp:80-98 */
    switch (tf) {
        case 0: return x;
        case 1: { double e = exp(x); return e / (1.0 + e); }
        case 2: return exp(x);
        case 3: return log(x);
        case 4: return log2(x);
        case 5: return log10(x);
        default: return sqrt(x);
    }
}

static double sv_pnorm(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* parameter set `m` of `P` (R layout: draws are rows) */
typedef struct sv_par {
    double betas[JMB_MAX_OUTCOMES][JMB_SVFT_MAX_PK];
    double sigma[JMB_MAX_OUTCOMES];
    double invD[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
    double Bs[JMB_MAX_BS], gam[JMB_MAX_GAMMAS], alp[JMB_MAX_ALPHAS];
} sv_par;

static void sv_load_par(const jmb_svft_data* d, const jmb_svft_params* P, int m, sv_par* o) {
    const int M = P->M, q = d->q;
    for (int k = 0; k < d->n_outcomes; ++k) {
        for (int j = 0; j < d->p_k[k]; ++j) o->betas[k][j] = P->betas[k][m + (size_t)j * M];
        o->sigma[k] = (d->fams[k] == JMB_FAM_GAUSSIAN && P->sigmas && P->sigmas[k]) ? P->sigmas[k][m] : 1.0;
    }
    for (int j = 0; j < q * q; ++j) o->invD[j] = P->invD[m + (size_t)j * M];
    for (int j = 0; j < d->n_Bs; ++j) o->Bs[j] = P->Bs_gammas[m + (size_t)j * M];
    for (int j = 0; j < d->n_gammas; ++j) o->gam[j] = P->gammas[m + (size_t)j * M];
    for (int j = 0; j < d->n_alphas; ++j) o->alp[j] = P->alphas[m + (size_t)j * M];
}

/* per-subject row ranges of the longitudinal outcomes (idL grouped by subject) */
typedef struct sv_index { int64_t* rp[JMB_MAX_OUTCOMES]; } sv_index;

static int sv_build_index(const jmb_svft_data* d, sv_index* ix) {
    for (int k = 0; k < d->n_outcomes; ++k) {
        int64_t* rp = (int64_t*)calloc((size_t)d->n + 1, sizeof(int64_t));
        if (!rp) return 1;
        int prev = -1;
        for (int64_t r = 0; r < d->N[k]; ++r) {
            int s = d->idL[k][r];
            if (s < 0 || s >= d->n || s < prev) { free(rp); return 2; }
            prev = s; rp[s + 1] += 1;
        }
        for (int i = 0; i < d->n; ++i) rp[i + 1] += rp[i];
        ix->rp[k] = rp;
    }
    return 0;
}
static void sv_free_index(const jmb_svft_data* d, sv_index* ix) { for (int k = 0; k < d->n_outcomes; ++k) free(ix->rp[k]); }

/* -sum_k Pw_k exp(W1s Bs + W2s gam + Wlongs alp) over the K rows starting at `row0` of the given designs:
 * lin_pred_matF_svft + rowsum_svft(Pw % exp(...)), src/survfitJM_mvJMbayes.cpp:68-101,158-160,181-184 */
static double sv_minus_H(const jmb_svft_data* d, const sv_par* par, const double* b, int64_t row0, int64_t rows,
                         const double* W1s, const double* W2s, const double* Pw,
                         const double* const* XXs, const double* const* ZZs, const double* const* Us) {
    const int K = d->K, A = d->n_alphas;
    double cs = 0.0;                                  /* cumsum over the subject's K rows */
    for (int k = 0; k < K; ++k) {
        const int64_t r = row0 + k;
        double wl[JMB_MAX_ALPHAS];
        for (int c = 0; c < A; ++c) wl[c] = 0.0;
        for (int t = 0; t < d->n_terms; ++t) {
            const double* bet = par->betas[d->outcome[t]];
            double xb = 0.0;                          /* Xbetas_calc: XXs[[t]] %*% betas[[outcome]][indFixed] */
            for (int f = 0; f < d->f_t[t]; ++f) xb += XXs[t][r + (int64_t)f * rows] * bet[d->indFixed[t][f]];
            double zb = 0.0;
            for (int j = 0; j < d->r_t[t]; ++j) zb += ZZs[t][r + (int64_t)j * rows] * b[d->RE_inds2[t][j]];
            const double v = sv_trans(d->trans_Funs[t], xb + zb);
            for (int u = 0; u < d->u_t[t]; ++u) wl[d->col_inds[t][u]] = Us[t][r + (int64_t)u * rows] * v;
        }
        double s1 = 0.0, s2 = 0.0, s3 = 0.0;
        for (int j = 0; j < d->n_Bs; ++j) s1 += W1s[r + (int64_t)j * rows] * par->Bs[j];
        for (int j = 0; j < d->n_gammas; ++j) s2 += W2s[r + (int64_t)j * rows] * par->gam[j];
        for (int c = 0; c < A; ++c) s3 += wl[c] * par->alp[c];
        cs += Pw[r] * exp((s1 + s2) + s3);
    }
    return 0.0 - cs;
}

/* log_post_RE_svft for subject i: log p(y_i | b) + log p(b) - H_i(0, last.time] */
static double sv_log_post(const jmb_svft_data* d, const sv_index* ix, const sv_par* par, int i, const double* b) {
    const int q = d->q;
    double log_pyb = 0.0;
    for (int k = 0; k < d->n_outcomes; ++k) {          /* lin_predF_svft + log_longF_svft, :57-66,103-135 */
        const int64_t r0 = ix->rp[k][i], r1 = ix->rp[k][i + 1], Nk = d->N[k];
        if (r1 == r0) continue;
        double cs = 0.0;
        for (int64_t r = r0; r < r1; ++r) {
            double xb = 0.0;
            for (int j = 0; j < d->p_k[k]; ++j) xb += d->X[k][r + (int64_t)j * Nk] * par->betas[k][j];
            double zb = 0.0;
            for (int j = 0; j < d->q_k[k]; ++j) zb += d->Z[k][r + (int64_t)j * Nk] * b[d->RE_inds[k][j]];
            const double eta = xb + zb, y = d->y[k][r];
            double ld;
            if (d->fams[k] == JMB_FAM_GAUSSIAN) { double t = (y - eta) / par->sigma[k]; ld = -0.5 * (t * t); }
            else if (d->fams[k] == JMB_FAM_BINOMIAL) {
                double pr;
                if (d->links[k] == JMB_LINK_LOGIT) { double e = exp(eta); pr = e / (1.0 + e); }
                else if (d->links[k] == JMB_LINK_PROBIT) pr = sv_pnorm(eta);
                else pr = -exp(-exp(eta)) + 1.0;
                ld = y * log(pr) + (1.0 - y) * log(1.0 - pr);
            } else { double mu = exp(eta); ld = y * log(mu) - mu; }
            cs += ld;
        }
        log_pyb += cs;
    }
    double quad = 0.0;                                 /* - 0.5 * sum((b %*% invD) * b), :157 */
    for (int j = 0; j < q; ++j) {
        double t = 0.0;
        for (int l = 0; l < q; ++l) t += b[l] * par->invD[l + j * q];
        quad += t * b[j];
    }
    const double log_pb = -0.5 * quad;
    const int64_t nK = (int64_t)d->n * d->K;
    const double log_ptb = sv_minus_H(d, par, b, (int64_t)i * d->K, nK, d->W1s, d->W2s, d->Pw, d->XXs, d->ZZs, d->Us);
    return (log_pyb + log_ptb) + log_pb;               /* :163 */
}

/* ---- small dense helpers ------------------------------------------------------------------- */
/* solve(A): Gauss-Jordan with partial pivoting; returns 1 when singular */
int orc_sv_inverse(const double* A, int n, double* inv) {
    double a[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
    for (int j = 0; j < n * n; ++j) { a[j] = A[j]; inv[j] = 0.0; }
    for (int j = 0; j < n; ++j) inv[j + j * n] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c; double best = fabs(a[c + c * n]);
        for (int r = c + 1; r < n; ++r) if (fabs(a[r + c * n]) > best) { best = fabs(a[r + c * n]); piv = r; }
        if (!(best > 0.0) || !isfinite(best)) return 1;
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                double t = a[c + j * n]; a[c + j * n] = a[piv + j * n]; a[piv + j * n] = t;
                t = inv[c + j * n]; inv[c + j * n] = inv[piv + j * n]; inv[piv + j * n] = t;
            }
        const double d = 1.0 / a[c + c * n];
        for (int j = 0; j < n; ++j) { a[c + j * n] *= d; inv[c + j * n] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r + c * n];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { a[r + j * n] -= f * a[c + j * n]; inv[r + j * n] -= f * inv[c + j * n]; }
        }
    }
    return 0;
}

/* eigen(S, symmetric = TRUE): cyclic Jacobi, 12 sweeps, values sorted decreasingly; vec is n x n column-major */
void orc_sv_eigen_sym(const double* S, int n, double* val, double* vec) {
    double a[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i + j * n] = 0.5 * (S[i + j * n] + S[j + i * n]); vec[i + j * n] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 12; ++sweep)
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p + q * n], app = a[p + p * n], aqq = a[q + q * n];
                if (fabs(apq) <= 1e-20 * (fabs(app) + fabs(aqq))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {           /* columns p, q */
                    const double akp = a[k + p * n], akq = a[k + q * n];
                    a[k + p * n] = c * akp - s * akq; a[k + q * n] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {           /* rows p, q */
                    const double apk = a[p + k * n], aqk = a[q + k * n];
                    a[p + k * n] = c * apk - s * aqk; a[q + k * n] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = vec[k + p * n], vkq = vec[k + q * n];
                    vec[k + p * n] = c * vkp - s * vkq; vec[k + q * n] = s * vkp + c * vkq;
                }
            }
    for (int j = 0; j < n; ++j) val[j] = a[j + j * n];
    for (int j = 0; j < n - 1; ++j) {                  /* selection sort, decreasing */
        int best = j;
        for (int l = j + 1; l < n; ++l) if (val[l] > val[best]) best = l;
        if (best != j) {
            double t = val[j]; val[j] = val[best]; val[best] = t;
            for (int k = 0; k < n; ++k) { t = vec[k + j * n]; vec[k + j * n] = vec[k + best * n]; vec[k + best * n] = t; }
        }
    }
}

/* ---- optim(method = "BFGS") ------------------------------------------------------------------ */
typedef struct sv_obj {
    const jmb_svft_data* d; const sv_index* ix; const sv_par* par; int i; const jmb_svft_control* ct;
} sv_obj;

/* ff(b) = -log_post_RE_svft(b) (R/survfitJM.mvJMbayes.R:297) */
static double sv_ff(const sv_obj* o, const double* b) { return 0.0 - sv_log_post(o->d, o->ix, o->par, o->i, b); }

/* gg(b) = cd(b, ff, eps = 1e-3) (R/cd.R) */
static void sv_gg(const sv_obj* o, const double* x, double* res) {
    const int n = o->d->q;
    const double eps = o->ct->cd_eps;
    double x1[JMB_SVFT_MAX_Q], x2[JMB_SVFT_MAX_Q];
    for (int i = 0; i < n; ++i) {
        const double ex = eps * (fabs(x[i]) + eps);
        for (int j = 0; j < n; ++j) { x1[j] = x[j]; x2[j] = x[j]; }
        x1[i] = x[i] + ex; x2[i] = x[i] - ex;
        const double diff_f = sv_ff(o, x1) - sv_ff(o, x2);
        const double diff_x = x1[i] - x2[i];
        res[i] = diff_f / diff_x;
    }
}

/* fminfn / fmingr of optim(): parameters are handed to the optimiser divided by parscale */
static double sv_fminfn(const sv_obj* o, const double* p) {
    double x[JMB_SVFT_MAX_Q];
    for (int i = 0; i < o->d->q; ++i) x[i] = p[i] * o->ct->parscale;
    return sv_ff(o, x);
}
static void sv_fmingr(const sv_obj* o, const double* p, double* df) {
    double x[JMB_SVFT_MAX_Q];
    for (int i = 0; i < o->d->q; ++i) x[i] = p[i] * o->ct->parscale;
    sv_gg(o, x, df);
    for (int i = 0; i < o->d->q; ++i) df[i] = df[i] * o->ct->parscale;
}

/* vmmin(), R src/appl/optim.c (abstol = -Inf, all parameters free).  b is in optimiser units. */
static int sv_vmmin(const sv_obj* o, double* b, double* Fmin, int* fncount, int* grcount) {
    const int n = o->d->q, maxit = o->ct->maxit;
    const double reltol = o->ct->reltol, stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    double g[JMB_SVFT_MAX_Q], t[JMB_SVFT_MAX_Q], X[JMB_SVFT_MAX_Q], c[JMB_SVFT_MAX_Q];
    double B[JMB_SVFT_MAX_Q][JMB_SVFT_MAX_Q];
    int count = 0, funcount, gradcount, ilast, iter = 0;
    double f = sv_fminfn(o, b), gradproj, s, steplength, D1, D2;
    if (!isfinite(f)) return 10;                      /* error("initial value in 'vmmin' is not finite") */
    *Fmin = f;
    funcount = gradcount = 1;
    sv_fmingr(o, b, g);
    iter++;
    ilast = gradcount;
    do {
        if (ilast == gradcount)
            for (int i = 0; i < n; ++i) { for (int j = 0; j < i; ++j) B[i][j] = 0.0; B[i][i] = 1.0; }
        for (int i = 0; i < n; ++i) { X[i] = b[i]; c[i] = g[i]; }
        gradproj = 0.0;
        for (int i = 0; i < n; ++i) {
            s = 0.0;
            for (int j = 0; j <= i; ++j) s -= B[i][j] * c[j];
            for (int j = i + 1; j < n; ++j) s -= B[j][i] * c[j];
            t[i] = s;
            gradproj += s * c[i];
        }
        if (gradproj < 0.0) {                          /* search direction is downhill */
            int accpoint = 0;
            steplength = 1.0;
            do {
                count = 0;
                for (int i = 0; i < n; ++i) {
                    b[i] = X[i] + steplength * t[i];
                    if (reltest + X[i] == reltest + b[i]) count++;
                }
                if (count < n) {
                    f = sv_fminfn(o, b);
                    funcount++;
                    accpoint = isfinite(f) && (f <= *Fmin + gradproj * steplength * acctol);
                    if (!accpoint) steplength *= stepredn;
                }
            } while (!(count == n || accpoint));
            const int enough = fabs(f - *Fmin) > reltol * (fabs(*Fmin) + reltol);   /* f > abstol = -Inf */
            if (!enough) { count = n; *Fmin = f; }
            if (count < n) {                           /* making progress */
                *Fmin = f;
                sv_fmingr(o, b, g);
                gradcount++;
                iter++;
                D1 = 0.0;
                for (int i = 0; i < n; ++i) { t[i] = steplength * t[i]; c[i] = g[i] - c[i]; D1 += t[i] * c[i]; }
                if (D1 > 0) {
                    D2 = 0.0;
                    for (int i = 0; i < n; ++i) {
                        s = 0.0;
                        for (int j = 0; j <= i; ++j) s += B[i][j] * c[j];
                        for (int j = i + 1; j < n; ++j) s += B[j][i] * c[j];
                        X[i] = s;
                        D2 += s * c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (int i = 0; i < n; ++i)
                        for (int j = 0; j <= i; ++j) B[i][j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
                } else {
                    ilast = gradcount;
                }
            } else {                                   /* no progress */
                if (ilast < gradcount) { count = 0; ilast = gradcount; }
            }
        } else {                                       /* uphill search */
            count = 0;
            if (ilast == gradcount) count = n; else ilast = gradcount;
        }
        if (iter >= maxit) break;
        if (gradcount - ilast > 2 * n) ilast = gradcount;   /* periodic restart */
    } while (count != n || ilast != gradcount);
    *fncount = funcount; *grcount = gradcount;
    return iter < maxit ? 0 : 1;
}

/* optimhess(): central differences of the (scaled) gradient, then symmetrised */
static void sv_optimhess(const sv_obj* o, const double* par, double* hess) {
    const int n = o->d->q;
    const double ps = o->ct->parscale, eps = o->ct->ndeps / ps;
    double dpar[JMB_SVFT_MAX_Q], df1[JMB_SVFT_MAX_Q], df2[JMB_SVFT_MAX_Q];
    for (int i = 0; i < n; ++i) dpar[i] = par[i] / ps;
    for (int i = 0; i < n; ++i) {
        dpar[i] = dpar[i] + eps;
        sv_fmingr(o, dpar, df1);
        dpar[i] = dpar[i] - 2 * eps;
        sv_fmingr(o, dpar, df2);
        for (int j = 0; j < n; ++j) hess[i * n + j] = (df1[j] - df2[j]) / (2 * eps * ps * ps);
        dpar[i] = dpar[i] + eps;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            const double tmp = 0.5 * (hess[i * n + j] + hess[j * n + i]);
            hess[i * n + j] = hess[j * n + i] = tmp;
        }
}

/* sort ascending (insertion sort; M is a few hundred) */
static void sv_sort(double* x, int n) {
    for (int i = 1; i < n; ++i) {
        double v = x[i]; int j = i - 1;
        while (j >= 0 && x[j] > v) { x[j + 1] = x[j]; --j; }
        x[j + 1] = v;
    }
}

/* Mean / Median / quantile(type = 7) with na.rm = TRUE (R/survfitJM.mvJMbayes.R:406-412) */
void orc_sv_summary(const double* x, int M, const double* probs, double* out4) {
    double* s = (double*)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    int cnt = 0; double sum = 0.0;
    for (int m = 0; m < M; ++m) if (!isnan(x[m])) { s[cnt++] = x[m]; sum += x[m]; }
    if (cnt == 0) { out4[0] = out4[1] = out4[2] = out4[3] = NAN; free(s); return; }
    sv_sort(s, cnt);
    out4[0] = sum / cnt;
    const int half = (cnt + 1) / 2;                    /* median.default */
    out4[1] = (cnt % 2 == 1) ? s[half - 1] : (s[half - 1] + s[half]) / 2.0;
    for (int k = 0; k < 2; ++k) {                      /* quantile.default, type 7 */
        const double index = 1.0 + (cnt - 1) * probs[k];
        const double lo = floor(index), hi = ceil(index);
        double qs = s[(int)lo - 1];
        const double xh = s[(int)hi - 1];
        if (index > lo && xh != qs) { const double h = index - lo; qs = (1.0 - h) * qs + h * xh; }
        out4[2 + k] = qs;
    }
    free(s);
}

/* ---- the survfitJM loops ----------------------------------------------------------------------- */
int orc_svft_modes_range(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_control* ct, jmb_svft_out* out,
                         int i_begin, int i_end) {
    if (d->q > JMB_SVFT_MAX_Q) return 3;
    sv_index ix; memset(&ix, 0, sizeof(ix));
    int rc = sv_build_index(d, &ix);
    if (rc) return rc;
    sv_par* par = (sv_par*)malloc(sizeof(sv_par));
    sv_load_par(d, pm, 0, par);
    const int q = d->q;
    for (int i = i_begin; i < i_end; ++i) {
        sv_obj o = {d, &ix, par, i, ct};
        double b[JMB_SVFT_MAX_Q], Fmin = 0.0, hess[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
        int fn = 0, gr = 0;
        for (int j = 0; j < q; ++j) b[j] = 0.0 / ct->parscale;            /* start = rep(0, q) */
        int conv = sv_vmmin(&o, b, &Fmin, &fn, &gr);
        for (int j = 0; j < q; ++j) b[j] = b[j] * ct->parscale;
        if (conv == 10) { for (int j = 0; j < q; ++j) b[j] = NAN; for (int j = 0; j < q * q; ++j) hess[j] = NAN; }
        else sv_optimhess(&o, b, hess);
        for (int j = 0; j < q; ++j) out->modes[i + (size_t)j * d->n] = b[j];
        for (int j = 0; j < q * q; ++j) out->hessians[(size_t)i * q * q + j] = hess[j];
        if (out->convergence) out->convergence[i] = conv;
        if (out->counts) { out->counts[2 * i] = fn; out->counts[2 * i + 1] = gr; }
    }
    free(par);
    sv_free_index(d, &ix);
    return 0;
}

int orc_svft_sample_range(const jmb_svft_data* d, const jmb_svft_params* pd, const jmb_svft_control* ct, jmb_svft_out* out,
                          int i_begin, int i_end) {
    if (d->q > JMB_SVFT_MAX_Q) return 3;
    sv_index ix; memset(&ix, 0, sizeof(ix));
    int rc = sv_build_index(d, &ix);
    if (rc) return rc;
    const int q = d->q, M = pd->M, L = d->L, K = d->K, n = d->n;
    const int64_t nH = (int64_t)n * L * K;
    sv_par* pars = (sv_par*)malloc(sizeof(sv_par) * (size_t)M);
    for (int m = 0; m < M; ++m) sv_load_par(d, pd, m, pars + m);
    double* S = (double*)malloc(sizeof(double) * (size_t)L * M);
    const uint32_t seed = ct->seed;
    for (int i = i_begin; i < i_end; ++i) {
        const uint32_t gid = (uint32_t)(d->subject_offset + i);
        double mode[JMB_SVFT_MAX_Q], Hs[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q], invV[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
        double V[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q], ev[JMB_SVFT_MAX_Q], evec[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
        double Af[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q], invS[JMB_SVFT_MAX_Q * JMB_SVFT_MAX_Q];
        for (int j = 0; j < q; ++j) mode[j] = out->modes[i + (size_t)j * n];
        for (int j = 0; j < q * q; ++j) Hs[j] = out->hessians[(size_t)i * q * q + j];
        int bad = 0;
        for (int j = 0; j < q; ++j) if (!isfinite(mode[j])) bad = 1;
        for (int j = 0; j < q * q; ++j) { if (!isfinite(Hs[j])) bad = 1; invV[j] = Hs[j] / ct->scale; }   /* invVars.b, :304 */
        if (!bad && orc_sv_inverse(Hs, q, V)) bad = 1;                                                 /* Vars.b, :305 */
        if (!bad) {
            for (int j = 0; j < q * q; ++j) V[j] *= ct->scale;
            orc_sv_eigen_sym(V, q, ev, evec);
            if (!(ev[q - 1] >= -1e-06 * fabs(ev[0]))) bad = 1;        /* dmvt: "'Sigma' is not positive definite" */
            for (int j = 0; j < q; ++j) if (!(ev[j] > 0.0)) bad = 1;  /* a zero eigenvalue makes invSigma infinite */
        }
        const int nh = d->n_horizons[i];
        if (bad) {
            for (int l = 0; l < L; ++l) {
                for (int k = 0; k < 4; ++k) out->summaries[(size_t)i * L * 4 + l + (size_t)k * L] = NAN;
                if (out->full) for (int m = 0; m < M; ++m) out->full[((size_t)i * M + m) * L + l] = NAN;
            }
            if (out->accept_rate) out->accept_rate[i] = NAN;
            if (out->b_last) for (int j = 0; j < q; ++j) out->b_last[i + (size_t)j * n] = NAN;
            if (out->convergence) out->convergence[i] = 20;
            continue;
        }
        for (int r = 0; r < q; ++r)
            for (int c = 0; c < q; ++c) {
                Af[r + c * q] = evec[r + c * q] * sqrt(ev[c] > 0.0 ? ev[c] : 0.0);   /* rmvt: evec * rep(sqrt(pmax(ev, 0)), each = p) */
                double s = 0.0;
                for (int k = 0; k < q; ++k) s += evec[r + k * q] * (evec[c + k * q] / ev[k]);   /* dmvt: evec %*% (t(evec) / ev) */
                invS[r + c * q] = s;
            }
        double b_old[JMB_SVFT_MAX_Q], b_new[JMB_SVFT_MAX_Q];
        for (int j = 0; j < q; ++j) b_old[j] = b_new[j] = mode[j];
        int n_acc = 0;
        for (int m = 0; m < M; ++m) {
            const sv_par* par = pars + m;
            /* proposal m of rmvt(M, mode, Vars.b, df) and its dmvt (:330-337) */
            double z[JMB_SVFT_MAX_Q + 1], pb[JMB_SVFT_MAX_Q];
            for (int s = 0; s < (q + 1) / 2; ++s) orc_rnorm_pair(seed, SV_KEY, gid, (uint32_t)m, (uint32_t)s, SV_ST_NORM, z + 2 * s);
            const double w = orc_rgamma(seed, SV_KEY, gid, (uint32_t)m, 0.5 * ct->df, 2.0);      /* rchisq(df) */
            const double den = sqrt(w / ct->df);
            for (int r = 0; r < q; ++r) {
                double s = 0.0;
                for (int c = 0; c < q; ++c) s += Af[r + c * q] * z[c];
                pb[r] = mode[r] + s / den;
            }
            double quad_p = 0.0, quad_o = 0.0;
            for (int c = 0; c < q; ++c) {
                double tp = 0.0, to = 0.0;
                for (int r = 0; r < q; ++r) { tp += (pb[r] - mode[r]) * invS[r + c * q]; to += (b_old[r] - mode[r]) * invV[r + c * q]; }
                quad_p += tp * (pb[c] - mode[c]); quad_o += to * (b_old[c] - mode[c]);
            }
            const double dmvt_prop = -0.5 * (ct->df + q) * log(1.0 + quad_p / ct->df);
            const double dmvt_old = -0.5 * (ct->df + q) * log(1.0 + quad_o / ct->df);
            /* MH step (:372-377) */
            const double lp_p = sv_log_post(d, &ix, par, i, pb), lp_o = sv_log_post(d, &ix, par, i, b_old);
            const double ea = exp(lp_p + dmvt_old - lp_o - dmvt_prop);
            uint32_t rr[4];
            orc_philox4x32(seed, SV_KEY, gid, (uint32_t)m, 0u, SV_ST_UNIF, rr);
            const double u = orc_u01(rr[0], rr[1]);
            if (!isnan(ea)) { const double a = ea < 1.0 ? ea : 1.0; if (u <= a) { for (int j = 0; j < q; ++j) b_new[j] = pb[j]; n_acc++; } }
            /* survival on each prediction interval (:378-395) */
            double cum = 0.0;
            for (int l = 0; l < L; ++l) {
                double v = NAN;
                if (l < nh) {
                    const double logS = sv_minus_H(d, par, b_new, ((int64_t)i * L + l) * K, nH, d->W1s_h, d->W2s_h, d->Pw_h,
                                                   d->XXs_h, d->ZZs_h, d->Us_h);
                    cum += logS;
                    v = ct->log ? logS : exp(cum);
                }
                S[l * M + m] = v;
                if (out->full) out->full[((size_t)i * M + m) * L + l] = v;
            }
            for (int j = 0; j < q; ++j) b_old[j] = b_new[j];
        }
        for (int l = 0; l < L; ++l) {
            double s4[4] = {NAN, NAN, NAN, NAN};
            if (l < nh) orc_sv_summary(S + (size_t)l * M, M, ct->CI_levels, s4);
            for (int k = 0; k < 4; ++k) out->summaries[(size_t)i * L * 4 + l + (size_t)k * L] = s4[k];
        }
        if (out->accept_rate) out->accept_rate[i] = (double)n_acc / (M > 0 ? M : 1);
        if (out->b_last) for (int j = 0; j < q; ++j) out->b_last[i + (size_t)j * n] = b_new[j];
    }
    free(S); free(pars);
    sv_free_index(d, &ix);
    return 0;
}

int orc_svft_modes(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_control* ct, jmb_svft_out* out) {
    return orc_svft_modes_range(d, pm, ct, out, 0, d->n);
}
int orc_svft_sample(const jmb_svft_data* d, const jmb_svft_params* pd, const jmb_svft_control* ct, jmb_svft_out* out) {
    return orc_svft_sample_range(d, pd, ct, out, 0, d->n);
}
/* subjects [i_begin, i_end) only: lets the CPU baseline run one worker thread per block of subjects */
int orc_survfitJM_range(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_params* pd,
                        const jmb_svft_control* ct, jmb_svft_out* out, int i_begin, int i_end) {
    int rc = orc_svft_modes_range(d, pm, ct, out, i_begin, i_end);
    if (rc) return rc;
    return orc_svft_sample_range(d, pd, ct, out, i_begin, i_end);
}
int orc_survfitJM(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_params* pd, const jmb_svft_control* ct,
                  jmb_svft_out* out) {
    return orc_survfitJM_range(d, pm, pd, ct, out, 0, d->n);
}

/* evaluators exposed for the tests: log_post_RE_svft / survPred_svft_2 of subject i under parameter set m */
double orc_svft_log_post(const jmb_svft_data* d, const jmb_svft_params* P, int m, int i, const double* b) {
    sv_index ix; memset(&ix, 0, sizeof(ix));
    if (sv_build_index(d, &ix)) return NAN;
    sv_par* par = (sv_par*)malloc(sizeof(sv_par));
    sv_load_par(d, P, m, par);
    const double v = sv_log_post(d, &ix, par, i, b);
    free(par); sv_free_index(d, &ix);
    return v;
}
double orc_svft_surv_pred(const jmb_svft_data* d, const jmb_svft_params* P, int m, int i, int l, const double* b) {
    sv_par* par = (sv_par*)malloc(sizeof(sv_par));
    sv_load_par(d, P, m, par);
    const int64_t nH = (int64_t)d->n * d->L * d->K;
    const double v = sv_minus_H(d, par, b, ((int64_t)i * d->L + l) * d->K, nH, d->W1s_h, d->W2s_h, d->Pw_h, d->XXs_h, d->ZZs_h, d->Us_h);
    free(par);
    return v;
}
