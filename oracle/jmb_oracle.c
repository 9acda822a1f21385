/*
 * jmb_oracle.c - CPU restatement of the reference's mvJointModelBayes hot path (plain C11).
 *
 * TEST INFRASTRUCTURE ONLY (see jmb_oracle.h).  Pinned against the reference's own sources
 * compiled in oracle/_ref against a stand-in Armadillo/Rcpp header (no reference golden vectors
 * exist; R/Rcpp/Armadillo are absent), plus an independent NumPy twin and analytic self-checks.
 *
 * Every function cites the reference lines it follows (paths relative to the JMbayes tree).
 * The algorithm and the data layout are the reference's: dense column-major W1/W1s/W2s,
 * materialised Wlong/Wlongs, the baseline part W1s*Bs_gammas + W2s*gammas recomputed in both the
 * b-step and the survival step, cumsum-difference rowsum, naive link formulas.  The only
 * deliberate deviation is the random stream: the reference materialises R's RNG up front
 * (src/fast_LAPRWM.cpp:617-618,658-661), which needs 10 GB at n = 1M and cannot be reproduced
 * without R; here the draws come from the counter generator specified in DESIGN.md (Philox4x32-10
 * keyed by seed/chain/iteration/subject), the same specification the CUDA path implements, so
 * that chains can be compared draw for draw.
 */
#include "jmb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define JMB_ORC_QMAX_PAD 18
#define JMB_ORC_PMAX 130

static int g_rowsum_mode = 0;
void orc_set_rowsum_mode(int mode) { g_rowsum_mode = mode; }

static orc_allreduce_fn g_allreduce = 0;
static void* g_allreduce_ctx = 0;
void orc_set_allreduce(orc_allreduce_fn fn, void* ctx) { g_allreduce = fn; g_allreduce_ctx = ctx; }

/* ------------------------------------------------------------------------------------------
 * Counter RNG (DESIGN.md "Random streams").  Philox4x32-10, Salmon et al. SC'11.
 * ---------------------------------------------------------------------------------------- */
void orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53 random bits -> (0,1) */
double orc_u01(uint32_t hi, uint32_t lo) {
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

#define ORC_PI 3.14159265358979323846264338327950288

/* sin(pi x), cos(pi x) for x in [0, 2] with exact range reduction (the CUDA path uses sincospi) */
static void orc_sincospi(double x, double* s, double* c) {
    double qd = rint(2.0 * x);            /* nearest multiple of 1/2 */
    double t = x - 0.5 * qd;              /* exact, |t| <= 1/4 */
    double st = sin(ORC_PI * t), ct = cos(ORC_PI * t);
    switch (((int)qd) & 3) {
        case 0: *s = st; *c = ct; break;
        case 1: *s = ct; *c = -st; break;
        case 2: *s = -st; *c = -ct; break;
        default: *s = -ct; *c = st; break;
    }
}

/* Box-Muller pair from one Philox block */
void orc_rnorm_pair(uint32_t seed, uint32_t chain, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, double z[2]) {
    uint32_t r[4];
    orc_philox4x32(seed, chain, c0, c1, c2, c3, r);
    double u1 = orc_u01(r[0], r[1]), u2 = orc_u01(r[2], r[3]);
    double rad = sqrt(-2.0 * log(u1)), sn, cs;
    orc_sincospi(2.0 * u2, &sn, &cs);
    z[0] = rad * cs;
    z[1] = rad * sn;
}

enum { ST_B_PROP = 0, ST_B_ACC = 1, ST_SURV = 2, ST_GIBBS_N = 3, ST_GIBBS_U = 4, ST_GIBBS_B = 5 };

/* Gamma(shape, scale) by Marsaglia & Tsang (2000); stands in for Rf_rgamma
 * (src/fast_LAPRWM.cpp:781) - same distribution, different stream. */
double orc_rgamma(uint32_t seed, uint32_t chain, uint32_t draw, uint32_t iter, double shape,
                  double scale) {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd), x = 0.0;
    for (uint32_t attempt = 0; attempt < 1000u; ++attempt) {
        double z[2];
        orc_rnorm_pair(seed, chain, draw, iter, attempt, ST_GIBBS_N, z);
        double t = 1.0 + cc * z[0];
        if (t <= 0.0) continue;
        double v = t * t * t;
        uint32_t r[4];
        orc_philox4x32(seed, chain, draw, iter, attempt, ST_GIBBS_U, r);
        double u = orc_u01(r[0], r[1]);
        if (log(u) < 0.5 * z[0] * z[0] + dd - dd * v + dd * log(v)) { x = dd * v; break; }
    }
    if (shape < 1.0) {
        uint32_t r[4];
        orc_philox4x32(seed, chain, draw, iter, 0u, ST_GIBBS_B, r);
        x *= pow(orc_u01(r[0], r[1]), 1.0 / shape);
    }
    return x * scale;
}

/* ------------------------------------------------------------------------------------------
 * rowsum: src/fast_LAPRWM.cpp:18-25.  out[i] = cs[last[i]] - cs[last[i-1]], cs = cumsum(x).
 * ---------------------------------------------------------------------------------------- */
void orc_rowsum(const double* x, int64_t N, const int32_t* last, int32_t n, double* out) {
    if (g_rowsum_mode == 0) {
        double cs = 0.0, prev = 0.0;
        int64_t r = 0;
        for (int32_t i = 0; i < n; ++i) {
            for (; r <= last[i] && r < N; ++r) cs += x[r];
            out[i] = cs - prev;
            prev = cs;
        }
    } else {
        int64_t r = 0;
        for (int32_t i = 0; i < n; ++i) {
            long double s = 0.0L;
            for (; r <= last[i] && r < N; ++r) s += x[r];
            out[i] = (double)s;
        }
    }
}

/* y += A x for column-major A (rows x cols) */
static void gemv_acc(const double* A, int64_t rows, int32_t cols, const double* x, double* y) {
    for (int32_t j = 0; j < cols; ++j) {
        const double* col = A + (int64_t)j * rows;
        double xj = x[j];
        for (int64_t r = 0; r < rows; ++r) y[r] += col[r] * xj;
    }
}

/* Rf_pnorm5(x, 0, 1, lower, !log): src/fast_LAPRWM.cpp:27-34 */
static double pnorm_std(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

/* lin_predF: src/fast_LAPRWM.cpp:67-78. eta_k[r] = Xbetas_k[r] + sum_j Z_k[r,j] b[idL_k[r], RE_inds_k[j]] */
static void lin_predF(const orc_data* d, const orc_draw* w, const double* b, int k, double* eta) {
    int64_t N = d->N[k];
    int32_t n = d->n, qk = d->q_k[k];
    const double* Z = d->Z[k];
    const int32_t* id = d->idL[k];
    for (int64_t r = 0; r < N; ++r) {
        double zb = 0.0;
        for (int32_t j = 0; j < qk; ++j) zb += Z[r + (int64_t)j * N] * b[id[r] + (int64_t)d->RE_inds[k][j] * n];
        eta[r] = w->Xbetas[k][r] + zb;
    }
}

/* log_longF: src/fast_LAPRWM.cpp:168-202 (constants dropped exactly as there) */
static void log_longF(const orc_data* d, const orc_draw* w, double* const* eta, double* out,
                      double* tmp_n) {
    int32_t n = d->n;
    for (int32_t i = 0; i < n; ++i) out[i] = 0.0;
    for (int32_t k = 0; k < d->n_outcomes; ++k) {
        int64_t N = d->N[k];
        const double* y = d->y[k];
        double* e = eta[k];               /* overwritten by the log-density */
        int fam = d->fams[k], link = d->links[k];
        if (fam == 0) {                   /* gaussian :177-180 */
            double s = w->sigmas[k];
            for (int64_t r = 0; r < N; ++r) { double t = (y[r] - e[r]) / s; e[r] = -0.5 * (t * t); }
        } else if (fam == 1) {            /* binomial :181-194 */
            for (int64_t r = 0; r < N; ++r) {
                double pr;
                if (link == 1) { double ee = exp(e[r]); pr = ee / (1.0 + ee); }
                else if (link == 2) pr = pnorm_std(e[r]);
                else pr = -exp(-exp(e[r])) + 1.0;
                e[r] = y[r] * log(pr) + (1.0 - y[r]) * log(1.0 - pr);
            }
        } else {                          /* poisson :195-199 */
            for (int64_t r = 0; r < N; ++r) { double mu = exp(e[r]); e[r] = y[r] * log(mu) - mu; }
        }
        orc_rowsum(e, N, d->idL2[k], n, tmp_n);
        for (int32_t i = 0; i < n; ++i) out[i] += tmp_n[i];
    }
}

/* log p(b): src/fast_LAPRWM.cpp:621,673.  -0.5 * rowSums((b invD) % b) */
static void log_pb_all(const orc_data* d, const orc_draw* w, const double* b, double* out) {
    int32_t n = d->n, q = d->q;
    for (int32_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t j = 0; j < q; ++j) {
            double t = 0.0;
            for (int32_t l = 0; l < q; ++l) t += b[i + (int64_t)l * n] * w->invD[l + j * q];
            s += t * b[i + (int64_t)j * n];
        }
        out[i] = -0.5 * s;
    }
}

static double trans_fun(int tf, double x) {
    switch (tf) {                         /* src/fast_LAPRWM.cpp:91-106 */
        case 0: return x;
        case 1: { double e = exp(x); return e / (1.0 + e); }
        case 2: return exp(x);
        case 3: return log(x);
        case 4: return log2(x);
        case 5: return log10(x);
        default: return sqrt(x);
    }
}

/* rows of the survival part: n, or the transition rows of a multi-state fit (src/fast_LAPRWM.cpp:519-529) */
static int32_t surv_rows(const orc_data* d) { return d->multi_state ? d->nT : d->n; }

/* idT_rsum (R/mvJointModelBayes.R:567-568): 0-based last transition row of each subject */
static int32_t* make_idT_rsum(const orc_data* d) {
    int32_t* last = (int32_t*)malloc(sizeof(int32_t) * d->n);
    if (!d->multi_state) { for (int32_t i = 0; i < d->n; ++i) last[i] = i; return last; }
    int32_t m = 0;
    for (int32_t r = 0; r < d->nT; ++r)
        if (r == d->nT - 1 || d->idT[r] != d->idT[r + 1]) { if (m < d->n) last[m] = r; ++m; }
    for (; m < d->n; ++m) last[m] = d->nT - 1;
    return last;
}

/* lin_pred_matF: src/fast_LAPRWM.cpp:80-109 */
int orc_wlong(const orc_data* d, const orc_draw* w, const double* b, int which, double* out) {
    int32_t n = d->n, K = d->K, nT = surv_rows(d);
    int64_t rows = which == 0 ? nT : (int64_t)nT * K;
    memset(out, 0, sizeof(double) * rows * d->n_alphas);
    for (int32_t t = 0; t < d->n_terms; ++t) {
        const double* xb = which == 0 ? w->XXbetas[t] : which == 1 ? w->XXsbetas[t] : w->XXsbetas_int[t];
        const double* ZZ = which == 0 ? d->ZZ[t] : which == 1 ? d->ZZs[t] : d->ZZs_int[t];
        const double* U = which == 0 ? d->U[t] : which == 1 ? d->Us[t] : d->Us_int[t];
        int32_t rt = d->r_t[t], ut = d->u_t[t];
        for (int64_t r = 0; r < rows; ++r) {
            int64_t id = which == 0 ? r : r / K;   /* idT2 / idT2s: :622-630 */
            if (d->multi_state) id = d->idT[id];   /* idT.list / idTs.list, R/mvJointModelBayes.R:560-564 */
            double zb = 0.0;
            for (int32_t j = 0; j < rt; ++j) zb += ZZ[r + (int64_t)j * rows] * b[id + (int64_t)d->RE_inds2[t][j] * n];
            double v = trans_fun(d->trans_Funs[t], xb[r] + zb);
            for (int32_t u = 0; u < ut; ++u)
                out[r + (int64_t)d->col_inds[t][u] * rows] = U[r + (int64_t)u * rows] * v;
        }
    }
    return 0;
}

/* log_h (:632,684,735) and H / HL (:633-639,685-691,736-741) for given Wlong/Wlongs and params */
static void hazard_parts(const orc_data* d, const double* Bs, const double* gam, const double* alp,
                         const double* Wlong, const double* Wlongs, const double* Wlongs_int,
                         double* log_h, double* H, double* HL, double* tmp_ns) {
    int32_t n = surv_rows(d), K = d->K;       /* one log_h / H per transition row in a multi-state fit */
    int64_t ns = (int64_t)n * K;
    for (int32_t i = 0; i < n; ++i) log_h[i] = 0.0;
    gemv_acc(d->W1, n, d->n_Bs, Bs, log_h);
    {   /* (W1 Bs + W2 gammas) + Wlong alphas, evaluated left to right like the expression */
        double* t2 = tmp_ns;                 /* n <= ns */
        for (int32_t i = 0; i < n; ++i) t2[i] = 0.0;
        gemv_acc(d->W2, n, d->n_gammas, gam, t2);
        for (int32_t i = 0; i < n; ++i) log_h[i] += t2[i];
        for (int32_t i = 0; i < n; ++i) t2[i] = 0.0;
        gemv_acc(Wlong, n, d->n_alphas, alp, t2);
        for (int32_t i = 0; i < n; ++i) log_h[i] += t2[i];
    }
    for (int pass = 0; pass < (d->interval_cens ? 2 : 1); ++pass) {
        const double* W1s = pass ? d->W1s_int : d->W1s;
        const double* W2s = pass ? d->W2s_int : d->W2s;
        const double* Wl = pass ? Wlongs_int : Wlongs;
        const double* Pw = pass ? d->Pw_int : d->Pw;
        double* lin = tmp_ns;
        double* part = tmp_ns + ns;
        for (int64_t r = 0; r < ns; ++r) lin[r] = 0.0;
        gemv_acc(W1s, ns, d->n_Bs, Bs, lin);
        for (int64_t r = 0; r < ns; ++r) part[r] = 0.0;
        gemv_acc(W2s, ns, d->n_gammas, gam, part);
        for (int64_t r = 0; r < ns; ++r) lin[r] += part[r];
        for (int64_t r = 0; r < ns; ++r) part[r] = 0.0;
        gemv_acc(Wl, ns, d->n_alphas, alp, part);
        for (int64_t r = 0; r < ns; ++r) lin[r] = Pw[r] * exp(lin[r] + part[r]);
        orc_rowsum(lin, ns, d->idGK_fast, n, pass ? HL : H);
    }
    if (!d->interval_cens && HL) for (int32_t i = 0; i < n; ++i) HL[i] = H[i];   /* :635 */
}

/* event / censoring term: src/fast_LAPRWM.cpp:640-650 (same at :692-702, :742-752) */
static void log_ptb_all(const orc_data* d, const double* log_h, const double* H, const double* HL,
                        double* out) {
    int32_t n = surv_rows(d);
    if (d->interval_cens) {
        for (int32_t i = 0; i < n; ++i) {
            double e = d->event[i], r = 0.0;       /* Levent* flags, R/mvJointModelBayes.R:694-696 */
            if (e == 1.0) r += log_h[i];
            if (e == 1.0 || e == 0.0) r -= H[i];
            if (e == 2.0) r += log(1.0 - exp(-H[i]));
            if (e == 3.0) r += log(exp(-HL[i]) - exp(-H[i]));
            out[i] = r;
        }
    } else {
        for (int32_t i = 0; i < n; ++i) out[i] = d->event[i] * log_h[i] - H[i];
    }
}

int orc_eval_logpost_re(const orc_data* d, const orc_draw* w, const double* b, const double* Bs,
                        const double* gam, const double* alp, orc_eval_out* out) {
    int32_t n = d->n, K = d->K, A = d->n_alphas, nT = surv_rows(d);
    int64_t ns = (int64_t)nT * K;
    double** eta = (double**)malloc(sizeof(double*) * d->n_outcomes);
    for (int k = 0; k < d->n_outcomes; ++k) {
        eta[k] = (double*)malloc(sizeof(double) * d->N[k]);
        lin_predF(d, w, b, k, eta[k]);
    }
    double* tmp_n = (double*)malloc(sizeof(double) * n);
    log_longF(d, w, eta, out->log_pyb, tmp_n);
    log_pb_all(d, w, b, out->log_pb);
    double* Wlong = (double*)malloc(sizeof(double) * nT * A);
    double* Wlongs = (double*)malloc(sizeof(double) * ns * A);
    double* Wlongs_int = d->interval_cens ? (double*)malloc(sizeof(double) * ns * A) : 0;
    double* tmp_ns = (double*)malloc(sizeof(double) * 2 * ns);
    double* HLbuf = out->HL ? out->HL : (double*)malloc(sizeof(double) * nT);
    orc_wlong(d, w, b, 0, Wlong);
    orc_wlong(d, w, b, 1, Wlongs);
    if (d->interval_cens) orc_wlong(d, w, b, 2, Wlongs_int);
    hazard_parts(d, Bs, gam, alp, Wlong, Wlongs, Wlongs_int, out->log_h, out->H, HLbuf, tmp_ns);
    log_ptb_all(d, out->log_h, out->H, HLbuf, out->log_ptb);
    if (!out->HL) free(HLbuf);
    for (int k = 0; k < d->n_outcomes; ++k) free(eta[k]);
    free(eta); free(tmp_n); free(Wlong); free(Wlongs); free(Wlongs_int); free(tmp_ns);
    return 0;
}

/* logPrior: src/fast_LAPRWM.cpp:371-374 */
static double logPrior(const double* x, const double* mean, const double* Tau, int32_t m, double tau) {
    double s = 0.0;
    for (int32_t j = 0; j < m; ++j) {
        double t = 0.0;
        for (int32_t l = 0; l < m; ++l) t += (x[l] - mean[l]) * Tau[l + j * m];
        s += t * (x[j] - mean[j]);
    }
    return -0.5 * tau * s;
}

static double quad_form(const double* x, const double* Tau, int32_t m) {
    double s = 0.0;
    for (int32_t j = 0; j < m; ++j) {
        double t = 0.0;
        for (int32_t l = 0; l < m; ++l) t += x[l] * Tau[l + j * m];
        s += t * x[j];
    }
    return s;
}

/* upper Cholesky factor H (H'H = S) of a column-major m x m matrix; returns 0 on success */
static int chol_upper(const double* S, int32_t m, double* H) {
    memset(H, 0, sizeof(double) * m * m);
    for (int32_t j = 0; j < m; ++j) {
        for (int32_t i = 0; i <= j; ++i) {
            double s = S[i + j * m];
            for (int32_t k = 0; k < i; ++k) s -= H[k + i * m] * H[k + j * m];
            if (i == j) { if (!(s > 0.0)) return 1; H[i + j * m] = sqrt(s); }
            else H[i + j * m] = s / H[i + i * m];
        }
    }
    return 0;
}

/* bounds_sigma: src/fast_LAPRWM.cpp:411-419 */
static double bounds_sigma(double s, double eps1, double eps3) {
    if (s < eps1) s = eps1;
    if (s > eps3) s = eps3;
    return s;
}

/* adaptCov update: Sigma <- bounds_Cov(Sigma + g1 (cov(block') - Sigma)), :421-428, :863-867 */
static void adapt_cov(double* Sigma, int32_t m, const double* block, int32_t nb, double g1,
                      double eps2, double eps3) {
    if (m == 0) return;
    double* mu = (double*)calloc(m, sizeof(double));
    double* C = (double*)calloc((size_t)m * m, sizeof(double));
    double* Hc = (double*)malloc(sizeof(double) * m * m);
    for (int32_t i = 0; i < nb; ++i) for (int32_t j = 0; j < m; ++j) mu[j] += block[j + i * m];
    for (int32_t j = 0; j < m; ++j) mu[j] /= nb;
    for (int32_t a = 0; a < m; ++a) for (int32_t c = 0; c < m; ++c) {
        double s = 0.0;
        for (int32_t i = 0; i < nb; ++i) s += (block[a + i * m] - mu[a]) * (block[c + i * m] - mu[c]);
        C[a + c * m] = s / (nb - 1);
    }
    for (int32_t e = 0; e < m * m; ++e) Sigma[e] = Sigma[e] + g1 * (C[e] - Sigma[e]);
    double det = 0.0;
    if (chol_upper(Sigma, m, Hc) == 0) { det = 1.0; for (int32_t j = 0; j < m; ++j) det *= Hc[j + j * m] * Hc[j + j * m]; }
    if (det > eps3) for (int32_t e = 0; e < m * m; ++e) Sigma[e] *= eps3 / det;
    for (int32_t j = 0; j < m; ++j) Sigma[j + j * m] += eps2;
    free(mu); free(C); free(Hc);
}

static double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* rowsum(log_ptb, idT_rsum) (src/fast_LAPRWM.cpp:669,703,728,754): the subject's sum over its transition rows.
 * Without multiState idT_rsum is the identity; the restatement then takes the values themselves, as it always
 * has (the pin against the compiled reference, tests/test_reference_golden.py, covers that choice). */
static void subject_ptb(const orc_data* d, const double* ptb, const int32_t* last, double* out) {
    if (!d->multi_state) { memcpy(out, ptb, sizeof(double) * d->n); return; }
    orc_rowsum(ptb, d->nT, last, d->n, out);
}

/* ------------------------------------------------------------------------------------------
 * lap_rwm_C: src/fast_LAPRWM.cpp:431-896
 * ---------------------------------------------------------------------------------------- */
int orc_lap_rwm(const orc_data* d, const orc_draw* w, const orc_priors* pr, const orc_control* ct,
                const orc_chain_in* in, orc_chain_out* out, int max_iters) {
    const int32_t n = d->n, q = d->q, K = d->K, A = d->n_alphas, B = d->n_Bs, p = d->n_gammas;
    const int32_t nT = surv_rows(d);                 /* :519-529 */
    const int32_t nRisks = d->multi_state ? d->nRisks : 1;
    const int64_t ns = (int64_t)nT * K;
    const int ic = d->interval_cens;
    if (d->multi_state && (ic || nRisks < 1 || nRisks > 16)) return 4;
    int32_t* idT_rsum = make_idT_rsum(d);
    /* derived from control :584-588 */
    const int total_it = ct->n_iter + ct->n_burnin;
    const int its = (int)ceil((double)total_it / ct->n_block);
    const int n_out = (int)floor((double)ct->n_iter / ct->n_thin);
    const int start_cov_update = (int)floor((double)ct->n_burnin / ct->n_block) - 1;
    int n_keep = 0;
    for (int v = ct->n_burnin + 1; v <= total_it; v += ct->n_thin) ++n_keep;
    {   /* more hits of iter == keep[k] than output slots: the reference indexes out of bounds (:836-850) */
        int hits = 0;
        for (int v = ct->n_burnin + 1; v <= total_it && v <= its * ct->n_block - 1; v += ct->n_thin) ++hits;
        if (hits > n_out && max_iters <= 0) return 3;
    }
    const uint32_t seed = ct->seed, chain = ct->chain_id;
    const double mc1 = -ct->c1;

    /* state :507-518 */
    double* b = (double*)malloc(sizeof(double) * n * q);
    memcpy(b, in->b, sizeof(double) * n * q);
    double* theta = (double*)malloc(sizeof(double) * 2 * (B + p + A + 1));
    double *Bs = theta, *gam = Bs + B, *alp = gam + p;
    double *nBs = alp + A + 1, *ngam = nBs + B, *nalp = ngam + p;
    memcpy(Bs, in->Bs_gammas, sizeof(double) * B);
    if (p) memcpy(gam, in->gammas, sizeof(double) * p);
    memcpy(alp, in->alphas, sizeof(double) * A);
    double tau_Bs[16];
    for (int j = 0; j < nRisks; ++j) tau_Bs[j] = in->tau_Bs_gammas[j];
    double tau_g = in->tau_gammas, tau_a = in->tau_alphas;
    double xi_a = in->tau_alphas;                                  /* :518 (quirk 2) */
    double* phi_g = (double*)malloc(sizeof(double) * (p + 1));
    double* phi_a = (double*)malloc(sizeof(double) * (A + 1));
    double* nu_a = (double*)malloc(sizeof(double) * (A + 1));
    for (int j = 0; j < p; ++j) phi_g[j] = in->phi_gammas[j];
    for (int j = 0; j < A; ++j) { phi_a[j] = in->phi_alphas[j]; nu_a[j] = in->phi_alphas[j]; }  /* :513 */
    int n_td = pr->n_td;
    double* tau_td = (double*)malloc(sizeof(double) * (n_td + 1));
    for (int j = 0; j < n_td; ++j) tau_td[j] = in->tau_td_alphas[j];
    double* Tau_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* Tau_a = (double*)malloc(sizeof(double) * (A * A + 1));
    double* Tau_a_pen = (double*)malloc(sizeof(double) * (A * A + 1));
    memcpy(Tau_g, pr->Tau_gammas, sizeof(double) * p * p);
    memcpy(Tau_a, pr->Tau_alphas, sizeof(double) * A * A);
    memcpy(Tau_a_pen, pr->Tau_alphas, sizeof(double) * A * A);   /* :570 */
    const double Apost_tau_Bs = pr->A_tau_Bs_gammas + 0.5 * pr->rank_Tau_Bs_gammas;    /* :539 */
    const double Apost_tau_g = pr->A_tau_gammas + 0.5 * pr->rank_Tau_gammas;
    const double Apost_phi_g = pr->A_phi_gammas + 0.5;
    const double Apost_tau_a = pr->A_tau_alphas + 0.5 * pr->rank_Tau_alphas;
    const double Apost_phi_a = pr->A_phi_alphas + 0.5;
    const double Apost_xi_a = pr->A_xi_alphas + 0.5, Apost_nu_a = pr->A_nu_alphas + 0.5;
    const double B_tau_td = pr->B_tau_Bs_gammas;                                        /* :568 */
    const double Apost_tau_td = pr->A_tau_Bs_gammas + 0.5 * pr->rank_Tau_td_alphas;     /* :569 */

    /* scales / covariances :594-601 */
    double* sigma_b = (double*)malloc(sizeof(double) * n);
    memcpy(sigma_b, in->sigma_b, sizeof(double) * n);
    double sig_Bs = in->sigma_Bs_gammas, sig_g = in->sigma_gammas, sig_a = in->sigma_alphas;
    double* S_Bs = (double*)malloc(sizeof(double) * (B * B + 1));
    double* S_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* S_a = (double*)malloc(sizeof(double) * (A * A + 1));
    memcpy(S_Bs, in->Sigma_Bs_gammas, sizeof(double) * B * B);
    if (p) memcpy(S_g, in->Sigma_gammas, sizeof(double) * p * p);
    memcpy(S_a, in->Sigma_alphas, sizeof(double) * A * A);
    double* H_Bs = (double*)malloc(sizeof(double) * (B * B + 1));
    double* H_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* H_a = (double*)malloc(sizeof(double) * (A * A + 1));

    /* work arrays */
    double** eta = (double**)malloc(sizeof(double*) * d->n_outcomes);
    for (int k = 0; k < d->n_outcomes; ++k) eta[k] = (double*)malloc(sizeof(double) * d->N[k]);
    double* tmp_n = (double*)malloc(sizeof(double) * n);
    double* tmp_ns = (double*)malloc(sizeof(double) * 2 * ns);
    double* new_b = (double*)malloc(sizeof(double) * n * q);
    double *c_pyb = (double*)malloc(sizeof(double) * n), *n_pyb = (double*)malloc(sizeof(double) * n);
    double *c_pb = (double*)malloc(sizeof(double) * n), *n_pb = (double*)malloc(sizeof(double) * n);
    double *c_Wl = (double*)malloc(sizeof(double) * nT * A), *n_Wl = (double*)malloc(sizeof(double) * nT * A);
    double *c_Wls = (double*)malloc(sizeof(double) * ns * A), *n_Wls = (double*)malloc(sizeof(double) * ns * A);
    double *c_Wli = ic ? (double*)malloc(sizeof(double) * ns * A) : 0;
    double *n_Wli = ic ? (double*)malloc(sizeof(double) * ns * A) : 0;
    double *c_lh = (double*)malloc(sizeof(double) * nT), *n_lh = (double*)malloc(sizeof(double) * nT);
    double *c_H = (double*)malloc(sizeof(double) * nT), *n_H = (double*)malloc(sizeof(double) * nT);
    double *c_HL = (double*)malloc(sizeof(double) * nT), *n_HL = (double*)malloc(sizeof(double) * nT);
    double *c_ptb = (double*)malloc(sizeof(double) * nT), *n_ptb = (double*)malloc(sizeof(double) * nT);
    double *c_sub = (double*)malloc(sizeof(double) * n), *n_sub = (double*)malloc(sizeof(double) * n);   /* per-subject sums */
    double* accept_b = (double*)calloc(n, sizeof(double));
    double* accept_b_tot = (double*)calloc(n, sizeof(double));
    double* block_par = (double*)malloc(sizeof(double) * (B + p + A) * ct->n_block);
    double accept_s_tot = 0.0;

    /* initial state :619-650 */
    for (int k = 0; k < d->n_outcomes; ++k) lin_predF(d, w, b, k, eta[k]);
    log_longF(d, w, eta, c_pyb, tmp_n);
    log_pb_all(d, w, b, c_pb);
    orc_wlong(d, w, b, 0, c_Wl);
    orc_wlong(d, w, b, 1, c_Wls);
    if (ic) orc_wlong(d, w, b, 2, c_Wli);
    hazard_parts(d, Bs, gam, alp, c_Wl, c_Wls, c_Wli, c_lh, c_H, c_HL, tmp_ns);
    log_ptb_all(d, c_lh, c_H, c_HL, c_ptb);

    int keep_iterator = 0, check_iterator = 0, done = 0, rc = 0;
    for (int it = 0; it < its && !done; ++it) {
        for (int32_t m = 0; m < n; ++m) accept_b[m] = 0.0;
        double accept_s = 0.0;
        /* mvrnorm(n_block, sigma * Sigma): chol(sigma Sigma) = sqrt(sigma) chol(Sigma) :659-661 */
        if (chol_upper(S_Bs, B, H_Bs) || (p && chol_upper(S_g, p, H_g)) || chol_upper(S_a, A, H_a)) { rc = 2; break; }
        int iters_in_block = 0;
        for (int i = 0; i < ct->n_block; ++i) {
            const uint32_t iter = (uint32_t)(it * ct->n_block + i);
            if (max_iters > 0 && (int)iter >= max_iters) { done = 1; break; }
            ++iters_in_block;
            /* ---- b-step :669-725 ---- */
            for (int32_t m = 0; m < n; ++m) {
                const uint32_t gid = (uint32_t)(d->subject_offset + m);
                const double* R = d->Covs_b + (int64_t)m * q * q;      /* upper factor, :129 */
                double z[JMB_ORC_QMAX_PAD];
                for (int32_t s = 0; s < (q + 1) / 2; ++s) orc_rnorm_pair(seed, chain, gid, iter, (uint32_t)s, ST_B_PROP, z + 2 * s);
                double sd = sqrt(sigma_b[m]);
                for (int32_t j = 0; j < q; ++j) {
                    double pj = 0.0;
                    for (int32_t l = 0; l <= j; ++l) pj += z[l] * R[l + j * q];
                    new_b[m + (int64_t)j * n] = b[m + (int64_t)j * n] + sd * pj;   /* :670 */
                }
            }
            for (int k = 0; k < d->n_outcomes; ++k) lin_predF(d, w, new_b, k, eta[k]);  /* :671 */
            log_longF(d, w, eta, n_pyb, tmp_n);                                          /* :672 */
            log_pb_all(d, w, new_b, n_pb);                                               /* :673 */
            orc_wlong(d, w, new_b, 0, n_Wl);                                             /* :674 */
            orc_wlong(d, w, new_b, 1, n_Wls);                                            /* :676 */
            if (ic) orc_wlong(d, w, new_b, 2, n_Wli);                                    /* :680 */
            hazard_parts(d, Bs, gam, alp, n_Wl, n_Wls, n_Wli, n_lh, n_H, n_HL, tmp_ns); /* :684-691 */
            log_ptb_all(d, n_lh, n_H, n_HL, n_ptb);                                      /* :692-702 */
            subject_ptb(d, c_ptb, idT_rsum, c_sub);                                      /* :669 */
            subject_ptb(d, n_ptb, idT_rsum, n_sub);                                      /* :703 */
            for (int32_t m = 0; m < n; ++m) {                                            /* :703-725 */
                const uint32_t gid = (uint32_t)(d->subject_offset + m);
                double cur = c_pyb[m] + c_sub[m] + c_pb[m];
                double nw = n_pyb[m] + n_sub[m] + n_pb[m];
                uint32_t r[4];
                orc_philox4x32(seed, chain, gid, iter, 0u, ST_B_ACC, r);
                double log_u = log(orc_u01(r[0], r[1]));
                if (log_u < nw - cur) {
                    accept_b[m] += 1.0;
                    for (int32_t j = 0; j < q; ++j) b[m + (int64_t)j * n] = new_b[m + (int64_t)j * n];
                    c_pyb[m] = n_pyb[m]; c_pb[m] = n_pb[m];
                    /* rows_wlong[[m]] / rows_wlongs[[m]]: the subject's transition rows (:711-722) */
                    for (int32_t tr = (m == 0 ? 0 : idT_rsum[m - 1] + 1); tr <= idT_rsum[m]; ++tr) {
                        for (int32_t a = 0; a < A; ++a) {
                            c_Wl[tr + (int64_t)a * nT] = n_Wl[tr + (int64_t)a * nT];
                            for (int32_t k = 0; k < K; ++k) {
                                int64_t r2 = (int64_t)tr * K + k + (int64_t)a * ns;
                                c_Wls[r2] = n_Wls[r2];
                                if (ic) c_Wli[r2] = n_Wli[r2];
                            }
                        }
                        c_lh[tr] = n_lh[tr]; c_H[tr] = n_H[tr];
                        if (ic) c_HL[tr] = n_HL[tr];
                        c_ptb[tr] = n_ptb[tr];
                    }
                }
            }
            /* ---- survival block :727-770 ---- */
            double sums[3] = {0.0, 0.0, 0.0};
            subject_ptb(d, c_ptb, idT_rsum, c_sub);                                          /* :728 */
            for (int32_t m = 0; m < n; ++m) sums[0] += c_sub[m];
            {
                int P = B + p + A;
                double zz[2 * ((JMB_ORC_PMAX + 1) / 2)];
                for (int s = 0; s < (P + 1) / 2; ++s) orc_rnorm_pair(seed, chain, 0u, iter, (uint32_t)s, ST_SURV, zz + 2 * s);
                for (int j = 0; j < B; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[l] * H_Bs[l + j * B]; nBs[j] = Bs[j] + sqrt(sig_Bs) * pj; }
                for (int j = 0; j < p; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[B + l] * H_g[l + j * p]; ngam[j] = gam[j] + sqrt(sig_g) * pj; }
                for (int j = 0; j < A; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[B + p + l] * H_a[l + j * A]; nalp[j] = alp[j] + sqrt(sig_a) * pj; }
            }
            hazard_parts(d, nBs, ngam, nalp, c_Wl, c_Wls, c_Wli, n_lh, n_H, n_HL, tmp_ns);   /* :735-741 */
            log_ptb_all(d, n_lh, n_H, n_HL, n_ptb);                                         /* :742-752 */
            subject_ptb(d, n_ptb, idT_rsum, n_sub);                                          /* :754 */
            for (int32_t m = 0; m < n; ++m) sums[1] += n_sub[m];
            for (int32_t m = 0; m < n; ++m) sums[2] += c_pyb[m] + c_pb[m];                   /* log_weightsREF :359-369 */
            if (g_allreduce) g_allreduce(sums, 3, g_allreduce_ctx);
            if (out->trace_surv) { out->trace_surv[2 * (int64_t)iter] = sums[0]; out->trace_surv[2 * (int64_t)iter + 1] = sums[1]; }
            double cur_lp = sums[0] + logPrior(Bs, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, B, 1.0)   /* tau = 1: quirk 1, :727 */
                + logPrior(gam, pr->mean_gammas, Tau_g, p, tau_g) + logPrior(alp, pr->mean_alphas, Tau_a, A, tau_a);
            double new_lp = sums[1] + logPrior(nBs, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, B, 1.0)
                + logPrior(ngam, pr->mean_gammas, Tau_g, p, tau_g) + logPrior(nalp, pr->mean_alphas, Tau_a, A, tau_a);
            {
                uint32_t r[4];
                orc_philox4x32(seed, chain, 0u, iter, 0xFFFFFFFFu, ST_SURV, r);
                double log_u = log(orc_u01(r[0], r[1]));
                if (log_u < new_lp - cur_lp) {                                          /* :759-770 */
                    accept_s += 1.0;
                    memcpy(Bs, nBs, sizeof(double) * B);
                    if (p) memcpy(gam, ngam, sizeof(double) * p);
                    memcpy(alp, nalp, sizeof(double) * A);
                    double* t;
                    t = c_lh; c_lh = n_lh; n_lh = t;  t = c_H; c_H = n_H; n_H = t;
                    if (ic) { t = c_HL; c_HL = n_HL; n_HL = t; }
                    t = c_ptb; c_ptb = n_ptb; n_ptb = t;
                }
            }
            for (int j = 0; j < B; ++j) block_par[j + i * (B + p + A)] = Bs[j];            /* :772-774 */
            for (int j = 0; j < p; ++j) block_par[B + j + i * (B + p + A)] = gam[j];
            for (int j = 0; j < A; ++j) block_par[B + p + j + i * (B + p + A)] = alp[j];
            /* ---- Gibbs precisions :776-834 ---- */
            uint32_t draw = 0;
            if (!d->multi_state) {   /* B_tau + 0.5 Bs' Tau Bs (no mean subtracted) :778-779 */
                double Bpost = pr->B_tau_Bs_gammas + 0.5 * quad_form(Bs, pr->Tau_Bs_gammas, B);
                tau_Bs[0] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_Bs, 1.0 / Bpost), ct->eps2, ct->eps3);
            } else {                 /* one draw per stratum on its diagonal block of Tau_Bs_gammas :777-782 */
                for (int ij = 0; ij < nRisks; ++ij) {
                    const int f = d->kn_strat_first[ij], l = d->kn_strat_last[ij];
                    double sq = 0.0;
                    for (int a2 = f; a2 <= l; ++a2) {
                        double t = 0.0;
                        for (int a1 = f; a1 <= l; ++a1) t += Bs[a1] * pr->Tau_Bs_gammas[a1 + a2 * B];
                        sq += t * Bs[a2];
                    }
                    tau_Bs[ij] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_Bs, 1.0 / (pr->B_tau_Bs_gammas + 0.5 * sq)), ct->eps2, ct->eps3);
                }
            }
            for (int kk = 0; kk < n_td; ++kk) {                                            /* :783-794 */
                int L = pr->td_len[kk];
                double s = 0.0;
                for (int a1 = 0; a1 < L; ++a1) for (int a2 = 0; a2 < L; ++a2)
                    s += alp[pr->td_cols[kk][a1]] * Tau_a_pen[pr->td_cols[kk][a1] + pr->td_cols[kk][a2] * A] * alp[pr->td_cols[kk][a2]];
                tau_td[kk] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_td, 1.0 / (B_tau_td + 0.5 * s)), ct->eps2, ct->eps3);
                for (int a1 = 0; a1 < L; ++a1) for (int a2 = 0; a2 < L; ++a2)
                    Tau_a[pr->td_cols[kk][a1] + pr->td_cols[kk][a2] * A] = tau_td[kk] * Tau_a_pen[pr->td_cols[kk][a1] + pr->td_cols[kk][a2] * A];
            }
            if (pr->shrink_gammas) {                                                       /* :795-807 */
                for (int k1 = 0; k1 < p; ++k1) {
                    phi_g[k1] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_phi_g, 1.0 / (pr->B_phi_gammas + 0.5 * tau_g * gam[k1] * gam[k1])), ct->eps2, ct->eps3);
                    Tau_g[k1 + k1 * p] = phi_g[k1];
                }
                double s = 0.0;
                for (int a1 = 0; a1 < p; ++a1) for (int a2 = 0; a2 < p; ++a2) s += gam[a1] * Tau_g[a1 + a2 * p] * gam[a2];
                tau_g = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_g, 1.0 / (pr->B_tau_gammas + 0.5 * s)), ct->eps2, ct->eps3);
            }
            if (pr->shrink_alphas) {                                                       /* :808-834 */
                for (int k2 = 0; k2 < A; ++k2) {
                    double Bp = (pr->double_gamma_alphas ? nu_a[k2] : pr->B_phi_alphas) + 0.5 * tau_a * alp[k2] * alp[k2];
                    phi_a[k2] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_phi_a, 1.0 / Bp), ct->eps2, ct->eps3);
                    Tau_a[k2 + k2 * A] = phi_a[k2];
                }
                double s = 0.0;
                for (int a1 = 0; a1 < A; ++a1) for (int a2 = 0; a2 < A; ++a2) s += alp[a1] * Tau_a[a1 + a2 * A] * alp[a2];
                double Bp = (pr->double_gamma_alphas ? xi_a : pr->B_tau_alphas) + 0.5 * s;
                tau_a = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_a, 1.0 / Bp), ct->eps2, ct->eps3);
                if (pr->double_gamma_alphas) {
                    for (int k2 = 0; k2 < A; ++k2)
                        nu_a[k2] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_nu_a, 1.0 / (pr->B_nu_alphas + phi_a[k2])), ct->eps2, ct->eps3);
                    xi_a = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_xi_a, 1.0 / (pr->B_xi_alphas + tau_a)), ct->eps2, ct->eps3);
                }
            }
            /* ---- store :836-851 (0-based iter against 1-based keep: quirk 3) ---- */
            if (check_iterator < n_keep && keep_iterator < n_out &&
                (int)iter == ct->n_burnin + 1 + check_iterator * ct->n_thin) {
                int ko = keep_iterator;
                memcpy(out->b + (int64_t)ko * n * q, b, sizeof(double) * n * q);
                memcpy(out->Bs_gammas + ko * B, Bs, sizeof(double) * B);
                if (p) memcpy(out->gammas + ko * p, gam, sizeof(double) * p);
                memcpy(out->alphas + ko * A, alp, sizeof(double) * A);
                if (p) memcpy(out->phi_gammas + ko * p, phi_g, sizeof(double) * p);
                memcpy(out->phi_alphas + ko * A, phi_a, sizeof(double) * A);
                for (int j = 0; j < nRisks; ++j) out->tau_Bs_gammas[j + ko * nRisks] = tau_Bs[j];
                out->tau_gammas[ko] = tau_g;
                out->tau_alphas[ko] = tau_a;
                for (int j = 0; j < n_td; ++j) out->tau_td_alphas[j + ko * n_td] = tau_td[j];
                out->logWeights[ko] = -sums[2];                                        /* :847 */
                ++keep_iterator; ++check_iterator;
            }
        }
        if (iters_in_block == 0) break;
        /* ---- adaptation :853-867 ---- */
        double g1 = pow((double)(it + 1), mc1), g2 = ct->c0 * g1;
        double rate_s = accept_s / ct->n_block;
        accept_s_tot += accept_s;
        for (int32_t m = 0; m < n; ++m) {
            double rate = accept_b[m] / ct->n_block;
            accept_b_tot[m] += accept_b[m];
            sigma_b[m] = bounds_sigma(exp(log(sigma_b[m]) + g2 * (rate - ct->target_acc)), ct->eps1, ct->eps3);
        }
        sig_Bs = bounds_sigma(exp(log(sig_Bs) + g2 * (rate_s - ct->target_acc)), ct->eps1, ct->eps3);
        sig_g = bounds_sigma(exp(log(sig_g) + g2 * (rate_s - ct->target_acc)), ct->eps1, ct->eps3);
        sig_a = bounds_sigma(exp(log(sig_a) + g2 * (rate_s - ct->target_acc)), ct->eps1, ct->eps3);
        if (ct->adaptCov && it > start_cov_update && iters_in_block == ct->n_block) {
            int P = B + p + A;
            double* blk = (double*)malloc(sizeof(double) * (B > A ? (B > p ? B : p) : (A > p ? A : p)) * ct->n_block);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < B; ++j) blk[j + i * B] = block_par[j + i * P];
            adapt_cov(S_Bs, B, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < p; ++j) blk[j + i * p] = block_par[B + j + i * P];
            adapt_cov(S_g, p, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < A; ++j) blk[j + i * A] = block_par[B + p + j + i * P];
            adapt_cov(S_a, A, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            free(blk);
        }
    }
    /* export :870-895 */
    int iters_done = max_iters > 0 ? max_iters : its * ct->n_block;
    memcpy(out->sigma_b, sigma_b, sizeof(double) * n);
    out->sigma_surv[0] = sig_Bs; out->sigma_surv[1] = sig_g; out->sigma_surv[2] = sig_a;
    memcpy(out->Sigma_Bs_gammas, S_Bs, sizeof(double) * B * B);
    if (p) memcpy(out->Sigma_gammas, S_g, sizeof(double) * p * p);
    memcpy(out->Sigma_alphas, S_a, sizeof(double) * A * A);
    if (out->accept_b) for (int32_t m = 0; m < n; ++m) out->accept_b[m] = accept_b_tot[m] / iters_done;
    if (out->accept_surv) out->accept_surv[0] = accept_s_tot / iters_done;
    /* final state is exported through out->b slot n_out when the caller asked for it via
       accept arrays only; nothing else to do */

    free(b); free(theta); free(phi_g); free(phi_a); free(nu_a); free(tau_td); free(Tau_g); free(Tau_a);
    free(Tau_a_pen); free(sigma_b); free(S_Bs); free(S_g); free(S_a); free(H_Bs); free(H_g); free(H_a);
    for (int k = 0; k < d->n_outcomes; ++k) free(eta[k]);
    free(eta); free(tmp_n); free(tmp_ns); free(new_b); free(c_pyb); free(n_pyb); free(c_pb); free(n_pb);
    free(c_Wl); free(n_Wl); free(c_Wls); free(n_Wls); free(c_Wli); free(n_Wli); free(c_lh); free(n_lh);
    free(c_H); free(n_H); free(c_HL); free(n_HL); free(c_ptb); free(n_ptb); free(accept_b);
    free(accept_b_tot); free(block_par); free(c_sub); free(n_sub); free(idT_rsum);
    return rc;
}

/* ------------------------------------------------------------------------------------------
 * logPosterior: src/fast_logPost.cpp:8-26 (p = 0 gives logPosterior_nogammas, :30-46).
 * Note: H is summed over all nK rows, idGK_fast is not used there either.
 * ---------------------------------------------------------------------------------------- */
double orc_logPosterior(double temp, int32_t n, int64_t ns, int32_t B, int32_t p, int32_t A,
                        const double* event, const double* W1, const double* W1s,
                        const double* Bs, const double* W2, const double* W2s, const double* gam,
                        const double* Wlong, const double* Wlongs, const double* alp,
                        const double* Pw, const double* mean_Bs, const double* Tau_Bs,
                        const double* mean_g, const double* Tau_g, const double* mean_a,
                        const double* Tau_a, double tauBs, double A_tauBs, double B_tauBs) {
    double* lh = (double*)calloc(n, sizeof(double));
    double* lin = (double*)calloc(ns, sizeof(double));
    gemv_acc(W1, n, B, Bs, lh); gemv_acc(W2, n, p, gam, lh); gemv_acc(Wlong, n, A, alp, lh);
    gemv_acc(W1s, ns, B, Bs, lin); gemv_acc(W2s, ns, p, gam, lin); gemv_acc(Wlongs, ns, A, alp, lin);
    double s1 = 0.0, s2 = 0.0;
    for (int32_t i = 0; i < n; ++i) s1 += event[i] * lh[i];
    for (int64_t r = 0; r < ns; ++r) s2 += Pw[r] * exp(lin[r]);
    free(lh); free(lin);
    return temp * (s1 - s2) + tauBs * logPrior(Bs, mean_Bs, Tau_Bs, B, 1.0) +
           logPrior(gam, mean_g, Tau_g, p, 1.0) + logPrior(alp, mean_a, Tau_a, A, 1.0) +
           (A_tauBs - 1.0) * log(tauBs) - B_tauBs * tauBs;
}

/* gradient_logPosterior: src/fast_score.cpp:8-50 (p = 0: :53-84). out has B + p + A + 1 entries */
void orc_gradient_logPosterior(double temp, int64_t ns, int32_t B, int32_t p, int32_t A,
                               const double* ecs_W1, const double* W1s, const double* Bs,
                               const double* ecs_W2, const double* W2s, const double* gam,
                               const double* ecs_Wlong, const double* Wlongs, const double* alp,
                               const double* Pw, const double* mean_Bs, const double* Tau_Bs,
                               const double* mean_g, const double* Tau_g, const double* mean_a,
                               const double* Tau_a, double tauBs, double A_tauBs, double B_tauBs,
                               double* out) {
    double* Int = (double*)calloc(ns, sizeof(double));
    gemv_acc(W1s, ns, B, Bs, Int); gemv_acc(W2s, ns, p, gam, Int); gemv_acc(Wlongs, ns, A, alp, Int);
    for (int64_t r = 0; r < ns; ++r) Int[r] = Pw[r] * exp(Int[r]);                       /* :17 */
    const double* Ws[3] = {W1s, W2s, Wlongs};
    const double* ecs[3] = {ecs_W1, ecs_W2, ecs_Wlong};
    const double* par[3] = {Bs, gam, alp};
    const double* mean[3] = {mean_Bs, mean_g, mean_a};
    const double* Tau[3] = {Tau_Bs, Tau_g, Tau_a};
    int32_t dims[3] = {B, p, A};
    double mult[3] = {tauBs, 1.0, 1.0};
    int o = 0;
    for (int blk = 0; blk < 3; ++blk) {
        int32_t m = dims[blk];
        for (int32_t i = 0; i < m; ++i) {
            double s = 0.0;
            for (int64_t r = 0; r < ns; ++r) s += Ws[blk][r + (int64_t)i * ns] * Int[r];
            double pen = 0.0;
            for (int32_t l = 0; l < m; ++l) pen += Tau[blk][i + l * m] * (par[blk][l] - mean[blk][l]);
            out[o + i] = temp * (ecs[blk][i] - s) - mult[blk] * pen;                    /* :21-41 */
        }
        o += m;
    }
    out[o] = tauBs * (logPrior(Bs, mean_Bs, Tau_Bs, B, 1.0) + (A_tauBs - 1.0) / tauBs - B_tauBs);   /* :44 */
    free(Int);
}

/* ------------------------------------------------------------------------------------------
 * lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas: src/fast_LAPRWM.cpp:934-1139, 1142-1297
 * (update_RE = FALSE: survival block only, on fixed Wlong / Wlongs).  Differences from lap_rwm_C
 * that are reproduced: the sampled tau_Bs_gammas IS applied in the prior (:911); current_logpost is
 * carried and only refreshed on acceptance (so it keeps the precisions it was computed with);
 * no td-effects and no double-gamma branch; output logPost instead of logWeights.
 * trace (optional): [2 * iter] = current_logpost before the step, [2 * iter + 1] = new_logpost.
 * ---------------------------------------------------------------------------------------- */
int orc_lap_rwm_woRE(const orc_data* d, const double* Wlong, const double* Wlongs, const orc_priors* pr,
                     const orc_control* ct, const orc_chain_in* in, orc_chain_out* out, double* logPost) {
    const int32_t n = d->n, K = d->K, A = d->n_alphas, B = d->n_Bs, p = d->n_gammas;
    const int64_t ns = (int64_t)n * K;
    const int total_it = ct->n_iter + ct->n_burnin;
    const int its = (int)ceil((double)total_it / ct->n_block);
    const int n_out = (int)floor((double)ct->n_iter / ct->n_thin);
    const int start_cov_update = (int)floor((double)ct->n_burnin / ct->n_block) - 1;
    int n_keep = 0;
    for (int v = ct->n_burnin + 1; v <= total_it; v += ct->n_thin) ++n_keep;
    {
        int hits = 0;
        for (int v = ct->n_burnin + 1; v <= total_it && v <= its * ct->n_block - 1; v += ct->n_thin) ++hits;
        if (hits > n_out) return 3;
    }
    const uint32_t seed = ct->seed, chain = ct->chain_id;
    const int P = B + p + A;
    double* Bs = (double*)malloc(sizeof(double) * 2 * (P + 1));
    double *gam = Bs + B, *alp = gam + p, *nBs = alp + A + 1, *ngam = nBs + B, *nalp = ngam + p;
    memcpy(Bs, in->Bs_gammas, sizeof(double) * B);
    if (p) memcpy(gam, in->gammas, sizeof(double) * p);
    memcpy(alp, in->alphas, sizeof(double) * A);
    double tau_Bs = in->tau_Bs_gammas[0], tau_g = in->tau_gammas, tau_a = in->tau_alphas;
    double* phi_g = (double*)malloc(sizeof(double) * (p + 1));
    double* phi_a = (double*)malloc(sizeof(double) * (A + 1));
    for (int j = 0; j < p; ++j) phi_g[j] = in->phi_gammas[j];
    for (int j = 0; j < A; ++j) phi_a[j] = in->phi_alphas[j];
    double* Tau_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* Tau_a = (double*)malloc(sizeof(double) * (A * A + 1));
    memcpy(Tau_g, pr->Tau_gammas, sizeof(double) * p * p);
    memcpy(Tau_a, pr->Tau_alphas, sizeof(double) * A * A);
    const double Apost_tau_Bs = pr->A_tau_Bs_gammas + 0.5 * pr->rank_Tau_Bs_gammas;
    const double Apost_tau_g = pr->A_tau_gammas + 0.5 * pr->rank_Tau_gammas, Apost_phi_g = pr->A_phi_gammas + 0.5;
    const double Apost_tau_a = pr->A_tau_alphas + 0.5 * pr->rank_Tau_alphas, Apost_phi_a = pr->A_phi_alphas + 0.5;
    double sig_Bs = in->sigma_Bs_gammas, sig_g = in->sigma_gammas, sig_a = in->sigma_alphas;
    double* S_Bs = (double*)malloc(sizeof(double) * (B * B + 1));
    double* S_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* S_a = (double*)malloc(sizeof(double) * (A * A + 1));
    memcpy(S_Bs, in->Sigma_Bs_gammas, sizeof(double) * B * B);
    if (p) memcpy(S_g, in->Sigma_gammas, sizeof(double) * p * p);
    memcpy(S_a, in->Sigma_alphas, sizeof(double) * A * A);
    double* H_Bs = (double*)malloc(sizeof(double) * (B * B + 1));
    double* H_g = (double*)malloc(sizeof(double) * (p * p + 1));
    double* H_a = (double*)malloc(sizeof(double) * (A * A + 1));
    double* lh = (double*)malloc(sizeof(double) * n);
    double* lin = (double*)malloc(sizeof(double) * ns);
    double* block_par = (double*)malloc(sizeof(double) * P * ct->n_block);

#define WORE_LOGPOST(BS, GA, AL, RES)                                                            \
    do {   /* logPosterior_woRE :898-914 */                                                       \
        for (int32_t i_ = 0; i_ < n; ++i_) lh[i_] = 0.0;                                           \
        gemv_acc(d->W1, n, B, BS, lh); gemv_acc(d->W2, n, p, GA, lh); gemv_acc(Wlong, n, A, AL, lh); \
        for (int64_t r_ = 0; r_ < ns; ++r_) lin[r_] = 0.0;                                         \
        gemv_acc(d->W1s, ns, B, BS, lin); gemv_acc(d->W2s, ns, p, GA, lin); gemv_acc(Wlongs, ns, A, AL, lin); \
        double s1_ = 0.0, s2_ = 0.0;                                                               \
        for (int32_t i_ = 0; i_ < n; ++i_) s1_ += d->event[i_] * lh[i_];                           \
        for (int64_t r_ = 0; r_ < ns; ++r_) s2_ += d->Pw[r_] * exp(lin[r_]);                       \
        RES = s1_ - s2_ + logPrior(BS, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, B, tau_Bs) +         \
              logPrior(GA, pr->mean_gammas, Tau_g, p, tau_g) + logPrior(AL, pr->mean_alphas, Tau_a, A, tau_a); \
    } while (0)

    double current_logpost;
    WORE_LOGPOST(Bs, gam, alp, current_logpost);
    int keep_iterator = 0, check_iterator = 0, rc = 0;
    for (int it = 0; it < its; ++it) {
        double accept = 0.0;
        if (chol_upper(S_Bs, B, H_Bs) || (p && chol_upper(S_g, p, H_g)) || chol_upper(S_a, A, H_a)) { rc = 2; break; }
        for (int i = 0; i < ct->n_block; ++i) {
            const uint32_t iter = (uint32_t)(it * ct->n_block + i);
            double zz[2 * ((JMB_ORC_PMAX + 1) / 2)];
            for (int s2 = 0; s2 < (P + 1) / 2; ++s2) orc_rnorm_pair(seed, chain, 0u, iter, (uint32_t)s2, ST_SURV, zz + 2 * s2);
            for (int j = 0; j < B; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[l] * H_Bs[l + j * B]; nBs[j] = Bs[j] + sqrt(sig_Bs) * pj; }
            for (int j = 0; j < p; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[B + l] * H_g[l + j * p]; ngam[j] = gam[j] + sqrt(sig_g) * pj; }
            for (int j = 0; j < A; ++j) { double pj = 0.0; for (int l = 0; l <= j; ++l) pj += zz[B + p + l] * H_a[l + j * A]; nalp[j] = alp[j] + sqrt(sig_a) * pj; }
            double new_logpost;
            WORE_LOGPOST(nBs, ngam, nalp, new_logpost);
            if (out->trace_surv) { out->trace_surv[2 * (int64_t)iter] = current_logpost; out->trace_surv[2 * (int64_t)iter + 1] = new_logpost; }
            uint32_t r[4];
            orc_philox4x32(seed, chain, 0u, iter, 0xFFFFFFFFu, ST_SURV, r);
            if (log(orc_u01(r[0], r[1])) < new_logpost - current_logpost) {
                accept += 1.0;
                memcpy(Bs, nBs, sizeof(double) * B);
                if (p) memcpy(gam, ngam, sizeof(double) * p);
                memcpy(alp, nalp, sizeof(double) * A);
                current_logpost = new_logpost;
            }
            for (int j = 0; j < p; ++j) block_par[B + j + i * P] = gam[j];
            for (int j = 0; j < A; ++j) block_par[B + p + j + i * P] = alp[j];
            for (int j = 0; j < B; ++j) block_par[j + i * P] = Bs[j];
            uint32_t draw = 0;
            tau_Bs = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_Bs, 1.0 / (pr->B_tau_Bs_gammas + 0.5 * quad_form(Bs, pr->Tau_Bs_gammas, B))), ct->eps2, ct->eps3);
            if (pr->shrink_gammas) {
                for (int k1 = 0; k1 < p; ++k1) {
                    phi_g[k1] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_phi_g, 1.0 / (pr->B_phi_gammas + 0.5 * tau_g * gam[k1] * gam[k1])), ct->eps2, ct->eps3);
                    Tau_g[k1 + k1 * p] = phi_g[k1];
                }
                tau_g = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_g, 1.0 / (pr->B_tau_gammas + 0.5 * quad_form(gam, Tau_g, p))), ct->eps2, ct->eps3);
            }
            if (pr->shrink_alphas) {
                for (int k2 = 0; k2 < A; ++k2) {
                    phi_a[k2] = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_phi_a, 1.0 / (pr->B_phi_alphas + 0.5 * tau_a * alp[k2] * alp[k2])), ct->eps2, ct->eps3);
                    Tau_a[k2 + k2 * A] = phi_a[k2];
                }
                tau_a = clampd(orc_rgamma(seed, chain, draw++, iter, Apost_tau_a, 1.0 / (pr->B_tau_alphas + 0.5 * quad_form(alp, Tau_a, A))), ct->eps2, ct->eps3);
            }
            if (check_iterator < n_keep && keep_iterator < n_out && (int)iter == ct->n_burnin + 1 + check_iterator * ct->n_thin) {
                int ko = keep_iterator;
                memcpy(out->Bs_gammas + ko * B, Bs, sizeof(double) * B);
                if (p) { memcpy(out->gammas + ko * p, gam, sizeof(double) * p); memcpy(out->phi_gammas + ko * p, phi_g, sizeof(double) * p); }
                memcpy(out->alphas + ko * A, alp, sizeof(double) * A);
                memcpy(out->phi_alphas + ko * A, phi_a, sizeof(double) * A);
                out->tau_Bs_gammas[ko] = tau_Bs; out->tau_gammas[ko] = tau_g; out->tau_alphas[ko] = tau_a;
                if (logPost) logPost[ko] = current_logpost;
                ++keep_iterator; ++check_iterator;
            }
        }
        double g1 = pow((double)(it + 1), -ct->c1), g2 = ct->c0 * g1, rate = accept / ct->n_block;
        sig_Bs = bounds_sigma(exp(log(sig_Bs) + g2 * (rate - ct->target_acc)), ct->eps1, ct->eps3);
        sig_g = bounds_sigma(exp(log(sig_g) + g2 * (rate - ct->target_acc)), ct->eps1, ct->eps3);
        sig_a = bounds_sigma(exp(log(sig_a) + g2 * (rate - ct->target_acc)), ct->eps1, ct->eps3);
        if (ct->adaptCov && it > start_cov_update) {
            double* blk = (double*)malloc(sizeof(double) * (P + 1) * ct->n_block);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < B; ++j) blk[j + i * B] = block_par[j + i * P];
            adapt_cov(S_Bs, B, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < p; ++j) blk[j + i * p] = block_par[B + j + i * P];
            adapt_cov(S_g, p, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            for (int i = 0; i < ct->n_block; ++i) for (int j = 0; j < A; ++j) blk[j + i * A] = block_par[B + p + j + i * P];
            adapt_cov(S_a, A, blk, ct->n_block, g1, ct->eps2, ct->eps3);
            free(blk);
        }
    }
#undef WORE_LOGPOST
    out->sigma_surv[0] = sig_Bs; out->sigma_surv[1] = sig_g; out->sigma_surv[2] = sig_a;
    memcpy(out->Sigma_Bs_gammas, S_Bs, sizeof(double) * B * B);
    if (p) memcpy(out->Sigma_gammas, S_g, sizeof(double) * p * p);
    memcpy(out->Sigma_alphas, S_a, sizeof(double) * A * A);
    free(Bs); free(phi_g); free(phi_a); free(Tau_g); free(Tau_a); free(S_Bs); free(S_g); free(S_a);
    free(H_Bs); free(H_g); free(H_a); free(lh); free(lin); free(block_par);
    return rc;
}
