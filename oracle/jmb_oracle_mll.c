/*
 * jmb_oracle_mll.c - CPU restatement of marglogLik2() (R/margLogLik2.R:1-140), the Laplace step that gives
 * mvJointModelBayes() its initial values and proposal covariances (R/mvJointModelBayes.R:716-725) and, with
 * update_RE = FALSE, runs once per draw (:792).
 *
 * TEST INFRASTRUCTURE ONLY (see jmb_oracle.h).  Restated from the reference: fn / gr (margLogLik2.R:34-73, on
 * orc_logPosterior / orc_gradient_logPosterior, which tests/test_reference_golden.py pins to the reference's own
 * C++), cd.vec (:74-87), the Laplace value (:110-111), nearPD (R/nearPD.R:1-47) and the Covs / inits split
 * (:113-137).  Third-party arithmetic absent from /root/reference and restated from its published algorithm:
 * optim(method = "BFGS", hessian = TRUE) = R's vmmin + optimhess (src/appl/optim.c, src/library/stats/src/optim.c),
 * solve / determinant (LAPACK LU -> Gaussian elimination with partial pivoting), eigen(symmetric = TRUE) (cyclic
 * Jacobi).  The try() fall-backs of :95-109 (jittered restart, Nelder-Mead) are not restated: a non-finite start is
 * reported as an error.  PARITY PIN: the pieces fn / gr are pinned; the R-level loop has no fixture (no R here):
 * "parity unpinned", checked against SciPy's BFGS, numpy.linalg and analytic properties in tests/.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "jmb_oracle.h"
#include "jmb_oracle_mll.h"

/* ---- generic helpers ------------------------------------------------------------------------------------- */
/* optim(par, fn, gr, method = "BFGS", hessian = TRUE, control = list(parscale, maxit, reltol, ndeps)) */
int orc_optim_bfgs(int n, const double* par, const double* parscale, orc_obj_fn fn, orc_obj_gr gr, void* ctx,
                   int maxit, double reltol, double ndeps, double* out_par, double* out_value, double* out_hessian,
                   int* counts) {
    const double stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    double* mem = (double*)calloc((size_t)(8 * n + n * n), sizeof(double));
    if (!mem) return 99;
    double *b = mem, *g = b + n, *t = g + n, *X = t + n, *c = X + n, *xs = c + n, *gs = xs + n, *df2 = gs + n, *B = df2 + n;
    int count = 0, funcount, gradcount, ilast, iter = 0, conv;
    /* fminfn / fmingr: the optimiser works on par / parscale */
#define FMINFN(z, res) do { for (int i_ = 0; i_ < n; ++i_) xs[i_] = (z)[i_] * parscale[i_]; res = fn(n, xs, ctx); } while (0)
#define FMINGR(z, out) do { for (int i_ = 0; i_ < n; ++i_) xs[i_] = (z)[i_] * parscale[i_]; gr(n, xs, gs, ctx); \
                            for (int i_ = 0; i_ < n; ++i_) (out)[i_] = gs[i_] * parscale[i_]; } while (0)
    for (int i = 0; i < n; ++i) b[i] = par[i] / parscale[i];
    double f, Fmin, gradproj, s, steplength, D1, D2;
    FMINFN(b, f);
    if (!isfinite(f)) { free(mem); return 10; }        /* error("initial value in 'vmmin' is not finite") */
    Fmin = f;
    funcount = gradcount = 1;
    FMINGR(b, g);
    iter++;
    ilast = gradcount;
    do {
        if (ilast == gradcount)
            for (int i = 0; i < n; ++i) { for (int j = 0; j < i; ++j) B[i * n + j] = 0.0; B[i * n + i] = 1.0; }
        for (int i = 0; i < n; ++i) { X[i] = b[i]; c[i] = g[i]; }
        gradproj = 0.0;
        for (int i = 0; i < n; ++i) {
            s = 0.0;
            for (int j = 0; j <= i; ++j) s -= B[i * n + j] * c[j];
            for (int j = i + 1; j < n; ++j) s -= B[j * n + i] * c[j];
            t[i] = s;
            gradproj += s * c[i];
        }
        if (gradproj < 0.0) {
            int accpoint = 0;
            steplength = 1.0;
            do {
                count = 0;
                for (int i = 0; i < n; ++i) {
                    b[i] = X[i] + steplength * t[i];
                    if (reltest + X[i] == reltest + b[i]) count++;
                }
                if (count < n) {
                    FMINFN(b, f);
                    funcount++;
                    accpoint = isfinite(f) && (f <= Fmin + gradproj * steplength * acctol);
                    if (!accpoint) steplength *= stepredn;
                }
            } while (!(count == n || accpoint));
            const int enough = fabs(f - Fmin) > reltol * (fabs(Fmin) + reltol);
            if (!enough) { count = n; Fmin = f; }
            if (count < n) {
                Fmin = f;
                FMINGR(b, g);
                gradcount++;
                iter++;
                D1 = 0.0;
                for (int i = 0; i < n; ++i) { t[i] = steplength * t[i]; c[i] = g[i] - c[i]; D1 += t[i] * c[i]; }
                if (D1 > 0) {
                    D2 = 0.0;
                    for (int i = 0; i < n; ++i) {
                        s = 0.0;
                        for (int j = 0; j <= i; ++j) s += B[i * n + j] * c[j];
                        for (int j = i + 1; j < n; ++j) s += B[j * n + i] * c[j];
                        X[i] = s;
                        D2 += s * c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (int i = 0; i < n; ++i)
                        for (int j = 0; j <= i; ++j) B[i * n + j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
                } else {
                    ilast = gradcount;
                }
            } else {
                if (ilast < gradcount) { count = 0; ilast = gradcount; }
            }
        } else {
            count = 0;
            if (ilast == gradcount) count = n; else ilast = gradcount;
        }
        if (iter >= maxit) break;
        if (gradcount - ilast > 2 * n) ilast = gradcount;
    } while (count != n || ilast != gradcount);
    conv = iter < maxit ? 0 : 1;
    for (int i = 0; i < n; ++i) out_par[i] = b[i] * parscale[i];
    *out_value = Fmin;
    if (counts) { counts[0] = funcount; counts[1] = gradcount; }
    if (out_hessian) {                                  /* optimhess */
        double* dpar = t; double* df1 = c;
        for (int i = 0; i < n; ++i) dpar[i] = out_par[i] / parscale[i];
        for (int i = 0; i < n; ++i) {
            const double eps = ndeps / parscale[i];
            dpar[i] = dpar[i] + eps;
            FMINGR(dpar, df1);
            dpar[i] = dpar[i] - 2 * eps;
            FMINGR(dpar, df2);
            for (int j = 0; j < n; ++j) out_hessian[i * n + j] = (df1[j] - df2[j]) / (2 * eps * parscale[i] * parscale[j]);
            dpar[i] = dpar[i] + eps;
        }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < i; ++j) {
                const double tmp = 0.5 * (out_hessian[i * n + j] + out_hessian[j * n + i]);
                out_hessian[i * n + j] = out_hessian[j * n + i] = tmp;
            }
    }
#undef FMINFN
#undef FMINGR
    free(mem);
    return conv;
}

/* LU with partial pivoting: inverse (may be NULL) and log|det|; returns 1 when singular */
int orc_lu_inverse(const double* A, int n, double* inv, double* logabsdet) {
    double* a = (double*)malloc(sizeof(double) * (size_t)n * n * 2);
    double* e = a + (size_t)n * n;
    for (int j = 0; j < n * n; ++j) { a[j] = A[j]; e[j] = 0.0; }
    for (int j = 0; j < n; ++j) e[j + j * n] = 1.0;
    double ld = 0.0;
    for (int c = 0; c < n; ++c) {
        int piv = c; double best = fabs(a[c + c * n]);
        for (int r = c + 1; r < n; ++r) if (fabs(a[r + c * n]) > best) { best = fabs(a[r + c * n]); piv = r; }
        if (!(best > 0.0) || !isfinite(best)) { free(a); return 1; }
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                double t = a[c + j * n]; a[c + j * n] = a[piv + j * n]; a[piv + j * n] = t;
                t = e[c + j * n]; e[c + j * n] = e[piv + j * n]; e[piv + j * n] = t;
            }
        ld += log(fabs(a[c + c * n]));
        const double d = 1.0 / a[c + c * n];
        for (int j = 0; j < n; ++j) { a[c + j * n] *= d; e[c + j * n] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r + c * n];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { a[r + j * n] -= f * a[c + j * n]; e[r + j * n] -= f * e[c + j * n]; }
        }
    }
    if (inv) memcpy(inv, e, sizeof(double) * (size_t)n * n);
    if (logabsdet) *logabsdet = ld;
    free(a);
    return 0;
}

/* eigen(S, symmetric = TRUE) for any n: cyclic Jacobi until the off-diagonal mass is negligible; values decreasing */
void orc_eigen_sym_n(const double* S, int n, double* val, double* vec) {
    double* a = (double*)malloc(sizeof(double) * (size_t)n * n);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i + j * n] = 0.5 * (S[i + j * n] + S[j + i * n]); vec[i + j * n] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < n; ++i) { dia += a[i + i * n] * a[i + i * n]; for (int j = i + 1; j < n; ++j) off += a[i + j * n] * a[i + j * n]; }
        if (off <= 1e-32 * dia || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p + q * n], app = a[p + p * n], aqq = a[q + q * n];
                if (fabs(apq) <= 1e-20 * (fabs(app) + fabs(aqq))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { const double akp = a[k + p * n], akq = a[k + q * n]; a[k + p * n] = c * akp - s * akq; a[k + q * n] = s * akp + c * akq; }
                for (int k = 0; k < n; ++k) { const double apk = a[p + k * n], aqk = a[q + k * n]; a[p + k * n] = c * apk - s * aqk; a[q + k * n] = s * apk + c * aqk; }
                for (int k = 0; k < n; ++k) { const double vkp = vec[k + p * n], vkq = vec[k + q * n]; vec[k + p * n] = c * vkp - s * vkq; vec[k + q * n] = s * vkp + c * vkq; }
            }
    }
    for (int j = 0; j < n; ++j) val[j] = a[j + j * n];
    for (int j = 0; j < n - 1; ++j) {
        int best = j;
        for (int l = j + 1; l < n; ++l) if (val[l] > val[best]) best = l;
        if (best != j) {
            double t = val[j]; val[j] = val[best]; val[best] = t;
            for (int k = 0; k < n; ++k) { t = vec[k + j * n]; vec[k + j * n] = vec[k + best * n]; vec[k + best * n] = t; }
        }
    }
    free(a);
}

/* nearPD(M, eig.tol = 1e-06, conv.tol = 1e-07, posd.tol = 1e-08, maxits = 100), R/nearPD.R */
void orc_nearPD(const double* M, int n, double* out) {
    const double eig_tol = 1e-06, conv_tol = 1e-07, posd_tol = 1e-08;
    const int maxits = 100;
    if (n == 1) { out[0] = fabs(M[0]); return; }
    const size_t nn = (size_t)n * n;
    double* mem = (double*)calloc(nn * 5 + (size_t)n, sizeof(double));
    double *U = mem, *X = U + nn, *Y = X + nn, *T = Y + nn, *Q = T + nn, *d = Q + nn;
    memcpy(X, M, sizeof(double) * nn);
    int iter = 0, converged = 0;
    while (iter < maxits && !converged) {
        memcpy(Y, X, sizeof(double) * nn);
        for (size_t j = 0; j < nn; ++j) T[j] = Y[j] - U[j];
        orc_eigen_sym_n(Y, n, d, Q);
        /* X <- QQ %*% D[p, p] %*% t(QQ) over the eigenvalues d > eig.tol * d[1] */
        for (size_t j = 0; j < nn; ++j) X[j] = 0.0;
        for (int k = 0; k < n; ++k) {
            if (!(d[k] > eig_tol * d[0])) continue;
            for (int c = 0; c < n; ++c) { const double w = d[k] * Q[c + (size_t)k * n]; for (int r = 0; r < n; ++r) X[r + (size_t)c * n] += Q[r + (size_t)k * n] * w; }
        }
        for (size_t j = 0; j < nn; ++j) U[j] = X[j] - T[j];
        for (int r = 0; r < n; ++r) for (int c = r + 1; c < n; ++c) { const double m = (X[r + (size_t)c * n] + X[c + (size_t)r * n]) / 2; X[r + (size_t)c * n] = m; X[c + (size_t)r * n] = m; }
        double num = 0.0, den = 0.0;                   /* inorm(x) = max(rowSums(abs(x))) */
        for (int r = 0; r < n; ++r) {
            double a = 0.0, b = 0.0;
            for (int c = 0; c < n; ++c) { a += fabs(Y[r + (size_t)c * n] - X[r + (size_t)c * n]); b += fabs(Y[r + (size_t)c * n]); }
            if (a > num) num = a;
            if (b > den) den = b;
        }
        iter++;
        converged = (num / den) <= conv_tol;
    }
    for (int r = 0; r < n; ++r) for (int c = r + 1; c < n; ++c) { const double m = (X[r + (size_t)c * n] + X[c + (size_t)r * n]) / 2; X[r + (size_t)c * n] = m; X[c + (size_t)r * n] = m; }
    orc_eigen_sym_n(X, n, d, Q);
    const double Eps = posd_tol * fabs(d[0]);
    if (d[n - 1] < Eps) {
        for (int k = 0; k < n; ++k) if (d[k] < Eps) d[k] = Eps;
        double* od = T;                                 /* o.diag */
        for (int j = 0; j < n; ++j) od[j] = X[j + (size_t)j * n];
        for (size_t j = 0; j < nn; ++j) X[j] = 0.0;      /* X <- Q %*% (d * t(Q)) */
        for (int k = 0; k < n; ++k)
            for (int c = 0; c < n; ++c) { const double w = d[k] * Q[c + (size_t)k * n]; for (int r = 0; r < n; ++r) X[r + (size_t)c * n] += Q[r + (size_t)k * n] * w; }
        double* D = Y;
        for (int j = 0; j < n; ++j) D[j] = sqrt((Eps > od[j] ? Eps : od[j]) / X[j + (size_t)j * n]);
        for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) X[r + (size_t)c * n] = D[r] * X[r + (size_t)c * n] * D[c];
    }
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) out[r + (size_t)c * n] = (X[r + (size_t)c * n] + X[c + (size_t)r * n]) / 2;
    free(mem);
}

/* ---- marglogLik2 ----------------------------------------------------------------------------------------------- */
typedef struct mll_ctx {
    const orc_mll_data* D; const orc_priors* pr; double temp; int fixed; double tau_fixed;
    const double* ecs;       /* event_colSums of [W1 | W2 | Wlong] */
    double* tmp;             /* B + p + A + 1 */
} mll_ctx;

static void mll_split(const mll_ctx* c, const double* th, const double** Bs, const double** g, const double** a, double* tau) {
    const orc_mll_data* D = c->D;
    *Bs = th; *g = th + D->B; *a = th + D->B + D->p;
    *tau = c->fixed ? c->tau_fixed : exp(th[D->B + D->p + D->A]);
}
static double mll_fn(int n, const double* th, void* vctx) {           /* R/margLogLik2.R:34-51 */
    (void)n;
    const mll_ctx* c = (const mll_ctx*)vctx;
    const orc_mll_data* D = c->D; const orc_priors* pr = c->pr;
    const double *Bs, *g, *a; double tau;
    mll_split(c, th, &Bs, &g, &a, &tau);
    return 0.0 - orc_logPosterior(c->temp, D->n, D->ns, D->B, D->p, D->A, D->event, D->W1, D->W1s, Bs, D->W2, D->W2s, g, D->Wlong, D->Wlongs,
                                  a, D->Pw, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, pr->mean_gammas, pr->Tau_gammas, pr->mean_alphas,
                                  pr->Tau_alphas, tau, pr->A_tau_Bs_gammas, pr->B_tau_Bs_gammas);
}
static void mll_gr(int n, const double* th, double* out, void* vctx) { /* :52-73 */
    const mll_ctx* c = (const mll_ctx*)vctx;
    const orc_mll_data* D = c->D; const orc_priors* pr = c->pr;
    const double *Bs, *g, *a; double tau;
    mll_split(c, th, &Bs, &g, &a, &tau);
    orc_gradient_logPosterior(c->temp, D->ns, D->B, D->p, D->A, c->ecs, D->W1s, Bs, c->ecs + D->B, D->W2s, g, c->ecs + D->B + D->p, D->Wlongs, a,
                              D->Pw, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, pr->mean_gammas, pr->Tau_gammas, pr->mean_alphas, pr->Tau_alphas,
                              tau, pr->A_tau_Bs_gammas, pr->B_tau_Bs_gammas, c->tmp);
    for (int j = 0; j < n; ++j) out[j] = 0.0 - c->tmp[j];           /* head(out, -1) when tau is fixed */
}

int orc_marglogLik2(const orc_mll_data* D, const orc_priors* pr, const double* thetas, double tau_Bs_gammas, double temp,
                    int fixed_tau_Bs_gammas, orc_mll_out* out) {
    const int B = D->B, p = D->p, A = D->A, P = B + p + A, d = fixed_tau_Bs_gammas ? P : P + 1;
    double* mem = (double*)calloc((size_t)(3 * (P + 1) + 3 * d + 4 * (size_t)d * d), sizeof(double));
    double *ecs = mem, *tmp = ecs + P + 1, *pscale = tmp + P + 1, *par = pscale + P + 1, *x1 = par + d, *g1 = x1 + d;
    double *H = g1 + d, *res = H + (size_t)d * d, *inv = res + (size_t)d * d, *g2 = inv + (size_t)d * d;
    /* event_colSums (R/mvJointModelBayes.R:677-681) */
    for (int j = 0; j < B; ++j) { double t = 0.0; for (int i = 0; i < D->n; ++i) t += D->event[i] * D->W1[i + (size_t)j * D->n]; ecs[j] = t; }
    for (int j = 0; j < p; ++j) { double t = 0.0; for (int i = 0; i < D->n; ++i) t += D->event[i] * D->W2[i + (size_t)j * D->n]; ecs[B + j] = t; }
    for (int j = 0; j < A; ++j) { double t = 0.0; for (int i = 0; i < D->n; ++i) t += D->event[i] * D->Wlong[i + (size_t)j * D->n]; ecs[B + p + j] = t; }
    mll_ctx c = {D, pr, temp, fixed_tau_Bs_gammas, tau_Bs_gammas, ecs, tmp};
    for (int j = 0; j < d; ++j) pscale[j] = 1.0;
    if (!fixed_tau_Bs_gammas) pscale[d - 1] = 0.1;                  /* :93 */
    int counts[2] = {0, 0};
    double value = 0.0;
    const int conv = orc_optim_bfgs(d, thetas, pscale, mll_fn, mll_gr, &c, 100, 1.490116119384765625e-8, 1e-3, par, &value, H, counts);
    if (conv == 10) { free(mem); return 10; }
    double logdet = 0.0;
    if (orc_lu_inverse(H, d, NULL, &logdet)) logdet = NAN;          /* determinant(opt$hessian)$modulus */
    out->value = 0.5 * (d * log(2.0 * 3.14159265358979323846) - logdet) - value;   /* :110-111 */
    out->opt_value = value; out->convergence = conv; out->counts[0] = counts[0]; out->counts[1] = counts[1];
    memcpy(out->par, par, sizeof(double) * d);
    memcpy(out->hessian, H, sizeof(double) * (size_t)d * d);
    if (fixed_tau_Bs_gammas) {
        /* hes(opt$par) = cd.vec(par, gr, eps = 0.001) (:74-87): column i = (gr(x1) - gr(x2)) / (x1_i - x2_i), symmetrised */
        for (int i = 0; i < d; ++i) {
            const double ex = fabs(par[i]) > 1.0 ? fabs(par[i]) : 1.0;
            memcpy(x1, par, sizeof(double) * d);
            x1[i] = par[i] + 0.001 * ex;
            mll_gr(d, x1, g1, &c);
            const double xp = x1[i];
            x1[i] = par[i] - 0.001 * ex;
            mll_gr(d, x1, g2, &c);
            const double dx = xp - x1[i];
            for (int j = 0; j < d; ++j) res[j + (size_t)i * d] = (g1[j] - g2[j]) / dx;
        }
        for (int r = 0; r < d; ++r) for (int cc = 0; cc < d; ++cc) H[r + (size_t)cc * d] = 0.5 * (res[r + (size_t)cc * d] + res[cc + (size_t)r * d]);
        if (orc_lu_inverse(H, d, inv, NULL)) { free(mem); return 11; }
        for (int r = 0; r < d; ++r) for (int cc = 0; cc < d; ++cc) res[r + (size_t)cc * d] = 0.5 * (inv[r + (size_t)cc * d] + inv[cc + (size_t)r * d]);
        orc_nearPD(res, d, inv);
        /* Covs: the diagonal blocks of invH (:116-131) */
        for (int r = 0; r < B; ++r) for (int cc = 0; cc < B; ++cc) out->Cov_Bs_gammas[r + (size_t)cc * B] = inv[r + (size_t)cc * d];
        for (int r = 0; r < p; ++r) for (int cc = 0; cc < p; ++cc) out->Cov_gammas[r + (size_t)cc * p] = inv[(B + r) + (size_t)(B + cc) * d];
        for (int r = 0; r < A; ++r) for (int cc = 0; cc < A; ++cc) out->Cov_alphas[r + (size_t)cc * A] = inv[(B + p + r) + (size_t)(B + p + cc) * d];
    }
    free(mem);
    return 0;
}
