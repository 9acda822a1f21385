// jmb_laplace.cu - marglogLik2() (R/margLogLik2.R:1-140) as a native driver around the device evaluators
// jmb_logPosterior / jmb_gradient_logPosterior: the BFGS iterations of optim(), its
// numerical Hessian, the Laplace value, cd.vec, solve and nearPD run here on the host in C++ - every function
// value and gradient they ask for is a full pass over the n + nK hazard rows on the GPU.  Product code: uses only
// the public C ABI of the library (no oracle, no CPU fallback for the evaluations).
//
// optim(method = "BFGS", hessian = TRUE) is R's own C code (vmmin in src/appl/optim.c; fminfn / fmingr / optimhess in
// src/library/stats/src/optim.c), absent from the JMbayes tree and restated here from the published algorithm, as are
// solve / determinant (LU with partial pivoting) and eigen(symmetric = TRUE) (cyclic Jacobi) behind nearPD
// (R/nearPD.R:1-47).  The try() fall-backs of R/margLogLik2.R:95-109 (jittered restart, Nelder-Mead) are not provided:
// a non-finite initial value triggers the reference's jittered restart (R/margLogLik2.R:95-100).
#include <cmath>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/jmb200.h"
#include "jmb_host.h"

namespace {

using Vec = std::vector<double>;
using Fn = std::function<int(const Vec&, double&)>;          // returns non-zero when the device call failed
using Gr = std::function<int(const Vec&, Vec&)>;

struct OptimResult { Vec par, hessian; double value = 0.0; int convergence = 0, fncount = 0, grcount = 0; };

// vmmin + optimhess on par / parscale; rc: 0 ok, 10 non-finite start, 1 device error
int optim_bfgs(const Vec& par0, const Vec& parscale, const Fn& fn, const Gr& gr, int maxit, double reltol, double ndeps,
               OptimResult& R) {
    const int n = (int)par0.size();
    const double stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    Vec b(n), g(n), t(n), X(n), c(n), xs(n), gs(n), B((size_t)n * n, 0.0);
    auto fminfn = [&](const Vec& z, double& out) { for (int i = 0; i < n; ++i) xs[i] = z[i] * parscale[i]; return fn(xs, out); };
    auto fmingr = [&](const Vec& z, Vec& out) {
        for (int i = 0; i < n; ++i) xs[i] = z[i] * parscale[i];
        if (gr(xs, gs)) return 1;
        for (int i = 0; i < n; ++i) out[i] = gs[i] * parscale[i];
        return 0;
    };
    for (int i = 0; i < n; ++i) b[i] = par0[i] / parscale[i];
    double f = 0.0;
    if (fminfn(b, f)) return 1;
    if (!std::isfinite(f)) return 10;
    double Fmin = f;
    int funcount = 1, gradcount = 1, iter = 0, count = 0;
    if (fmingr(b, g)) return 1;
    iter++;
    int ilast = gradcount;
    do {
        if (ilast == gradcount)
            for (int i = 0; i < n; ++i) { for (int j = 0; j < i; ++j) B[(size_t)i * n + j] = 0.0; B[(size_t)i * n + i] = 1.0; }
        for (int i = 0; i < n; ++i) { X[i] = b[i]; c[i] = g[i]; }
        double gradproj = 0.0;
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = 0; j <= i; ++j) s -= B[(size_t)i * n + j] * c[j];
            for (int j = i + 1; j < n; ++j) s -= B[(size_t)j * n + i] * c[j];
            t[i] = s;
            gradproj += s * c[i];
        }
        if (gradproj < 0.0) {                          // search direction is downhill
            double steplength = 1.0;
            bool accpoint = false;
            do {
                count = 0;
                for (int i = 0; i < n; ++i) {
                    b[i] = X[i] + steplength * t[i];
                    if (reltest + X[i] == reltest + b[i]) count++;
                }
                if (count < n) {
                    if (fminfn(b, f)) return 1;
                    funcount++;
                    accpoint = std::isfinite(f) && (f <= Fmin + gradproj * steplength * acctol);
                    if (!accpoint) steplength *= stepredn;
                }
            } while (!(count == n || accpoint));
            const bool enough = std::fabs(f - Fmin) > reltol * (std::fabs(Fmin) + reltol);
            if (!enough) { count = n; Fmin = f; }
            if (count < n) {                           // making progress
                Fmin = f;
                if (fmingr(b, g)) return 1;
                gradcount++;
                iter++;
                double D1 = 0.0;
                for (int i = 0; i < n; ++i) { t[i] = steplength * t[i]; c[i] = g[i] - c[i]; D1 += t[i] * c[i]; }
                if (D1 > 0) {
                    double D2 = 0.0;
                    for (int i = 0; i < n; ++i) {
                        double s = 0.0;
                        for (int j = 0; j <= i; ++j) s += B[(size_t)i * n + j] * c[j];
                        for (int j = i + 1; j < n; ++j) s += B[(size_t)j * n + i] * c[j];
                        X[i] = s;
                        D2 += s * c[i];
                    }
                    D2 = 1.0 + D2 / D1;
                    for (int i = 0; i < n; ++i)
                        for (int j = 0; j <= i; ++j) B[(size_t)i * n + j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
                } else {
                    ilast = gradcount;
                }
            } else if (ilast < gradcount) {            // no progress
                count = 0;
                ilast = gradcount;
            }
        } else {                                       // uphill search
            count = 0;
            if (ilast == gradcount) count = n; else ilast = gradcount;
        }
        if (iter >= maxit) break;
        if (gradcount - ilast > 2 * n) ilast = gradcount;
    } while (count != n || ilast != gradcount);
    R.convergence = iter < maxit ? 0 : 1;
    R.par.resize(n);
    for (int i = 0; i < n; ++i) R.par[i] = b[i] * parscale[i];
    R.value = Fmin; R.fncount = funcount; R.grcount = gradcount;
    // optimhess: central differences of the scaled gradient, symmetrised
    R.hessian.assign((size_t)n * n, 0.0);
    Vec dpar(n), df1(n), df2(n);
    for (int i = 0; i < n; ++i) dpar[i] = R.par[i] / parscale[i];
    for (int i = 0; i < n; ++i) {
        const double eps = ndeps / parscale[i];
        dpar[i] = dpar[i] + eps;
        if (fmingr(dpar, df1)) return 1;
        dpar[i] = dpar[i] - 2 * eps;
        if (fmingr(dpar, df2)) return 1;
        for (int j = 0; j < n; ++j) R.hessian[(size_t)i * n + j] = (df1[j] - df2[j]) / (2 * eps * parscale[i] * parscale[j]);
        dpar[i] = dpar[i] + eps;
    }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < i; ++j) {
            const double tmp = 0.5 * (R.hessian[(size_t)i * n + j] + R.hessian[(size_t)j * n + i]);
            R.hessian[(size_t)i * n + j] = tmp; R.hessian[(size_t)j * n + i] = tmp;
        }
    return 0;
}

// LU with partial pivoting (Gauss-Jordan form): inverse and log|det|; false when singular
bool lu_inverse(const Vec& A, int n, Vec* inv, double* logabsdet) {
    Vec a = A, e((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j) e[j + (size_t)j * n] = 1.0;
    double ld = 0.0;
    for (int c = 0; c < n; ++c) {
        int piv = c; double best = std::fabs(a[c + (size_t)c * n]);
        for (int r = c + 1; r < n; ++r) if (std::fabs(a[r + (size_t)c * n]) > best) { best = std::fabs(a[r + (size_t)c * n]); piv = r; }
        if (!(best > 0.0) || !std::isfinite(best)) return false;
        if (piv != c)
            for (int j = 0; j < n; ++j) { std::swap(a[c + (size_t)j * n], a[piv + (size_t)j * n]); std::swap(e[c + (size_t)j * n], e[piv + (size_t)j * n]); }
        ld += std::log(std::fabs(a[c + (size_t)c * n]));
        const double d = 1.0 / a[c + (size_t)c * n];
        for (int j = 0; j < n; ++j) { a[c + (size_t)j * n] *= d; e[c + (size_t)j * n] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r + (size_t)c * n];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { a[r + (size_t)j * n] -= f * a[c + (size_t)j * n]; e[r + (size_t)j * n] -= f * e[c + (size_t)j * n]; }
        }
    }
    if (inv) *inv = e;
    if (logabsdet) *logabsdet = ld;
    return true;
}

// eigen(S, symmetric = TRUE): cyclic Jacobi until the off-diagonal mass is negligible; values decreasing
void eigen_sym(const Vec& S, int n, Vec& val, Vec& vec) {
    Vec a((size_t)n * n);
    vec.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) { a[i + (size_t)j * n] = 0.5 * (S[i + (size_t)j * n] + S[j + (size_t)i * n]); if (i == j) vec[i + (size_t)j * n] = 1.0; }
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < n; ++i) { dia += a[i + (size_t)i * n] * a[i + (size_t)i * n]; for (int j = i + 1; j < n; ++j) off += a[i + (size_t)j * n] * a[i + (size_t)j * n]; }
        if (off <= 1e-32 * dia || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p + (size_t)q * n], app = a[p + (size_t)p * n], aqq = a[q + (size_t)q * n];
                if (std::fabs(apq) <= 1e-20 * (std::fabs(app) + std::fabs(aqq))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) { const double x = a[k + (size_t)p * n], y = a[k + (size_t)q * n]; a[k + (size_t)p * n] = c * x - s * y; a[k + (size_t)q * n] = s * x + c * y; }
                for (int k = 0; k < n; ++k) { const double x = a[p + (size_t)k * n], y = a[q + (size_t)k * n]; a[p + (size_t)k * n] = c * x - s * y; a[q + (size_t)k * n] = s * x + c * y; }
                for (int k = 0; k < n; ++k) { const double x = vec[k + (size_t)p * n], y = vec[k + (size_t)q * n]; vec[k + (size_t)p * n] = c * x - s * y; vec[k + (size_t)q * n] = s * x + c * y; }
            }
    }
    val.resize(n);
    for (int j = 0; j < n; ++j) val[j] = a[j + (size_t)j * n];
    for (int j = 0; j < n - 1; ++j) {
        int best = j;
        for (int l = j + 1; l < n; ++l) if (val[l] > val[best]) best = l;
        if (best != j) {
            std::swap(val[j], val[best]);
            for (int k = 0; k < n; ++k) std::swap(vec[k + (size_t)j * n], vec[k + (size_t)best * n]);
        }
    }
}

// nearPD(M, eig.tol = 1e-06, conv.tol = 1e-07, posd.tol = 1e-08, maxits = 100), R/nearPD.R
Vec near_pd(const Vec& M, int n) {
    const double eig_tol = 1e-06, conv_tol = 1e-07, posd_tol = 1e-08;
    if (n == 1) return Vec{std::fabs(M[0])};
    const size_t nn = (size_t)n * n;
    Vec U(nn, 0.0), X = M, Y, T(nn), Q, d;
    auto rebuild = [&](Vec& out, bool all) {           // Q[, p] %*% diag(d[p]) %*% t(Q[, p])
        std::fill(out.begin(), out.end(), 0.0);
        for (int k = 0; k < n; ++k) {
            if (!all && !(d[k] > eig_tol * d[0])) continue;
            for (int c = 0; c < n; ++c) { const double w = d[k] * Q[c + (size_t)k * n]; for (int r = 0; r < n; ++r) out[r + (size_t)c * n] += Q[r + (size_t)k * n] * w; }
        }
    };
    auto symmetrise = [&](Vec& A) {
        for (int r = 0; r < n; ++r) for (int c = r + 1; c < n; ++c) { const double m = (A[r + (size_t)c * n] + A[c + (size_t)r * n]) / 2; A[r + (size_t)c * n] = m; A[c + (size_t)r * n] = m; }
    };
    bool converged = false;
    for (int iter = 0; iter < 100 && !converged; ++iter) {
        Y = X;
        for (size_t j = 0; j < nn; ++j) T[j] = Y[j] - U[j];
        eigen_sym(Y, n, d, Q);
        rebuild(X, false);
        for (size_t j = 0; j < nn; ++j) U[j] = X[j] - T[j];
        symmetrise(X);
        double num = 0.0, den = 0.0;                   // inorm(x) = max(rowSums(abs(x)))
        for (int r = 0; r < n; ++r) {
            double a = 0.0, b = 0.0;
            for (int c = 0; c < n; ++c) { a += std::fabs(Y[r + (size_t)c * n] - X[r + (size_t)c * n]); b += std::fabs(Y[r + (size_t)c * n]); }
            num = std::max(num, a); den = std::max(den, b);
        }
        converged = (num / den) <= conv_tol;
    }
    symmetrise(X);
    eigen_sym(X, n, d, Q);
    const double Eps = posd_tol * std::fabs(d[0]);
    if (d[n - 1] < Eps) {
        for (int k = 0; k < n; ++k) if (d[k] < Eps) d[k] = Eps;
        Vec od(n);
        for (int j = 0; j < n; ++j) od[j] = X[j + (size_t)j * n];
        rebuild(X, true);
        Vec D(n);
        for (int j = 0; j < n; ++j) D[j] = std::sqrt(std::max(Eps, od[j]) / X[j + (size_t)j * n]);
        for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) X[r + (size_t)c * n] = D[r] * X[r + (size_t)c * n] * D[c];
    }
    Vec out(nn);
    for (int r = 0; r < n; ++r) for (int c = 0; c < n; ++c) out[r + (size_t)c * n] = (X[r + (size_t)c * n] + X[c + (size_t)r * n]) / 2;
    return out;
}

}  // namespace

extern "C" int jmb_marglogLik2(jmb_handle h, const double* Bs_gammas, const double* gammas, const double* alphas,
                               double tau_Bs_gammas, const jmb_priors* pr, double temp, int fixed_tau_Bs_gammas,
                               jmb_mll_out* out) {
    if (!h || !Bs_gammas || !alphas || !pr || !out || !out->par || !out->hessian)
        return jmb_set_error("jmb_marglogLik2: null argument");
    int32_t B = 0, p = 0, A = 0;
    if (jmb_param_dims(h, &B, &p, &A)) return 1;
    if (p && !gammas) return jmb_set_error("jmb_marglogLik2: gammas required");
    const int P = B + p + A, d = fixed_tau_Bs_gammas ? P : P + 1;
    if (fixed_tau_Bs_gammas && (!out->Cov_Bs_gammas || !out->Cov_alphas || (p && !out->Cov_gammas)))
        return jmb_set_error("jmb_marglogLik2: the Cov_* outputs are required with fixed_tau_Bs_gammas");
    Vec th(d);
    memcpy(th.data(), Bs_gammas, sizeof(double) * B);
    if (p) memcpy(th.data() + B, gammas, sizeof(double) * p);
    memcpy(th.data() + B + p, alphas, sizeof(double) * A);
    if (!fixed_tau_Bs_gammas) th[P] = tau_Bs_gammas;                     // log scale (R/margLogLik2.R:39-40)
    Vec score(P + 1);
    auto tau_of = [&](const Vec& x) { return fixed_tau_Bs_gammas ? tau_Bs_gammas : std::exp(x[P]); };
    Fn fn = [&](const Vec& x, double& v) {                               // :34-51
        double lp = 0.0;
        if (jmb_logPosterior(h, temp, x.data(), x.data() + B, x.data() + B + p, pr, tau_of(x), pr->A_tau_Bs_gammas, pr->B_tau_Bs_gammas, &lp)) return 1;
        v = 0.0 - lp;
        return 0;
    };
    Gr gr = [&](const Vec& x, Vec& g) {                                  // :52-73
        if (jmb_gradient_logPosterior(h, temp, x.data(), x.data() + B, x.data() + B + p, pr, tau_of(x), pr->A_tau_Bs_gammas, pr->B_tau_Bs_gammas, score.data())) return 1;
        for (int j = 0; j < d; ++j) g[j] = 0.0 - score[j];               // head(out, -1) when tau is fixed
        return 0;
    };
    Vec pscale(d, 1.0);
    if (!fixed_tau_Bs_gammas) pscale[d - 1] = 0.1;                        // :93
    OptimResult R;
    int rc = optim_bfgs(th, pscale, fn, gr, 100, 1.490116119384765625e-8, 1e-3, R);
    out->restarts = 0;
    if (rc == 10) {
        // R/margLogLik2.R:95-100: optim(BFGS) raised ("initial value in 'vmmin' is not finite") -> one restart from
        // vec_thetas + rnorm(d, sd = 0.1).  The reference draws from R's global stream; here the normals are a fixed
        // SplitMix64 / Box-Muller stream (the perturbation only has to leave the non-finite region).  The third attempt of the
        // reference, Nelder-Mead from the ORIGINAL vec_thetas (:101-104), evaluates the same non-finite value first and stops
        // ("function cannot be evaluated at initial parameters"), so after a failed restart the reference's
        // stop("failed to find initial values.") is what remains.
        Vec th2 = th;
        uint64_t st = 0x4D4C4C32u;                                            // "MLL2"
        auto next_u = [&]() { st += 0x9E3779B97F4A7C15ull; uint64_t z = st; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                              z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0); };
        for (int j = 0; j < d; j += 2) {
            const double u1 = next_u(), u2 = next_u(), rad = std::sqrt(-2.0 * std::log(u1));
            th2[j] += 0.1 * rad * std::cos(2.0 * 3.14159265358979323846 * u2);
            if (j + 1 < d) th2[j + 1] += 0.1 * rad * std::sin(2.0 * 3.14159265358979323846 * u2);
        }
        out->restarts = 1;
        rc = optim_bfgs(th2, pscale, fn, gr, 100, 1.490116119384765625e-8, 1e-3, R);
        if (rc == 10) return jmb_set_error("jmb_marglogLik2: failed to find initial values (non-finite log-posterior at the starting values and "
                                           "at the jittered restart, R/margLogLik2.R:95-107)");
    }
    if (rc) return 1;
    double logdet = 0.0;
    if (!lu_inverse(R.hessian, d, nullptr, &logdet)) logdet = NAN;
    out->value = 0.5 * (d * std::log(2.0 * 3.14159265358979323846) - logdet) - R.value;   // :110-111
    out->opt_value = R.value; out->convergence = R.convergence; out->counts[0] = R.fncount; out->counts[1] = R.grcount;
    memcpy(out->par, R.par.data(), sizeof(double) * d);
    memcpy(out->hessian, R.hessian.data(), sizeof(double) * (size_t)d * d);
    if (fixed_tau_Bs_gammas) {
        // hes(opt$par) = cd.vec(par, gr, eps = 0.001) (:74-87), solve, nearPD, diagonal blocks (:113-131)
        Vec res((size_t)d * d), x1(d), g1(d), g2(d);
        for (int i = 0; i < d; ++i) {
            const double ex = std::max(std::fabs(R.par[i]), 1.0);
            x1 = R.par;
            x1[i] = R.par[i] + 0.001 * ex;
            if (gr(x1, g1)) return 1;
            const double xp = x1[i];
            x1[i] = R.par[i] - 0.001 * ex;
            if (gr(x1, g2)) return 1;
            const double dx = xp - x1[i];
            for (int j = 0; j < d; ++j) res[j + (size_t)i * d] = (g1[j] - g2[j]) / dx;
        }
        Vec H((size_t)d * d), inv;
        for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) H[r + (size_t)c * d] = 0.5 * (res[r + (size_t)c * d] + res[c + (size_t)r * d]);
        if (!lu_inverse(H, d, &inv, nullptr)) return jmb_set_error("jmb_marglogLik2: the Hessian at the mode is singular");
        for (int r = 0; r < d; ++r) for (int c = 0; c < d; ++c) res[r + (size_t)c * d] = 0.5 * (inv[r + (size_t)c * d] + inv[c + (size_t)r * d]);
        const Vec iH = near_pd(res, d);
        for (int r = 0; r < B; ++r) for (int c = 0; c < B; ++c) out->Cov_Bs_gammas[r + (size_t)c * B] = iH[r + (size_t)c * d];
        for (int r = 0; r < p; ++r) for (int c = 0; c < p; ++c) out->Cov_gammas[r + (size_t)c * p] = iH[(B + r) + (size_t)(B + c) * d];
        for (int r = 0; r < A; ++r) for (int c = 0; c < A; ++c) out->Cov_alphas[r + (size_t)c * A] = iH[(B + p + r) + (size_t)(B + p + c) * d];
    }
    return 0;
}
