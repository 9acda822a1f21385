/*
 * jmb_oracle_mll.h - CPU restatement of marglogLik2() (see jmb_oracle_mll.c).  TEST INFRASTRUCTURE ONLY.
 */
#ifndef JMB_ORACLE_MLL_H
#define JMB_ORACLE_MLL_H
#include "jmb_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef double (*orc_obj_fn)(int n, const double* x, void* ctx);
typedef void (*orc_obj_gr)(int n, const double* x, double* g, void* ctx);

/* the fields of `Data` that marglogLik2 reads (R/margLogLik2.R:3-14), dense column-major */
typedef struct orc_mll_data {
    int32_t n; int64_t ns; int32_t B, p, A;
    const double* event; const double* W1; const double* W1s; const double* W2; const double* W2s;
    const double* Wlong; const double* Wlongs; const double* Pw;
} orc_mll_data;

/* d = B + p + A (+ 1 when tau_Bs_gammas is free; then thetas[d-1] = log tau_Bs_gammas) */
typedef struct orc_mll_out {
    double value;            /* the Laplace approximation returned by marglogLik2 (:110-111) */
    double opt_value;        /* opt$value = -logPosterior at the mode */
    int32_t convergence; int32_t counts[2];
    double* par;             /* [d] opt$par */
    double* hessian;         /* d x d opt$hessian */
    double* Cov_Bs_gammas; double* Cov_gammas; double* Cov_alphas;   /* attr(, "Covs") when fixed_tau_Bs_gammas */
} orc_mll_out;

int orc_optim_bfgs(int n, const double* par, const double* parscale, orc_obj_fn fn, orc_obj_gr gr, void* ctx,
                   int maxit, double reltol, double ndeps, double* out_par, double* out_value, double* out_hessian,
                   int* counts);
int orc_lu_inverse(const double* A, int n, double* inv, double* logabsdet);
void orc_eigen_sym_n(const double* S, int n, double* val, double* vec);
void orc_nearPD(const double* M, int n, double* out);
int orc_marglogLik2(const orc_mll_data* D, const orc_priors* pr, const double* thetas, double tau_Bs_gammas, double temp,
                    int fixed_tau_Bs_gammas, orc_mll_out* out);
#ifdef __cplusplus
}
#endif
#endif
