/*
 * jmb200_survfit.h - C ABI of the B200-native survfitJM.mvJMbayes() hot loop (BASELINE config 5).
 *
 * Replaces, for all new subjects at once (three kernel launches: mode search, sampling, summaries), the per-subject R loops of
 * R/survfitJM.mvJMbayes.R (paths relative to the JMbayes source tree):
 *   :277-305  mode search  optim(start = 0, ff = -log_post_RE_svft, gg = cd(ff, eps = 1e-3), method = "BFGS",
 *             hessian = TRUE, control = list(maxit = 200, parscale = 0.1))
 *   :330-337  proposals    rmvt(M, mu = mode, Sigma = scale * solve(hessian), df = 4) and their dmvt()
 *   :338-398  for m in 1..M (sequential, b.old carries over): independence Metropolis-Hastings step on
 *             log_post_RE_svft under the parameters of posterior draw m, then survPred_svft_2 on each
 *             horizon interval, S = exp(cumsum(logS))
 *   :400-414  Mean / Median / Lower / Upper over the M draws
 * i.e. every .Call to log_post_RE_svft / survPred_svft_2 (src/survfitJM_mvJMbayes.cpp:190-271) that
 * survfitJM makes.  The per-draw linear predictors (Xbetas_calc, R/supportFuns_mvJointModelBayes.R:43-55;
 * called at :340,356,380) are evaluated on the device from the uploaded fixed-effects designs.
 *
 * Conventions as in jmb200.h: plain pointers and sizes, column-major doubles, 0-based int32 indices,
 * caller-allocated outputs, 0 on success / jmb_last_error() otherwise, no CPU fallback.
 *
 * Random numbers: R's stream (set.seed(seed), rnorm / rchisq / runif, :306-335,375) cannot be reproduced
 * without R; the draws come from the Philox4x32-10 streams specified in DESIGN.md section 4 (key
 * (seed, 0x53564654), counter (global subject, draw m, slot, stream)), so results agree with the
 * reference in distribution and with the oracle exactly.
 */
#ifndef JMB200_SURVFIT_H
#define JMB200_SURVFIT_H

#include "jmb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define JMB_SVFT_MAX_Q 8        /* random effects per subject (north_star: q <= 8) */
#define JMB_SVFT_MAX_PK 8       /* columns of one X_long[[k]] */
#define JMB_SVFT_MAX_FT 8       /* columns of one XXs[[t]] */

typedef struct jmb_svft_handle_s* jmb_svft_handle;

/*
 * survMats.last[[i]] (:149-245) and survMats[[i]][[l]] (:246-276) of all new subjects, flattened.
 * Longitudinal rows are grouped by subject (idL); a subject may have no rows in an outcome (its
 * measurements were NA, :159-161; log_longF_svft skips empty outcomes, src/survfitJM_mvJMbayes.cpp:109).
 * Quadrature rows on (0, last.time]: row i*K + k.  Horizon rows: row (i*L + l)*K + k for the l-th
 * prediction interval (times.to.pred_lower[l], times.to.pred_upper[l]] of subject i; only l <
 * n_horizons[i] is read.
 */
typedef struct jmb_svft_data {
    int32_t n;              /* new subjects held by this handle (this rank's shard) */
    int32_t n_outcomes, q, n_terms, n_alphas, n_Bs, n_gammas;
    int32_t K;              /* quadrature points (length(sk)) */
    int32_t L;              /* maximum number of prediction intervals per subject */
    int64_t subject_offset; /* global index of local subject 0 (keys the counter RNG) */
    /* longitudinal part, per outcome k */
    const int64_t* N;               /* [n_outcomes] rows */
    const int32_t* fams;            /* JMB_FAM_* */
    const int32_t* links;           /* JMB_LINK_* */
    const int32_t* q_k;             /* length(RE_inds[[k]]) */
    const int32_t* const* RE_inds;  /* [n_outcomes][q_k] */
    const int32_t* p_k;             /* ncol(X_long[[k]]) = length(betas[[k]]) */
    const double* const* y;         /* [n_outcomes][N_k]       y_long */
    const double* const* X;         /* [n_outcomes] N_k x p_k  X_long */
    const double* const* Z;         /* [n_outcomes] N_k x q_k  Z_long */
    const int32_t* const* idL;      /* [n_outcomes][N_k] subject of each row */
    /* association terms t */
    const int32_t* r_t;             /* ncol(ZZs[[t]]) = length(RE_inds2[[t]]) */
    const int32_t* const* RE_inds2;
    const int32_t* u_t;             /* ncol(Us[[t]]) = length(col_inds[[t]]) */
    const int32_t* const* col_inds;
    const int32_t* trans_Funs;      /* JMB_TF_* */
    const int32_t* outcome;         /* [n_terms] outcome whose betas term t uses (0-based) */
    const int32_t* f_t;             /* ncol(XXs[[t]]) = length(indFixed[[t]]) */
    const int32_t* const* indFixed; /* [n_terms][f_t] positions in betas[[outcome[t]]] (0-based) */
    /* quadrature designs on (0, last.time] (GK_points_postRE), nK = n*K rows */
    const double* W1s; const double* W2s; const double* Pw;   /* nK x n_Bs, nK x n_gammas, [nK] */
    const double* const* XXs; const double* const* ZZs; const double* const* Us;  /* [n_terms] nK x f_t / r_t / u_t */
    /* horizon designs (GK_points_CumHaz), nH = n*L*K rows */
    const int32_t* n_horizons;      /* [n] length(times.to.pred_upper[[i]]) <= L */
    const double* W1s_h; const double* W2s_h; const double* Pw_h;
    const double* const* XXs_h; const double* const* ZZs_h; const double* const* Us_h;
    /* Alternative to the dense bases: when W1s (resp. W1s_h) is NULL, the rows
     * splines::splineDesign(knots, x, ord, outer.ok = TRUE) (R/survfitJM.mvJMbayes.R:205,258) are evaluated by
     * the library from the quadrature times themselves - the dense nK x n_Bs / nH x n_Bs matrices (20 GB for the
     * horizons of 1M subjects) never exist.  n_knots = n_Bs + ord. */
    const double* knots; int32_t n_knots; int32_t ord;
    const double* st;               /* [nK]  c(t(GK_points_postRE)) */
    const double* st_h;             /* [nH]  GK_points_CumHaz rows, row (i*L + l)*K + k */
} jmb_svft_data;

/* Parameters: either the posterior means (M = 1; :117-128) or the M sub-sampled posterior draws
 * (:318-329), in R's own layout: draws are the ROWS of the mcmc matrices (leading dimension M). */
typedef struct jmb_svft_params {
    int32_t M;
    const double* const* betas;  /* [n_outcomes] M x p_k */
    const double* const* sigmas; /* [n_outcomes] [M] (NULL unless gaussian) */
    const double* invD;          /* M x q x q  (mcmc$inv_D) */
    const double* Bs_gammas;     /* M x n_Bs */
    const double* gammas;        /* M x n_gammas */
    const double* alphas;        /* M x n_alphas */
} jmb_svft_params;

typedef struct jmb_svft_control {
    double scale;        /* 1.6 */
    double df;           /* 4 */
    int32_t log;         /* log = FALSE: S = exp(cumsum(logS)); TRUE: logS per interval (:395) */
    double CI_levels[2]; /* 0.025, 0.975 */
    uint32_t seed;       /* seed = 1L */
    /* optim(): maxit = 200, parscale = 0.1, reltol = sqrt(.Machine$double.eps), ndeps = 1e-3; cd(): eps = 1e-3 */
    int32_t maxit; double parscale, reltol, ndeps, cd_eps;
} jmb_svft_control;

typedef struct jmb_svft_out {
    double* modes;       /* n x q      modes.b (:303) */
    double* hessians;    /* q x q x n  opt$hessian (:304-305); may be NULL in jmb_survfitJM (only Vars.b / invVars.b use it) */
    int32_t* convergence;/* [n] optim()$convergence: 0, 1 (maxit reached); 10 + x: could not be started / not PD */
    int32_t* counts;     /* 2 x n      optim()$counts (function, gradient) */
    double* summaries;   /* L x 4 x n  Mean, Median, Lower, Upper (:406-412); NaN beyond n_horizons[i] */
    double* full;        /* L x M x n  full.results (:397); may be NULL */
    double* accept_rate; /* [n] share of accepted proposals (additive; may be NULL) */
    double* b_last;      /* n x q  b.new after draw M (additive; may be NULL) */
} jmb_svft_out;

int jmb_svft_create(const jmb_svft_data* data, int device, jmb_svft_handle* out);
int jmb_svft_destroy(jmb_svft_handle h);

/* :277-305 only: modes.b and the optim Hessians under the posterior-mean parameters. */
int jmb_svft_modes(jmb_svft_handle h, const jmb_svft_params* post_means, const jmb_svft_control* control,
                   jmb_svft_out* out);
/* :306-414 only, from given modes / Hessians (out->modes, out->hessians are INPUTS here). */
int jmb_svft_sample(jmb_svft_handle h, const jmb_svft_params* draws, const jmb_svft_control* control,
                    jmb_svft_out* out);
/* The whole loop (:277-414): mode-search kernel, sampling kernel, summary kernel, back to back on one stream. */
int jmb_survfitJM(jmb_svft_handle h, const jmb_svft_params* post_means, const jmb_svft_params* draws,
                  const jmb_svft_control* control, jmb_svft_out* out);

/* device time of the last call's kernels (ms), launches so far, record bytes per subject */
int jmb_svft_info(jmb_svft_handle h, float* last_kernel_ms, int64_t* launches, double* record_bytes_per_subject);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_SURVFIT_H */
