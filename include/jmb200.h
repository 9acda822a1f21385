/*
 * jmb200.h - C ABI of the B200-native mvJointModelBayes MCMC hot path.
 *
 * This is the drop-in boundary an Rcpp shim binds instead of the reference's compiled bodies.
 * Each entry point cites the reference interface it replaces (paths relative to the JMbayes
 * source tree).  Conventions:
 *   - plain pointers and sizes only; every matrix is column-major double (R layout);
 *   - every index is 0-based int32 (the shim subtracts 1 where R is 1-based, exactly where
 *     List2Field_uvec does, src/fast_LAPRWM.cpp:45-56);
 *   - inputs are read-only and only need to stay valid for the duration of the call
 *     (the reference deep-copies its SEXPs, src/fast_LAPRWM.cpp:434-505);
 *   - outputs are caller-allocated;
 *   - every function returns 0 on success, non-zero on failure; jmb_last_error() gives the
 *     message (the shim turns it into Rcpp::stop, mirroring BEGIN_RCPP/END_RCPP,
 *     src/RcppExports.cpp:12,25).  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails.
 */
#ifndef JMB200_H
#define JMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Data$fams / Data$links (src/fast_LAPRWM.cpp:177-199) */
enum { JMB_FAM_GAUSSIAN = 0, JMB_FAM_BINOMIAL = 1, JMB_FAM_POISSON = 2 };
enum { JMB_LINK_IDENTITY = 0, JMB_LINK_LOGIT = 1, JMB_LINK_PROBIT = 2, JMB_LINK_CLOGLOG = 3,
       JMB_LINK_LOG = 4 };
/* Data$trans_Funs (src/fast_LAPRWM.cpp:91-106) */
enum { JMB_TF_IDENTITY = 0, JMB_TF_EXPIT = 1, JMB_TF_EXP = 2, JMB_TF_LOG = 3, JMB_TF_LOG2 = 4,
       JMB_TF_LOG10 = 5, JMB_TF_SQRT = 6 };

/* compile-time limits of the device model descriptor */
#define JMB_MAX_OUTCOMES 8
#define JMB_MAX_TERMS 16
#define JMB_MAX_Q 16
#define JMB_MAX_RT 8     /* columns of one ZZ[[t]] */
#define JMB_MAX_UT 8     /* columns of one U[[t]]  */
#define JMB_MAX_BS 64
#define JMB_MAX_GAMMAS 32
#define JMB_MAX_ALPHAS 32
#define JMB_MAX_RISKS 16  /* strata of the baseline hazard (multi-state fits) */

typedef struct jmb_handle_s* jmb_handle;

/*
 * The static part of the reference's `Data` list: everything that is identical for all the
 * M chains of one fit (built at R/mvJointModelBayes.R:651-714, read at
 * src/fast_LAPRWM.cpp:434-505).  Rows of each longitudinal outcome must be grouped by subject
 * (as extractFrames produces them, R/extractFrames.R:19), every subject needs >= 1 row per
 * outcome (R/mv_glmer.R:21); jmb_create() validates idL/idL2/idGK_fast against that.
 */
typedef struct jmb_data {
    int32_t n;              /* subjects held by this handle (this rank's shard) */
    int32_t n_outcomes;     /* length(Data$y) */
    int32_t q;              /* ncol(initials$b) */
    int32_t n_terms;        /* length(Data$XXbetas) */
    int32_t n_alphas;       /* ncol(Wlong) */
    int32_t n_Bs;           /* ncol(W1) */
    int32_t n_gammas;       /* ncol(W2) (0 -> the *_nogammas flavour) */
    int32_t K;              /* quadrature points per subject (15 or 7) */
    int32_t interval_cens;  /* lap_rwm_C argument `interval_cens` */
    int32_t multi_state;    /* lap_rwm_C argument `multiState`: the survival part has nT transition rows, several
                               per subject (see the multi-state fields at the end of this struct) */
    int64_t subject_offset; /* global index of local subject 0: keys the counter RNG so that a
                               sharded run reproduces the single-GPU chain */
    /* longitudinal part, per outcome k */
    const int64_t* N;               /* [n_outcomes] rows of outcome k */
    const int32_t* fams;            /* [n_outcomes] JMB_FAM_* */
    const int32_t* links;           /* [n_outcomes] JMB_LINK_* */
    const int32_t* q_k;             /* [n_outcomes] length(RE_inds[[k]]) */
    const int32_t* const* RE_inds;  /* [n_outcomes][q_k] columns of b used by outcome k */
    const double* const* y;         /* [n_outcomes][N_k] */
    const double* const* Z;         /* [n_outcomes] N_k x q_k */
    const int32_t* const* idL;      /* [n_outcomes][N_k] subject of each row */
    const int32_t* const* idL2;     /* [n_outcomes][n] last row of each subject */
    /* survival part */
    const double* event;            /* [n] 0/1, or 0/1/2/3 with interval censoring */
    const double* W1;               /* n x n_Bs */
    const double* W1s;              /* nK x n_Bs */
    const double* W1s_int;          /* nK x n_Bs (interval_cens only, else NULL) */
    const double* W2;               /* n x n_gammas */
    const double* W2s;              /* nK x n_gammas */
    const double* W2s_int;          /* nK x n_gammas (interval_cens only) */
    const double* Pw;               /* [nK] */
    const double* Pw_int;           /* [nK] (interval_cens only) */
    const int32_t* idGK_fast;       /* [n] last quadrature row of each subject, (i+1)K-1 */
    /* association terms t */
    const int32_t* r_t;             /* [n_terms] ncol(ZZ[[t]]) = length(RE_inds2[[t]]) */
    const int32_t* const* RE_inds2; /* [n_terms][r_t] */
    const int32_t* u_t;             /* [n_terms] ncol(U[[t]]) = length(col_inds[[t]]) */
    const int32_t* const* col_inds; /* [n_terms][u_t] columns of Wlong owned by term t */
    const int32_t* trans_Funs;      /* [n_terms] JMB_TF_* */
    const double* const* ZZ;        /* [n_terms] n x r_t */
    const double* const* ZZs;       /* [n_terms] nK x r_t */
    const double* const* ZZs_int;   /* [n_terms] nK x r_t (interval_cens only) */
    const double* const* U;         /* [n_terms] n x u_t */
    const double* const* Us;        /* [n_terms] nK x u_t */
    const double* const* Us_int;    /* [n_terms] nK x u_t (interval_cens only) */
    /* Covs$b: q x q x n upper Cholesky factors (R/mvJointModelBayes.R:728-729) */
    const double* Covs_b;
    /* Alternative to the dense bases: when W1 is NULL, the rows of
     * splines::splineDesign(knots, x, ord, outer.ok = TRUE) (R/mvJointModelBayes.R:382-383,389-392) are evaluated by
     * the library from Time / st / st_int - the dense n x n_Bs and nK x n_Bs matrices (2 x 2 GB at 1M subjects) never
     * exist.  n_knots = n_Bs + ord.  W1s / W1s_int are then ignored. */
    const double* knots; int32_t n_knots; int32_t ord;
    const double* Time;             /* [n]  event times */
    const double* st;               /* [nK] c(t(st)): quadrature times */
    const double* st_int;           /* [nK] (interval_cens only) */
    /* Multi-state fits (multi_state != 0; src/fast_LAPRWM.cpp:519-529, R/mvJointModelBayes.R:76-101,559-570,651-670):
     * every array documented above with n (or nK) survival rows - event, W1, W2, ZZ[[t]], U[[t]], XXbetas[[t]] and
     * W1s, W2s, Pw, ZZs[[t]], Us[[t]], XXsbetas[[t]], idGK_fast - then has nT (or nT * K) rows, one per transition a
     * subject is at risk for, grouped by subject in subject order; n stays the number of subjects (rows of b).
     * Ignored when multi_state == 0.  interval_cens and the knots alternative cannot be combined with it. */
    int32_t nT;                     /* length(Data$event) */
    int32_t nRisks;                 /* Data$nRisks: strata of the baseline hazard */
    const int32_t* idT;             /* [nT] subject of each transition row = Data$idT2[[t]] - 1 (the same for every t);
                                       rows_wlong / rows_wlongs / idT2s / idT_rsum follow from it */
    const int32_t* kn_strat_first;  /* [nRisks] first / last column of W1 owned by each stratum (0-based, as R builds */
    const int32_t* kn_strat_last;   /*          them, R/mvJointModelBayes.R:668-669) */
} jmb_data;

/* The per-draw part of `Data` (R/mvJointModelBayes.R:760-770). */
typedef struct jmb_draw {
    const double* const* Xbetas;       /* [n_outcomes][N_k] */
    const double* const* XXbetas;      /* [n_terms][n] */
    const double* const* XXsbetas;     /* [n_terms][nK] */
    const double* const* XXsbetas_int; /* [n_terms][nK] (interval_cens only) */
    const double* sigmas;              /* [n_outcomes] (ignored unless gaussian) */
    const double* invD;                /* q x q */
} jmb_draw;

/* Fixed-effects designs of the fit (Data$X, Data$XX, Data$XXs, Data$XXs_int; R/mvJointModelBayes.R:234,407-467) for
 * jmb_set_xdesigns / jmb_set_draw_betas: the per-draw linear predictors of Xbetas_calc
 * (R/supportFuns_mvJointModelBayes.R:43-55, called for every draw at R/mvJointModelBayes.R:760-768) are then
 * computed on the device from the draw's betas instead of being sent as N-vectors. */
typedef struct jmb_xdesigns {
    const int32_t* p_k;             /* [n_outcomes] ncol(X[[k]]) = length(betas[[k]]) */
    const double* const* X;         /* [n_outcomes] N_k x p_k */
    const int32_t* outcome;         /* [n_terms] outcome whose betas term t uses (0-based) */
    const int32_t* f_t;             /* [n_terms] ncol(XX[[t]]) = length(indFixed[[t]]) */
    const int32_t* const* indFixed; /* [n_terms][f_t] positions in betas[[outcome[t]]] (0-based) */
    const double* const* XX;        /* [n_terms] n x f_t */
    const double* const* XXs;       /* [n_terms] nK x f_t */
    const double* const* XXs_int;   /* [n_terms] nK x f_t (interval_cens only) */
} jmb_xdesigns;

/* `priors` (src/fast_LAPRWM.cpp:531-569) */
typedef struct jmb_priors {
    const double* mean_Bs_gammas; const double* Tau_Bs_gammas;  /* [B], B x B */
    const double* mean_gammas;    const double* Tau_gammas;     /* [p], p x p */
    const double* mean_alphas;    const double* Tau_alphas;     /* [A], A x A */
    double A_tau_Bs_gammas, B_tau_Bs_gammas, rank_Tau_Bs_gammas;
    int32_t shrink_gammas;
    double A_tau_gammas, B_tau_gammas, rank_Tau_gammas, A_phi_gammas, B_phi_gammas;
    int32_t shrink_alphas, double_gamma_alphas;
    double A_tau_alphas, B_tau_alphas, rank_Tau_alphas, A_phi_alphas, B_phi_alphas;
    double A_xi_alphas, B_xi_alphas, A_nu_alphas, B_nu_alphas;
    int32_t n_td;                    /* length(td_cols) */
    const int32_t* td_len;           /* [n_td] */
    const int32_t* const* td_cols;   /* [n_td][td_len] */
    double rank_Tau_td_alphas;
} jmb_priors;

/* `control` (src/fast_LAPRWM.cpp:572-582) plus the RNG key */
typedef struct jmb_control {
    int32_t n_iter, n_burnin, n_block, n_thin;
    double target_acc, eps1, eps2, eps3, c0, c1;
    int32_t adaptCov;
    uint32_t seed;       /* control$seed */
    uint32_t chain_id;   /* index of the mvglmer draw: second half of the Philox key */
} jmb_control;

/* `initials` + `scales` + `Covs` (src/fast_LAPRWM.cpp:507-518,594-601) */
typedef struct jmb_chain_in {
    const double* b;              /* n x q */
    const double* Bs_gammas; const double* gammas; const double* alphas;
    const double* tau_Bs_gammas;  /* [1]; [nRisks] for a multi-state fit (R/mvJointModelBayes.R:723-724) */
    double tau_gammas, tau_alphas;
    const double* phi_gammas;     /* [p] */
    const double* phi_alphas;     /* [A] */
    const double* tau_td_alphas;  /* [n_td] */
    const double* sigma_b;        /* scales$b [n] */
    double sigma_Bs_gammas, sigma_gammas, sigma_alphas;
    const double* Sigma_Bs_gammas; const double* Sigma_gammas; const double* Sigma_alphas;
} jmb_chain_in;

/* The list lap_rwm_C returns (src/fast_LAPRWM.cpp:870-895); n_out = floor(n_iter/n_thin). */
typedef struct jmb_chain_out {
    double* b;               /* n x q x n_out */
    double* Bs_gammas;       /* B x n_out */
    double* gammas;          /* p x n_out */
    double* alphas;          /* A x n_out */
    double* phi_gammas;      /* p x n_out */
    double* phi_alphas;      /* A x n_out */
    double* tau_Bs_gammas;   /* 1 x n_out; nRisks x n_out for a multi-state fit (src/fast_LAPRWM.cpp:607) */
    double* tau_gammas;      /* [n_out] */
    double* tau_alphas;      /* [n_out] */
    double* tau_td_alphas;   /* n_td x n_out */
    double* logWeights;      /* [n_out] */
    double* sigma_b;         /* [n] adapted scales */
    double* sigma_surv;      /* [3] sigma_Bs_gammas, sigma_gammas, sigma_alphas */
    double* Sigma_Bs_gammas; double* Sigma_gammas; double* Sigma_alphas;
    /* additive extras (not in the reference's list; may be NULL) */
    double* accept_b;        /* [n] overall acceptance rate of each subject's b-step */
    double* accept_surv;     /* [1] overall acceptance rate of the survival block */
    double* trace_surv;      /* [2 * total iterations] (sum log_ptb current, proposed) per iteration */
} jmb_chain_out;

/* Per-subject pieces of the b-step target for given (b, Bs_gammas, gammas, alphas): the
 * quantities initialised at src/fast_LAPRWM.cpp:619-650 (== log_postREF, :219-248). */
typedef struct jmb_eval_out {
    double* log_pyb;  /* [n] */
    double* log_pb;   /* [n] */
    double* log_h;    /* [n]; [nT] for a multi-state fit (one value per transition row) */
    double* H;        /* [n]; [nT] */
    double* HL;       /* [n] (interval_cens only; may be NULL); [nT] */
    double* log_ptb;  /* [n]; [nT] - the b-step uses rowsum(log_ptb, idT_rsum), src/fast_LAPRWM.cpp:669 */
} jmb_eval_out;

const char* jmb_last_error(void);
int jmb_device_count(int* count);

/* Upload the static designs of one fit (replaces the unpacking at src/fast_LAPRWM.cpp:434-505,
 * done once per fit instead of once per chain).  `device` is the CUDA ordinal. */
int jmb_create(const jmb_data* data, int device, jmb_handle* out);
int jmb_destroy(jmb_handle h);

/* Per-draw vectors of one chain (R/mvJointModelBayes.R:760-770). */
int jmb_set_draw(jmb_handle h, const jmb_draw* draw);

/* Upload the fixed-effects designs once per fit; afterwards jmb_set_draw_betas (betas[[k]] of the draw, sigmas,
 * invD) is equivalent to Xbetas_calc on the host + jmb_set_draw, with a few hundred bytes of host-to-device
 * traffic per draw instead of the N-vectors. */
int jmb_set_xdesigns(jmb_handle h, const jmb_xdesigns* x);
int jmb_set_draw_betas(jmb_handle h, const double* const* betas, const double* sigmas, const double* invD);

/* lap_rwm_C (src/fast_LAPRWM.cpp:431-896) for the current draw. */
int jmb_lap_rwm(jmb_handle h, const jmb_priors* priors, const jmb_control* control,
                const jmb_chain_in* in, jmb_chain_out* out);

/* Per-subject log-posterior pieces at (b, params) for the current draw
 * (src/fast_LAPRWM.cpp:619-650); the 1e-10 parity probe. */
int jmb_eval_logpost_re(jmb_handle h, const double* b, const double* Bs_gammas,
                        const double* gammas, const double* alphas, jmb_eval_out* out);

/* Survival-block Laplace step on fixed Wlong (n x A) / Wlongs (nK x A), the functions marglogLik2
 * drives through optim(BFGS) (R/margLogLik2.R:42-65,95):
 *   jmb_set_wlong           uploads the two association designs (Data$Wlong, Data$Wlongs);
 *   jmb_logPosterior        logPosterior[_nogammas]         (src/fast_logPost.cpp:8-26,30-46)
 *   jmb_gradient_logPosterior  gradient_logPosterior[_nogammas] (src/fast_score.cpp:8-50,53-84),
 *                           out has n_Bs + n_gammas + n_alphas + 1 entries.
 * event_colSums* are derived from the handle's own event/W1/W2/Wlong (R/mvJointModelBayes.R:677-681).
 * Only the mean_* / Tau_* fields of `priors` are read.  With several ranks the data sums are
 * all-reduced, every rank returns the full-cohort value. */
int jmb_set_wlong(jmb_handle h, const double* Wlong, const double* Wlongs);
int jmb_logPosterior(jmb_handle h, double temp, const double* Bs_gammas, const double* gammas,
                     const double* alphas, const jmb_priors* priors, double tauBs, double A_tauBs,
                     double B_tauBs, double* out);
int jmb_gradient_logPosterior(jmb_handle h, double temp, const double* Bs_gammas,
                              const double* gammas, const double* alphas, const jmb_priors* priors,
                              double tauBs, double A_tauBs, double B_tauBs, double* out);

/* marglogLik2(thetas, Data, priors, temp, fixed_tau_Bs_gammas) (R/margLogLik2.R:1-140; called at
 * R/mvJointModelBayes.R:716-725 for the initial values / proposal covariances and, with update_RE = FALSE, once per draw
 * at :792): optim(BFGS, hessian = TRUE) on -logPosterior with the analytic score, the Laplace value, and - with
 * fixed_tau_Bs_gammas - cd.vec + solve + nearPD for attr(, "Covs").  The BFGS driver runs natively on the host; every
 * function value / gradient is a pass over the designs on the device (jmb_set_wlong first).  d = n_Bs + n_gammas + n_alphas
 * (+ 1 when tau_Bs_gammas is free; then `tau_Bs_gammas` is on the log scale, :39-40).  SURVEY.md section 8(f)2. */
typedef struct jmb_mll_out {
    double value;            /* the Laplace approximation marglogLik2 returns (:110-111) */
    double opt_value;        /* opt$value */
    int32_t convergence;     /* opt$convergence */
    int32_t counts[2];       /* opt$counts */
    double* par;             /* [d]   opt$par (attr(, "inits") = its Bs_gammas | gammas | alphas blocks) */
    double* hessian;         /* d x d opt$hessian */
    double* Cov_Bs_gammas; double* Cov_gammas; double* Cov_alphas;   /* attr(, "Covs") (fixed_tau_Bs_gammas only) */
    int32_t restarts;        /* 1 when optim(BFGS) raised at the starting values and the jittered restart of :95-100 was used */
} jmb_mll_out;
int jmb_marglogLik2(jmb_handle h, const double* Bs_gammas, const double* gammas, const double* alphas,
                    double tau_Bs_gammas, const jmb_priors* priors, double temp, int fixed_tau_Bs_gammas,
                    jmb_mll_out* out);
int jmb_param_dims(jmb_handle h, int32_t* n_Bs, int32_t* n_gammas, int32_t* n_alphas);

/* lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (src/fast_LAPRWM.cpp:934-1139,1142-1297; control$update_RE =
 * FALSE): the survival block alone on the association designs of jmb_set_wlong.  `in->b`, `in->sigma_b`
 * and the b-related outputs are not used; logPost [n_out] receives the list element `logPost`. */
int jmb_lap_rwm_woRE(jmb_handle h, const jmb_priors* priors, const jmb_control* control,
                     const jmb_chain_in* in, jmb_chain_out* out, double* logPost);

/* survfitJM.mvJMbayes evaluators (R/survfitJM.mvJMbayes.R:297,372-373,393), batched over the subjects
 * of a handle built from the new subjects' data (longitudinal rows up to last.time, quadrature designs
 * on the interval of interest; event / W1 / W2 / ZZ / U of the event-time row are not used and may be
 * zero; subjects may have no rows in some outcomes):
 *   jmb_log_post_RE_svft  log_post_RE_svft (src/survfitJM_mvJMbayes.cpp:190-238): out[i] =
 *                         log p(y_i|b_i) + log p(b_i) - H_i, parameters of the current draw
 *   jmb_survPred_svft_2   survPred_svft_2 (:241-271): out[i] = -H_i on the handle's quadrature grid */
int jmb_log_post_RE_svft(jmb_handle h, const double* b, const double* Bs_gammas, const double* gammas,
                         const double* alphas, double* out);
int jmb_survPred_svft_2(jmb_handle h, const double* b, const double* Bs_gammas, const double* gammas,
                        const double* alphas, double* out);

/* Layout facts for the parity tests: the CSR row pointers the device layout was built with
 * (must equal idL2-derived segments bit for bit) and bytes moved per subject-iteration. */
int jmb_get_row_ptr(jmb_handle h, int outcome, int64_t* row_ptr /* [n+1] */);
int jmb_layout_info(jmb_handle h, double* shared_bytes_per_subject,
                    double* chain_bytes_per_subject, int32_t* w1_band, int32_t* group_lanes);

/* Layout inspection without a device, for the CPU tests of the record packer (never part of a compute path; a handle
 * made by jmb_debug_pack_host holds no device state and must only be passed to these functions and jmb_destroy).  jmb_debug_layout: ints[0..15] = RP,
 * o_V, o_W2c, o_Pw, o_W1off, o_W1v, o_W2, o_long, c_long, HB, state_stride, WB, G, w2_const, MS, S; ints[16+t] = o_ZZ[t];
 * ints[32+t] = o_U[t]; ints[48+t] bit 0 = first ZZ column not stored (all ones), bit 1 = U not stored (all ones), and for
 * outcome o < 8, ints[56+o] bit 8 = first Z column not stored, bits 16.. = stored doubles per longitudinal row
 * (jmbayes_b200/csrc/jmb_model.h describes the records). */
int jmb_debug_pack_host(const jmb_data* data, jmb_handle* out);
int jmb_debug_layout(jmb_handle h, int32_t* ints /* [64] */, int64_t* doff /* [n+1] */, int64_t* coff /* [n+1] */);
int jmb_debug_records(jmb_handle h, const jmb_draw* draw, double* drec, double* crec);

/* Multi-GPU: one process per GPU, contiguous subject shards, one all-reduce of the
 * survival-block partial sums per iteration.  The 128-byte NCCL unique id
 * is created on rank 0 and distributed by the host language (torch.distributed / Rmpi). */
int jmb_comm_unique_id(char id[128]);
int jmb_comm_init(jmb_handle h, int nranks, int rank, const char id[128]);
/* Fused compute + collective: after jmb_p2p_attach the finalize kernel all-reduces the three
 * per-iteration sums itself through NVLink peer memory (each rank stores into every peer's mailbox and
 * waits for the peers' stores), replacing the reduce kernel + ncclAllReduce of the per-iteration path.
 * jmb_p2p_export returns the 64-byte CUDA IPC handle of this rank's mailbox; `handles` holds the
 * nranks handles in rank order (gathered by the host language).  jmb_comm_init is still required (the
 * Laplace-step sums use NCCL). */
int jmb_p2p_export(jmb_handle h, int nranks, char handle[64]);
int jmb_p2p_attach(jmb_handle h, int nranks, int rank, const char* handles);

/* Benchmark hooks: run `iters` iterations of the current chain state entirely on the device
 * (no host copies) and report the device time of the loop; jmb_chain_begin uploads state. */
int jmb_chain_begin(jmb_handle h, const jmb_priors* priors, const jmb_control* control,
                    const jmb_chain_in* in);
int jmb_chain_iterate(jmb_handle h, int iters, float* elapsed_ms, float* step_kernel_ms);
/* Same, with `lead` untimed iterations enqueued right before the timed ones: on several ranks the timed window then
 * starts after the all-reduce of the last lead iteration (a device-side rendezvous), not after a host barrier. */
int jmb_chain_iterate_after(jmb_handle h, int lead, int iters, float* elapsed_ms);
int jmb_chain_end(jmb_handle h, jmb_chain_out* out);
int jmb_launch_count(jmb_handle h, int64_t* launches);
int jmb_launch_count_timed(jmb_handle h, int64_t* launches);   /* kernels launched between the two events of the last jmb_chain_iterate[_after] */

/* Several chains (mvglmer draws) of one fit in flight on one device - the reference's chain
 * parallelism (foreach over blocks of draws, R/mvJointModelBayes.R:860-905).  jmb_chain_slots makes
 * `count` chain slots (slot 0 always exists); jmb_select_chain chooses the slot that jmb_set_draw,
 * jmb_chain_begin, jmb_chain_iterate, jmb_chain_end and jmb_lap_rwm act on; jmb_chains_iterate
 * advances slots 0..nslots-1 together, round-robin on their own streams. */
int jmb_chain_slots(jmb_handle h, int count);
int jmb_select_chain(jmb_handle h, int slot);
int jmb_chains_iterate(jmb_handle h, int nslots, int iters, float* elapsed_ms);
/* "static:<shape>" when a compile-time-specialised step kernel serves this model, else "generic" */
const char* jmb_kernel_name(jmb_handle h);
void* jmb_stream(jmb_handle h);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_H */
