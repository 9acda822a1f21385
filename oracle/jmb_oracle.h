/*
 * jmb_oracle.h - CPU restatement of the reference's algorithm for the mvJointModelBayes hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or
 * call it, and only as the checker / the CPU baseline.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4 / 8c) and R, Rcpp and RcppArmadillo are absent from the image.  The pin is therefore
 * built here: the reference's OWN sources (/root/reference/src/fast_LAPRWM.cpp, fast_logPost.cpp,
 * fast_score.cpp) are compiled, unmodified and where they lie, against a stand-in for the missing
 * Armadillo/Rcpp headers (oracle/ref_shim/RcppArmadillo.h) into oracle/_ref/, and driven with the
 * same random streams; this restatement reproduces that build's lap_rwm_C chains, logPosterior
 * and score to <= 1e-11 (tests/test_reference_golden.py, fixtures in tests/golden/).  What the pin
 * does not cover: the real Armadillo/BLAS summation order and R's RNG (both unavailable).  It is
 * additionally checked against an independently written NumPy twin (oracle/np_twin.py), Random123
 * Philox known-answer vectors and analytic self-checks.
 *
 * The input structs mirror the reference's `Data`/`priors`/`initials`/`scales`/`Covs`/`control`
 * lists in the reference's own dense column-major layout (the oracle keeps that layout: dense
 * W1/W1s, materialised Wlong/Wlongs, cumsum-difference rowsum).  Their field order is kept
 * identical to include/jmb200.h so one marshaller can fill both.
 */
#ifndef JMB_ORACLE_H
#define JMB_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_data {
    int32_t n, n_outcomes, q, n_terms, n_alphas, n_Bs, n_gammas, K, interval_cens, multi_state;
    int64_t subject_offset;
    const int64_t* N; const int32_t* fams; const int32_t* links; const int32_t* q_k;
    const int32_t* const* RE_inds; const double* const* y; const double* const* Z;
    const int32_t* const* idL; const int32_t* const* idL2;
    const double* event; const double* W1; const double* W1s; const double* W1s_int;
    const double* W2; const double* W2s; const double* W2s_int; const double* Pw;
    const double* Pw_int; const int32_t* idGK_fast;
    const int32_t* r_t; const int32_t* const* RE_inds2; const int32_t* u_t;
    const int32_t* const* col_inds; const int32_t* trans_Funs;
    const double* const* ZZ; const double* const* ZZs; const double* const* ZZs_int;
    const double* const* U; const double* const* Us; const double* const* Us_int;
    const double* Covs_b;
    /* fields of jmb_data the oracle does not use (kept so that the two structs stay layout-compatible) */
    const double* knots; int32_t n_knots; int32_t ord; const double* Time; const double* st; const double* st_int;
    /* multi-state fits (multi_state != 0): nT transition rows in the survival part, src/fast_LAPRWM.cpp:519-529 */
    int32_t nT, nRisks; const int32_t* idT; const int32_t* kn_strat_first; const int32_t* kn_strat_last;
} orc_data;

typedef struct orc_draw {
    const double* const* Xbetas; const double* const* XXbetas; const double* const* XXsbetas;
    const double* const* XXsbetas_int; const double* sigmas; const double* invD;
} orc_draw;

typedef struct orc_priors {
    const double* mean_Bs_gammas; const double* Tau_Bs_gammas;
    const double* mean_gammas; const double* Tau_gammas;
    const double* mean_alphas; const double* Tau_alphas;
    double A_tau_Bs_gammas, B_tau_Bs_gammas, rank_Tau_Bs_gammas;
    int32_t shrink_gammas;
    double A_tau_gammas, B_tau_gammas, rank_Tau_gammas, A_phi_gammas, B_phi_gammas;
    int32_t shrink_alphas, double_gamma_alphas;
    double A_tau_alphas, B_tau_alphas, rank_Tau_alphas, A_phi_alphas, B_phi_alphas;
    double A_xi_alphas, B_xi_alphas, A_nu_alphas, B_nu_alphas;
    int32_t n_td; const int32_t* td_len; const int32_t* const* td_cols;
    double rank_Tau_td_alphas;
} orc_priors;

typedef struct orc_control {
    int32_t n_iter, n_burnin, n_block, n_thin;
    double target_acc, eps1, eps2, eps3, c0, c1;
    int32_t adaptCov; uint32_t seed; uint32_t chain_id;
} orc_control;

typedef struct orc_chain_in {
    const double* b; const double* Bs_gammas; const double* gammas; const double* alphas;
    const double* tau_Bs_gammas; double tau_gammas, tau_alphas;
    const double* phi_gammas; const double* phi_alphas; const double* tau_td_alphas;
    const double* sigma_b; double sigma_Bs_gammas, sigma_gammas, sigma_alphas;
    const double* Sigma_Bs_gammas; const double* Sigma_gammas; const double* Sigma_alphas;
} orc_chain_in;

typedef struct orc_chain_out {
    double* b; double* Bs_gammas; double* gammas; double* alphas; double* phi_gammas;
    double* phi_alphas; double* tau_Bs_gammas; double* tau_gammas; double* tau_alphas;
    double* tau_td_alphas; double* logWeights; double* sigma_b; double* sigma_surv;
    double* Sigma_Bs_gammas; double* Sigma_gammas; double* Sigma_alphas;
    double* accept_b; double* accept_surv; double* trace_surv;
} orc_chain_out;

typedef struct orc_eval_out {
    double* log_pyb; double* log_pb; double* log_h; double* H; double* HL; double* log_ptb;
} orc_eval_out;

/* rowsum flavour: 0 = literal cumsum difference (src/fast_LAPRWM.cpp:18-25),
 *                 1 = exact per-segment long-double sum (SURVEY section 7 hard part (a)) */
void orc_set_rowsum_mode(int mode);

/* counter RNG shared (by specification, not by code) with the CUDA path; see DESIGN.md */
void orc_philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                    uint32_t out[4]);
double orc_u01(uint32_t hi, uint32_t lo);
void orc_rnorm_pair(uint32_t seed, uint32_t chain, uint32_t c0, uint32_t c1, uint32_t c2,
                    uint32_t c3, double z[2]);
double orc_rgamma(uint32_t seed, uint32_t chain, uint32_t draw, uint32_t iter, double shape,
                  double scale);

void orc_rowsum(const double* x, int64_t N, const int32_t* last, int32_t n, double* out);
int orc_eval_logpost_re(const orc_data* d, const orc_draw* w, const double* b,
                        const double* Bs_gammas, const double* gammas, const double* alphas,
                        orc_eval_out* out);
/* fn(ctx, iter) is called after every iteration when non-NULL (sharded-run tests hook it) */
int orc_lap_rwm(const orc_data* d, const orc_draw* w, const orc_priors* pr, const orc_control* ct,
                const orc_chain_in* in, orc_chain_out* out, int max_iters);
/* hook for sharded runs: combine (sum over ranks) `len` doubles in place */
typedef void (*orc_allreduce_fn)(double* buf, int len, void* ctx);
void orc_set_allreduce(orc_allreduce_fn fn, void* ctx);

double orc_logPosterior(double temp, int32_t n, int64_t ns, int32_t B, int32_t p, int32_t A,
                        const double* event, const double* W1, const double* W1s,
                        const double* Bs_gammas, const double* W2, const double* W2s,
                        const double* gammas, const double* Wlong, const double* Wlongs,
                        const double* alphas, const double* Pw, const double* mean_Bs,
                        const double* Tau_Bs, const double* mean_g, const double* Tau_g,
                        const double* mean_a, const double* Tau_a, double tauBs, double A_tauBs,
                        double B_tauBs);
void orc_gradient_logPosterior(double temp, int64_t ns, int32_t B, int32_t p, int32_t A,
                               const double* ecs_W1, const double* W1s, const double* Bs_gammas,
                               const double* ecs_W2, const double* W2s, const double* gammas,
                               const double* ecs_Wlong, const double* Wlongs,
                               const double* alphas, const double* Pw, const double* mean_Bs,
                               const double* Tau_Bs, const double* mean_g, const double* Tau_g,
                               const double* mean_a, const double* Tau_a, double tauBs,
                               double A_tauBs, double B_tauBs, double* out);
/* lap_rwm_C_woRE[_nogammas] (src/fast_LAPRWM.cpp:934-1297): survival block only, fixed Wlong / Wlongs;
 * `logPost` receives out_logPost [n_out]; out->b, logWeights, sigma_b, accept_b are not touched */
int orc_lap_rwm_woRE(const orc_data* d, const double* Wlong, const double* Wlongs, const orc_priors* pr,
                     const orc_control* ct, const orc_chain_in* in, orc_chain_out* out, double* logPost);
/* Wlong (rows x A) for given b: lin_pred_matF, src/fast_LAPRWM.cpp:80-109.
 * which: 0 event-time designs, 1 quadrature designs, 2 interval quadrature designs */
int orc_wlong(const orc_data* d, const orc_draw* w, const double* b, int which, double* out);

#ifdef __cplusplus
}
#endif
#endif
