// jmb_device.cuh - device-side model descriptor, counter RNG and small helpers (sm_100a).
//
// Part of the B200-native mvJointModelBayes hot path; see DESIGN.md for the layout and the
// random-stream specification.  No code here comes from the reference: the reference's
// semantics are cited per function (paths relative to the JMbayes source tree).
#pragma once
#include <cstdint>
#include <utility>
#include "../../include/jmb200.h"
#include "jmb_model.h"

namespace jmb {

template <class F, int... I>
__device__ __forceinline__ void sfor_impl(F&& f, std::integer_sequence<int, I...>) {
    (f(std::integral_constant<int, I>{}), ...);
}
// compile-time loop: f(integral_constant<int, 0>) ... f(integral_constant<int, N-1>)
template <int N, class F>
__device__ __forceinline__ void sfor(F&& f) {
    sfor_impl(static_cast<F&&>(f), std::make_integer_sequence<int, N>{});
}


// ---- model descriptor passed to the kernels as a __grid_constant__ parameter -----------------
struct ModelMeta {
    ModelDesc d;                    // what the model is (jmb_model.h)
    RecordLayout L;                 // where each field lives inside a subject's records
    int B;                          // n_Bs (only shifts indices into the parameter block)
    int P;                          // B + p + A
};

// ---- per-chain parameter block in global memory (all doubles; offsets in ParamLayout) ---------
struct ParamLayout {
    int cur, prop;                  // [P] each: Bs_gammas | gammas | alphas
    int invD, sigmas;               // [q*q], [O]
    int tau_Bs, tau_g, tau_a, xi_a; // scalars (tau_Bs: [JMB_MAX_RISKS], one per stratum of a multi-state fit)
    int phi_g, phi_a, nu_a, tau_td; // [p], [A], [A], [n_td]
    int Tau_g, Tau_a, Tau_a_pen;    // [p*p], [A*A], [A*A]
    int sig;                        // [3] sigma_Bs, sigma_gammas, sigma_alphas
    int S_Bs, S_g, S_a;             // covariance matrices
    int H_Bs, H_g, H_a;             // their upper Cholesky factors
    int mean_Bs, Tau_Bs, mean_g, mean_a;  // priors
    int cur_lp;                     // carried current_logpost of the woRE flavour
    int block_par;                  // [P * n_block] states of the current block (adaptCov)
    int sums;                       // [3] reduced partial sums (multi-GPU path)
    int total;
};

struct Ctl {                        // device-resident loop control (ints)
    int iter, it, i, keep_iterator, check_iterator, accept_s_block, accept_s_total, error;
    int last_accept;                // 1 when the previous iteration accepted the survival proposal
};

struct ChainConst {                 // scalar constants of one chain (kernel parameter)
    int n_iter, n_burnin, n_block, n_thin, n_out, n_keep, start_cov_update, adaptCov;
    double target_acc, eps1, eps2, eps3, c0, mc1;
    uint32_t seed, chain_id;
    // priors (src/fast_LAPRWM.cpp:537-569)
    double B_tau_Bs, Apost_tau_Bs;
    int shrink_gammas; double B_tau_g, Apost_tau_g, B_phi_g, Apost_phi_g;
    int shrink_alphas, double_gamma_alphas;
    double B_tau_a, Apost_tau_a, B_phi_a, Apost_phi_a, B_xi_a, Apost_xi_a, B_nu_a, Apost_nu_a;
    int n_td; double B_tau_td, Apost_tau_td;
    int td_len[JMB_MAX_TERMS]; int td_cols[JMB_MAX_TERMS][JMB_MAX_UT];
    // strata of the baseline hazard (multi-state fits, src/fast_LAPRWM.cpp:489-491,777-782); nRisks = 0: unstratified
    int nRisks; int kn_first[JMB_MAX_RISKS]; int kn_last[JMB_MAX_RISKS];
};

enum { ST_B_PROP = 0, ST_B_ACC = 1, ST_SURV = 2, ST_GIBBS_N = 3, ST_GIBBS_U = 4, ST_GIBBS_B = 5 };

// ---- Philox4x32-10 (Salmon et al., SC'11); key = (seed, chain), counter = (c0,c1,c2,c3) -------
__device__ __forceinline__ void philox4x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3, uint32_t (&out)[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ double u01(uint32_t hi, uint32_t lo) {
    unsigned long long x = ((static_cast<unsigned long long>(hi) << 32) | lo) >> 11;
    return (static_cast<double>(x) + 0.5) * (1.0 / 9007199254740992.0);
}

// Box-Muller pair from one Philox block
__device__ __forceinline__ void rnorm_pair(uint32_t seed, uint32_t chain, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3, double& z0, double& z1) {
    uint32_t r[4];
    philox4x32(seed, chain, c0, c1, c2, c3, r);
    double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    z0 = rad * c;
    z1 = rad * s;
}

__device__ __forceinline__ double log_unif(uint32_t seed, uint32_t chain, uint32_t c0, uint32_t c1,
                                           uint32_t c2, uint32_t c3) {
    uint32_t r[4];
    philox4x32(seed, chain, c0, c1, c2, c3, r);
    return log(u01(r[0], r[1]));
}

// Gamma(shape, scale), Marsaglia & Tsang (2000); replaces Rf_rgamma (src/fast_LAPRWM.cpp:781)
__device__ inline double rgamma_dev(uint32_t seed, uint32_t chain, uint32_t draw, uint32_t iter,
                                    double shape, double scale) {
    double a = shape < 1.0 ? shape + 1.0 : shape;
    double dd = a - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd), x = 0.0;
    for (uint32_t attempt = 0; attempt < 1000u; ++attempt) {
        double z0, z1;
        rnorm_pair(seed, chain, draw, iter, attempt, ST_GIBBS_N, z0, z1);
        double t = 1.0 + cc * z0;
        if (t <= 0.0) continue;
        double v = t * t * t;
        double lu = log_unif(seed, chain, draw, iter, attempt, ST_GIBBS_U);
        if (lu < 0.5 * z0 * z0 + dd - dd * v + dd * log(v)) { x = dd * v; break; }
    }
    if (shape < 1.0) {
        uint32_t r[4];
        philox4x32(seed, chain, draw, iter, 0u, ST_GIBBS_B, r);
        x *= pow(u01(r[0], r[1]), 1.0 / shape);
    }
    return x * scale;
}

// Data$trans_Funs, src/fast_LAPRWM.cpp:91-106
__device__ __forceinline__ double trans_fun(int tf, double x) {
    switch (tf) {
        case JMB_TF_IDENTITY: return x;
        case JMB_TF_EXPIT: { double e = exp(x); return e / (1.0 + e); }
        case JMB_TF_EXP: return exp(x);
        case JMB_TF_LOG: return log(x);
        case JMB_TF_LOG2: return log2(x);
        case JMB_TF_LOG10: return log10(x);
        default: return sqrt(x);
    }
}

// per-row log-density with constants dropped, src/fast_LAPRWM.cpp:177-199 (naive formulas kept)
__device__ __forceinline__ double row_logdens(int fam, int link, double y, double eta, double inv_sigma) {
    if (fam == JMB_FAM_GAUSSIAN) {
        double t = (y - eta) * inv_sigma;
        return -0.5 * (t * t);
    } else if (fam == JMB_FAM_BINOMIAL) {
        double pr;
        if (link == JMB_LINK_LOGIT) { double e = exp(eta); pr = e / (1.0 + e); }
        else if (link == JMB_LINK_PROBIT) pr = 0.5 * erfc(-eta * 0.70710678118654752440);
        else pr = -exp(-exp(eta)) + 1.0;
        return y * log(pr) + (1.0 - y) * log(1.0 - pr);
    } else {
        double mu = exp(eta);
        return y * log(mu) - mu;
    }
}

// event / censoring term, src/fast_LAPRWM.cpp:640-650
__device__ __forceinline__ double log_ptb_term(int S, double event, double lh, double H, double HL) {
    if (S == 2) {
        double r = 0.0;
        if (event == 1.0) r += lh;
        if (event == 1.0 || event == 0.0) r -= H;
        if (event == 2.0) r += log(1.0 - exp(-H));
        if (event == 3.0) r += log(exp(-HL) - exp(-H));
        return r;
    }
    return event * lh - H;
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned mask) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) v += __shfl_xor_sync(mask, v, off, G);
    return v;
}

}  // namespace jmb
