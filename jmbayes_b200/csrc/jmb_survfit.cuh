// jmb_survfit.cuh - the survfitJM.mvJMbayes() loop of one new subject, fused into one kernel (sm_100a).
//
// Reference (paths in the JMbayes tree): R/survfitJM.mvJMbayes.R:277-414 drives, per subject, optim(BFGS)
// on -log_post_RE_svft with the central-difference gradient of R/cd.R, optim's numerical Hessian, M
// independence Metropolis-Hastings steps with t proposals (R/rmvt.R, R/dmvt.R) under the parameters of M
// posterior draws, and after each step one survPred_svft_2 per prediction interval; every evaluation is
// a .Call into src/survfitJM_mvJMbayes.cpp:140-187.  optim's BFGS (vmmin) and optimhess are R's own C code
// (src/appl/optim.c, src/library/stats/src/optim.c), restated here from the published algorithm.
//
// Mapping: one 16-lane group per subject (two subjects per warp); all 16 lanes evaluate log_post_RE_svft /
// survPred_svft_2 together - one Gauss-Kronrod node per lane, the visit rows lane-strided.
//   jmb_svft_modes_*  : mode search.  Lane 0 of the group is the subject's control thread: a resumable vmmin state
//                       machine that posts coarse requests (one function value, or a whole central-difference
//                       gradient = 2q values), so both groups of a warp always meet in the same evaluation loop no
//                       matter where each subject is in its own line search; groups pull subjects from an atomic
//                       counter.
//   jmb_svft_sample_* : the M Metropolis-Hastings steps and the horizon integrals as straight-line code; the
//                       vector work (proposal, quadratic forms, state copies) is spread over the lanes and the two
//                       subjects of a warp stay in lock-step (every step costs 2 + n_horizons evaluations).
//   jmb_svft_summary  : per subject and interval, sort of the M values and R's type-7 summaries.
// A subject's design record is read from HBM once per kernel and stays L1/L2-resident for its ~3000 evaluations:
// the path is bound by fp64 issue, not by HBM.
#pragma once
#include "../../include/jmb200_survfit.h"
#include "jmb_device.cuh"

namespace jmb {

constexpr int kSvQ = JMB_SVFT_MAX_Q, kSvPK = JMB_SVFT_MAX_PK, kSvFT = JMB_SVFT_MAX_FT;
constexpr int kSvG = 16;                     // lanes per subject
constexpr uint32_t kSvKey = 0x53564654u;     // second half of the Philox key of the survfit streams

// model shape of a survfit handle; constexpr-constructible so that precompiled shapes fold away
struct SvDesc {
    int O, q, T, A, p, K, WB, w2_const, NB;  // NB = sum of p_k
    int NC;                                  // distinct non-constant XXs / ZZs columns stored per quadrature row
    int fam[kMaxO], link[kMaxO], qk[kMaxO], z_ones0[kMaxO], pk[kMaxO], x_ones0[kMaxO], boff[kMaxO];
    int re_inds[kMaxO][kSvQ];
    int rt[kMaxT], ut[kMaxT], tf[kMaxT], u_ones[kMaxT], ft[kMaxT], outc[kMaxT];
    int re2[kMaxT][kMaxRT], col[kMaxT][kMaxUT], indf[kMaxT][kSvFT];
    // where column f of XXs[[t]] / column j of ZZs[[t]] lives: -1 = all ones (not stored), else pooled column
    int xx_col[kMaxT][kSvFT], zz_col[kMaxT][kMaxRT];
};

// Record of one subject (doubles): int32 visit counts [O] + number of prediction intervals | W2 row (when the
// W2s rows repeat) | nh + 1 quadrature blocks (block 0: (0, last.time], block l + 1: interval l), each field
// stored [column][16 lanes] | longitudinal rows per outcome: y[v] | stored X columns | stored Z columns.
// All-ones columns of XXs / ZZs / Us and all-ones first columns of X / Z are not stored, and XXs / ZZs columns
// that are identical over all rows (e.g. the time variable of every value term) are stored once (decided from
// the data by jmb_svft_create, so it never changes results).
struct SvLayout {
    int o_V, o_W2c, o_blk0, blk;
    int b_Pw, b_W1off, b_W1v, b_W2, b_col, b_U[kMaxT];
    int long_cols[kMaxO];
    int ps_betas, ps_isig, ps_gam, ps_alp, ps_invD, ps_Bs;   // parameter set: ... | invD q x q | Bs_gammas [B]
};

JMB_HD constexpr SvLayout sv_make_layout(const SvDesc& d) {
    SvLayout L{};
    L.o_V = 0;
    int o = (d.O + 1 + 1) / 2;
    L.o_W2c = o;
    if (d.w2_const) o += d.p;
    L.o_blk0 = o;
    int b = 0;
    L.b_Pw = b; b += kSvG;
    L.b_W1off = b; b += kSvG / 2;
    L.b_W1v = b; b += d.WB * kSvG;
    L.b_W2 = b;
    if (!d.w2_const) b += d.p * kSvG;
    L.b_col = b; b += d.NC * kSvG;
    for (int t = 0; t < d.T; ++t) { L.b_U[t] = b; b += (d.u_ones[t] ? 0 : d.ut[t]) * kSvG; }
    L.blk = b;
    for (int k = 0; k < d.O; ++k) L.long_cols[k] = 1 + (d.pk[k] - d.x_ones0[k]) + (d.qk[k] - d.z_ones0[k]);
    int s = 0;
    L.ps_betas = s; s += d.NB;
    L.ps_isig = s; s += d.O;
    L.ps_gam = s; s += d.p;
    L.ps_alp = s; s += d.A;
    L.ps_invD = s; s += d.q * d.q;
    L.ps_Bs = s;
    return L;
}

JMB_HD constexpr bool sv_same_desc(const SvDesc& a, const SvDesc& b) {
    if (a.O != b.O || a.q != b.q || a.T != b.T || a.A != b.A || a.p != b.p || a.K != b.K || a.WB != b.WB ||
        a.w2_const != b.w2_const || a.NB != b.NB || a.NC != b.NC) return false;
    for (int k = 0; k < a.O; ++k) {
        if (a.fam[k] != b.fam[k] || a.link[k] != b.link[k] || a.qk[k] != b.qk[k] || a.z_ones0[k] != b.z_ones0[k] ||
            a.pk[k] != b.pk[k] || a.x_ones0[k] != b.x_ones0[k] || a.boff[k] != b.boff[k]) return false;
        for (int j = 0; j < a.qk[k]; ++j) if (a.re_inds[k][j] != b.re_inds[k][j]) return false;
    }
    for (int t = 0; t < a.T; ++t) {
        if (a.rt[t] != b.rt[t] || a.ut[t] != b.ut[t] || a.tf[t] != b.tf[t] || a.u_ones[t] != b.u_ones[t] ||
            a.ft[t] != b.ft[t] || a.outc[t] != b.outc[t]) return false;
        for (int j = 0; j < a.rt[t]; ++j) if (a.re2[t][j] != b.re2[t][j] || a.zz_col[t][j] != b.zz_col[t][j]) return false;
        for (int u = 0; u < a.ut[t]; ++u) if (a.col[t][u] != b.col[t][u]) return false;
        for (int f = 0; f < a.ft[t]; ++f) if (a.indf[t][f] != b.indf[t][f] || a.xx_col[t][f] != b.xx_col[t][f]) return false;
    }
    return true;
}

// BASELINE config 5 (and any fit of the config 3 / 4 shape): gaussian + binomial(logit) + poisson(log), random
// intercept + slope each (q = 6), value + slope association per outcome, X = Z = [1, time]
JMB_HD constexpr SvDesc sv_desc_three_outcomes() {
    SvDesc d{};
    d.O = 3; d.q = 6; d.T = 6; d.A = 6; d.p = 2; d.K = 15; d.WB = 4; d.w2_const = 1; d.NB = 6;
    d.NC = 1;                                 // the time variable, shared by the three value terms
    d.fam[0] = JMB_FAM_GAUSSIAN; d.link[0] = JMB_LINK_IDENTITY;
    d.fam[1] = JMB_FAM_BINOMIAL; d.link[1] = JMB_LINK_LOGIT;
    d.fam[2] = JMB_FAM_POISSON; d.link[2] = JMB_LINK_LOG;
    for (int k = 0; k < 3; ++k) {
        d.qk[k] = 2; d.z_ones0[k] = 1; d.pk[k] = 2; d.x_ones0[k] = 1; d.boff[k] = 2 * k;
        d.re_inds[k][0] = 2 * k; d.re_inds[k][1] = 2 * k + 1;
        const int tv = 2 * k, ts = 2 * k + 1;
        d.rt[tv] = 2; d.ut[tv] = 1; d.tf[tv] = JMB_TF_IDENTITY; d.u_ones[tv] = 1;
        d.ft[tv] = 2; d.outc[tv] = k; d.indf[tv][0] = 0; d.indf[tv][1] = 1;
        d.xx_col[tv][0] = -1; d.xx_col[tv][1] = 0; d.zz_col[tv][0] = -1; d.zz_col[tv][1] = 0;
        d.re2[tv][0] = 2 * k; d.re2[tv][1] = 2 * k + 1; d.col[tv][0] = tv;
        d.rt[ts] = 1; d.ut[ts] = 1; d.tf[ts] = JMB_TF_IDENTITY; d.u_ones[ts] = 1;
        d.ft[ts] = 1; d.outc[ts] = k; d.indf[ts][0] = 1; d.xx_col[ts][0] = -1; d.zz_col[ts][0] = -1;
        d.re2[ts][0] = 2 * k + 1; d.col[ts][0] = ts;
    }
    return d;
}

// one gaussian outcome, random intercept + slope, current-value association (configs 1 / 2)
JMB_HD constexpr SvDesc sv_desc_value_gaussian() {
    SvDesc d{};
    d.O = 1; d.q = 2; d.T = 1; d.A = 1; d.p = 2; d.K = 15; d.WB = 4; d.w2_const = 1; d.NB = 2; d.NC = 1;
    d.fam[0] = JMB_FAM_GAUSSIAN; d.link[0] = JMB_LINK_IDENTITY; d.qk[0] = 2; d.z_ones0[0] = 1; d.pk[0] = 2; d.x_ones0[0] = 1;
    d.re_inds[0][0] = 0; d.re_inds[0][1] = 1;
    d.rt[0] = 2; d.ut[0] = 1; d.tf[0] = JMB_TF_IDENTITY; d.u_ones[0] = 1; d.ft[0] = 2;
    d.xx_col[0][0] = -1; d.xx_col[0][1] = 0; d.zz_col[0][0] = -1; d.zz_col[0][1] = 0;
    d.indf[0][0] = 0; d.indf[0][1] = 1; d.re2[0][0] = 0; d.re2[0][1] = 1; d.col[0][0] = 0;
    return d;
}

struct SvSpecThree { static constexpr SvDesc d = sv_desc_three_outcomes(); static constexpr SvLayout L = sv_make_layout(d); };
struct SvSpecValue { static constexpr SvDesc d = sv_desc_value_gaussian(); static constexpr SvLayout L = sv_make_layout(d); };

struct SvMeta { SvDesc d; SvLayout L; };

struct SvConst {            // jmb_svft_control + sizes
    double scale, df, ci_lo, ci_hi, parscale, reltol, ndeps, cd_eps;
    int log, maxit;
    uint32_t seed;
};

struct SvArgs {
    const double* rec; const long long* off;       // subject records and their offsets [n + 1]
    const double* par_mean; const double* par_draws; // [ps_len], [M][ps_len]
    int n, M, L, B, ps_len, ncoef;
    long long subject_offset;
    int do_modes, do_sample;
    SvConst c;
    double* modes; double* hess; int* conv; int* counts;   // [n][q], [n][q*q], [n], [n][2]
    double* S;                                     // [n][M][L]
    double* Cbuf;                                  // [n][M][NC + 1] linear-predictor coefficients of b.new per draw, or NULL
    double* acc_rate; double* b_last;              // [n], [n][q]
    int* work;                                     // atomic subject counter
};

// ---- per-group shared state --------------------------------------------------------------------------
enum { SC_FMIN = 0, SC_F, SC_GP, SC_STEP, SC_F1, SC_X1, SC_W, SC_U, SC_EPS_H, SC_N };
enum { IC_SUBJ = 0, IC_STATE, IC_NEV, IC_RET, IC_COUNT, IC_FUN, IC_GRC, IC_ILAST, IC_ITER, IC_HI, IC_CONV, IC_BAD, IC_N };
// states of the mode-search control thread: ST_* consume the result of the posted request, PC_* are internal
enum { ST_FETCH = 0, ST_F0, ST_GRAD, ST_LS, PC_LOOP_TOP, PC_LS_TRY, PC_LS_DONE, PC_LOOP_END, PC_VM_DONE, PC_MODE_DONE };
enum { RET_INIT = 0, RET_LS, RET_H1, RET_H2 };

struct SvG {                // views into one group's shared memory
    double *par, *btry, *x, *grad, *bz, *X, *c, *t, *g, *Bm, *dpar, *df1, *hess, *mode, *Af, *invS, *invV, *Wa, *Wb,
           *b_old, *b_new, *pb, *z, *sc;
    int* ic;
};

__host__ __device__ inline int sv_group_doubles(int q, int ps_len) {
    const int zq = (q + 2) / 2 * 2;
    return ps_len + 14 * q + zq + 7 * q * q + SC_N + (IC_N + 1) / 2;
}

__device__ __forceinline__ void sv_bind(SvG& G, double* base, int q, int ps_len) {
    const int zq = (q + 2) / 2 * 2, qq = q * q;
    double* p = base;
    G.par = p; p += ps_len;
    G.btry = p; p += q; G.x = p; p += q; G.grad = p; p += q; G.bz = p; p += q; G.X = p; p += q; G.c = p; p += q;
    G.t = p; p += q; G.g = p; p += q; G.dpar = p; p += q; G.df1 = p; p += q; G.mode = p; p += q;
    G.b_old = p; p += q; G.b_new = p; p += q; G.pb = p; p += q; G.z = p; p += zq;
    G.Bm = p; p += qq; G.hess = p; p += qq; G.Af = p; p += qq; G.invS = p; p += qq; G.invV = p; p += qq;
    G.Wa = p; p += qq; G.Wb = p; p += qq;
    G.sc = p; p += SC_N;
    G.ic = reinterpret_cast<int*>(p);
}

// ---- log_post_RE_svft / survPred_svft_2 by the 16 lanes of a group ---------------------------------------
// runtime-branch twin of logdens_static (jmb_step_static.cuh): folds when fam / link are compile-time constants
__device__ __forceinline__ double sv_logdens(int fam, int link, double y, double eta, double inv_sigma) {
    if (fam == JMB_FAM_GAUSSIAN) {
        const double t = (y - eta) * inv_sigma;
        return -0.5 * (t * t);
    } else if (fam == JMB_FAM_BINOMIAL) {
        double pr;
        if (link == JMB_LINK_LOGIT) { const double e = exp(eta); pr = e / (1.0 + e); }
        else if (link == JMB_LINK_PROBIT) pr = 0.5 * erfc(-eta * 0.70710678118654752440);
        else pr = -exp(-exp(eta)) + 1.0;
        const bool one = (y == 1.0), zero = (y == 0.0);
        if (one || zero) {
            const double ld = log(one ? pr : 1.0 - pr);
            const bool dead_inf = one ? (pr == 1.0) : (pr == 0.0);      // 0 * log(0) in the naive formula
            return dead_inf ? __longlong_as_double(0x7ff8000000000000LL) : ld;
        }
        return y * log(pr) + (1.0 - y) * log(1.0 - pr);
    } else {
        if (link == JMB_LINK_LOG) {
            const double mu = exp(eta);
            double ld = y * eta - mu;
            if (mu == 0.0) ld = (y == 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : (y > 0.0 ? -INFINITY : INFINITY);
            return ld;
        }
        return row_logdens(fam, link, y, eta, inv_sigma);
    }
}

// responses of a binomial outcome that are neither 0 nor 1 (not produced by the reference's factor2numeric, R/extractFrames.R:27-29):
// kept out of line so that its two logarithms do not sit in the hot instruction stream
__device__ __noinline__ double sv_binom_general(double y, double pr) { return y * log(pr) + (1.0 - y) * log(1.0 - pr); }

// sv_logdens with E = exp(eta) supplied by the caller (the same call, hoisted out of the family branches)
__device__ __forceinline__ double sv_logdens_e(int fam, int link, double y, double eta, double E, double inv_sigma) {
    if (fam == JMB_FAM_GAUSSIAN) {
        const double t = (y - eta) * inv_sigma;
        return -0.5 * (t * t);
    } else if (fam == JMB_FAM_BINOMIAL) {
        double pr;
        if (link == JMB_LINK_LOGIT) pr = E / (1.0 + E);
        else if (link == JMB_LINK_PROBIT) pr = 0.5 * erfc(-eta * 0.70710678118654752440);
        else pr = -exp(-E) + 1.0;
        const bool one = (y == 1.0), zero = (y == 0.0);
        if (one || zero) {
            const double ld = log(one ? pr : 1.0 - pr);
            const bool dead_inf = one ? (pr == 1.0) : (pr == 0.0);
            return dead_inf ? __longlong_as_double(0x7ff8000000000000LL) : ld;
        }
        return sv_binom_general(y, pr);
    } else {
        if (link == JMB_LINK_LOG) {
            double ld = y * eta - E;
            if (E == 0.0) ld = (y == 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : (y > 0.0 ? -INFINITY : INFINITY);
            return ld;
        }
        return row_logdens(fam, link, y, eta, inv_sigma);
    }
}

// value of the posted request: full -> log_post_RE_svft(btry) on block 0 (src/survfitJM_mvJMbayes.cpp:140-166),
// else survPred_svft_2(btry) = -H on block `blk_index` (:168-187).  Generic flavour: runtime descriptor.
__device__ __forceinline__ double sv_eval_generic(const SvDesc& d, const SvLayout& L, const double* __restrict__ rec,
                                                  const int blk_index, const bool full, const double* sb, const double* sp,
                                                  const int lane, const unsigned gmask) {
    // ---- the lane's quadrature row (lin_pred_matF_svft + the linear predictor of the hazard) ----
    const double* __restrict__ blk = rec + L.o_blk0 + (long long)blk_index * L.blk;
    const int off = reinterpret_cast<const int*>(blk + L.b_W1off)[lane];
    double s1 = 0.0;
    for (int j = 0; j < d.WB; ++j) s1 += blk[L.b_W1v + j * kSvG + lane] * sp[L.ps_Bs + off + j];
    double s2 = 0.0;
    for (int j = 0; j < d.p; ++j) s2 += (d.w2_const ? rec[L.o_W2c + j] : blk[L.b_W2 + j * kSvG + lane]) * sp[L.ps_gam + j];
    double s3 = 0.0;
    for (int t = 0; t < d.T; ++t) {
        const int bo = d.boff[d.outc[t]];
        double xb = 0.0;
        for (int f = 0; f < d.ft[t]; ++f) {
            const int c = d.xx_col[t][f];
            xb += (c < 0 ? 1.0 : blk[L.b_col + c * kSvG + lane]) * sp[L.ps_betas + bo + d.indf[t][f]];
        }
        double zb = 0.0;
        for (int j = 0; j < d.rt[t]; ++j) {
            const int c = d.zz_col[t][j];
            zb += (c < 0 ? 1.0 : blk[L.b_col + c * kSvG + lane]) * sb[d.re2[t][j]];
        }
        const double v = trans_fun(d.tf[t], xb + zb);
        for (int u = 0; u < d.ut[t]; ++u) {
            const double w = d.u_ones[t] ? v : blk[L.b_U[t] + u * kSvG + lane] * v;
            s3 += w * sp[L.ps_alp + d.col[t][u]];
        }
    }
    const double e = (lane < d.K) ? blk[L.b_Pw + lane] * exp((s1 + s2) + s3) : 0.0;
    const double H = group_sum<kSvG>(e, gmask);
    if (!full) return 0.0 - H;

    // ---- log p(y | b) over the visit rows (lin_predF_svft + log_longF_svft, :57-66,103-135) ----
    const int* hdr = reinterpret_cast<const int*>(rec + L.o_V);
    const int nh = hdr[d.O];
    const double* __restrict__ lrec = rec + L.o_blk0 + (long long)(nh + 1) * L.blk;
    double acc = 0.0;
    for (int k = 0; k < d.O; ++k) {
        const int v = hdr[k];
        const int xc = d.pk[k] - d.x_ones0[k];
        for (int r = lane; r < v; r += kSvG) {
            double xb = d.x_ones0[k] ? sp[L.ps_betas + d.boff[k]] : 0.0;
            for (int c = d.x_ones0[k]; c < d.pk[k]; ++c) xb += lrec[(1 + c - d.x_ones0[k]) * v + r] * sp[L.ps_betas + d.boff[k] + c];
            double zb = d.z_ones0[k] ? sb[d.re_inds[k][0]] : 0.0;
            for (int j = d.z_ones0[k]; j < d.qk[k]; ++j) zb += lrec[(1 + xc + j - d.z_ones0[k]) * v + r] * sb[d.re_inds[k][j]];
            acc += sv_logdens(d.fam[k], d.link[k], lrec[r], xb + zb, sp[L.ps_isig + k]);
        }
        lrec += L.long_cols[k] * v;
    }
    // ---- log p(b) = -0.5 b' invD b (:157): lane j takes column j ----
    double qf = 0.0;
    if (lane < d.q) {
        double t = 0.0;
        for (int l = 0; l < d.q; ++l) t += sb[l] * sp[L.ps_invD + l + lane * d.q];
        qf = t * sb[lane];
    }
    const double log_pyb = group_sum<kSvG>(acc, gmask);
    const double log_pb = -0.5 * group_sum<kSvG>(qf, gmask);
    return (log_pyb + (0.0 - H)) + log_pb;                                     // :163
}

// Static flavour: the model shape is the constexpr descriptor Spec::d, every loop over outcomes / terms /
// columns is unrolled at compile time and b stays in registers.  Same arithmetic, same order.
// b-independent part of the lane's block-0 hazard row under fixed parameters (the mode search evaluates it ~500 times)
template <class Spec>
struct SvHazCache { double s12; double xb[Spec::d.T]; };

template <class Spec>
__device__ __forceinline__ void sv_haz_prepare(const double* __restrict__ rec, const double* sp, const int lane, SvHazCache<Spec>& hc) {
    constexpr int T = Spec::d.T, p = Spec::d.p, WB = Spec::d.WB;
    constexpr bool W2C = Spec::d.w2_const != 0;
    constexpr int oBlk0 = Spec::L.o_blk0, psB = Spec::L.ps_betas, psG = Spec::L.ps_gam, psBs = Spec::L.ps_Bs;
    constexpr int bW1off = Spec::L.b_W1off, bW1v = Spec::L.b_W1v, oW2c = Spec::L.o_W2c, bW2 = Spec::L.b_W2, bCol = Spec::L.b_col;
    const double* __restrict__ blk = rec + oBlk0;
    const int off = reinterpret_cast<const int*>(blk + bW1off)[lane];
    double s1 = 0.0;
    sfor<WB>([&](auto j) { s1 += blk[bW1v + j * kSvG + lane] * sp[psBs + off + j]; });
    double s2 = 0.0;
    sfor<p>([&](auto j) { s2 += (W2C ? rec[oW2c + j] : blk[bW2 + j * kSvG + lane]) * sp[psG + j]; });
    hc.s12 = s1 + s2;
    sfor<T>([&](auto tc) {
        constexpr int t = tc;
        constexpr int bo = Spec::d.boff[Spec::d.outc[t]];
        double xb = 0.0;
        sfor<Spec::d.ft[t]>([&](auto fc) {
            constexpr int f = fc;
            constexpr int idx = Spec::d.indf[t][f], c = Spec::d.xx_col[t][f];
            if constexpr (c < 0) xb += sp[psB + bo + idx];
            else xb += blk[bCol + c * kSvG + lane] * sp[psB + bo + idx];
        });
        hc.xb[t] = xb;
    });
}

// NB blocks x NP points in one pass (one of them is 1): slot s evaluates block (NB == 2 ? s : 0) at point (NP == 2 ? s : 0).
// Two slots give every lane two independent exp / log chains (the kernels are bound by fixed-latency fp64 dependencies),
// and two points on the same block share the b-independent part of the row.  Arithmetic per slot is that of one slot.
template <class Spec, bool CACHED, int NB, int NP>
__device__ __forceinline__ void sv_eval_static_n(const double* __restrict__ rec, const int (&blk_index)[NB], const bool full,
                                                 const double* const (&sb)[NP], const double* sp, const int lane,
                                                 const unsigned gmask, const SvHazCache<Spec>& hc,
                                                 double (&out)[(NB > NP ? NB : NP)]) {
    constexpr int S = (NB > NP ? NB : NP);
    static_assert(NB == 1 || NP == 1, "two blocks or two points, not both");
    constexpr int q = Spec::d.q, O = Spec::d.O, T = Spec::d.T, p = Spec::d.p, K = Spec::d.K, WB = Spec::d.WB, NC = Spec::d.NC;
    constexpr bool W2C = Spec::d.w2_const != 0;
    constexpr int oBlk0 = Spec::L.o_blk0, BLK = Spec::L.blk, psB = Spec::L.ps_betas, psG = Spec::L.ps_gam, psA = Spec::L.ps_alp;
    constexpr int psBs = Spec::L.ps_Bs, psD = Spec::L.ps_invD, psS = Spec::L.ps_isig;
    constexpr int bW1off = Spec::L.b_W1off, bW1v = Spec::L.b_W1v, oW2c = Spec::L.o_W2c, bW2 = Spec::L.b_W2, bPw = Spec::L.b_Pw;
    constexpr int oV = Spec::L.o_V, bCol = Spec::L.b_col;
    double b[NP][q];
    sfor<NP>([&](auto pc) { sfor<q>([&](auto j) { b[pc][j] = sb[pc][j]; }); });

    const double* __restrict__ blk[NB];
    double s12[NB], colv[NB][NC > 0 ? NC : 1], xbv[NB][T], pw[NB];
    sfor<NB>([&](auto bc) {
        constexpr int ib = bc;
        blk[ib] = rec + oBlk0 + (long long)blk_index[ib] * BLK;
        pw[ib] = blk[ib][bPw + lane];
        sfor<NC>([&](auto c) { colv[ib][c] = blk[ib][bCol + c * kSvG + lane]; });
        if constexpr (CACHED) s12[ib] = hc.s12;
        else {
            const int off = reinterpret_cast<const int*>(blk[ib] + bW1off)[lane];
            double s1 = 0.0;
            sfor<WB>([&](auto j) { s1 += blk[ib][bW1v + j * kSvG + lane] * sp[psBs + off + j]; });
            double s2 = 0.0;
            sfor<p>([&](auto j) { s2 += (W2C ? rec[oW2c + j] : blk[ib][bW2 + j * kSvG + lane]) * sp[psG + j]; });
            s12[ib] = s1 + s2;
        }
        sfor<T>([&](auto tc) {
            constexpr int t = tc;
            [[maybe_unused]] constexpr int bo = Spec::d.boff[Spec::d.outc[t]];
            double xb = 0.0;
            if constexpr (CACHED) xb = hc.xb[t];
            else sfor<Spec::d.ft[t]>([&](auto fc) {
                constexpr int f = fc;
                constexpr int idx = Spec::d.indf[t][f], c = Spec::d.xx_col[t][f];
                if constexpr (c < 0) xb += sp[psB + bo + idx];               // 1 * beta is exact
                else xb += colv[ib][c] * sp[psB + bo + idx];
            });
            xbv[ib][t] = xb;
        });
    });
    double H[S];
    sfor<S>([&](auto sc) {
        constexpr int sl = sc;
        constexpr int ib = (NB == 2) ? sl : 0, ip = (NP == 2) ? sl : 0;
        double s3 = 0.0;
        sfor<T>([&](auto tc) {
            constexpr int t = tc;
            constexpr int bU = Spec::L.b_U[t], tf = Spec::d.tf[t];
            double zb = 0.0;
            sfor<Spec::d.rt[t]>([&](auto jc) {
                constexpr int j = jc;
                constexpr int idx = Spec::d.re2[t][j], c = Spec::d.zz_col[t][j];
                if constexpr (c < 0) zb += b[ip][idx];
                else zb += colv[ib][c] * b[ip][idx];
            });
            double v = xbv[ib][t] + zb;
            if constexpr (tf != JMB_TF_IDENTITY) v = trans_fun(tf, v);
            sfor<Spec::d.ut[t]>([&](auto uc) {
                constexpr int u = uc;
                double w = v;
                constexpr int cidx = Spec::d.col[t][u];
                if constexpr (Spec::d.u_ones[t] == 0) w = blk[ib][bU + u * kSvG + lane] * v;
                s3 += w * sp[psA + cidx];
            });
        });
        H[sl] = (lane < K) ? pw[ib] * exp(s12[ib] + s3) : 0.0;
    });
    sfor<S>([&](auto sc) { H[sc] = group_sum<kSvG>(H[sc], gmask); });
    if (!full) {
        sfor<S>([&](auto sc) { out[sc] = 0.0 - H[sc]; });
        return;
    }

    // visit rows of all outcomes as one lane-strided list (a subject has ~n_i rows per outcome, far fewer than 16):
    // each lane finds the outcome of its row, the exponential is shared by the logit / cloglog / log tails
    const int* hdr = reinterpret_cast<const int*>(rec + oV);
    const int nh = hdr[O];
    const double* __restrict__ lrec = rec + oBlk0 + (long long)(nh + 1) * BLK;
    int vk[O], total = 0;
    sfor<O>([&](auto k) { vk[k] = hdr[k]; total += vk[k]; });
    double acc[NP];
    sfor<NP>([&](auto pc) { acc[pc] = 0.0; });
    for (int r = lane; r < total; r += kSvG) {
        double eta[NP], y = 0.0, isg = 1.0;
        sfor<NP>([&](auto pc) { eta[pc] = 0.0; });
        int mine = -1, c0 = 0, o0 = 0;
        sfor<O>([&](auto kc) {
            constexpr int k = kc;
            constexpr int x0 = Spec::d.x_ones0[k], z0 = Spec::d.z_ones0[k], pk = Spec::d.pk[k], qk = Spec::d.qk[k], xc = pk - x0;
            constexpr int bo = Spec::d.boff[k], ri0 = Spec::d.re_inds[k][0], lcols = Spec::L.long_cols[k];
            const int v = vk[k];
            if (r >= c0 && r < c0 + v) {
                const double* __restrict__ lr = lrec + o0;
                const int rr = r - c0;
                double xb = 0.0;
                if constexpr (x0 == 1) xb = sp[psB + bo];
                sfor<pk - x0>([&](auto cc) { constexpr int c = cc + x0; xb += lr[(1 + c - x0) * v + rr] * sp[psB + bo + c]; });
                sfor<NP>([&](auto pc) {
                    constexpr int ip = pc;
                    double zb = 0.0;
                    if constexpr (z0 == 1) zb = b[ip][ri0];
                    sfor<qk - z0>([&](auto jc) { constexpr int j = jc + z0; constexpr int idx = Spec::d.re_inds[k][j]; zb += lr[(1 + xc + j - z0) * v + rr] * b[ip][idx]; });
                    eta[ip] = xb + zb;
                });
                y = lr[rr]; isg = sp[psS + k]; mine = k;
            }
            c0 += v; o0 += lcols * v;
        });
        double E[NP];
        sfor<NP>([&](auto pc) { E[pc] = exp(eta[pc]); });
        sfor<O>([&](auto kc) {
            constexpr int k = kc;
            constexpr int fam = Spec::d.fam[k], link = Spec::d.link[k];
            if (mine == k) sfor<NP>([&](auto pc) { acc[pc] += sv_logdens_e(fam, link, y, eta[pc], E[pc], isg); });
        });
    }
    sfor<NP>([&](auto pc) {
        constexpr int ip = pc;
        double qf = 0.0;
        if (lane < q) {
            double t = 0.0;
            sfor<q>([&](auto l) { t += b[ip][l] * sp[psD + l + lane * q]; });
            qf = t * sb[ip][lane];
        }
        const double log_pyb = group_sum<kSvG>(acc[ip], gmask);
        const double log_pb = -0.5 * group_sum<kSvG>(qf, gmask);
        out[ip] = (log_pyb + (0.0 - H[ip])) + log_pb;
    });
}

// Models whose association terms are all untransformed with the default interaction design (tf = identity, Us = 1):
// sum_t alpha_t (XXs_t beta + ZZs_t b) is then linear in the pooled design columns, sum_c col_c * C_c + C_0, with
// coefficients that depend only on (draw, b).  The horizon integrals of one draw (same b, same parameters, nh blocks)
// compute C once and pay NC + 1 multiply-adds per quadrature row instead of ~40; the value differs from the term-by-term
// sum only by fp64 rounding (<= 1e-15 relative; the parity gate is 1e-9).
template <class Spec>
constexpr bool sv_is_linear() {
    for (int t = 0; t < Spec::d.T; ++t)
        if (Spec::d.tf[t] != JMB_TF_IDENTITY || Spec::d.u_ones[t] == 0 || Spec::d.ut[t] != 1) return false;
    return Spec::d.w2_const != 0;
}

template <class Spec>
struct SvLinCoef { double c0; double cc[Spec::d.NC > 0 ? Spec::d.NC : 1]; };

// C_0 (incl. W2 gamma, constant within the subject) and C_c for the point sb under the parameter set sp
template <class Spec>
__device__ __forceinline__ void sv_lin_coefs(const double* __restrict__ rec, const double* sb, const double* sp, SvLinCoef<Spec>& L) {
    constexpr int T = Spec::d.T, p = Spec::d.p, NC = Spec::d.NC, q = Spec::d.q;
    constexpr int psB = Spec::L.ps_betas, psG = Spec::L.ps_gam, psA = Spec::L.ps_alp, oW2c = Spec::L.o_W2c;
    double b[q];
    sfor<q>([&](auto j) { b[j] = sb[j]; });
    double c0 = 0.0;
    sfor<p>([&](auto j) { c0 += rec[oW2c + j] * sp[psG + j]; });
    sfor<NC>([&](auto c) { L.cc[c] = 0.0; });
    sfor<T>([&](auto tc) {
        constexpr int t = tc;
        constexpr int bo = Spec::d.boff[Spec::d.outc[t]], cidx = Spec::d.col[t][0];
        const double al = sp[psA + cidx];
        double k0 = 0.0;
        double kc[NC > 0 ? NC : 1];
        sfor<NC>([&](auto c) { kc[c] = 0.0; });
        sfor<Spec::d.ft[t]>([&](auto fc) {
            constexpr int f = fc;
            constexpr int idx = Spec::d.indf[t][f], c = Spec::d.xx_col[t][f];
            if constexpr (c < 0) k0 += sp[psB + bo + idx]; else kc[c] += sp[psB + bo + idx];
        });
        sfor<Spec::d.rt[t]>([&](auto jc) {
            constexpr int j = jc;
            constexpr int idx = Spec::d.re2[t][j], c = Spec::d.zz_col[t][j];
            if constexpr (c < 0) k0 += b[idx]; else kc[c] += b[idx];
        });
        c0 += al * k0;
        sfor<NC>([&](auto c) { L.cc[c] += al * kc[c]; });
    });
    L.c0 = c0;
}

// -H of two blocks at the point / parameters behind the coefficients L
template <class Spec>
__device__ __forceinline__ void sv_block2_lin(const double* __restrict__ rec, const int blkA, const int blkB, const double* sp,
                                              const SvLinCoef<Spec>& L, const int lane, const unsigned gmask, double& vA, double& vB) {
    constexpr int K = Spec::d.K, WB = Spec::d.WB, NC = Spec::d.NC;
    constexpr int oBlk0 = Spec::L.o_blk0, BLK = Spec::L.blk, psBs = Spec::L.ps_Bs;
    constexpr int bW1off = Spec::L.b_W1off, bW1v = Spec::L.b_W1v, bPw = Spec::L.b_Pw, bCol = Spec::L.b_col;
    double e[2];
    sfor<2>([&](auto sc) {
        constexpr int sl = sc;
        const double* __restrict__ blk = rec + oBlk0 + (long long)(sl == 0 ? blkA : blkB) * BLK;
        const int off = reinterpret_cast<const int*>(blk + bW1off)[lane];
        double s1 = 0.0;
        sfor<WB>([&](auto j) { s1 += blk[bW1v + j * kSvG + lane] * sp[psBs + off + j]; });
        double a = L.c0;
        sfor<NC>([&](auto c) { a += blk[bCol + c * kSvG + lane] * L.cc[c]; });
        e[sl] = (lane < K) ? blk[bPw + lane] * exp(s1 + a) : 0.0;
    });
    vA = 0.0 - group_sum<kSvG>(e[0], gmask);
    vB = 0.0 - group_sum<kSvG>(e[1], gmask);
}

template <class Spec, bool CACHED>
__device__ __forceinline__ double sv_eval_static(const double* __restrict__ rec, const int blk_index, const bool full,
                                                 const double* sb, const double* sp, const int lane, const unsigned gmask,
                                                 const SvHazCache<Spec>& hc) {
    const int bi[1] = {blk_index};
    const double* const pts[1] = {sb};
    double out[1];
    sv_eval_static_n<Spec, CACHED, 1, 1>(rec, bi, full, pts, sp, lane, gmask, hc, out);
    return out[0];
}

// rare path of the chi-square draw (shape < 1, or four Marsaglia-Tsang rejections in a row): out of the hot instruction stream
__device__ __noinline__ double sv_rgamma_slow(uint32_t seed, uint32_t key, uint32_t gid, uint32_t m, double shape, double scale) {
    return rgamma_dev(seed, key, gid, m, shape, scale);
}

// ---- small dense helpers of the control thread (same algorithms as oracle/jmb_oracle_svft.c) ---------------
// solve(A): Gauss-Jordan with partial pivoting; a is destroyed; returns 1 when singular
__device__ inline int sv_inverse(double* a, int n, double* inv) {
    for (int j = 0; j < n * n; ++j) inv[j] = 0.0;
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
        const double dd = 1.0 / a[c + c * n];
        for (int j = 0; j < n; ++j) { a[c + j * n] *= dd; inv[c + j * n] *= dd; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = a[r + c * n];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) { a[r + j * n] -= f * a[c + j * n]; inv[r + j * n] -= f * inv[c + j * n]; }
        }
    }
    return 0;
}

// eigen(S, symmetric = TRUE): cyclic Jacobi (12 sweeps), values decreasing.  a: symmetrised copy of S (destroyed)
__device__ inline void sv_eigen_sym(double* a, int n, double* val, double* vec) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) vec[i + j * n] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep)
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = a[p + q * n], app = a[p + p * n], aqq = a[q + q * n];
                if (fabs(apq) <= 1e-20 * (fabs(app) + fabs(aqq))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = a[k + p * n], akq = a[k + q * n];
                    a[k + p * n] = c * akp - s * akq; a[k + q * n] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = a[p + k * n], aqk = a[q + k * n];
                    a[p + k * n] = c * apk - s * aqk; a[q + k * n] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = vec[k + p * n], vkq = vec[k + q * n];
                    vec[k + p * n] = c * vkp - s * vkq; vec[k + q * n] = s * vkp + c * vkq;
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
}

#ifndef JMB_SVFT_THREADS
#define JMB_SVFT_THREADS 256
#endif
#ifndef JMB_SVFT_MIN_CTAS
#define JMB_SVFT_MIN_CTAS 3     // 80 registers, no spills; A/B on B200: 2 -> 40.8 ms, 3 -> 36.1 ms, 4 -> 37.4 ms per 25k subjects
#endif

struct SvEvalGeneric {
    static constexpr bool kSplit = false;            // the horizon integrals stay inside the sampling kernel
    const SvDesc& d; const SvLayout& L;
    __device__ __forceinline__ double operator()(const double* rec, int blk, bool full, const double* sb, const double* sp,
                                                 int lane, unsigned gmask) const {
        return sv_eval_generic(d, L, rec, blk, full, sb, sp, lane, gmask);
    }
    // the mode search evaluates block 0 under fixed parameters; the generic flavour keeps no per-lane cache
    __device__ __forceinline__ void prepare(const double*, const double*, int) {}
    __device__ __forceinline__ double full0(const double* rec, const double* sb, const double* sp, int lane, unsigned gmask) const {
        return sv_eval_generic(d, L, rec, 0, true, sb, sp, lane, gmask);
    }
    __device__ __forceinline__ void full2c(const double* rec, const double* sbA, const double* sbB, const double* sp, int lane,
                                           unsigned gmask, double& vA, double& vB) const {
        if (sbA == sbB) { vA = sv_eval_generic(d, L, rec, 0, true, sbA, sp, lane, gmask); vB = vA; }
        else full2(rec, sbA, sbB, sp, lane, gmask, vA, vB);
    }
    // two points on block 0 / two blocks at one point: the generic flavour simply evaluates twice
    __device__ __forceinline__ void full2(const double* rec, const double* sbA, const double* sbB, const double* sp, int lane,
                                          unsigned gmask, double& vA, double& vB) const {
        vA = sv_eval_generic(d, L, rec, 0, true, sbA, sp, lane, gmask);
        vB = sv_eval_generic(d, L, rec, 0, true, sbB, sp, lane, gmask);
    }
    __device__ __forceinline__ void block2(const double* rec, int blkA, int blkB, const double* sb, const double* sp, int lane,
                                           unsigned gmask, double& vA, double& vB) const {
        vA = sv_eval_generic(d, L, rec, blkA, false, sb, sp, lane, gmask);
        vB = sv_eval_generic(d, L, rec, blkB, false, sb, sp, lane, gmask);
    }
    // horizon integrals of one draw: nothing to hoist in the generic flavour
    __device__ __forceinline__ void begin_horizons(const double*, const double*, const double*) {}
    __device__ __forceinline__ void store_coefs(double*) const {}
    __device__ __forceinline__ void block2h(const double* rec, int blkA, int blkB, const double* sb, const double* sp, int lane,
                                            unsigned gmask, double& vA, double& vB) const {
        block2(rec, blkA, blkB, sb, sp, lane, gmask, vA, vB);
    }
};
template <class Spec>
struct SvEvalStatic {
    static constexpr bool kSplit = sv_is_linear<Spec>();   // the horizon integrals run in jmb_svft_horizons<Spec>
    SvHazCache<Spec> hc;
    __device__ __forceinline__ double operator()(const double* rec, int blk, bool full, const double* sb, const double* sp,
                                                 int lane, unsigned gmask) const {
        return sv_eval_static<Spec, false>(rec, blk, full, sb, sp, lane, gmask, hc);
    }
    __device__ __forceinline__ void prepare(const double* rec, const double* sp, int lane) { sv_haz_prepare<Spec>(rec, sp, lane, hc); }
    __device__ __forceinline__ double full0(const double* rec, const double* sb, const double* sp, int lane, unsigned gmask) const {
        return sv_eval_static<Spec, true>(rec, 0, true, sb, sp, lane, gmask, hc);
    }
    __device__ __forceinline__ void full2(const double* rec, const double* sbA, const double* sbB, const double* sp, int lane,
                                          unsigned gmask, double& vA, double& vB) const {
        const int bi[1] = {0};
        const double* const pts[2] = {sbA, sbB};
        double out[2];
        sv_eval_static_n<Spec, false, 1, 2>(rec, bi, true, pts, sp, lane, gmask, hc, out);
        vA = out[0]; vB = out[1];
    }
    // the same with the cached b-independent row part (mode search: both points of a central difference in one pass)
    __device__ __forceinline__ void full2c(const double* rec, const double* sbA, const double* sbB, const double* sp, int lane,
                                           unsigned gmask, double& vA, double& vB) const {
        const int bi[1] = {0};
        const double* const pts[2] = {sbA, sbB};
        double out[2];
        sv_eval_static_n<Spec, true, 1, 2>(rec, bi, true, pts, sp, lane, gmask, hc, out);
        vA = out[0]; vB = out[1];
    }
    __device__ __forceinline__ void block2(const double* rec, int blkA, int blkB, const double* sb, const double* sp, int lane,
                                           unsigned gmask, double& vA, double& vB) const {
        const int bi[2] = {blkA, blkB};
        const double* const pts[1] = {sb};
        double out[2];
        sv_eval_static_n<Spec, false, 2, 1>(rec, bi, false, pts, sp, lane, gmask, hc, out);
        vA = out[0]; vB = out[1];
    }
    // horizon integrals of one draw at one point: hoist the (draw, b) part of the linear predictor when the model allows
    SvLinCoef<Spec> lc;
    __device__ __forceinline__ void begin_horizons(const double* rec, const double* sb, const double* sp) {
        if constexpr (sv_is_linear<Spec>()) sv_lin_coefs<Spec>(rec, sb, sp, lc);
    }
    __device__ __forceinline__ void store_coefs(double* dst) const {      // [NC + 1]: c0, cc[0..NC-1]
        if constexpr (sv_is_linear<Spec>()) { dst[0] = lc.c0; sfor<Spec::d.NC>([&](auto c) { dst[1 + c] = lc.cc[c]; }); }
    }
    __device__ __forceinline__ void block2h(const double* rec, int blkA, int blkB, const double* sb, const double* sp, int lane,
                                            unsigned gmask, double& vA, double& vB) const {
        if constexpr (sv_is_linear<Spec>()) sv_block2_lin<Spec>(rec, blkA, blkB, sp, lc, lane, gmask, vA, vB);
        else block2(rec, blkA, blkB, sb, sp, lane, gmask, vA, vB);
    }
};

// Mode search of all subjects (R/survfitJM.mvJMbayes.R:277-305): vmmin (R src/appl/optim.c), fminfn / fmingr /
// optimhess (R src/library/stats/src/optim.c) and cd (R/cd.R) as a resumable state machine that all 16 lanes of the
// group execute together.  Scalars are replicated in registers and computed by every lane with vmmin's own sequential
// arithmetic; vectors live in shared memory, lane i owning element / row i.  A request is `nev` evaluations of
// -log_post_RE_svft: 1 = the point btry, 2q = the central-difference gradient at x (written to grad[], already
// multiplied by parscale as fmingr does).  Both groups of a warp always meet in the evaluation loop.
template <class Eval>
__device__ __forceinline__ void sv_modes_run(const int q, const SvArgs& a, Eval& eval) {
    extern __shared__ double sv_smem[];
    const int tid = threadIdx.x, lane = tid & (kSvG - 1), grp = tid / kSvG;
    const unsigned gmask = 0xFFFFu << (16 * ((tid & 31) >> 4));
    const int gdn = sv_group_doubles(q, a.ps_len), gd = (gdn + 1) / 2 * 2;
    // views carved from the shared-memory window itself (same order as sv_bind), so that loads compile to LDS
    double* gbase = sv_smem + (size_t)grp * gd;
    const double* sp = gbase;                        // parameter set (posterior means)
    double* v0 = gbase + a.ps_len;
    double *btry = v0, *x = v0 + q, *grad = v0 + 2 * q, *bz = v0 + 3 * q, *X = v0 + 4 * q, *c = v0 + 5 * q, *t = v0 + 6 * q,
           *g = v0 + 7 * q, *dpar = v0 + 8 * q, *df1 = v0 + 9 * q, *mode = v0 + 10 * q, *bt2 = v0 + 13 * q;
    const int zq = (q + 2) / 2 * 2;
    double* Bm = v0 + 14 * q + zq;
    double* hess = Bm + q * q;
    const int n = q;
    const bool own = lane < n;                       // this lane owns element / row `lane`
    const double ps = a.c.parscale, eps = a.c.cd_eps, stepredn = 0.2, acctol = 0.0001, reltest = 10.0;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int j = lane; j < a.ps_len; j += kSvG) gbase[j] = a.par_mean[j];
    __syncwarp();

    // replicated control state
    int st = ST_FETCH, subj = -1, nev = 0, ret = 0, count = 0, fun = 0, grc = 0, ilast = 0, iter = 0, hi = 0, conv = 0;
    double Fmin = 0.0, f = 0.0, gradproj = 0.0, step = 0.0, epsH = 0.0, val = 0.0;
    const double* rec = a.rec;
    auto gsync = [&]() { __syncwarp(gmask); };
    auto post_grad = [&](const double* z, int r) {   // fmingr at the optimiser point z: x = z * parscale
        if (own) x[lane] = z[lane] * ps;
        gsync();
        ret = r; nev = 2 * n; st = ST_GRAD;
    };
    bool finished = false;                           // no more subjects for this group
    int it = 0, nit = 0;                             // pass of the posted request / passes it needs
    for (;;) {
        // ---- control: when this group's request is complete, consume `val` (or grad[]) and post the next one; the other
        //      group of the warp is not waited for (requests are 1 or q passes long and rarely aligned) ----
        bool posted = finished || it < nit;
        while (!posted) {
            switch (st) {
            case ST_FETCH: {
                int s = 0;
                if (lane == 0) s = atomicAdd(a.work, 1);
                s = __shfl_sync(gmask, s, 0, kSvG);
                if (s >= a.n) { subj = -1; nev = 0; finished = true; posted = true; break; }
                subj = s; conv = 0;
                rec = a.rec + a.off[s];
                eval.prepare(rec, sp, lane);
                if (own) { bz[lane] = 0.0 / ps; btry[lane] = bz[lane] * ps; }       // start = rep(0, q)
                gsync();
                nev = 1; st = ST_F0; posted = true;
                break;
            }
            case ST_F0: {
                f = 0.0 - val;
                if (!isfinite(f)) {                  // error("initial value in 'vmmin' is not finite")
                    conv = 10; fun = 1; grc = 0;
                    if (own) mode[lane] = nan;
                    for (int j = lane; j < n * n; j += kSvG) hess[j] = nan;
                    gsync();
                    st = PC_MODE_DONE;
                    break;
                }
                Fmin = f;
                fun = 1; grc = 1; iter = 0; count = 0;
                post_grad(bz, RET_INIT); posted = true;
                break;
            }
            case ST_GRAD: {
                if (ret == RET_INIT) {
                    if (own) g[lane] = grad[lane];
                    gsync();
                    iter = 1; ilast = grc;
                    st = PC_LOOP_TOP;
                } else if (ret == RET_LS) {
                    grc += 1; iter += 1;
                    if (own) { t[lane] = step * t[lane]; c[lane] = grad[lane] - c[lane]; g[lane] = grad[lane]; }
                    gsync();
                    double D1 = 0.0;
                    for (int i = 0; i < n; ++i) D1 += t[i] * c[i];
                    if (D1 > 0) {
                        if (own) {
                            const int i = lane;
                            double s = 0.0;
                            for (int j = 0; j <= i; ++j) s += Bm[i * n + j] * c[j];
                            for (int j = i + 1; j < n; ++j) s += Bm[j * n + i] * c[j];
                            X[i] = s;
                        }
                        gsync();
                        double D2 = 0.0;
                        for (int i = 0; i < n; ++i) D2 += X[i] * c[i];
                        D2 = 1.0 + D2 / D1;
                        if (own) {
                            const int i = lane;
                            for (int j = 0; j <= i; ++j) Bm[i * n + j] += (D2 * t[i] * t[j] - X[i] * t[j] - t[i] * X[j]) / D1;
                        }
                        gsync();
                    } else {
                        ilast = grc;
                    }
                    st = PC_LOOP_END;
                } else if (ret == RET_H1) {          // optimhess: df1 at dpar + eps e_i
                    if (own) df1[lane] = grad[lane];
                    if (lane == hi) dpar[hi] = dpar[hi] - 2 * epsH;
                    gsync();
                    post_grad(dpar, RET_H2); posted = true;
                } else {                             // RET_H2: df2 at dpar - eps e_i
                    if (own) hess[hi * n + lane] = (df1[lane] - grad[lane]) / (2 * epsH * ps * ps);
                    if (lane == hi) dpar[hi] = dpar[hi] + epsH;
                    gsync();
                    if (hi + 1 < n) {
                        hi += 1;
                        if (lane == hi) dpar[hi] = dpar[hi] + epsH;
                        gsync();
                        post_grad(dpar, RET_H1); posted = true;
                        break;
                    }
                    if (own) {
                        const int i = lane;
                        for (int j = 0; j < i; ++j) {
                            const double tmp = 0.5 * (hess[i * n + j] + hess[j * n + i]);
                            hess[i * n + j] = tmp; hess[j * n + i] = tmp;
                        }
                    }
                    gsync();
                    st = PC_MODE_DONE;
                }
                break;
            }
            case PC_LOOP_TOP: {
                if (ilast == grc && own) {
                    const int i = lane;
                    for (int j = 0; j < i; ++j) Bm[i * n + j] = 0.0;
                    Bm[i * n + i] = 1.0;
                }
                if (own) { X[lane] = bz[lane]; c[lane] = g[lane]; }
                gsync();
                if (own) {
                    const int i = lane;
                    double s = 0.0;
                    for (int j = 0; j <= i; ++j) s -= Bm[i * n + j] * c[j];
                    for (int j = i + 1; j < n; ++j) s -= Bm[j * n + i] * c[j];
                    t[i] = s;
                }
                gsync();
                gradproj = 0.0;
                for (int i = 0; i < n; ++i) gradproj += t[i] * c[i];
                if (gradproj < 0.0) {                // search direction is downhill
                    step = 1.0;
                    st = PC_LS_TRY;
                } else {                             // uphill search
                    count = 0;
                    if (ilast == grc) count = n; else ilast = grc;
                    st = PC_LOOP_END;
                }
                break;
            }
            case PC_LS_TRY: {
                bool same = false;
                if (own) {
                    const double nb = X[lane] + step * t[lane];
                    bz[lane] = nb;
                    same = (reltest + X[lane] == reltest + nb);
                    btry[lane] = nb * ps;
                }
                count = __popc(__ballot_sync(gmask, same));
                gsync();
                if (count < n) { nev = 1; st = ST_LS; posted = true; }
                else st = PC_LS_DONE;
                break;
            }
            case ST_LS: {
                f = 0.0 - val;
                fun += 1;
                const bool accpoint = isfinite(f) && (f <= Fmin + gradproj * step * acctol);
                if (!accpoint) { step *= stepredn; st = PC_LS_TRY; }
                else st = PC_LS_DONE;
                break;
            }
            case PC_LS_DONE: {
                const bool enough = fabs(f - Fmin) > a.c.reltol * (fabs(Fmin) + a.c.reltol);   // abstol = -Inf
                if (!enough) { count = n; Fmin = f; }
                if (count < n) {                     // making progress
                    Fmin = f;
                    post_grad(bz, RET_LS); posted = true;
                    break;
                }
                if (ilast < grc) { count = 0; ilast = grc; }                              // no progress
                st = PC_LOOP_END;
                break;
            }
            case PC_LOOP_END: {
                if (iter >= a.c.maxit) { st = PC_VM_DONE; break; }
                if (grc - ilast > 2 * n) ilast = grc;                                     // periodic restart
                st = (count != n || ilast != grc) ? PC_LOOP_TOP : PC_VM_DONE;
                break;
            }
            case PC_VM_DONE: {
                conv = (iter < a.c.maxit) ? 0 : 1;
                epsH = a.c.ndeps / ps;                                                    // optimhess
                if (own) { mode[lane] = bz[lane] * ps; dpar[lane] = mode[lane] / ps; }    // opt$par
                hi = 0;
                if (lane == 0) dpar[0] = dpar[0] + epsH;
                gsync();
                post_grad(dpar, RET_H1); posted = true;
                break;
            }
            default: {                               // PC_MODE_DONE
                const long long s = subj;
                if (own) a.modes[s * n + lane] = mode[lane];
                for (int j = lane; j < n * n; j += kSvG) a.hess[s * n * n + j] = hess[j];
                if (lane == 0) { a.conv[s] = conv; a.counts[2 * s] = fun; a.counts[2 * s + 1] = grc; }
                st = ST_FETCH;
                break;
            }
            }
        }
        if (it >= nit) { nit = (nev > 1) ? nev / 2 : nev; it = 0; }
        if (__all_sync(0xFFFFFFFFu, finished)) break;
        // ---- one evaluation pass.  A gradient request evaluates the two points x +- ex e_i of a central difference
        //      (R/cd.R:5-10) in one pass; a single value goes through the same two-slot code (both slots on the point), so
        //      that the two groups of the warp stay converged whatever each of them is asking for ----
        if (!finished) {
            const bool grad_req = nev > 1;
            if (grad_req && own) {
                double xa = x[lane], xb2 = xa;
                if (lane == it) { const double ex = eps * (fabs(xa) + eps); xb2 = xa - ex; xa = xa + ex; }
                btry[lane] = xa; bt2[lane] = xb2;
            }
            __syncwarp(gmask);
            double vA, vB;
            eval.full2c(rec, btry, grad_req ? bt2 : btry, sp, lane, gmask, vA, vB);
            if (grad_req) {
                if (lane == 0) grad[it] = (((0.0 - vA) - (0.0 - vB)) / (btry[it] - bt2[it])) * ps;   // diff.f / diff.x, fmingr scaling
            } else {
                val = vA;
            }
            __syncwarp(gmask);
            it += 1;
        }
    }
}

// per-subject set-up of the sampler: invVars.b = hessian / scale, Vars.b = scale * solve(hessian) (:304-305) and
// eigen(Vars.b) for rmvt (R/rmvt.R) and dmvt (R/dmvt.R).  The inverse is taken by the control lane; the Jacobi rotations and
// the two factor matrices are spread over the lanes of the group (lane k owns row / column element k of each update, so
// every element sees the arithmetic of the serial algorithm in oracle/jmb_oracle_svft.c).  Returns 1 when the Hessian is
// unusable (uniform over the group).
__device__ __noinline__ int sv_sample_setup(const SvArgs& a, const int n, SvG& G, const long long s, const int lane,
                                            const unsigned gmask) {
    if (lane == 0) {
        for (int j = 0; j < n; ++j) G.mode[j] = a.modes[s * n + j];
        for (int j = 0; j < n * n; ++j) G.hess[j] = a.hess[s * n * n + j];
        bool bad = false;
        for (int j = 0; j < n; ++j) if (!isfinite(G.mode[j])) bad = true;
        for (int j = 0; j < n * n; ++j) {
            if (!isfinite(G.hess[j])) bad = true;
            G.invV[j] = G.hess[j] / a.c.scale;
            G.Wa[j] = G.hess[j];
        }
        if (!bad && sv_inverse(G.Wa, n, G.Wb)) bad = true;
        if (!bad)
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) G.Wa[i + j * n] = 0.5 * (G.Wb[i + j * n] * a.c.scale + G.Wb[j + i * n] * a.c.scale);
        G.ic[IC_BAD] = bad ? 1 : 0;
    }
    __syncwarp(gmask);
    if (G.ic[IC_BAD]) return 1;
    // eigen(S, symmetric = TRUE): cyclic Jacobi, 12 sweeps (sv_eigen_sym), rotations applied by lane k to element k
    double* am = G.Wa; double* vec = G.Wb; double* val = G.X;
    const bool own = lane < n;
    if (own) for (int j = 0; j < n; ++j) vec[lane + j * n] = (lane == j) ? 1.0 : 0.0;
    __syncwarp(gmask);
    for (int sweep = 0; sweep < 12; ++sweep)
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = am[p + q * n], app = am[p + p * n], aqq = am[q + q * n];
                if (fabs(apq) <= 1e-20 * (fabs(app) + fabs(aqq))) continue;
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                __syncwarp(gmask);                   // everybody has read the pivot entries
                if (own) { const double x = am[lane + p * n], y = am[lane + q * n]; am[lane + p * n] = c * x - sn * y; am[lane + q * n] = sn * x + c * y; }
                __syncwarp(gmask);
                if (own) { const double x = am[p + lane * n], y = am[q + lane * n]; am[p + lane * n] = c * x - sn * y; am[q + lane * n] = sn * x + c * y; }
                if (own) { const double x = vec[lane + p * n], y = vec[lane + q * n]; vec[lane + p * n] = c * x - sn * y; vec[lane + q * n] = sn * x + c * y; }
                __syncwarp(gmask);
            }
    if (lane == 0) {
        for (int j = 0; j < n; ++j) val[j] = am[j + j * n];
        for (int j = 0; j < n - 1; ++j) {
            int best = j;
            for (int l = j + 1; l < n; ++l) if (val[l] > val[best]) best = l;
            if (best != j) {
                double t = val[j]; val[j] = val[best]; val[best] = t;
                for (int k = 0; k < n; ++k) { t = vec[k + j * n]; vec[k + j * n] = vec[k + best * n]; vec[k + best * n] = t; }
            }
        }
        bool bad = false;
        if (!(val[n - 1] >= -1e-06 * fabs(val[0]))) bad = true;      // dmvt: "'Sigma' is not positive definite"
        for (int j = 0; j < n; ++j) if (!(val[j] > 0.0)) bad = true;
        G.ic[IC_BAD] = bad ? 1 : 0;
    }
    __syncwarp(gmask);
    if (G.ic[IC_BAD]) return 1;
    if (own) {
        const int r = lane;
        for (int c = 0; c < n; ++c) {
            G.Af[r + c * n] = vec[r + c * n] * sqrt(val[c] > 0.0 ? val[c] : 0.0);    // evec * rep(sqrt(pmax(ev, 0)), each = p)
            double t = 0.0;
            for (int k = 0; k < n; ++k) t += vec[r + k * n] * (vec[c + k * n] / val[k]);   // evec %*% (t(evec) / ev)
            G.invS[r + c * n] = t;
        }
        G.b_old[r] = G.mode[r]; G.b_new[r] = G.mode[r];
    }
    __syncwarp(gmask);
    return 0;
}

// the M Metropolis-Hastings steps and the horizon integrals of all subjects (R/survfitJM.mvJMbayes.R:330-398)
template <class Eval>
__device__ __forceinline__ void sv_sample_run(const int q, const int O, const SvArgs& a, Eval& eval) {
    extern __shared__ double sv_smem[];
    const int tid = threadIdx.x, lane = tid & (kSvG - 1), grp = tid / kSvG;
    const unsigned gmask = 0xFFFFu << (16 * ((tid & 31) >> 4));
    const int gdn = sv_group_doubles(q, a.ps_len), gd = (gdn + 1) / 2 * 2;
    double* gbase = sv_smem + (size_t)grp * gd;
    const int zq = (q + 2) / 2 * 2, qq = q * q;
    // the same carving as sv_bind, from the shared-memory window (LDS / STS)
    const double* sp = gbase;
    double* v0 = gbase + a.ps_len;
    const double* smode = v0 + 10 * q;
    double* sb_old = v0 + 11 * q;
    double* sb_new = v0 + 12 * q;
    double* spb = v0 + 13 * q;
    double* sz = v0 + 14 * q;
    const double* sAf = sz + zq + 2 * qq;
    const double* sinvS = sAf + qq;
    const double* sinvV = sinvS + qq;
    double* ssc = gbase + (gdn - (IC_N + 1) / 2 - SC_N);
    SvG G;
    sv_bind(G, gbase, q, a.ps_len);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const int groups = blockDim.x / kSvG;
    const long long stride = (long long)gridDim.x * groups;
    for (long long s = (long long)blockIdx.x * groups + grp;; s += stride) {
        const bool have = s < a.n;
        if (!__any_sync(0xFFFFFFFFu, have)) break;
        const double* rec = a.rec + a.off[have ? s : 0];
        const int nh = have ? reinterpret_cast<const int*>(rec)[O] : 0;
        int bad_subject = 1;
        if (have) bad_subject = sv_sample_setup(a, q, G, s, lane, gmask);
        __syncwarp();
        const bool live = have && bad_subject == 0;
        if (have && !live) {                         // unusable Hessian: reported, not propagated
            for (long long k = lane; k < (long long)a.M * a.L; k += kSvG) a.S[s * a.M * a.L + k] = nan;
            if (lane < q) a.b_last[s * q + lane] = nan;
            if (lane == 0) { a.acc_rate[s] = nan; a.conv[s] = 20; }
        }
        const int nhl = live ? nh : 0;
        const int nh_max = max(nhl, __shfl_xor_sync(0xFFFFFFFFu, nhl, 16));
        const uint32_t gid = (uint32_t)(a.subject_offset + s);
        int n_acc = 0;
        for (int m = 0; m < a.M; ++m) {
            if (live) {
                // draws of step m (streams of DESIGN.md section 4, key (seed, 'SVFT')) and parameter set m.  Every lane
                // runs the same Philox block + Box-Muller on its own counter: lanes 0-3 the proposal normals (slot =
                // lane), lanes 4-7 / 8-11 the normal / log-uniform of attempts 0-3 of the Marsaglia-Tsang gamma
                // (rchisq(df) = rgamma(df / 2, scale 2); same streams and attempt order as rgamma_dev), lane 12 the
                // acceptance uniform.
                {
                    const uint32_t c2 = (lane < 4) ? (uint32_t)lane : ((lane < 12) ? (uint32_t)(lane & 3) : 0u);
                    const uint32_t c3 = (lane < 4) ? 0u : ((lane < 8) ? (uint32_t)ST_GIBBS_N : ((lane < 12) ? (uint32_t)ST_GIBBS_U : 1u));
                    uint32_t r[4];
                    philox4x32(a.c.seed, kSvKey, gid, (uint32_t)m, c2, c3, r);
                    const double u1 = u01(r[0], r[1]), u2 = u01(r[2], r[3]);
                    const double lg = log(u1);
                    const double rad = sqrt(-2.0 * lg);
                    double sn, cs;
                    sincospi(2.0 * u2, &sn, &cs);
                    const double z0 = rad * cs, z1 = rad * sn;
                    if (lane < (q + 1) / 2) { sz[2 * lane] = z0; sz[2 * lane + 1] = z1; }
                    if (lane == 12) ssc[SC_U] = u1;
                    const double shape = 0.5 * a.c.df;
                    const double lu = __shfl_sync(gmask, lg, (lane & 3) + 8, kSvG);     // attempt's log-uniform
                    bool ok = false;
                    double xg = 0.0;
                    if (shape >= 1.0 && lane >= 4 && lane < 8) {
                        const double dd = shape - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * dd);
                        const double t = 1.0 + cc * z0;
                        if (t > 0.0) {
                            const double v = t * t * t;
                            if (lu < 0.5 * z0 * z0 + dd - dd * v + dd * log(v)) { ok = true; xg = dd * v; }
                        }
                    }
                    const unsigned hit = (__ballot_sync(gmask, ok) >> (16 * ((tid & 31) >> 4))) & 0xF0u;
                    if (hit) {
                        const int first = __ffs(hit) - 1;
                        if (lane == first) ssc[SC_W] = xg * 2.0;
                    } else if (lane == 4) {
                        ssc[SC_W] = sv_rgamma_slow(a.c.seed, kSvKey, gid, (uint32_t)m, shape, 2.0);   // shape < 1 or four rejections
                    }
                }
                const double* src = a.par_draws + (size_t)m * a.ps_len;
                for (int j = lane; j < a.ps_len; j += kSvG) gbase[j] = src[j];
            }
            __syncwarp();
            if (live && lane < q) {                  // proposal m of rmvt(M, mode, Vars.b, df) (:330-331)
                const double den = sqrt(ssc[SC_W] / a.c.df);
                double t = 0.0;
                for (int c = 0; c < q; ++c) t += sAf[lane + c * q] * sz[c];
                spb[lane] = smode[lane] + t / den;
            }
            __syncwarp();
            bool accept = false;
            if (live) {
                double tp = 0.0, to = 0.0;           // dmvt(prop = TRUE) of the proposal and of b.old (:334-337,354)
                if (lane < q) {
                    for (int r = 0; r < q; ++r) {
                        tp += (spb[r] - smode[r]) * sinvS[r + lane * q];
                        to += (sb_old[r] - smode[r]) * sinvV[r + lane * q];
                    }
                    tp *= (spb[lane] - smode[lane]); to *= (sb_old[lane] - smode[lane]);
                }
                const double quad_p = group_sum<kSvG>(tp, gmask), quad_o = group_sum<kSvG>(to, gmask);
                const double dmvt_p = -0.5 * (a.c.df + q) * log(1.0 + quad_p / a.c.df);
                const double dmvt_o = -0.5 * (a.c.df + q) * log(1.0 + quad_o / a.c.df);
                double lp_p, lp_o;                   // both points in one pass over block 0 and the visit rows
                eval.full2(rec, spb, sb_old, sp, lane, gmask, lp_p, lp_o);
                const double ea = exp(lp_p + dmvt_o - lp_o - dmvt_p);               // :372-377
                accept = !isnan(ea) && ssc[SC_U] <= (ea < 1.0 ? ea : 1.0);
                if (accept && lane < q) sb_new[lane] = spb[lane];
                n_acc += accept ? 1 : 0;
            }
            __syncwarp();
            double cum = 0.0, mine = 0.0;            // survival on each prediction interval (:378-395)
            double* Srow = a.S + ((size_t)(have ? s : 0) * a.M + m) * a.L;
            constexpr bool split = Eval::kSplit;      // the horizon integrals run in jmb_svft_horizons from the coefficients
            if constexpr (split) {
                if (have && a.Cbuf != nullptr) {
                    if (live) eval.begin_horizons(rec, sb_new, sp);
                    if (lane == 0) {
                        double* cdst = a.Cbuf + ((size_t)s * a.M + m) * a.ncoef;
                        if (live) eval.store_coefs(cdst); else for (int j = 0; j < a.ncoef; ++j) cdst[j] = nan;
                    }
                }
            } else {
                for (int l = 0; l < nh_max; l += 2) {    // two intervals per pass
                    if (live && l < nh) {
                        const bool two = l + 1 < nh;
                        double v0, v1;
                        eval.block2(rec, l + 1, two ? l + 2 : l + 1, sb_new, sp, lane, gmask, v0, v1);
                        cum += v0;
                        if (lane == (l & (kSvG - 1))) mine = a.c.log ? v0 : cum;
                        if (two) {
                            cum += v1;
                            if (lane == ((l + 1) & (kSvG - 1))) mine = a.c.log ? v1 : cum;
                        }
                    }
                    if (((l + 1) & (kSvG - 1)) == kSvG - 1 || l + 2 >= nh_max) {  // 16 intervals: exp(cumsum(logS)) by lane
                        const int ll = (l & ~(kSvG - 1)) + lane;
                        if (live && ll < nh) Srow[ll] = a.c.log ? mine : exp(mine);
                    }
                }
            }
            if (live) {
                if (!split) for (int l = nh + lane; l < a.L; l += kSvG) Srow[l] = nan;   // intervals this subject does not have
                if (lane < q) sb_old[lane] = sb_new[lane];
            }
            __syncwarp();
        }
        if (live) {
            if (lane == 0) a.acc_rate[s] = (double)n_acc / (a.M > 0 ? a.M : 1);
            if (lane < q) a.b_last[s * q + lane] = sb_new[lane];
        }
        __syncwarp();
    }
}

#ifndef JMB_SVFT_MODES_MIN_CTAS
#define JMB_SVFT_MODES_MIN_CTAS 2  // the mode search keeps more scalars live: 128 registers, no spills (A/B: 29.2 vs 30.0 ms)
#endif
template <class Spec>
__global__ void __launch_bounds__(JMB_SVFT_THREADS, JMB_SVFT_MODES_MIN_CTAS) jmb_svft_modes_static(const __grid_constant__ SvArgs a) {
    SvEvalStatic<Spec> ev{};
    sv_modes_run(Spec::d.q, a, ev);
}
__global__ void __launch_bounds__(JMB_SVFT_THREADS, JMB_SVFT_MODES_MIN_CTAS)
jmb_svft_modes_generic(const __grid_constant__ SvMeta mm, const __grid_constant__ SvArgs a) {
    SvEvalGeneric ev{mm.d, mm.L};
    sv_modes_run(mm.d.q, a, ev);
}
template <class Spec>
__global__ void __launch_bounds__(JMB_SVFT_THREADS, JMB_SVFT_MIN_CTAS) jmb_svft_sample_static(const __grid_constant__ SvArgs a) {
    SvEvalStatic<Spec> ev{};
    sv_sample_run(Spec::d.q, Spec::d.O, a, ev);
}
__global__ void __launch_bounds__(JMB_SVFT_THREADS, JMB_SVFT_MIN_CTAS)
jmb_svft_sample_generic(const __grid_constant__ SvMeta mm, const __grid_constant__ SvArgs a) {
    SvEvalGeneric ev{mm.d, mm.L};
    sv_sample_run(mm.d.q, mm.d.O, a, ev);
}

// ---- horizon integrals of linear models: one THREAD per posterior draw ---------------------------------------------------
// For the models of sv_is_linear() the linear predictor of a quadrature row is s1(row, Bs_m) + C_0(m) + sum_c col_c(row)
// C_c(m), where the coefficients C(m) of b.new after draw m come from jmb_svft_sample_* (a.Cbuf).  A CTA walks subjects;
// thread m keeps C(m) in registers, the Bs_gammas of all draws sit in shared memory ([M][B], conflict-free for odd B), the K
// rows of an interval are staged in shared memory and broadcast; every thread sums its K rows in row order (the order of the
// reference's cumsum, src/survfitJM_mvJMbayes.cpp:7-14), accumulates cumsum(logS) over the intervals and writes
// S = exp(cumsum) (R/survfitJM.mvJMbayes.R:378-395).  No reductions, no divergence, 32 of 32 lanes busy.
template <class Spec>
__global__ void __launch_bounds__(256) jmb_svft_horizons(const __grid_constant__ SvArgs a) {
    constexpr int K = Spec::d.K, WB = Spec::d.WB, NC = Spec::d.NC, O = Spec::d.O;
    constexpr int oBlk0 = Spec::L.o_blk0, BLK = Spec::L.blk, psBs = Spec::L.ps_Bs;
    constexpr int bW1off = Spec::L.b_W1off, bW1v = Spec::L.b_W1v, bPw = Spec::L.b_Pw, bCol = Spec::L.b_col;
    constexpr int RW = 2 + WB + NC;                  // staged row: Pw, band offset, band values, pooled columns
    extern __shared__ double hz_smem[];
    const int B = a.B, tid = threadIdx.x, MC = blockDim.x;
    double* s_Bs = hz_smem;                          // [MC][B]
    double* s_row = hz_smem + (size_t)MC * B;        // [K][RW]
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int m0 = 0; m0 < a.M; m0 += MC) {           // chunks of blockDim.x draws (one chunk when M <= 256)
        const int m = m0 + tid;
        const bool on = m < a.M;
        __syncthreads();
        for (int e = tid; e < MC * B; e += MC) {
            const int mm = e / B, j = e - mm * B;
            s_Bs[e] = (m0 + mm < a.M) ? a.par_draws[(size_t)(m0 + mm) * a.ps_len + psBs + j] : 0.0;
        }
        const double* myBs = s_Bs + (size_t)tid * B;
        for (int s = blockIdx.x; s < a.n; s += gridDim.x) {
            const double* __restrict__ rec = a.rec + a.off[s];
            const int nh = reinterpret_cast<const int*>(rec)[O];
            double c0 = 0.0, cc[NC > 0 ? NC : 1];
            if (on) {
                const double* cs = a.Cbuf + ((size_t)s * a.M + m) * a.ncoef;
                c0 = cs[0];
#pragma unroll
                for (int c = 0; c < NC; ++c) cc[c] = cs[1 + c];
            }
            double cum = 0.0;
            double* Srow = a.S + ((size_t)s * a.M + (on ? m : 0)) * a.L;
            for (int l = 0; l < nh; ++l) {
                const double* __restrict__ blk = rec + oBlk0 + (long long)(l + 1) * BLK;
                __syncthreads();
                for (int e = tid; e < K * RW; e += MC) {
                    const int k = e / RW, f = e - k * RW;
                    double v;
                    if (f == 0) v = blk[bPw + k];
                    else if (f == 1) v = (double)reinterpret_cast<const int*>(blk + bW1off)[k];
                    else if (f < 2 + WB) v = blk[bW1v + (f - 2) * kSvG + k];
                    else v = blk[bCol + (f - 2 - WB) * kSvG + k];
                    s_row[e] = v;
                }
                __syncthreads();
                if (on) {
                    double H = 0.0;
#pragma unroll 5
                    for (int k = 0; k < K; ++k) {
                        const double* r = s_row + k * RW;
                        const int off = (int)r[1];
                        double s1 = 0.0;
#pragma unroll
                        for (int j = 0; j < WB; ++j) s1 += r[2 + j] * myBs[off + j];
                        double al = c0;
#pragma unroll
                        for (int c = 0; c < NC; ++c) al += r[2 + WB + c] * cc[c];
                        H += r[0] * exp(s1 + al);
                    }
                    const double logS = 0.0 - H;
                    cum += logS;
                    Srow[l] = a.c.log ? logS : exp(cum);
                }
            }
            if (on) for (int l = nh; l < a.L; ++l) Srow[l] = nan;
        }
    }
}

// ---- Mean / Median / Lower / Upper, M <= 32 * EPL: one warp per (subject, interval) -------------------------------------
// The M values live in registers (element e = lane * EPL + i), sorted by a bitonic network whose partner distances below
// EPL are register swaps and the others lane shuffles; NaN is removed (na.rm = TRUE) by sorting it last as +inf.
template <int EPL>
__global__ void __launch_bounds__(128) jmb_svft_summary_warp(const double* __restrict__ S, const long long* __restrict__ off,
                                                             const double* __restrict__ rec, const int O, const int n,
                                                             const int M, const int L, const double p_lo, const double p_hi,
                                                             double* __restrict__ out) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)n * L) return;
    const long long i = w / L;
    const int l = (int)(w % L);
    const int nh = reinterpret_cast<const int*>(rec + off[i])[O];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    double* o4 = out + (size_t)i * 4 * L + l;
    if (l >= nh) {
        if (lane < 4) o4[(size_t)lane * L] = nan;
        return;
    }
    double x[EPL];
    int cnt = 0;
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < EPL; ++k) {
        const int m = lane * EPL + k;
        double v = (m < M) ? S[((size_t)i * M + m) * L + l] : INFINITY;
        if (m < M && !isnan(v)) { cnt++; sum += v; } else v = INFINITY;
        x[k] = v;
    }
    // rowMeans(na.rm = TRUE): per-lane partial sums in draw order, then lanes in order (fixed, data-independent order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o); }
    double tot = 0.0;
    for (int src = 0; src < 32; ++src) tot += __shfl_sync(0xFFFFFFFFu, sum, src);
#pragma unroll
    for (int k = 2; k <= 32 * EPL; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= EPL) {                          // partner in another lane
                const int lj = j / EPL;
#pragma unroll
                for (int t = 0; t < EPL; ++t) {
                    const int e = lane * EPL + t;
                    const double y = __shfl_xor_sync(0xFFFFFFFFu, x[t], lj);
                    const bool up = ((e & k) == 0), lower = ((lane & lj) == 0);
                    const double mn = fmin(x[t], y), mx = fmax(x[t], y);
                    x[t] = (up == lower) ? mn : mx;
                }
            } else {                                 // partner in this lane's registers
#pragma unroll
                for (int t = 0; t < EPL; ++t) {
                    const int u = t ^ j;
                    if (u > t) {
                        const int e = lane * EPL + t;
                        const bool up = ((e & k) == 0);
                        const double a = x[t], b = x[u];
                        const double mn = fmin(a, b), mx = fmax(a, b);
                        x[t] = up ? mn : mx; x[u] = up ? mx : mn;
                    }
                }
            }
        }
    }
    // sorted ascending: element e in lane e / EPL, register e % EPL
    auto at = [&](int e) -> double {
        double v = 0.0;
#pragma unroll
        for (int t = 0; t < EPL; ++t) if ((e % EPL) == t) v = x[t];
        return __shfl_sync(0xFFFFFFFFu, v, e / EPL);
    };
    if (cnt == 0) {
        if (lane < 4) o4[(size_t)lane * L] = nan;
        return;
    }
    const int half = (cnt + 1) / 2;
    const double m0 = at(half - 1), m1 = at(min(half, 32 * EPL - 1));
    const double pr[2] = {p_lo, p_hi};
    double qv[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        const double index = 1.0 + (cnt - 1) * pr[k];
        const double lo = floor(index), hi = ceil(index);
        double qs = at((int)lo - 1);
        const double xh = at((int)hi - 1);
        if (index > lo && xh != qs) { const double h = index - lo; qs = (1.0 - h) * qs + h * xh; }
        qv[k] = qs;
    }
    if (lane == 0) {
        o4[0] = tot / cnt;
        o4[L] = (cnt % 2 == 1) ? m0 : (m0 + m1) / 2.0;
        o4[2 * (size_t)L] = qv[0];
        o4[3 * (size_t)L] = qv[1];
    }
}

// ---- Mean / Median / Lower / Upper over the M draws (R/survfitJM.mvJMbayes.R:400-414) ----------------------------
// one CTA per subject; per interval a bitonic sort of the M values in shared memory (NaN removed, na.rm = TRUE),
// median.default and quantile(type = 7).  out: [n][4][L] as L x 4 per subject (column-major).
__global__ void __launch_bounds__(128) jmb_svft_summary(const double* __restrict__ S, const long long* __restrict__ off,
                                                        const double* __restrict__ rec, const int O, const int n,
                                                        const int M, const int L, const int Mp, const double p_lo,
                                                        const double p_hi, double* __restrict__ out) {
    extern __shared__ double s_x[];                  // [Mp]
    __shared__ int s_cnt;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int nh = reinterpret_cast<const int*>(rec + off[i])[O];
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (int l = 0; l < L; ++l) {
        double* o4 = out + (size_t)i * 4 * L + l;
        if (l >= nh) {
            if (tid < 4) o4[(size_t)tid * L] = nan;
            continue;
        }
        if (tid == 0) s_cnt = 0;
        __syncthreads();
        int local = 0;
        for (int m = tid; m < Mp; m += blockDim.x) {
            double v = (m < M) ? S[((size_t)i * M + m) * L + l] : INFINITY;
            if (m < M && isnan(v)) v = INFINITY; else if (m < M) local++;
            s_x[m] = v;
        }
        atomicAdd(&s_cnt, local);
        __syncthreads();
        for (int k = 2; k <= Mp; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int t = tid; t < Mp; t += blockDim.x) {
                    const int ixj = t ^ j;
                    if (ixj > t) {
                        const double x = s_x[t], y = s_x[ixj];
                        const bool up = ((t & k) == 0);
                        if ((x > y) == up) { s_x[t] = y; s_x[ixj] = x; }
                    }
                }
                __syncthreads();
            }
        if (tid == 0) {
            const int cnt = s_cnt;                   // +inf paddings sort last; genuine +inf values are counted
            if (cnt == 0) { o4[0] = nan; o4[L] = nan; o4[2 * (size_t)L] = nan; o4[3 * (size_t)L] = nan; }
            else {
                double sum = 0.0;                    // rowMeans(na.rm = TRUE) in draw order
                for (int m = 0; m < M; ++m) { const double v = S[((size_t)i * M + m) * L + l]; if (!isnan(v)) sum += v; }
                o4[0] = sum / cnt;
                const int half = (cnt + 1) / 2;
                o4[L] = (cnt % 2 == 1) ? s_x[half - 1] : (s_x[half - 1] + s_x[half]) / 2.0;
                const double pr[2] = {p_lo, p_hi};
                for (int k = 0; k < 2; ++k) {
                    const double index = 1.0 + (cnt - 1) * pr[k];
                    const double lo = floor(index), hi = ceil(index);
                    double qs = s_x[(int)lo - 1];
                    const double xh = s_x[(int)hi - 1];
                    if (index > lo && xh != qs) { const double h = index - lo; qs = (1.0 - h) * qs + h * xh; }
                    o4[(size_t)(2 + k) * L] = qs;
                }
            }
        }
        __syncthreads();
    }
}

}  // namespace jmb
