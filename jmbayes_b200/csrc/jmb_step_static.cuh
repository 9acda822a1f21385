// jmb_step_static.cuh - compile-time-specialised step kernel (sm_100a).
//
// Same work as jmb_step_kernel (one full b-step + the subject's share of the survival step per
// subject and iteration; src/fast_LAPRWM.cpp:666-757), but the model shape (outcomes, families,
// terms, column maps, compaction flags) is a constexpr descriptor (jmb_model.h), so every loop
// over outcomes / terms / columns is unrolled, b_i (old and proposed) lives in registers from
// proposal to acceptance, and the only runtime loop is the ragged visit segment.  One lane per
// hazard row (event time + K quadrature nodes [+ K lower-interval nodes]); G = 16 or 32 lanes
// per subject.  Reductions are xor-butterflies inside the lane group.
#pragma once
#include <utility>

#include "jmb_kernels.cuh"

#ifndef JMB_LOGIT_FAST
#define JMB_LOGIT_FAST 1
#endif

namespace jmb {

template <int TF>
__device__ __forceinline__ double trans_static(double x) {
    if constexpr (TF == JMB_TF_IDENTITY) return x;
    else return trans_fun(TF, x);
}

// Per-row log-density (src/fast_LAPRWM.cpp:177-199).  For 0/1 responses the logit/probit/cloglog
// branch evaluates only the live logarithm and reproduces the naive formula's NaN when the dead
// term would be 0 * (-inf); the Poisson branch uses y*eta for y*log(exp(eta)) with the same
// overflow / underflow outcomes.  Anything else goes through the literal formulas.
// responses other than 0 / 1 of a binomial outcome: the literal formula, kept out of line (the specialised kernels' hot loops must
// fit the instruction cache: two inlined logarithms per call site were 700 instructions of a kernel that never executes them)
__device__ __noinline__ double binomial_general(double y, double pr) { return y * log(pr) + (1.0 - y) * log(1.0 - pr); }

#ifndef JMB_EXP3
#define JMB_EXP3 1
#endif
// Exponentials of the hot loops.  For |x| < 700 - every value a chain ever sees - exp(x) = 2^n exp(r), n = rint(x log2 e),
// r = x - n ln 2 (two-term Cody-Waite reduction, |r| <= 0.347), exp(r) by the degree-13 Taylor polynomial (truncation 4e-18),
// 2^n added to the exponent field: about 1 ulp worst case over [-700, 700] against the correctly rounded value (measured on the
// host with the same arithmetic: 0.88 ulp Horner, 1.04 ulp split).  The polynomial is evaluated as 1 + r + r^2 (Qe(r^2) + r Qo(r^2)):
// two chains of five fused multiply-adds instead of one of thirteen (the loops wait on fixed-latency dependencies, not on issue
// slots); the coefficients come from the constant bank (libm's exp materialises its 64-bit coefficients with two moves each, per
// call).  Anything else (overflow, gradual underflow, NaN) goes through libm, out of line.
#ifndef JMB_EXP_SPLIT
#define JMB_EXP_SPLIT 1
#endif
__constant__ double kExpTaylor[12] = {1.0 / 6227020800.0, 1.0 / 479001600.0, 1.0 / 39916800.0, 1.0 / 3628800.0, 1.0 / 362880.0,
                                      1.0 / 40320.0, 1.0 / 5040.0, 1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5};   // [j] = 1 / (13 - j)!
// out-of-line code takes and returns scalars only: a reference or pointer to a local would move that local to the stack everywhere
__device__ __noinline__ double exp_libm(const double x) { return exp(x); }
__device__ __forceinline__ double exp_reduce(const double x, double& t) {
    const double L2E = 1.44269504088896338700e+00, SH = 6755399441055744.0;
    const double LN2H = 6.93147180559945286227e-01, LN2L = 2.31904681384629955842e-17;
    t = fma(x, L2E, SH);
    const double n = t - SH;
    return fma(n, -LN2L, fma(n, -LN2H, x));
}
__device__ __forceinline__ double exp_poly(const double r) {
#if JMB_EXP_SPLIT
    const double s = r * r;
    double qe = kExpTaylor[1], qo = kExpTaylor[0];           // 1/12!, 1/13!
#pragma unroll
    for (int k = 3; k < 12; k += 2) { qe = fma(qe, s, kExpTaylor[k]); qo = fma(qo, s, kExpTaylor[k - 1]); }      // ... 1/2!, 1/3!
    return 1.0 + fma(s, fma(r, qo, qe), r);
#else
    double p = kExpTaylor[0];
#pragma unroll
    for (int k = 1; k < 12; ++k) p = fma(p, r, kExpTaylor[k]);
    p = fma(p, r, 1.0);
    return fma(p, r, 1.0);
#endif
}
__device__ __forceinline__ double exp_scale(const double p, const double t) {
    return __hiloint2double(__double2hiint(p) + (__double2loint(t) << 20), __double2loint(p));
}
// three at once: the three hazard values of a quadrature row
__device__ __forceinline__ void exp3(const double x0, const double x1, const double x2, double& e0, double& e1, double& e2) {
#if JMB_EXP3
    if ((fabs(x0) < 700.0) && (fabs(x1) < 700.0) && (fabs(x2) < 700.0)) {
        double t0, t1, t2;
        const double r0 = exp_reduce(x0, t0), r1 = exp_reduce(x1, t1), r2 = exp_reduce(x2, t2);
        const double p0 = exp_poly(r0), p1 = exp_poly(r1), p2 = exp_poly(r2);
        e0 = exp_scale(p0, t0); e1 = exp_scale(p1, t1); e2 = exp_scale(p2, t2);
    } else
#endif
    { e0 = exp_libm(x0); e1 = exp_libm(x1); e2 = exp_libm(x2); }
}
// one: visit rows (logit, Poisson), censoring terms
__device__ __forceinline__ double exp1(const double x) {
#if JMB_EXP3
    if (fabs(x) < 700.0) { double t; const double r = exp_reduce(x, t); return exp_scale(exp_poly(r), t); }
#endif
    return exp_libm(x);
}

// logit rows outside |eta| < 30: the literal formula with its 0 * log(0) = NaN and log(0) = -inf outcomes (src/fast_LAPRWM.cpp:184-188)
__device__ __noinline__ double logit_naive(double y, double eta) {
    const double e = exp(eta), pr = e / (1.0 + e);
    return y * log(pr) + (1.0 - y) * log(1.0 - pr);
}
// y log(pr) + (1 - y) log(1 - pr) with pr = e / (1 + e), e = exp(eta): log(pr) = eta - log(1 + e) and log(1 - pr) = -log(1 + e), so the
// row is y eta - log(1 + e) - one exponential and one logarithm, no division.  For |eta| < 30 pr is neither 0 nor 1 in double
// precision, so none of the literal formula's special outcomes can occur and the two agree to the last few bits (the rewritten form
// is the more accurate one); everything else takes the literal formula.
__device__ __forceinline__ double logit_row(double y, double eta) {
    if (fabs(eta) < 30.0) return y * eta - log(1.0 + exp1(eta));
    return logit_naive(y, eta);
}

template <int FAM, int LINK>
__device__ __forceinline__ double logdens_static(double y, double eta, double inv_sigma) {
    if constexpr (FAM == JMB_FAM_GAUSSIAN) {
        const double t = (y - eta) * inv_sigma;
        return -0.5 * (t * t);
    } else if constexpr (FAM == JMB_FAM_BINOMIAL && LINK == JMB_LINK_LOGIT && JMB_LOGIT_FAST) {
        return logit_row(y, eta);
    } else if constexpr (FAM == JMB_FAM_BINOMIAL) {
        double pr;
        if constexpr (LINK == JMB_LINK_LOGIT) { const double e = exp(eta); pr = e / (1.0 + e); }
        else if constexpr (LINK == JMB_LINK_PROBIT) pr = 0.5 * erfc(-eta * 0.70710678118654752440);
        else pr = -exp(-exp(eta)) + 1.0;
        const bool one = (y == 1.0), zero = (y == 0.0);
        if (one || zero) {
            const double ld = log(one ? pr : 1.0 - pr);
            const bool dead_inf = one ? (pr == 1.0) : (pr == 0.0);      // 0 * log(0) in the naive formula
            return dead_inf ? __longlong_as_double(0x7ff8000000000000LL) : ld;
        }
        return binomial_general(y, pr);
    } else {
        if constexpr (LINK == JMB_LINK_LOG) {
            const double mu = exp1(eta);
            double ld = y * eta - mu;
            if (mu == 0.0) ld = (y == 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : (y > 0.0 ? -INFINITY : INFINITY);
            return ld;                                                   // mu == inf gives inf - inf = NaN already
        } else {
            return row_logdens(FAM, LINK, y, eta, inv_sigma);
        }
    }
}

// event / censoring term for a lane group (uniform across its lanes), src/fast_LAPRWM.cpp:640-650
template <int S>
__device__ __forceinline__ double log_ptb_static(double event, double lh, double H, double HL) {
    if constexpr (S == 2) {
        if (event == 1.0) return lh - H;
        if (event == 0.0) return 0.0 - H;
        if (event == 2.0) return log(1.0 - exp(-H));
        if (event == 3.0) return log(exp(-HL) - exp(-H));
        return 0.0;
    } else {
        return event * lh - H;
    }
}

#ifndef JMB_STATIC_MIN_CTAS
#define JMB_STATIC_MIN_CTAS 2
#endif
// launch shape of the TMA-staged kernel: threads per CTA and CTAs per SM (registers are capped to fit)
#ifndef JMB_TMA_THREADS
#define JMB_TMA_THREADS 256
#endif
#ifndef JMB_TMA_CTAS
#define JMB_TMA_CTAS 2
#endif
constexpr int kTmaThreads = JMB_TMA_THREADS, kTmaCtas = JMB_TMA_CTAS, kTmaWarps = JMB_TMA_THREADS / 32;

// loop-invariant parameters a thread keeps in registers
template <class Spec>
struct StepRegs {
    double gc[Spec::d.p > 0 ? Spec::d.p : 1], gp[Spec::d.p > 0 ? Spec::d.p : 1];
    double alc[Spec::d.A], alp[Spec::d.A], isig[Spec::d.O];
};

// One b-step of one subject by one lane group, plus the subject's share of the survival step.
// rec / crec / st_in may point to global or to (TMA-staged) shared memory.
template <class Spec>
__device__ __forceinline__ void subject_step(const double* __restrict__ rec, const double* __restrict__ crec,
                                             const double* __restrict__ st_in, const double* __restrict__ rg,
                                             double* __restrict__ st_out,
                                             const double* s_cur, const double* s_prop, const double* s_invD,
                                             const StepRegs<Spec>& pr,
                                             const bool last_acc, const int lane, const unsigned gmask,
                                             double& acc_cur, double& acc_prop, double& acc_lw) {
    constexpr int G = Spec::d.G, q = Spec::d.q, O = Spec::d.O, T = Spec::d.T, p = Spec::d.p;
    constexpr int K = Spec::d.K, S = Spec::d.S, WB = Spec::d.WB, RP = Spec::L.RP, QS = q + kStateExtra;
    constexpr int oV = Spec::L.o_V, oLong = Spec::L.o_long, cLong = Spec::L.c_long, oW1off = Spec::L.o_W1off;
    constexpr int oW1v = Spec::L.o_W1v, oW2c = Spec::L.o_W2c, oW2 = Spec::L.o_W2, oPw = Spec::L.o_Pw;
    constexpr bool W2C = Spec::d.w2_const != 0;
    static_assert(RP == G && 1 + K * S <= G && QS <= G, "one hazard row per lane");
    const int cls = (lane == 0) ? 0 : (lane <= K ? 1 : ((S == 2 && lane <= 2 * K) ? 2 : 3));
    const bool quad = (cls == 1 || cls == 2);

    // ---- current state (uniform loads) ----
    double bo[q], bn[q];
    sfor<q>([&](auto j) { bo[j] = st_in[j]; });
    const double pyb_old = st_in[q + 1], pb_old = st_in[q + 2];
    double accb = st_in[q + 3], acct = st_in[q + 4];
    // log p(T | b_old) under the current survival parameters, carried from the previous iteration:
    // slot q+5 if its survival proposal was rejected, slot q+6 (evaluated at that proposal) if accepted
    const double ptb_co = last_acc ? st_in[q + 6] : st_in[q + 5];
    // proposal increment sqrt(sigma_b) z'R_i and log u of this iteration (jmb_block_prep_kernel; :658,670,706)
    sfor<q>([&](auto j) { bn[j] = bo[j] + rg[j]; });
    const double log_u = rg[q];

    // ---- log p(b) at the proposal (:673) ----
    double pb_new = 0.0;
    sfor<q>([&](auto jc) {
        constexpr int j = jc;
        double t = 0.0;
        sfor<q>([&](auto l) { t += bn[l] * s_invD[l + j * q]; });
        pb_new += t * bn[j];
    });
    pb_new *= -0.5;

    // ---- GLMM log-density over the visit segment (lin_predF + log_longF, :67-78,168-202) ----
    double v_pyb = 0.0;
    {
        const double* __restrict__ lrec = rec + oLong;
        const double* __restrict__ lcrec = crec + cLong;
        const int* vcount = reinterpret_cast<const int*>(rec + oV);
        sfor<O>([&](auto kc) {
            constexpr int k = kc;
            constexpr int ones0 = Spec::d.z_ones0[k], qk = Spec::d.qk[k];
            const int v = vcount[k];
            for (int r = lane; r < v; r += G) {
                double eta = lcrec[r];
                if constexpr (ones0 == 1) { constexpr int c0 = Spec::d.re_inds[k][0]; eta += bn[c0]; }
                sfor<qk - ones0>([&](auto jc) {
                    constexpr int j = jc;
                    constexpr int cj = Spec::d.re_inds[k][j + ones0];
                    eta += lrec[(1 + j) * v + r] * bn[cj];
                });
                v_pyb += logdens_static<Spec::d.fam[k], Spec::d.link[k]>(lrec[r], eta, pr.isig[k]);
            }
            constexpr int lcols = Spec::L.long_cols[k];
            lrec += lcols * v;
            lcrec += v;
        });
    }

    // ---- hazard row of this lane under current / proposed parameters and old / new b
    //      (lin_pred_matF :80-109, log_h :684, H/HL :685-691, survival step :735-741) ----
    double l_cn, l_po, l_pn;
    {
        const int off = reinterpret_cast<const int*>(rec + oW1off)[lane];
        double s1c = 0.0, s1p = 0.0;
        sfor<WB>([&](auto j) {
            const double wv = rec[oW1v + j * RP + lane];
            s1c += wv * s_cur[off + j];
            s1p += wv * s_prop[off + j];
        });
        double s2c = 0.0, s2p = 0.0;
        sfor<p>([&](auto j) {
            const double wv = W2C ? rec[oW2c + j] : rec[oW2 + j * RP + lane];
            s2c += wv * pr.gc[j];
            s2p += wv * pr.gp[j];
        });
        double a_cn = 0.0, a_po = 0.0, a_pn = 0.0;
        sfor<T>([&](auto tc) {
            constexpr int t = tc;
            constexpr int ones0 = Spec::d.zz_ones0[t], rt = Spec::d.rt[t];
            const double xxb = crec[t * RP + lane];
            double zo = 0.0, zn = 0.0;
            if constexpr (ones0 == 1) { constexpr int c0 = Spec::d.re2[t][0]; zo = bo[c0]; zn = bn[c0]; }
            constexpr int oZZ = Spec::L.o_ZZ[t], oU = Spec::L.o_U[t];
            sfor<rt - ones0>([&](auto jc) {
                constexpr int j = jc;
                constexpr int cj = Spec::d.re2[t][j + ones0];
                const double zz = rec[oZZ + j * RP + lane];
                zo += zz * bo[cj];
                zn += zz * bn[cj];
            });
            const double vo = trans_static<Spec::d.tf[t]>(xxb + zo), vn = trans_static<Spec::d.tf[t]>(xxb + zn);
            sfor<Spec::d.ut[t]>([&](auto uc) {
                constexpr int u = uc;
                constexpr int c = Spec::d.col[t][u];
                double wo = vo, wn = vn;
                if constexpr (Spec::d.u_ones[t] == 0) {
                    const double uu = rec[oU + u * RP + lane];
                    wo *= uu; wn *= uu;
                }
                a_cn += wn * pr.alc[c];
                a_po += wo * pr.alp[c]; a_pn += wn * pr.alp[c];
            });
        });
        l_cn = (s1c + s2c) + a_cn;
        l_po = (s1p + s2p) + a_po; l_pn = (s1p + s2p) + a_pn;
    }
    const double pw = quad ? rec[oPw + lane] : 0.0;
    const double event = rec[0];

    // ---- reductions: visit segment over the group; quadrature rows inside 16-lane halves ----
    double e_cn = quad ? pw * exp(l_cn) : 0.0;
    double H_cn, HL_cn = 0.0;
    if constexpr (S == 1 || K == 15) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) e_cn += __shfl_xor_sync(gmask, e_cn, o, 16);
        if constexpr (S == 2) { H_cn = __shfl_sync(gmask, e_cn, 0, G); HL_cn = __shfl_sync(gmask, e_cn, 16, G); }
        else H_cn = e_cn;
    } else {
        H_cn = group_sum<G>(cls == 1 ? e_cn : 0.0, gmask); HL_cn = group_sum<G>(cls == 2 ? e_cn : 0.0, gmask);
    }
    const double pyb_new = group_sum<G>(v_pyb, gmask);
    const double h_cn = __shfl_sync(gmask, l_cn, 0, G);
    const double h_po = __shfl_sync(gmask, l_po, 0, G), h_pn = __shfl_sync(gmask, l_pn, 0, G);

    const double ptb_cn = log_ptb_static<S>(event, h_cn, H_cn, HL_cn);
    const double cur_lp = pyb_old + ptb_co + pb_old;                 // :669
    const double new_lp = pyb_new + ptb_cn + pb_new;                 // :703
    const bool accept = log_u < (new_lp - cur_lp);                   // :706 (NaN rejects)

    // ---- proposed survival parameters at the accepted b (:735-752) ----
    double e_p = quad ? pw * exp(accept ? l_pn : l_po) : 0.0;
    double Hp, HLp = 0.0;
    if constexpr (S == 1 || K == 15) {
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) e_p += __shfl_xor_sync(gmask, e_p, o, 16);
        if constexpr (S == 2) { Hp = __shfl_sync(gmask, e_p, 0, G); HLp = __shfl_sync(gmask, e_p, 16, G); }
        else Hp = e_p;
    } else {
        Hp = group_sum<G>(cls == 1 ? e_p : 0.0, gmask); HLp = group_sum<G>(cls == 2 ? e_p : 0.0, gmask);
    }
    const double ptb_c = accept ? ptb_cn : ptb_co;
    const double ptb_p = log_ptb_static<S>(event, accept ? h_pn : h_po, Hp, HLp);
    const double pyb_acc = accept ? pyb_new : pyb_old;
    const double pb_acc = accept ? pb_new : pb_old;
    acc_cur += ptb_c; acc_prop += ptb_p; acc_lw += pyb_acc + pb_acc;

    // ---- state write-back (:707-724); sigma_b is adapted by jmb_adapt_sigma_kernel at block ends ----
    accb += accept ? 1.0 : 0.0;
    acct += accept ? 1.0 : 0.0;
    double outv = 0.0;
    sfor<q>([&](auto j) { if (lane == j) outv = accept ? bn[j] : bo[j]; });
    if (lane == q + 1) outv = pyb_acc;
    if (lane == q + 2) outv = pb_acc;
    if (lane == q + 3) outv = accb;
    if (lane == q + 4) outv = acct;
    if (lane == q + 5) outv = ptb_c;
    if (lane == q + 6) outv = ptb_p;
    if (lane < QS && lane != q && (accept || lane > q)) st_out[lane] = outv;     // slot q (sigma_b) is the adapt kernel's
}

template <class Spec>
__device__ __forceinline__ void load_step_regs(StepRegs<Spec>& pr, const double* s_cur, const double* s_prop,
                                               const double* par, const ParamLayout& pl, int B) {
    constexpr int p = Spec::d.p, A = Spec::d.A, O = Spec::d.O;
    sfor<O>([&](auto kc) { pr.isig[kc] = 1.0 / par[pl.sigmas + kc]; });
    sfor<p>([&](auto j) { pr.gc[j] = s_cur[B + j]; pr.gp[j] = s_prop[B + j]; });
    sfor<A>([&](auto j) { pr.alc[j] = s_cur[B + p + j]; pr.alp[j] = s_prop[B + p + j]; });
}

// cold path of the TMA kernel (records larger than a stage): kept out of line; takes and returns
// everything by value so that the caller's registers are not forced into local memory
struct StepSums { double cur, prop, lw; };
template <class Spec>
__device__ __noinline__ StepSums subject_step_direct(const double* rec, const double* crec, double* st, const double* rg,
                                                     const double* s_cur, const double* s_prop, const double* s_invD,
                                                     const double* par, const int o_sigmas, const int B,
                                                     const bool last_acc, const int lane, const unsigned gmask) {
    StepRegs<Spec> pr;
    constexpr int p = Spec::d.p, A = Spec::d.A, O = Spec::d.O;
    sfor<O>([&](auto kc) { pr.isig[kc] = 1.0 / par[o_sigmas + kc]; });
    sfor<p>([&](auto j) { pr.gc[j] = s_cur[B + j]; pr.gp[j] = s_prop[B + j]; });
    sfor<A>([&](auto j) { pr.alc[j] = s_cur[B + p + j]; pr.alp[j] = s_prop[B + p + j]; });
    StepSums r{0.0, 0.0, 0.0};
    subject_step<Spec>(rec, crec, st, rg, st, s_cur, s_prop, s_invD, pr, last_acc, lane, gmask, r.cur, r.prop, r.lw);
    return r;
}

// -------------------------------------------------------------------------------------------
// direct variant: every lane group reads its subject's records straight from global memory
// -------------------------------------------------------------------------------------------
template <class Spec>
__global__ void __launch_bounds__(256, JMB_STATIC_MIN_CTAS)
jmb_step_static(const int B, const __grid_constant__ ParamLayout pl, const __grid_constant__ ChainConst cc,
                const __grid_constant__ StepArgs a) {
    constexpr int G = Spec::d.G, q = Spec::d.q, p = Spec::d.p, A = Spec::d.A, SS = Spec::L.state_stride;
    constexpr int THREADS = 256, GROUPS = THREADS / G;
    const int P = B + p + A;
    extern __shared__ double smem[];
    double* s_cur = smem;                 // [P] current Bs_gammas | gammas | alphas
    double* s_prop = s_cur + P;           // [P] proposed
    double* s_invD = s_prop + P;          // [q*q]
    double* s_red = s_invD + q * q;       // [GROUPS][3]

    const int tid = threadIdx.x;
    for (int j = tid; j < P; j += THREADS) { s_cur[j] = a.par[pl.cur + j]; s_prop[j] = a.par[pl.prop + j]; }
    for (int j = tid; j < q * q; j += THREADS) s_invD[j] = a.par[pl.invD + j];
    const int i_blk = a.ctl->i;
    const bool last_acc = a.ctl->last_accept != 0;
    __syncthreads();
    StepRegs<Spec> pr;
    load_step_regs<Spec>(pr, s_cur, s_prop, a.par, pl, B);

    const int grp = tid / G, lane = tid % G;
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (0xFFFFu << (16 * ((tid & 31) >> 4)));
    const double* rng_it = a.rng + (long long)(i_blk % a.rng_chunk) * a.n * a.rng_stride;
    double acc_cur = 0.0, acc_prop = 0.0, acc_lw = 0.0;
    const int s_begin = blockIdx.x * a.subjects_per_cta;
    const int s_end = min(a.n, s_begin + a.subjects_per_cta);
    for (int s = s_begin + grp; s < s_end; s += GROUPS) {
        double* st = a.state + (long long)s * SS;
        subject_step<Spec>(a.drec + a.doff[s], a.crec + a.coff[s], st, rng_it + (long long)s * a.rng_stride, st,
                           s_cur, s_prop, s_invD, pr, last_acc, lane, gmask, acc_cur, acc_prop, acc_lw);
    }
    if (lane == 0) { s_red[grp * 3] = acc_cur; s_red[grp * 3 + 1] = acc_prop; s_red[grp * 3 + 2] = acc_lw; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int g = 0; g < GROUPS; ++g) t += s_red[g * 3 + tid];
        a.partials[blockIdx.x * 3 + tid] = t;
    }
}

// -------------------------------------------------------------------------------------------
// TMA variant: each warp runs a two-stage cp.async.bulk + mbarrier pipeline over its contiguous
// run of subjects; while it computes subject(s) t the records and state of t+1 are already in
// flight into its other shared-memory stage.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

struct TmaArgs {
    int cap_d, cap_c;        // per-stage capacity (doubles) for the design / per-draw records of one warp step
    int subjects_per_cta;    // multiple of 8 * SPW
};

template <class Spec>
__global__ void __launch_bounds__(JMB_TMA_THREADS, JMB_TMA_CTAS)
jmb_step_tma(const int B, const __grid_constant__ ParamLayout pl, const __grid_constant__ ChainConst cc,
             const __grid_constant__ StepArgs a, const TmaArgs ta) {
    constexpr int G = Spec::d.G, q = Spec::d.q, p = Spec::d.p, A = Spec::d.A, SS = Spec::L.state_stride;
    constexpr int THREADS = kTmaThreads, WARPS = THREADS / 32, GROUPS = THREADS / G, SPW = 32 / G;
    const int P = B + p + A;
    extern __shared__ double smem[];
    double* s_cur = smem;                          // [P]
    double* s_prop = s_cur + P;                    // [P]
    double* s_invD = s_prop + P;                   // [q*q]
    double* s_red = s_invD + q * q;                // [GROUPS][3]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + GROUPS * 3);   // [WARPS][2]
    const int head = (2 * P + q * q + GROUPS * 3 + WARPS * 2 + 1) / 2 * 2;   // keep 16-byte alignment
    constexpr int RS = (q + 2) / 2 * 2;            // rng stride: q increments + log u, padded to even
    const int stage_doubles = ta.cap_d + ta.cap_c + SPW * (SS + RS);
    double* s_stage = smem + head;                 // [WARPS][2][stage_doubles]

    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    // Programmatic dependent launch: this kernel may be scheduled while the previous finalize kernel still runs (launch
    // latency and CTA set-up overlap it); pdl_wait() returns once the parameter block and the loop control are final.
    if (tid < WARPS * 2) mbar_init(&s_bar[tid], 1);
    pdl_wait();
    pdl_trigger();             // only now (the previous finalize kernel is complete) may the next finalize kernel be
                               // scheduled: it preloads the parameter block before its own wait
    for (int j = tid; j < P; j += THREADS) { s_cur[j] = a.par[pl.cur + j]; s_prop[j] = a.par[pl.prop + j]; }
    for (int j = tid; j < q * q; j += THREADS) s_invD[j] = a.par[pl.invD + j];
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const int i_blk = a.ctl->i;
    const bool last_acc = a.ctl->last_accept != 0;
    __syncthreads();
    StepRegs<Spec> pr;
    load_step_regs<Spec>(pr, s_cur, s_prop, a.par, pl, B);
    const double* rng_it = a.rng + (long long)(i_blk % a.rng_chunk) * a.n * RS;

    const int half = (G == 16) ? (lane32 >> 4) : 0, lane = lane32 % G;
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (0xFFFFu << (16 * half));
    double acc_cur = 0.0, acc_prop = 0.0, acc_lw = 0.0;

    // this warp's contiguous run of subjects
    const int c_begin = blockIdx.x * ta.subjects_per_cta;
    const int per_warp = ta.subjects_per_cta / WARPS;
    const int w_begin = min(a.n, c_begin + warp * per_warp);
    const int w_end = min(a.n, w_begin + per_warp);
    const int steps = (w_end - w_begin + SPW - 1) / SPW;
    uint64_t* bar = s_bar + warp * 2;
    double* stage0 = s_stage + (size_t)warp * 2 * stage_doubles;

    // Offsets of warp step t (subjects [sA, sB)): loaded two steps ahead by every lane (uniform
    // loads that complete under the compute of the steps in between).
    struct Off { long long d0, dm, d1, c0, cm, c1; int sA, sB; };
    auto load_off = [&](int t) {
        Off o;
        o.sA = w_begin + t * SPW; o.sB = min(w_end, o.sA + SPW);
        const int sM = min(o.sB, o.sA + 1);
        o.d0 = a.doff[o.sA]; o.dm = a.doff[sM]; o.d1 = a.doff[o.sB];
        o.c0 = a.coff[o.sA]; o.cm = a.coff[sM]; o.c1 = a.coff[o.sB];
        return o;
    };
    auto issue = [&](const Off& o, int stg) {     // lane32 == 0 only
        double* dst = stage0 + (size_t)stg * stage_doubles;
        const uint32_t bd = (uint32_t)(o.d1 - o.d0) * 8u, bc = (uint32_t)(o.c1 - o.c0) * 8u, bs = (uint32_t)(o.sB - o.sA) * SS * 8u;
        const uint32_t br = (uint32_t)(o.sB - o.sA) * RS * 8u;
        mbar_expect_tx(&bar[stg], bd + bc + bs + br);
        bulk_g2s(dst, a.drec + o.d0, bd, &bar[stg]);
        bulk_g2s(dst + ta.cap_d, a.crec + o.c0, bc, &bar[stg]);
        bulk_g2s(dst + ta.cap_d + ta.cap_c, a.state + (long long)o.sA * SS, bs, &bar[stg]);
        bulk_g2s(dst + ta.cap_d + ta.cap_c + SPW * SS, rng_it + (long long)o.sA * RS, br, &bar[stg]);
    };
    // A step whose records exceed the stage capacity (rare subjects with very many visits) is not
    // staged: its lane groups read global memory directly.
    auto fits = [&](const Off& o) { return (o.d1 - o.d0) <= ta.cap_d && (o.c1 - o.c0) <= ta.cap_c; };
    struct Pending { long long d0, c0; int dmid, cmid; bool staged; };
    auto pend = [&](const Off& o) { return Pending{o.d0, o.c0, (int)(o.dm - o.d0), (int)(o.cm - o.c0), fits(o)}; };
    Pending pd0{}, pd1{};
    uint32_t ph0 = 0, ph1 = 0;                      // mbarrier phase parity per stage
    if (steps > 0) { const Off o = load_off(0); pd0 = pend(o); if (lane32 == 0 && pd0.staged) issue(o, 0); }
    if (steps > 1) { const Off o = load_off(1); pd1 = pend(o); if (lane32 == 0 && pd1.staged) issue(o, 1); }
    for (int t = 0; t < steps; ++t) {
        const int stg = t & 1;
        const bool more = (t + 2 < steps);
        Off nxt{};
        if (more) nxt = load_off(t + 2);
        const Pending cur = stg ? pd1 : pd0;
        const int s = w_begin + t * SPW + half;
        const int dmid = (half == 1) ? cur.dmid : 0, cmid = (half == 1) ? cur.cmid : 0;
        if (cur.staged) {
            mbar_wait(&bar[stg], stg ? ph1 : ph0);
            if (stg) ph1 ^= 1u; else ph0 ^= 1u;
            // addresses derived from the shared-memory window so that the loads compile to LDS
            const double* stage = smem + head + ((size_t)warp * 2 + stg) * stage_doubles;
            if (s < w_end)
                subject_step<Spec>(stage + dmid, stage + ta.cap_d + cmid, stage + ta.cap_d + ta.cap_c + half * SS,
                                   stage + ta.cap_d + ta.cap_c + SPW * SS + half * RS, a.state + (long long)s * SS,
                                   s_cur, s_prop, s_invD, pr, last_acc, lane, gmask, acc_cur, acc_prop, acc_lw);
        } else if (s < w_end) {
            const StepSums r = subject_step_direct<Spec>(a.drec + cur.d0 + dmid, a.crec + cur.c0 + cmid,
                                                         a.state + (long long)s * SS, rng_it + (long long)s * RS, s_cur, s_prop,
                                                         s_invD, a.par, pl.sigmas, B, last_acc, lane, gmask);
            acc_cur += r.cur; acc_prop += r.prop; acc_lw += r.lw;
        }
        __syncwarp();
        if (more) {
            const Pending np = pend(nxt);
            if (stg) pd1 = np; else pd0 = np;
            if (lane32 == 0 && np.staged) {
                // the warp's generic-proxy reads of this stage (ordered before this point by the __syncwarp above) must be ordered
                // before the async-proxy writes of the bulk copies that re-fill it
#if !defined(JMB_EMU)
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
                issue(nxt, stg);
            }
        }
    }
    const int grp = tid / G;
    if (lane == 0) { s_red[grp * 3] = acc_cur; s_red[grp * 3 + 1] = acc_prop; s_red[grp * 3 + 2] = acc_lw; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int g = 0; g < GROUPS; ++g) t += s_red[g * 3 + tid];
        a.partials[blockIdx.x * 3 + tid] = t;
    }
}

}  // namespace jmb
