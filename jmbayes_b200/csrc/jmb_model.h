// jmb_model.h - model descriptor and per-subject record layout, usable at compile time (static
// step kernels) and at run time (host packer, generic step kernel).
//
// A subject's designs live in two contiguous records (units: doubles):
//
//   design record (shared by all chains of a fit)
//     [0] event  [o_V ..] int32 visit counts per outcome
//     [o_W2c ..] W2 row when the W2s rows repeat within the subject (w2_const)
//     hazard rows, one per lane r = 0..RP-1 (0: event time, 1..K: quadrature, K+1..2K: lower-
//     interval quadrature), each field stored as [column][RP]:
//       Pw | int32 band offsets of the B-spline rows | WB band values | W2 columns (unless
//       w2_const) | per term: stored ZZ columns, stored U columns
//     longitudinal rows, per outcome: y[v] | stored Z columns [v] each
//   per-draw record: XXbetas/XXsbetas rows [T][RP] | Xbetas rows per outcome
//
// The upper Cholesky factors of Covs$b (packed by column, q(q+1)/2 doubles per subject) live in
// their own array: only the once-per-block proposal kernel reads them.
//
// Multi-state fits (ModelDesc::MS; src/fast_LAPRWM.cpp:519-529): a subject has ntr >= 1 transition rows, each with its
// own event-time row and K quadrature rows.  The hazard-row part of the design record ([o_Pw, o_long), HB doubles) is
// then repeated ntr times, followed by the longitudinal rows; rec[0] holds ntr instead of the event indicator, and the
// event indicator of a transition sits in its block's Pw slot of lane 0 (the event-time row has no quadrature weight).
// The per-draw record repeats its [T][RP] part ntr times in the same way.  w2_const is never set for these fits.
//
// "Stored" columns: an all-ones first column of Z / ZZ (random intercept) and an all-ones U
// (default interaction design ~1) are not stored; the flags below say so.  jmb_create() decides
// the flags by scanning the data, so they never change results.
#pragma once
#include "../../include/jmb200.h"

#if defined(__CUDACC__)
#define JMB_HD __host__ __device__
#else
#define JMB_HD
#endif

namespace jmb {

constexpr int kMaxO = JMB_MAX_OUTCOMES, kMaxT = JMB_MAX_TERMS, kMaxQ = JMB_MAX_Q;
constexpr int kMaxRT = JMB_MAX_RT, kMaxUT = JMB_MAX_UT;
// per-subject per-chain state after b[q]: sigma_b, log_pyb, log_pb, accepts in the current block,
// accepts in total, log p(T|b) under the current and under the last proposed survival parameters
constexpr int kStateExtra = 7;

struct ModelDesc {
    int O, q, T, A, p, K, S;        // S = 1, or 2 with interval censoring
    int WB;                         // stored band of the B-spline basis rows (n_Bs when dense)
    int G;                          // lanes per subject (16 or 32); RP == G
    int w2_const;                   // W2s rows equal the W2 row within each subject
    int MS;                         // multi-state fit: variable number of hazard-row blocks per subject
    int fam[kMaxO], link[kMaxO], qk[kMaxO], z_ones0[kMaxO];
    int re_inds[kMaxO][kMaxQ];
    int rt[kMaxT], ut[kMaxT], tf[kMaxT], zz_ones0[kMaxT], u_ones[kMaxT];
    int re2[kMaxT][kMaxRT];
    int col[kMaxT][kMaxUT];
};

struct RecordLayout {
    int RP, o_V, o_W2c, o_Pw, o_W1off, o_W1v, o_W2, o_ZZ[kMaxT], o_U[kMaxT], o_long, c_long;
    int long_cols[kMaxO];           // stored doubles per longitudinal row of outcome k (y + Z cols)
    int state_stride;               // q + kStateExtra, padded to even
    int HB;                         // doubles of one hazard-row block = o_long - o_Pw (stride between transition blocks)
};

JMB_HD constexpr RecordLayout make_layout(const ModelDesc& d) {
    RecordLayout L{};
    L.RP = d.G;
    L.o_V = 1;
    int o = 1 + (d.O + 1) / 2;
    L.o_W2c = o;
    if (d.w2_const) o += d.p;
    o = (o + 3) / 4 * 4;            // hazard rows start on a 32-byte sector (records are 32-byte aligned too)
    L.o_Pw = o; o += L.RP;
    L.o_W1off = o; o += L.RP / 2;
    L.o_W1v = o; o += d.WB * L.RP;
    L.o_W2 = o;
    if (!d.w2_const) o += d.p * L.RP;
    for (int t = 0; t < d.T; ++t) {
        L.o_ZZ[t] = o; o += (d.rt[t] - d.zz_ones0[t]) * L.RP;
        L.o_U[t] = o; o += (d.u_ones[t] ? 0 : d.ut[t]) * L.RP;
    }
    L.o_long = o;
    L.HB = L.o_long - L.o_Pw;
    L.c_long = d.T * L.RP;
    for (int k = 0; k < d.O; ++k) L.long_cols[k] = 1 + d.qk[k] - d.z_ones0[k];
    L.state_stride = (d.q + kStateExtra + 1) / 2 * 2;
    return L;
}

JMB_HD constexpr bool same_desc(const ModelDesc& a, const ModelDesc& b) {
    if (a.O != b.O || a.q != b.q || a.T != b.T || a.A != b.A || a.p != b.p || a.K != b.K || a.S != b.S ||
        a.WB != b.WB || a.G != b.G || a.w2_const != b.w2_const || a.MS != b.MS) return false;
    for (int k = 0; k < a.O; ++k) {
        if (a.fam[k] != b.fam[k] || a.link[k] != b.link[k] || a.qk[k] != b.qk[k] || a.z_ones0[k] != b.z_ones0[k]) return false;
        for (int j = 0; j < a.qk[k]; ++j) if (a.re_inds[k][j] != b.re_inds[k][j]) return false;
    }
    for (int t = 0; t < a.T; ++t) {
        if (a.rt[t] != b.rt[t] || a.ut[t] != b.ut[t] || a.tf[t] != b.tf[t] || a.zz_ones0[t] != b.zz_ones0[t] ||
            a.u_ones[t] != b.u_ones[t]) return false;
        for (int j = 0; j < a.rt[t]; ++j) if (a.re2[t][j] != b.re2[t][j]) return false;
        for (int u = 0; u < a.ut[t]; ++u) if (a.col[t][u] != b.col[t][u]) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// Precompiled model shapes (BASELINE.json configs).  Any other model runs the generic kernel.
// ---------------------------------------------------------------------------------------------
// configs 1 and 2: one gaussian outcome, random intercept + slope, current-value association
JMB_HD constexpr ModelDesc desc_value_gaussian() {
    ModelDesc d{};
    d.O = 1; d.q = 2; d.T = 1; d.A = 1; d.p = 2; d.K = 15; d.S = 1; d.WB = 4; d.G = 16; d.w2_const = 1;
    d.fam[0] = JMB_FAM_GAUSSIAN; d.link[0] = JMB_LINK_IDENTITY; d.qk[0] = 2; d.z_ones0[0] = 1;
    d.re_inds[0][0] = 0; d.re_inds[0][1] = 1;
    d.rt[0] = 2; d.ut[0] = 1; d.tf[0] = JMB_TF_IDENTITY; d.zz_ones0[0] = 1; d.u_ones[0] = 1;
    d.re2[0][0] = 0; d.re2[0][1] = 1; d.col[0][0] = 0;
    return d;
}

// configs 3 and 4: gaussian + binomial(logit) + poisson(log), q = 6, value + slope per outcome
JMB_HD constexpr ModelDesc desc_three_outcomes(int S) {
    ModelDesc d{};
    d.O = 3; d.q = 6; d.T = 6; d.A = 6; d.p = 2; d.K = 15; d.S = S; d.WB = 4; d.G = (S == 2) ? 32 : 16;
    d.w2_const = 1;
    d.fam[0] = JMB_FAM_GAUSSIAN; d.link[0] = JMB_LINK_IDENTITY;
    d.fam[1] = JMB_FAM_BINOMIAL; d.link[1] = JMB_LINK_LOGIT;
    d.fam[2] = JMB_FAM_POISSON; d.link[2] = JMB_LINK_LOG;
    for (int k = 0; k < 3; ++k) {
        d.qk[k] = 2; d.z_ones0[k] = 1; d.re_inds[k][0] = 2 * k; d.re_inds[k][1] = 2 * k + 1;
        const int tv = 2 * k, ts = 2 * k + 1;
        d.rt[tv] = 2; d.ut[tv] = 1; d.tf[tv] = JMB_TF_IDENTITY; d.zz_ones0[tv] = 1; d.u_ones[tv] = 1;
        d.re2[tv][0] = 2 * k; d.re2[tv][1] = 2 * k + 1; d.col[tv][0] = tv;
        d.rt[ts] = 1; d.ut[ts] = 1; d.tf[ts] = JMB_TF_IDENTITY; d.zz_ones0[ts] = 1; d.u_ones[ts] = 1;
        d.re2[ts][0] = 2 * k + 1; d.col[ts][0] = ts;
    }
    return d;
}

// multi-state illness-death shape of bench.py --workload multistate (cohorts.make_multistate_cohort): gaussian + binomial(logit),
// q = 4, value of outcome 1 interacting with the transition (U = 3 stratum indicators) + expit(value of outcome 2),
// transition-specific W2 (p = 3), stratified baseline hazard (3 x 7 B-spline coefficients, band of 4 inside a stratum)
JMB_HD constexpr ModelDesc desc_multistate_illness_death() {
    ModelDesc d{};
    d.O = 2; d.q = 4; d.T = 2; d.A = 4; d.p = 3; d.K = 15; d.S = 1; d.WB = 4; d.G = 16; d.w2_const = 0; d.MS = 1;
    d.fam[0] = JMB_FAM_GAUSSIAN; d.link[0] = JMB_LINK_IDENTITY;
    d.fam[1] = JMB_FAM_BINOMIAL; d.link[1] = JMB_LINK_LOGIT;
    for (int k = 0; k < 2; ++k) {
        d.qk[k] = 2; d.z_ones0[k] = 1; d.re_inds[k][0] = 2 * k; d.re_inds[k][1] = 2 * k + 1;
        d.rt[k] = 2; d.zz_ones0[k] = 1; d.re2[k][0] = 2 * k; d.re2[k][1] = 2 * k + 1;
    }
    d.ut[0] = 3; d.tf[0] = JMB_TF_IDENTITY; d.u_ones[0] = 0; d.col[0][0] = 0; d.col[0][1] = 1; d.col[0][2] = 2;
    d.ut[1] = 1; d.tf[1] = JMB_TF_EXPIT; d.u_ones[1] = 1; d.col[1][0] = 3;
    return d;
}

struct SpecValueGaussian { static constexpr ModelDesc d = desc_value_gaussian(); static constexpr RecordLayout L = make_layout(d); };
struct SpecThreeOutcomes { static constexpr ModelDesc d = desc_three_outcomes(1); static constexpr RecordLayout L = make_layout(d); };
struct SpecThreeOutcomesIC { static constexpr ModelDesc d = desc_three_outcomes(2); static constexpr RecordLayout L = make_layout(d); };
struct SpecMultiState { static constexpr ModelDesc d = desc_multistate_illness_death(); static constexpr RecordLayout L = make_layout(d); };

}  // namespace jmb
