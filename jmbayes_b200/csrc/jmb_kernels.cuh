// jmb_kernels.cuh - the sm_100a kernels of the mvJointModelBayes MCMC iteration.
//
// One MCMC iteration (src/fast_LAPRWM.cpp:666-851) is two launches:
//   jmb_step_kernel     - per subject: proposal, GLMM log-density over the visit segment, log-hazard
//                         and Gauss-Kronrod cumulative hazard under the current AND the proposed
//                         survival parameters, accept/reject, scale adaptation, and the subject's
//                         contribution to the three global sums.  b_i lives in shared memory /
//                         registers between proposal and acceptance; nothing but the new state
//                         (q + 5 doubles) is written back.
//   jmb_finalize_kernel - one CTA: deterministic reduction of the per-CTA partial sums, the
//                         survival-block accept/reject, the Gibbs draws of the precisions, output
//                         of kept iterations, adaptation at block ends, next survival proposal.
// No host round trip is needed inside a chain.
#pragma once
#include "jmb_device.cuh"

namespace jmb {

enum { MODE_STEP = 0, MODE_INIT = 1, MODE_EVAL = 2 };

struct StepArgs {
    const double* drec;      // per-subject design records (shared by all chains of a fit)
    const long long* doff;   // [n+1] offsets into drec (doubles)
    const double* crec;      // per-subject per-draw records
    const long long* coff;   // [n+1]
    const int* rowptr;       // [O][n+1] CSR row pointers of the longitudinal outcomes
    double* state;           // [n][state_stride]
    const double* par;       // chain parameter block
    const Ctl* ctl;
    double* partials;        // [gridDim.x][3]
    double* eval_out;        // MODE_EVAL: [6][n] log_pyb, log_pb, log_h, H, HL, log_ptb
    int n;
    long long subject_offset;
    int mode;
    int subjects_per_cta;
    const double* rng;       // [rng_chunk][n][rng_stride]: proposal increments b_new - b (q) and log u
    int rng_chunk;           // iterations covered by the buffer (the block position i indexes it)
    int rng_stride;          // q + 1 rounded up to even
    // multi-state fits: first transition row of each subject (global row index, for MODE_EVAL) and the
    // distance between the six MODE_EVAL output vectors (max(n, nT); n otherwise)
    const int* trow;         // [n+1]
    long long eval_stride;
};

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still running; pdl_wait() blocks until the predecessor has completed and flushed,
// pdl_trigger() lets the successor start.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// -------------------------------------------------------------------------------------------
// generic step kernel: runtime model descriptor, G lanes per subject.  Fallback for model shapes
// without a precompiled specialisation (jmb_step_static.cuh); same results.
// -------------------------------------------------------------------------------------------
// MS = true (multi-state fits, src/fast_LAPRWM.cpp:519-529,669,703): the hazard-row block is repeated once per
// transition row of the subject (jmb_model.h); log p(T|b) is the sum of the blocks' terms, i.e.
// rowsum(log_ptb, idT_rsum).  The survival-step sums are then taken for both the old and the proposed b inside
// the block loop (two more exp per quadrature row) so that nothing per block has to outlive the accept decision.
template <int G, bool MS>
__global__ void __launch_bounds__(256)
jmb_step_kernel(const __grid_constant__ ModelMeta mm, const __grid_constant__ ParamLayout pl,
                const __grid_constant__ ChainConst cc, const __grid_constant__ StepArgs a) {
    constexpr int THREADS = 256;
    constexpr int GROUPS = THREADS / G;
    const ModelDesc& md = mm.d;
    const RecordLayout& L = mm.L;
    // [cur P][prop P][invD q*q][inv_sigma O]
    extern __shared__ double smem[];
    double* s_cur = smem;
    double* s_prop = s_cur + mm.P;
    double* s_invD = s_prop + mm.P;
    double* s_isig = s_invD + md.q * md.q;
    double* s_grp = s_isig + md.O;                         // per group: b_old[QS], b_new[QS], z[QS]
    const int QS = md.q + kStateExtra;
    double* s_red = s_grp + GROUPS * 3 * QS;                // [GROUPS][3]

    const int tid = threadIdx.x;
    for (int j = tid; j < mm.P; j += THREADS) { s_cur[j] = a.par[pl.cur + j]; s_prop[j] = a.par[pl.prop + j]; }
    for (int j = tid; j < md.q * md.q; j += THREADS) s_invD[j] = a.par[pl.invD + j];
    for (int j = tid; j < md.O; j += THREADS) s_isig[j] = 1.0 / a.par[pl.sigmas + j];
    const int i_blk = a.ctl->i, last_acc = a.ctl->last_accept;
    __syncthreads();

    const int grp = tid / G, lane = tid % G;
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (0xFFFFu << (16 * ((tid & 31) >> 4)));
    double* bo = s_grp + grp * 3 * QS;
    double* bn = bo + QS;
    const int q = md.q, RP = L.RP, K = md.K, S = md.S;
    double acc_cur = 0.0, acc_prop = 0.0, acc_lw = 0.0;     // lane 0 of each group
    const int s_begin = blockIdx.x * a.subjects_per_cta;
    const int s_end = min(a.n, s_begin + a.subjects_per_cta);

    for (int s = s_begin + grp; s < s_end; s += GROUPS) {
        const double* rec = a.drec + a.doff[s];
        const double* crec = a.crec + a.coff[s];
        double* st = a.state + (long long)s * L.state_stride;
        const int ntr = MS ? (int)rec[0] : 1;               // transition rows of this subject
        // ---- state + proposal (mvrnorm_cube / extract_b, src/fast_LAPRWM.cpp:123-146,670) ----
        // the draws of this iteration were produced by jmb_block_prep_kernel
        const double* rg = a.rng + ((long long)(i_blk % a.rng_chunk) * a.n + s) * a.rng_stride;
        for (int j = lane; j < QS; j += G) bo[j] = st[j];
        __syncwarp(gmask);
        const double pyb_old = bo[q + 1], pb_old = bo[q + 2];
        for (int j = lane; j < q; j += G) bn[j] = (a.mode == MODE_STEP) ? bo[j] + rg[j] : bo[j];   // :670
        __syncwarp(gmask);

        // ---- log p(b) for the proposal: -0.5 b' invD b (src/fast_LAPRWM.cpp:673) ----
        double v_pb = 0.0;
        for (int j = lane; j < q; j += G) {
            double t = 0.0;
            for (int l = 0; l < q; ++l) t += bn[l] * s_invD[l + j * q];
            v_pb += t * bn[j];
        }

        // ---- GLMM log-density over the visit segment (lin_predF + log_longF, :67-78,168-202) ----
        double v_pyb = 0.0;
        {
            const double* lrec = rec + L.o_long + (ntr - 1) * L.HB;
            const double* lcrec = crec + L.c_long * ntr;
            for (int k = 0; k < md.O; ++k) {
                const int* rp = a.rowptr + (long long)k * (a.n + 1);
                const int v = rp[s + 1] - rp[s];
                const int qk = md.qk[k], fam = md.fam[k], link = md.link[k], ones0 = md.z_ones0[k];
                const double isg = s_isig[k];
                for (int r = lane; r < v; r += G) {
                    double eta_z = ones0 ? bn[md.re_inds[k][0]] : 0.0;
                    for (int j = ones0; j < qk; ++j) eta_z += lrec[(1 + j - ones0) * v + r] * bn[md.re_inds[k][j]];
                    v_pyb += row_logdens(fam, link, lrec[r], lcrec[r] + eta_z, isg);
                }
                lrec += L.long_cols[k] * v;
                lcrec += v;
            }
        }

        // ---- hazard rows: row 0 = event time, 1..K quadrature, K+1..2K lower-interval quadrature
        //      (lin_pred_matF :80-109, log_h :684, H/HL :685-691; survival step :735-741) ----
        // One row per lane (RP == G is enforced at jmb_create).
        const int r = lane;
        const int cls = (r == 0) ? 0 : (r <= K ? 1 : ((S == 2 && r <= 2 * K) ? 2 : 3));
        const int src0 = (tid & 31) - lane;                 // lane 0 of this group inside the warp
        double l_co, l_cn, l_po, l_pn;                       // current/proposed params x old/new b
        double pw = 0.0, H_co = 0.0, H_cn = 0.0, HL_co = 0.0, HL_cn = 0.0, h_co, h_cn, h_po, h_pn;
        double ms_co = 0.0, ms_cn = 0.0, ms_po = 0.0, ms_pn = 0.0;   // MS: sums over the subject's transitions
        for (int tb = 0; tb < ntr; ++tb) {
            const double* hrec = rec + tb * L.HB;            // this transition's hazard rows (tb == 0 unless MS)
            const double* hcrec = crec + tb * L.c_long;
            const int off = reinterpret_cast<const int*>(hrec + L.o_W1off)[r];
            double s1c = 0.0, s1p = 0.0;
            for (int j = 0; j < md.WB; ++j) {
                const double wv = hrec[L.o_W1v + j * RP + r];
                s1c += wv * s_cur[off + j];
                s1p += wv * s_prop[off + j];
            }
            double s2c = 0.0, s2p = 0.0;
            for (int j = 0; j < md.p; ++j) {
                const double wv = md.w2_const ? rec[L.o_W2c + j] : hrec[L.o_W2 + j * RP + r];
                s2c += wv * s_cur[mm.B + j];
                s2p += wv * s_prop[mm.B + j];
            }
            double a_co = 0.0, a_cn = 0.0, a_po = 0.0, a_pn = 0.0;
            const double* alc = s_cur + mm.B + md.p;
            const double* alp = s_prop + mm.B + md.p;
            for (int t = 0; t < md.T; ++t) {
                const double xxb = hcrec[t * RP + r];
                const int ones0 = md.zz_ones0[t];
                double zo = ones0 ? bo[md.re2[t][0]] : 0.0, zn = ones0 ? bn[md.re2[t][0]] : 0.0;
                for (int j = ones0; j < md.rt[t]; ++j) {
                    const double zz = hrec[L.o_ZZ[t] + (j - ones0) * RP + r];
                    zo += zz * bo[md.re2[t][j]];
                    zn += zz * bn[md.re2[t][j]];
                }
                const double vo = trans_fun(md.tf[t], xxb + zo), vn = trans_fun(md.tf[t], xxb + zn);
                for (int u = 0; u < md.ut[t]; ++u) {
                    const double uu = md.u_ones[t] ? 1.0 : hrec[L.o_U[t] + u * RP + r];
                    const int c = md.col[t][u];
                    const double wo = uu * vo, wn = uu * vn;
                    a_co += wo * alc[c]; a_cn += wn * alc[c];
                    a_po += wo * alp[c]; a_pn += wn * alp[c];
                }
            }
            l_co = (s1c + s2c) + a_co; l_cn = (s1c + s2c) + a_cn;
            l_po = (s1p + s2p) + a_po; l_pn = (s1p + s2p) + a_pn;
            pw = (cls == 1 || cls == 2) ? hrec[L.o_Pw + r] : 0.0;
            H_co = 0.0; H_cn = 0.0; HL_co = 0.0; HL_cn = 0.0;
            if (cls == 1) { H_co = pw * exp(l_co); H_cn = pw * exp(l_cn); }
            else if (cls == 2) { HL_co = pw * exp(l_co); HL_cn = pw * exp(l_cn); }
            h_co = l_co; h_cn = l_cn; h_po = l_po; h_pn = l_pn;          // meaningful on lane 0 only
            H_co = group_sum<G>(H_co, gmask);
            H_cn = group_sum<G>(H_cn, gmask);
            if (S == 2) { HL_co = group_sum<G>(HL_co, gmask); HL_cn = group_sum<G>(HL_cn, gmask); }
            h_co = __shfl_sync(gmask, h_co, src0); h_cn = __shfl_sync(gmask, h_cn, src0);
            h_po = __shfl_sync(gmask, h_po, src0); h_pn = __shfl_sync(gmask, h_pn, src0);
            if (MS) {
                // event indicator of this transition: the Pw slot of the event-time row (lane 0 of the block)
                const double ev_tb = hrec[L.o_Pw];
                const double Hp_o = group_sum<G>((cls == 1) ? pw * exp(l_po) : 0.0, gmask);
                const double Hp_n = group_sum<G>((cls == 1) ? pw * exp(l_pn) : 0.0, gmask);
                const double t_cn = log_ptb_term(1, ev_tb, h_cn, H_cn, 0.0);
                ms_co += log_ptb_term(1, ev_tb, h_co, H_co, 0.0);
                ms_cn += t_cn;
                ms_po += log_ptb_term(1, ev_tb, h_po, Hp_o, 0.0);
                ms_pn += log_ptb_term(1, ev_tb, h_pn, Hp_n, 0.0);
                if (a.mode == MODE_EVAL && lane == 0) {      // one value per transition row
                    const long long tr = (long long)a.trow[s] + tb;
                    a.eval_out[2LL * a.eval_stride + tr] = h_cn; a.eval_out[3LL * a.eval_stride + tr] = H_cn;
                    a.eval_out[4LL * a.eval_stride + tr] = H_cn; a.eval_out[5LL * a.eval_stride + tr] = t_cn;
                }
            }
        }
        v_pyb = group_sum<G>(v_pyb, gmask);
        v_pb = -0.5 * group_sum<G>(v_pb, gmask);

        const double event = MS ? 0.0 : rec[0];
        // log p(T|b_old) under the current parameters is carried in the state: slot q+5 when the
        // last survival proposal was rejected, slot q+6 (evaluated at that proposal) when accepted
        const double ptb_co = (a.mode == MODE_STEP) ? (last_acc ? bo[q + 6] : bo[q + 5])
                                                    : (MS ? ms_co : log_ptb_term(S, event, h_co, H_co, HL_co));
        const double ptb_cn = MS ? ms_cn : log_ptb_term(S, event, h_cn, H_cn, HL_cn);

        bool accept;
        double pyb_new = v_pyb, pb_new = v_pb;
        if (a.mode == MODE_STEP) {
            const double cur_lp = pyb_old + ptb_co + pb_old;          // :669
            const double new_lp = pyb_new + ptb_cn + pb_new;          // :703
            const double lu = rg[q];
            accept = lu < (new_lp - cur_lp);                          // :706 (NaN rejects)
        } else {
            accept = true;
        }
        // ---- proposed survival parameters at the accepted b (:735-752) ----
        double ptb_p;
        if (MS) {
            ptb_p = accept ? ms_pn : ms_po;
        } else {
            double Hp = 0.0, HLp = 0.0;
            if (cls == 1) Hp = pw * exp(accept ? l_pn : l_po);
            else if (cls == 2) HLp = pw * exp(accept ? l_pn : l_po);
            Hp = group_sum<G>(Hp, gmask);
            if (S == 2) HLp = group_sum<G>(HLp, gmask);
            ptb_p = log_ptb_term(S, event, accept ? h_pn : h_po, Hp, HLp);
        }
        const double ptb_c = accept ? ptb_cn : ptb_co;
        const double pyb_acc = accept ? pyb_new : pyb_old;
        const double pb_acc = accept ? pb_new : pb_old;
        acc_cur += ptb_c; acc_prop += ptb_p; acc_lw += pyb_acc + pb_acc;

        // ---- write back (:707-724) + per-subject scale adaptation at block ends (:853-859) ----
        if (a.mode == MODE_EVAL) {
            if (lane == 0) {
                a.eval_out[s] = pyb_new; a.eval_out[a.eval_stride + s] = pb_new;
                if (!MS) {
                    a.eval_out[2LL * a.eval_stride + s] = h_cn; a.eval_out[3LL * a.eval_stride + s] = H_cn;
                    a.eval_out[4LL * a.eval_stride + s] = (S == 2) ? HL_cn : H_cn;
                    a.eval_out[5LL * a.eval_stride + s] = ptb_cn;
                }
            }
        } else {
            double accb = bo[q + 3] + ((a.mode == MODE_STEP && accept) ? 1.0 : 0.0);
            double acct = bo[q + 4] + ((a.mode == MODE_STEP && accept) ? 1.0 : 0.0);
            for (int j = lane; j < QS; j += G) {
                double val;
                if (j < q) val = accept ? bn[j] : bo[j];
                else if (j == q) val = bo[q];
                else if (j == q + 1) val = pyb_acc;
                else if (j == q + 2) val = pb_acc;
                else if (j == q + 3) val = accb;
                else if (j == q + 4) val = acct;
                else if (j == q + 5) val = ptb_c;
                else val = (a.mode == MODE_INIT) ? ptb_c : ptb_p;
                st[j] = val;
            }
        }
        __syncwarp(gmask);
    }

    // ---- deterministic per-CTA partial sums ----
    if (lane == 0) { s_red[grp * 3] = acc_cur; s_red[grp * 3 + 1] = acc_prop; s_red[grp * 3 + 2] = acc_lw; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int g = 0; g < GROUPS; ++g) t += s_red[g * 3 + tid];
        a.partials[blockIdx.x * 3 + tid] = t;
    }
}

// -------------------------------------------------------------------------------------------
// helpers for the finalize kernel (single CTA, thread 0 does the scalar logic)
// -------------------------------------------------------------------------------------------
__device__ inline double quad_form_dev(const double* x, const double* mean, const double* Tau, int m) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) {
        double t = 0.0;
        for (int l = 0; l < m; ++l) t += (x[l] - (mean ? mean[l] : 0.0)) * Tau[l + j * m];
        s += t * (x[j] - (mean ? mean[j] : 0.0));
    }
    return s;
}

__device__ inline int chol_upper_dev(const double* Sg, int m, double* H) {
    for (int e = 0; e < m * m; ++e) H[e] = 0.0;
    for (int j = 0; j < m; ++j)
        for (int i = 0; i <= j; ++i) {
            double s = Sg[i + j * m];
            for (int k = 0; k < i; ++k) s -= H[k + i * m] * H[k + j * m];
            if (i == j) { if (!(s > 0.0)) return 1; H[i + j * m] = sqrt(s); }
            else H[i + j * m] = s / H[i + i * m];
        }
    return 0;
}

// Sigma <- bounds_Cov(Sigma + g1 (cov(block') - Sigma)): src/fast_LAPRWM.cpp:421-428,863-867
__device__ inline void adapt_cov_dev(double* Sg, double* Hwork, int m, const double* block, int stride,
                                     int nb, double g1, double eps2, double eps3) {
    if (m == 0) return;
    for (int a = 0; a < m; ++a)
        for (int c = 0; c < m; ++c) {
            double ma = 0.0, mc = 0.0;
            for (int i = 0; i < nb; ++i) { ma += block[a + i * stride]; mc += block[c + i * stride]; }
            ma /= nb; mc /= nb;
            double s = 0.0;
            for (int i = 0; i < nb; ++i) s += (block[a + i * stride] - ma) * (block[c + i * stride] - mc);
            const double cv = s / (nb - 1);
            Hwork[a + c * m] = Sg[a + c * m] + g1 * (cv - Sg[a + c * m]);
        }
    for (int e = 0; e < m * m; ++e) Sg[e] = Hwork[e];
    double det = 0.0;
    if (chol_upper_dev(Sg, m, Hwork) == 0) { det = 1.0; for (int j = 0; j < m; ++j) det *= Hwork[j + j * m] * Hwork[j + j * m]; }
    if (det > eps3) for (int e = 0; e < m * m; ++e) Sg[e] *= eps3 / det;
    for (int j = 0; j < m; ++j) Sg[j + j * m] += eps2;
}

struct FinalizeArgs {
    double* par;
    Ctl* ctl;
    const double* partials;   // [n_partials][3]; ignored when use_reduced
    int n_partials;
    int use_reduced;          // multi-GPU: par[pl.sums..] already holds the all-reduced sums
    int phase;                // 0: regular iteration, 1: chain start (Cholesky + first proposal only)
    int use_smem;             // the parameter block fits the kernel's dynamic shared memory
    // one-shot all-reduce of the three sums over NVLink peer memory (use_reduced == 2): every rank
    // stores its sums into every peer's mailbox and then waits for all the peers' stores
    double* const* peer_mbox; // [nranks] mailbox base of each rank as mapped into this process
    double* my_mbox;          // this rank's own mailbox: [slots][2][nranks][4] (3 sums + stamp)
    int nranks, rank, slot;
    unsigned long long stamp; // strictly increasing per slot; parity selects the mailbox half
    int published;            // the step kernel's last CTA has already stored this rank's sums into the mailboxes (jmb_step_rows)
    // lap_rwm_C_woRE flavour (use_reduced == 3): sp_out[0] - sp_out[1] is the data term at the proposal
    // (phase 1: at the initial values); the current log-posterior is carried in par[pl.cur_lp]
    int wore; const double* sp_out;
    // outputs of kept iterations (device buffers laid out like jmb_chain_out)
    double* o_Bs; double* o_g; double* o_a; double* o_phi_g; double* o_phi_a; double* o_tau_Bs;
    double* o_tau_g; double* o_tau_a; double* o_tau_td; double* o_lw; double* trace;
};

// survival proposal for iteration `iter`: cur + sqrt(sigma) * z' chol(Sigma)  (:659-661,732-734)
__device__ inline void proposal_normals(const ModelMeta& mm, const ChainConst& cc, int iter, int tid, int nthreads, double* s_z) {
    for (int s = tid; s < (mm.P + 1) / 2; s += nthreads) {
        double z0, z1;
        rnorm_pair(cc.seed, cc.chain_id, 0u, (uint32_t)iter, (uint32_t)s, ST_SURV, z0, z1);
        s_z[2 * s] = z0; s_z[2 * s + 1] = z1;
    }
}
__device__ inline void make_proposal(const ModelMeta& mm, const ParamLayout& pl, const ChainConst& cc,
                                     double* par, int iter, int tid, int nthreads, double* s_z, bool have_z = false) {
    const int B = mm.B, p = mm.d.p, A = mm.d.A;
    if (!have_z) proposal_normals(mm, cc, iter, tid, nthreads, s_z);
    __syncthreads();
    for (int j = tid; j < mm.P; j += nthreads) {
        int base, m, jj, Ho, sg;
        if (j < B) { base = 0; m = B; jj = j; Ho = pl.H_Bs; sg = 0; }
        else if (j < B + p) { base = B; m = p; jj = j - B; Ho = pl.H_g; sg = 1; }
        else { base = B + p; m = A; jj = j - B - p; Ho = pl.H_a; sg = 2; }
        double pj = 0.0;
        for (int l = 0; l <= jj; ++l) pj += s_z[base + l] * par[Ho + l + jj * m];
        par[pl.prop + j] = par[pl.cur + j] + sqrt(par[pl.sig + sg]) * pj;
    }
}

// Mailbox entry of one rank's three sums: six 8-byte words, each (stamp << 32) | one half of a double.  An 8-byte store is one
// NVLink transaction, so a word whose upper half carries the expected stamp carries its data too: neither a system-scope fence
// between data and flag on the sender (a round trip over NVLink) nor a separate flag is needed.
constexpr int kMboxEntry = 8;            // doubles per (slot, parity, rank) entry: six words used
__device__ __forceinline__ void mbox_send(double* entry, const double v0, const double v1, const double v2, const unsigned long long stamp) {
    volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(entry);
    const unsigned long long hi = (stamp & 0xFFFFFFFFull) << 32;
    const double v[3] = {v0, v1, v2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned long long b = (unsigned long long)__double_as_longlong(v[k]);
        w[2 * k] = hi | (b & 0xFFFFFFFFull);
        w[2 * k + 1] = hi | (b >> 32);
    }
}
// one word per thread: word k (0..5) of an entry
__device__ __forceinline__ void mbox_send_word(double* entry, const int k, const double v, const unsigned long long stamp) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    reinterpret_cast<volatile unsigned long long*>(entry)[k] = ((stamp & 0xFFFFFFFFull) << 32) | ((k & 1) ? (b >> 32) : (b & 0xFFFFFFFFull));
}
__device__ __forceinline__ bool mbox_recv_word(const double* entry, const int k, const unsigned long long stamp, unsigned int& half) {
    const volatile unsigned long long* w = reinterpret_cast<const volatile unsigned long long*>(entry) + k;
    const unsigned long long want = stamp & 0xFFFFFFFFull;
    unsigned long long x;
    long long spins = 0;
    while (((x = *w) >> 32) != want) { if (++spins > 200000000LL) return false; }
    half = (unsigned int)(x & 0xFFFFFFFFull);
    return true;
}
__global__ void __launch_bounds__(128)
jmb_finalize_kernel(const __grid_constant__ ModelMeta mm, const __grid_constant__ ParamLayout pl,
                    const __grid_constant__ ChainConst cc, const __grid_constant__ FinalizeArgs a) {
    __shared__ double s_sum[128 * 3];
    __shared__ double s_tot[3];
    __shared__ int s_next_iter, s_mats_changed;
    __shared__ double s_z[JMB_MAX_BS + JMB_MAX_GAMMAS + JMB_MAX_ALPHAS + 2];
    __shared__ double s_q[4][JMB_MAX_BS + JMB_MAX_GAMMAS + JMB_MAX_ALPHAS];
    const int tid = threadIdx.x;
    double* gpar = a.par;
    const int B = mm.B, p = mm.d.p, A = mm.d.A;
    // work on a shared-memory copy of the parameter block (everything before the adaptCov history):
    // the scalar logic below makes dozens of dependent accesses, each a 600-cycle trip in global memory
    extern __shared__ double s_par[];
    double* par = a.use_smem ? s_par : gpar;
    if (a.use_smem) {          // the parameter block is written by finalize kernels only: safe before pdl_wait()
        for (int j = tid; j < pl.block_par; j += blockDim.x) s_par[j] = gpar[j];
        __syncthreads();
    }
    // Everything random in this kernel depends only on (seed, chain, iteration), and the loop control is written by
    // finalize kernels only: draw it now, under the tail of the step kernel.  Gamma draws are taken at unit scale (the
    // data-dependent scale multiplies the same number afterwards: bitwise the value rgamma_dev would return).
    __shared__ double s_lu, s_gam[4 + JMB_MAX_RISKS + JMB_MAX_TERMS + JMB_MAX_GAMMAS + 2 * JMB_MAX_ALPHAS];
    const int nR = cc.nRisks > 0 ? cc.nRisks : 1;       // Gibbs draws of tau_Bs_gammas: one per stratum (:777-782)
    const int iter_pre = a.ctl->iter;
    if (a.phase == 0) {
        int nd = nR + cc.n_td;
        const int d_g = nd; if (cc.shrink_gammas) nd += p + 1;
        const int d_a = nd; if (cc.shrink_alphas) nd += A + 1 + (cc.double_gamma_alphas ? A + 1 : 0);
        for (int dd = tid; dd < nd; dd += blockDim.x) {
            double shape;
            if (dd < nR) shape = cc.Apost_tau_Bs;
            else if (dd < d_g) shape = cc.Apost_tau_td;
            else if (dd < d_a) shape = (dd - d_g < p) ? cc.Apost_phi_g : cc.Apost_tau_g;
            else {
                const int k = dd - d_a;
                shape = (k < A) ? cc.Apost_phi_a : (k == A ? cc.Apost_tau_a : (k < 2 * A + 1 ? cc.Apost_nu_a : cc.Apost_xi_a));
            }
            s_gam[dd] = rgamma_dev(cc.seed, cc.chain_id, (uint32_t)dd, (uint32_t)iter_pre, shape, 1.0);
        }
        if (tid == blockDim.x - 1) s_lu = log_unif(cc.seed, cc.chain_id, 0u, (uint32_t)iter_pre, 0xFFFFFFFFu, ST_SURV);
        proposal_normals(mm, cc, iter_pre + 1, tid, blockDim.x, s_z);
    }
    // ---- quadratic forms of the three Gaussian priors, one coordinate per thread ----
    // thread j < P handles coordinate j of (Bs_gammas | gammas | alphas) against its block's Tau:
    //   s_q[0][j] = ((cur - mean)' Tau)_j (cur - mean)_j   s_q[1][j]: same at the proposal
    //   s_q[2][j] / s_q[3][j]: (x' Tau_Bs)_j x_j without the mean, for the tau_Bs_gammas draw
    // then the eight block sums (s_qs: current / proposed x three blocks, the two raw Bs sums), in coordinate order.
    // They depend on the parameter block only (written by finalize kernels), so on one rank they run before the dependency wait,
    // under the tail of the step kernel.
    __shared__ double s_qs[8];
    auto quad_forms = [&]() {
        double* cur = par + pl.cur;
        double* prop = par + pl.prop;
        if (tid < mm.P) {
            int base, m, To, Mo;
            if (tid < B) { base = 0; m = B; To = pl.Tau_Bs; Mo = pl.mean_Bs; }
            else if (tid < B + p) { base = B; m = p; To = pl.Tau_g; Mo = pl.mean_g; }
            else { base = B + p; m = A; To = pl.Tau_a; Mo = pl.mean_a; }
            const int jj = tid - base;
            // stratified baseline hazard: the tau_Bs_gammas draw of a stratum uses its diagonal block of Tau_Bs_gammas only
            int sf = 0, sl = m - 1;
            if (cc.nRisks > 0 && tid < B)
                for (int k = 0; k < cc.nRisks; ++k) if (jj >= cc.kn_first[k] && jj <= cc.kn_last[k]) { sf = cc.kn_first[k]; sl = cc.kn_last[k]; }
            double tc = 0.0, tp = 0.0, rc = 0.0, rp = 0.0;
            for (int l = 0; l < m; ++l) {
                const double tau = par[To + l + jj * m], mu = par[Mo + l];
                const double xc = cur[base + l], xp = prop[base + l];
                tc += (xc - mu) * tau; tp += (xp - mu) * tau;
                if (l >= sf && l <= sl) { rc += xc * tau; rp += xp * tau; }
            }
            const double mu = par[Mo + jj];
            s_q[0][tid] = tc * (cur[tid] - mu); s_q[1][tid] = tp * (prop[tid] - mu);
            s_q[2][tid] = rc * cur[tid]; s_q[3][tid] = rp * prop[tid];
        }
        __syncthreads();
        if (tid < 8) {
            // 0..2: current, blocks Bs | gammas | alphas; 3..5: proposed; 6 / 7: raw Bs sums (current / proposed)
            const int blk = tid < 6 ? tid % 3 : 0, row = tid < 6 ? tid / 3 : tid - 4;
            const int lo = blk == 0 ? 0 : (blk == 1 ? B : B + p), hi = blk == 0 ? B : (blk == 1 ? B + p : mm.P);
            double t = 0.0;
            for (int j = lo; j < hi; ++j) t += s_q[row][j];
            s_qs[tid] = t;
        }
    };
    if (a.phase == 0) quad_forms();
    pdl_wait();                // the step kernel (partials, state) has completed
    pdl_trigger();             // the next step kernel may start its prologue (records, state and draws are final)

    if (a.phase == 1) {
        if (tid == 0) {
            int e = chol_upper_dev(par + pl.S_Bs, B, par + pl.H_Bs);
            if (p) e |= chol_upper_dev(par + pl.S_g, p, par + pl.H_g);
            e |= chol_upper_dev(par + pl.S_a, A, par + pl.H_a);
            if (e) a.ctl->error = 2;
        }
        __syncthreads();
        if (a.wore && tid == 0) {     // current_logpost at the initial values: logPosterior_woRE, :898-914,1024-1028
            const double* cur = par + pl.cur;
            par[pl.cur_lp] = (a.sp_out[0] - a.sp_out[1])
                - 0.5 * par[pl.tau_Bs] * quad_form_dev(cur, par + pl.mean_Bs, par + pl.Tau_Bs, B)
                - 0.5 * par[pl.tau_g] * quad_form_dev(cur + B, par + pl.mean_g, par + pl.Tau_g, p)
                - 0.5 * par[pl.tau_a] * quad_form_dev(cur + B + p, par + pl.mean_a, par + pl.Tau_a, A);
        }
        make_proposal(mm, pl, cc, par, a.ctl->iter, tid, blockDim.x, s_z);
        if (a.use_smem) {
            __syncthreads();
            for (int j = tid; j < pl.block_par; j += blockDim.x) gpar[j] = s_par[j];
        }
        return;
    }

    // ---- deterministic reduction of the per-CTA partials ----
    if (a.use_reduced == 1) {
        if (tid < 3) s_tot[tid] = gpar[pl.sums + tid];
    } else if (a.use_reduced == 3) {
        if (tid == 0) { s_tot[0] = 0.0; s_tot[1] = a.sp_out[0] - a.sp_out[1]; s_tot[2] = 0.0; }
    } else if (!(a.use_reduced == 2 && a.published)) {
        // warp c sums component c: lane-strided partial sums (independent loads), then a butterfly - fixed order, one barrier
        if (tid < 96) {
            const int c = tid >> 5, ln = tid & 31;
            double t = 0.0;
            for (int i = ln; i < a.n_partials; i += 32) t += a.partials[i * 3 + c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            if (ln == 0) s_tot[c] = t;
        }
    }
    __syncthreads();
    if (a.use_reduced == 2) {
        // ---- fused all-reduce over peer memory: compute (local sums) and collective in one kernel ----
        const int par2 = (int)(a.stamp & 1ull);
        const long long base = ((long long)(a.slot * 2 + par2) * a.nranks);
        // one thread per (peer, word): six independent 8-byte stores per peer
        if (!a.published)
            for (int w = tid; w < 6 * a.nranks; w += blockDim.x)
                mbox_send_word(a.peer_mbox[w / 6] + (base + a.rank) * kMboxEntry, w % 6, s_tot[(w % 6) >> 1], a.stamp);
        // the peers' sums are collected after the quadratic forms below, which do not depend on them: their arithmetic hides
        // part of the NVLink flight
    }

    if (a.use_reduced == 2) {
        // ---- second half of the fused all-reduce: wait for every rank's stamp in this rank's mailbox, add in rank order ----
        const int par2 = (int)(a.stamp & 1ull);
        const long long base = ((long long)(a.slot * 2 + par2) * a.nranks);
        // one thread per (rank, word): the six words of an entry are awaited side by side, not one L2 round trip after the other
        unsigned int* s_half = reinterpret_cast<unsigned int*>(s_sum + 3 * 64);          // [nranks][6] (s_sum has 384 doubles; nranks <= 64)
        for (int w = tid; w < 6 * a.nranks; w += blockDim.x) {
            unsigned int half = 0u;
            if (!mbox_recv_word(a.my_mbox + (base + w / 6) * kMboxEntry, w % 6, a.stamp, half)) a.ctl->error = 3;   // a peer died: give up instead of hanging
            s_half[w] = half;
        }
        __syncthreads();
        for (int v = tid; v < 3 * a.nranks; v += blockDim.x)      // value v = (rank, component): its two halves are words 2v, 2v + 1
            s_sum[v] = __longlong_as_double((long long)((unsigned long long)s_half[2 * v] | ((unsigned long long)s_half[2 * v + 1] << 32)));
        __syncthreads();
        if (tid < 3) {
            double t = 0.0;
            for (int r = 0; r < a.nranks; ++r) t += s_sum[r * 3 + tid];     // rank order: identical on every rank
            s_tot[tid] = t;
        }
        __syncthreads();
    }

    if (tid == 0) {
        s_mats_changed = 0;
        Ctl* ctl = a.ctl;
        const int iter = ctl->iter;
        double* cur = par + pl.cur;
        double* prop = par + pl.prop;
        if (a.trace && !a.wore) { a.trace[2LL * iter] = s_tot[0]; a.trace[2LL * iter + 1] = s_tot[1]; }
        // ---- survival-block RWM (:727-770); tau = 1 for Bs_gammas (reference behaviour) ----
        const double tau_g = par[pl.tau_g], tau_a = par[pl.tau_a];
        const double qc[3] = {s_qs[0], s_qs[1], s_qs[2]}, qp[3] = {s_qs[3], s_qs[4], s_qs[5]}, rawc = s_qs[6], rawp = s_qs[7];
        // lap_rwm_C: tau = 1 for Bs_gammas (:727,753); woRE: the sampled tau_Bs_gammas (:911) and a carried
        // current value that is only refreshed on acceptance (:1031-1053)
        const double tBs = a.wore ? par[pl.tau_Bs] : 1.0;
        const double cur_lp = a.wore ? par[pl.cur_lp] : s_tot[0] - 0.5 * qc[0] - 0.5 * tau_g * qc[1] - 0.5 * tau_a * qc[2];
        const double new_lp = s_tot[1] - 0.5 * tBs * qp[0] - 0.5 * tau_g * qp[1] - 0.5 * tau_a * qp[2];
        if (a.wore && a.trace) { a.trace[2LL * iter] = cur_lp; a.trace[2LL * iter + 1] = new_lp; }
        const double lu = s_lu;
        ctl->last_accept = 0;
        double raw_Bs = rawc;
        int raw_row = 2;                                   // s_q row holding (x' Tau_Bs)_j x_j of the state kept
        if (lu < new_lp - cur_lp) {
            for (int j = 0; j < mm.P; ++j) cur[j] = prop[j];
            ctl->accept_s_block += 1; ctl->accept_s_total += 1;
            ctl->last_accept = 1;
            raw_Bs = rawp; raw_row = 3;
            if (a.wore) par[pl.cur_lp] = new_lp;
        }
        if (cc.adaptCov) for (int j = 0; j < mm.P; ++j) gpar[pl.block_par + j + ctl->i * mm.P] = cur[j];
        // ---- Gibbs draws of the precisions (:776-834) ----
        uint32_t draw = 0;
        if (cc.nRisks == 0) {
            const double Bpost = cc.B_tau_Bs + 0.5 * raw_Bs;           // B_tau + 0.5 Bs' Tau Bs (:778-779)
            par[pl.tau_Bs] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / Bpost))));
        } else {
            for (int k = 0; k < cc.nRisks; ++k) {                      // one draw per stratum (:777-782)
                double sq = 0.0;
                for (int j = cc.kn_first[k]; j <= cc.kn_last[k]; ++j) sq += s_q[raw_row][j];
                par[pl.tau_Bs + k] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_tau_Bs + 0.5 * sq)))));
            }
        }
        double* alp = cur + B + p;
        double* gam = cur + B;
        for (int kk = 0; kk < cc.n_td; ++kk) {
            const int L = cc.td_len[kk];
            double s = 0.0;
            for (int a1 = 0; a1 < L; ++a1) for (int a2 = 0; a2 < L; ++a2)
                s += alp[cc.td_cols[kk][a1]] * par[pl.Tau_a_pen + cc.td_cols[kk][a1] + cc.td_cols[kk][a2] * A] * alp[cc.td_cols[kk][a2]];
            const double tv = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_tau_td + 0.5 * s)))));
            par[pl.tau_td + kk] = tv;
            for (int a1 = 0; a1 < L; ++a1) for (int a2 = 0; a2 < L; ++a2)
                par[pl.Tau_a + cc.td_cols[kk][a1] + cc.td_cols[kk][a2] * A] = tv * par[pl.Tau_a_pen + cc.td_cols[kk][a1] + cc.td_cols[kk][a2] * A];
        }
        if (cc.shrink_gammas) {
            for (int k1 = 0; k1 < p; ++k1) {
                const double v = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_phi_g + 0.5 * par[pl.tau_g] * gam[k1] * gam[k1])))));
                par[pl.phi_g + k1] = v; par[pl.Tau_g + k1 + k1 * p] = v;
            }
            const double s = quad_form_dev(gam, nullptr, par + pl.Tau_g, p);
            par[pl.tau_g] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_tau_g + 0.5 * s)))));
        }
        if (cc.shrink_alphas) {
            for (int k2 = 0; k2 < A; ++k2) {
                const double Bp = (cc.double_gamma_alphas ? par[pl.nu_a + k2] : cc.B_phi_a) + 0.5 * par[pl.tau_a] * alp[k2] * alp[k2];
                const double v = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / Bp))));
                par[pl.phi_a + k2] = v; par[pl.Tau_a + k2 + k2 * A] = v;
            }
            const double s = quad_form_dev(alp, nullptr, par + pl.Tau_a, A);
            const double Bp = (cc.double_gamma_alphas ? par[pl.xi_a] : cc.B_tau_a) + 0.5 * s;
            par[pl.tau_a] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / Bp))));
            if (cc.double_gamma_alphas) {
                for (int k2 = 0; k2 < A; ++k2)
                    par[pl.nu_a + k2] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_nu_a + par[pl.phi_a + k2])))));
                par[pl.xi_a] = fmin(cc.eps3, fmax(cc.eps2, (s_gam[draw++] * (1.0 / (cc.B_xi_a + par[pl.tau_a])))));
            }
        }
        // ---- kept iterations (:836-851; 0-based iter against the 1-based keep sequence) ----
        if (ctl->check_iterator < cc.n_keep && ctl->keep_iterator < cc.n_out &&
            iter == cc.n_burnin + 1 + ctl->check_iterator * cc.n_thin) {
            const int ko = ctl->keep_iterator;
            for (int j = 0; j < B; ++j) a.o_Bs[j + ko * B] = cur[j];
            for (int j = 0; j < p; ++j) { a.o_g[j + ko * p] = gam[j]; a.o_phi_g[j + ko * p] = par[pl.phi_g + j]; }
            for (int j = 0; j < A; ++j) { a.o_a[j + ko * A] = alp[j]; a.o_phi_a[j + ko * A] = par[pl.phi_a + j]; }
            for (int k = 0; k < nR; ++k) a.o_tau_Bs[k + ko * nR] = par[pl.tau_Bs + k];
            a.o_tau_g[ko] = par[pl.tau_g]; a.o_tau_a[ko] = par[pl.tau_a];
            for (int j = 0; j < cc.n_td; ++j) a.o_tau_td[j + ko * cc.n_td] = par[pl.tau_td + j];
            a.o_lw[ko] = a.wore ? par[pl.cur_lp] : -s_tot[2];        // log_weightsREF :359-369,847 / out_logPost :1107
            ctl->keep_iterator += 1; ctl->check_iterator += 1;
        }
        // ---- block end: adapt the global scales (:853-867) ----
        if (ctl->i == cc.n_block - 1) {
            const double g1 = pow((double)(ctl->it + 1), cc.mc1), g2 = cc.c0 * g1;
            const double rate_s = (double)ctl->accept_s_block / cc.n_block;
            for (int k = 0; k < 3; ++k) {
                double sv = exp(log(par[pl.sig + k]) + g2 * (rate_s - cc.target_acc));
                par[pl.sig + k] = fmin(fmax(sv, cc.eps1), cc.eps3);
            }
            if (cc.adaptCov && ctl->it > cc.start_cov_update) {
                // H_* double as scratch here; they are rebuilt right below
                adapt_cov_dev(par + pl.S_Bs, par + pl.H_Bs, B, gpar + pl.block_par, mm.P, cc.n_block, g1, cc.eps2, cc.eps3);
                adapt_cov_dev(par + pl.S_g, par + pl.H_g, p, gpar + pl.block_par + B, mm.P, cc.n_block, g1, cc.eps2, cc.eps3);
                adapt_cov_dev(par + pl.S_a, par + pl.H_a, A, gpar + pl.block_par + B + p, mm.P, cc.n_block, g1, cc.eps2, cc.eps3);
                int e = chol_upper_dev(par + pl.S_Bs, B, par + pl.H_Bs);
                if (p) e |= chol_upper_dev(par + pl.S_g, p, par + pl.H_g);
                e |= chol_upper_dev(par + pl.S_a, A, par + pl.H_a);
                if (e) ctl->error = 2;
                s_mats_changed = 1;
            }
            ctl->accept_s_block = 0; ctl->i = 0; ctl->it += 1;
        } else {
            ctl->i += 1;
        }
        ctl->iter = iter + 1;
        s_next_iter = iter + 1;
    }
    __syncthreads();
    make_proposal(mm, pl, cc, par, s_next_iter, tid, blockDim.x, s_z, /*have_z=*/true);   // normals of iter + 1 drawn above
    if (a.use_smem) {
        // back to global memory: the part an iteration changes (parameters, precisions, scales - everything before the proposal
        // covariances) and, after a covariance adaptation only, the matrices behind it
        __syncthreads();
        const int hi = s_mats_changed ? pl.block_par : pl.S_Bs;
        for (int j = tid; j < hi; j += blockDim.x) gpar[j] = s_par[j];
        if (tid == 0 && pl.cur_lp >= hi) gpar[pl.cur_lp] = s_par[pl.cur_lp];
    }
}

// -------------------------------------------------------------------------------------------
// Once per block of iterations (src/fast_LAPRWM.cpp:651-661): the random draws of the b-step.
// The reference materialises rand_b = mvrnorm_cube(n_block, Sigma_b, sigma_b) per block (:658) and
// all accept uniforms up front (:618); here one thread per (iteration-in-chunk, subject) writes
// the proposal increment sqrt(sigma_b[i]) z'R_i (q doubles) and log u into the rng buffer that the
// step kernels stream.  sigma_b only changes at block ends, so a chunk never straddles a block.
// -------------------------------------------------------------------------------------------
struct PrepArgs {
    const double* Rchol;     // [n][q(q+1)/2] upper Cholesky factors of Covs$b, packed by column
    const double* state; double* rng;
    int n, q, state_stride, rng_stride, iter0, count;
    long long subject_offset;
};

template <int Q>   // Q > 0: q known at compile time (draws stay in registers); Q == 0: runtime q
__global__ void __launch_bounds__(256)
jmb_block_prep_kernel(const __grid_constant__ ChainConst cc, const __grid_constant__ PrepArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)a.n * a.count) return;
    const int q = Q > 0 ? Q : a.q;
    const int k = (int)(t / a.n), s = (int)(t % a.n);
    const uint32_t gid = (uint32_t)(a.subject_offset + s), iter = (uint32_t)(a.iter0 + k);
    const double* R = a.Rchol + (long long)s * (q * (q + 1) / 2);
    const double sd = sqrt(a.state[(long long)s * a.state_stride + q]);
    double* out = a.rng + ((long long)k * a.n + s) * a.rng_stride;
    double z[Q > 0 ? Q + 1 : JMB_MAX_Q];
#pragma unroll
    for (int sl = 0; sl < (Q > 0 ? (Q + 1) / 2 : JMB_MAX_Q / 2); ++sl) {
        if (sl < (q + 1) / 2) rnorm_pair(cc.seed, cc.chain_id, gid, iter, (uint32_t)sl, ST_B_PROP, z[2 * sl], z[2 * sl + 1]);
    }
#pragma unroll
    for (int j = 0; j < (Q > 0 ? Q : JMB_MAX_Q); ++j) {
        if (j < q) {
            const double* Rj = R + (j * (j + 1)) / 2;      // column j of the upper factor
            double pj = 0.0;
#pragma unroll
            for (int l = 0; l < (Q > 0 ? Q : JMB_MAX_Q); ++l) if (l <= j) pj += z[l] * Rj[l];
            out[j] = sd * pj;
        }
    }
    out[q] = log_unif(cc.seed, cc.chain_id, gid, iter, 0u, ST_B_ACC);
}

// per-subject Robbins-Monro update of sigma_b at the end of block `it` (:853-859, bounds_sigma :411-419)
__global__ void __launch_bounds__(256)
jmb_adapt_sigma_kernel(const __grid_constant__ ChainConst cc, double* state, int n, int q, int stride, const Ctl* ctl) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    // launched after the finalize kernel of the block's last iteration: ctl->it already counts it
    const double g2 = cc.c0 * pow((double)ctl->it, cc.mc1);
    double* st = state + (long long)s * stride;
    const double rate = st[q + 3] / cc.n_block;
    double sg = exp(log(st[q]) + g2 * (rate - cc.target_acc));
    st[q] = fmin(fmax(sg, cc.eps1), cc.eps3);
    st[q + 3] = 0.0;
}

// -------------------------------------------------------------------------------------------
// Survival-block log-posterior data terms and score sums for FIXED Wlong / Wlongs:
// logPosterior (src/fast_logPost.cpp:8-26) and gradient_logPosterior (src/fast_score.cpp:8-50).
//   out[0] = sum_i event_i * log_h_i        out[1] = sum_k Pw_k exp(lin_k)   (all nK rows, no
//   per-subject rowsum: the reference does not segment-sum H here either)
//   out[2 + j] = sum_k W[k, j] * Pw_k exp(lin_k),  W = [W1s | W2s | Wlongs]
// One lane per hazard row like the step kernel; lower-interval rows are not part of these
// functions.  Per-warp accumulators are updated in lane order and combined in warp / CTA order, so
// the result is reproducible.
// -------------------------------------------------------------------------------------------
struct SurvPostArgs {
    const double* drec; const long long* doff;
    const double* Wlong;     // n x A column-major
    const double* Wlongs;    // nK x A column-major
    const double* theta;     // [P] Bs_gammas | gammas | alphas
    double* partials;        // [gridDim.x][P + 2]
    int n, subjects_per_cta;
    int value_only;          // skip the score sums (woRE chain: only the log-posterior is needed)
    // multi-state fits: Wlong has nT rows, Wlongs nT * K; trow[s] = first transition row of subject s
    const int* trow; int nT; int ms;
};

template <int G>
__global__ void __launch_bounds__(256)
jmb_surv_post_kernel(const __grid_constant__ ModelMeta mm, const __grid_constant__ SurvPostArgs a) {
    constexpr int THREADS = 256, GROUPS = THREADS / G;
    const ModelDesc& md = mm.d;
    const RecordLayout& L = mm.L;
    const int P = mm.P, B = mm.B, p = md.p, A = md.A, K = md.K, RP = L.RP, NA = P + 2;
    extern __shared__ double smem[];
    double* s_theta = smem;              // [P]
    double* s_acc = s_theta + P;         // [GROUPS][NA]
    const int tid = threadIdx.x, grp = tid / G, lane = tid % G;
    const unsigned gmask = (G == 32) ? 0xFFFFFFFFu : (0xFFFFu << (16 * ((tid & 31) >> 4)));
    for (int j = tid; j < P; j += THREADS) s_theta[j] = a.theta[j];
    for (int j = tid; j < GROUPS * NA; j += THREADS) s_acc[j] = 0.0;
    __syncthreads();
    double* acc = s_acc + grp * NA;
    const long long nT = a.ms ? a.nT : a.n;           // rows of Wlong
    const long long ns = nT * K;
    const int s_begin = blockIdx.x * a.subjects_per_cta, s_end = min(a.n, s_begin + a.subjects_per_cta);
    for (int s = s_begin + grp; s < s_end; s += GROUPS) {
      const double* rec0 = a.drec + a.doff[s];
      const int ntr = a.ms ? (int)rec0[0] : 1;       // one hazard-row block per transition row (jmb_model.h)
      for (int tb = 0; tb < ntr; ++tb) {
        const double* rec = rec0 + tb * L.HB;
        const long long tr = a.ms ? (long long)a.trow[s] + tb : s;
        const int r = lane;
        const int cls = (r == 0) ? 0 : (r <= K ? 1 : 3);
        const int off = reinterpret_cast<const int*>(rec + L.o_W1off)[r];
        double lin = 0.0, w1[8], wl[JMB_MAX_ALPHAS];
        if (cls != 3) {
            for (int j = 0; j < md.WB && md.WB <= 8; ++j) { w1[j] = rec[L.o_W1v + j * RP + r]; lin += w1[j] * s_theta[off + j]; }
            if (md.WB > 8) for (int j = 0; j < md.WB; ++j) lin += rec[L.o_W1v + j * RP + r] * s_theta[off + j];
            double l2 = 0.0;
            for (int j = 0; j < p; ++j) l2 += (md.w2_const ? rec0[L.o_W2c + j] : rec[L.o_W2 + j * RP + r]) * s_theta[B + j];
            double l3 = 0.0;
            for (int c = 0; c < A; ++c) {
                wl[c] = (cls == 0) ? a.Wlong[tr + (long long)c * nT] : a.Wlongs[tr * K + (r - 1) + (long long)c * ns];
                l3 += wl[c] * s_theta[B + p + c];
            }
            lin = (lin + l2) + l3;
        }
        const double event = a.ms ? rec[L.o_Pw] : rec0[0];
        const double pw = (cls == 1) ? rec[L.o_Pw + r] : 0.0;
        const double Int = (cls == 1) ? pw * exp(lin) : 0.0;
        // value terms
        const double Hsum = group_sum<G>(Int, gmask);
        const double lh = __shfl_sync(gmask, lin, 0, G);
        if (lane == 0) { acc[0] += event * lh; acc[1] += Hsum; }
        if (a.value_only) continue;
        // score sums: baseline-hazard band, lane by lane in fixed order
        for (int l = 1; l <= K; ++l) {
            const int offl = __shfl_sync(gmask, off, l, G);
            if (md.WB <= 8) {
                for (int j = 0; j < md.WB; ++j) {
                    const double v = __shfl_sync(gmask, w1[j] * Int, l, G);
                    if (lane == 0) acc[2 + offl + j] += v;
                }
            } else {
                for (int j = 0; j < md.WB; ++j) {
                    const double v = __shfl_sync(gmask, (lane == l) ? rec[L.o_W1v + j * RP + r] * Int : 0.0, l, G);
                    if (lane == 0) acc[2 + offl + j] += v;
                }
            }
        }
        for (int j = 0; j < p; ++j) {
            const double wv = md.w2_const ? rec0[L.o_W2c + j] : rec[L.o_W2 + j * RP + r];
            const double v = group_sum<G>(wv * Int, gmask);
            if (lane == 0) acc[2 + B + j] += v;
        }
        for (int c = 0; c < A; ++c) {
            const double v = group_sum<G>((cls == 1) ? wl[c] * Int : 0.0, gmask);
            if (lane == 0) acc[2 + B + p + c] += v;
        }
      }
    }
    __syncthreads();
    for (int j = tid; j < NA; j += THREADS) {
        double t = 0.0;
        for (int g = 0; g < GROUPS; ++g) t += s_acc[g * NA + j];
        a.partials[(long long)blockIdx.x * NA + j] = t;
    }
}

// fixed-order column sums of a [rows][cols] partials matrix
__global__ void __launch_bounds__(128)
jmb_colsum_kernel(const double* partials, int rows, int cols, double* out) {
    const int j = blockIdx.x;
    __shared__ double s[128];
    double t = 0.0;
    for (int i = threadIdx.x; i < rows; i += 128) t += partials[(long long)i * cols + j];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) { if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) out[j] = s[0];
}

// sum the per-CTA partials into par[pl.sums] (multi-GPU path, followed by an NCCL all-reduce)
__global__ void __launch_bounds__(128)
jmb_reduce_partials_kernel(const double* partials, int n_partials, double* sums) {
    __shared__ double s_sum[128 * 3];
    const int tid = threadIdx.x;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int i = tid; i < n_partials; i += 128) { t0 += partials[i * 3]; t1 += partials[i * 3 + 1]; t2 += partials[i * 3 + 2]; }
    s_sum[tid] = t0; s_sum[128 + tid] = t1; s_sum[256 + tid] = t2;
    __syncthreads();
    for (int w = 64; w > 0; w >>= 1) {
        if (tid < w) { s_sum[tid] += s_sum[tid + w]; s_sum[128 + tid] += s_sum[128 + tid + w]; s_sum[256 + tid] += s_sum[256 + tid + w]; }
        __syncthreads();
    }
    if (tid < 3) sums[tid] = s_sum[tid * 128];
}

// out_b slice (n x q column-major) <- state (kept iterations, :837)
__global__ void jmb_extract_b_kernel(const double* state, int stride, int n, int q, int col0, double* out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * q) return;
    const int i = (int)(idx % n), j = (int)(idx / n);
    out[idx] = state[(long long)i * stride + col0 + j];
}

// state <- initial b (n x q column-major) and scales$b
// Xbetas_calc on the device (R/supportFuns_mvJointModelBayes.R:43-55): fills a chain's per-draw records from the
// fixed-effects designs of the fit and the betas of the draw.  The x-record of a subject mirrors its per-draw
// record: per term [f_t][RP] design columns, then per outcome [p_k][v] design columns.  One warp per subject.
struct XbDesc {
    int T, O, RP;
    int ft[kMaxT], tbase[kMaxT], tboff[kMaxT];          // columns, first column in the x-record, offset of betas[[outcome]]
    int indf[kMaxT][JMB_MAX_RT];
    int pk[kMaxO], boff[kMaxO];
};
// trow != nullptr (multi-state fits): the term part of both records is repeated once per transition row of the subject.
__global__ void jmb_xbetas_kernel(const __grid_constant__ XbDesc xd, const double* __restrict__ xrec,
                                  const long long* __restrict__ xoff, const long long* __restrict__ coff,
                                  const int* __restrict__ rowptr, const double* __restrict__ betas, const int n,
                                  double* __restrict__ crec, const int* __restrict__ trow) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const int i = (int)w;
    const double* __restrict__ xr = xrec + xoff[i];
    double* __restrict__ cr = crec + coff[i];
    const int RP = xd.RP, nt = xd.T * RP;
    const int ntr = trow ? trow[i + 1] - trow[i] : 1;
    int xcols = 0;
    for (int t = 0; t < xd.T; ++t) xcols += xd.ft[t];
    for (int e = lane; e < nt * ntr; e += 32) {
        const int tb = e / nt, et = e - tb * nt;
        const int t = et / RP, r = et - t * RP;
        const double* __restrict__ xb = xr + (long long)tb * xcols * RP;
        double acc = 0.0;
        for (int f = 0; f < xd.ft[t]; ++f) acc += xb[(xd.tbase[t] + f) * RP + r] * betas[xd.tboff[t] + xd.indf[t][f]];
        cr[e] = acc;
    }
    int xo = xcols * RP * ntr, co = nt * ntr;
    for (int k = 0; k < xd.O; ++k) {
        const int v = rowptr[(long long)k * (n + 1) + i + 1] - rowptr[(long long)k * (n + 1) + i];
        for (int r = lane; r < v; r += 32) {
            double acc = 0.0;
            for (int c = 0; c < xd.pk[k]; ++c) acc += xr[xo + c * v + r] * betas[xd.boff[k] + c];
            cr[co + r] = acc;
        }
        xo += xd.pk[k] * v;
        co += v;
    }
}

__global__ void jmb_load_state_kernel(double* state, int stride, int n, int q, const double* b,
                                      const double* sigma_b) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)n * stride) return;
    const int i = (int)(idx / stride), j = (int)(idx % stride);
    double v = 0.0;
    if (j < q) v = b[(long long)j * n + i];
    else if (j == q) v = sigma_b[i];
    state[idx] = v;
}

}  // namespace jmb
