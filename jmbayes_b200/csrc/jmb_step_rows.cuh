// jmb_step_rows.cuh - row-strided step kernel (sm_100a), round 2.
//
// Same per-subject work as jmb_step_static / jmb_step_tma (one b-step + the subject's share of the survival
// step, src/fast_LAPRWM.cpp:666-757) on the same records, but a subject is served by a group of G = 2..16
// lanes (32 / G subjects per warp step) instead of one lane per hazard row:
//   * lane l of a group walks the hazard rows l, l + G, l + 2G, ... (row 0 = event time, 1..K quadrature,
//     K+1..2K lower-interval quadrature) and the visit rows l, l + G, ... of every outcome, keeps its own
//     partial sums, and the group combines them with log2(G) xor-shuffles;
//   * everything that is per subject and not per row (state, proposal, log p(b), accept, write-back) is executed once
//     per warp for 32 / G subjects instead of once per subject; the censoring terms, which are per subject AND
//     expensive, are spread over the lanes of the group (JMB_ROWS_PTBL).
// With one lane per row (round 1) 1 100 warp instructions were issued per subject-iteration of config 4, most of
// them replicated per-subject work and visit loops with 10 of 32 lanes active; here 550 (profiles/r2_rows_kernel.md,
// profiles/r2b_rows_kernel.md).
//
// Work distribution: the host deals the warp steps to the (CTA, warp) slots by estimated cost (RowsArgs::order); the
// kernel walks its slot's list, so the per-slot partial sums keep a fixed order.
//
// Memory pipeline (a warp step covers 32 / G whole records, 18 - 37 KB: too much to double-buffer per warp) - everything
// from HBM arrives in shared memory ahead of its use and nothing of it passes through the L1:
//   * hazard rows (80 % of the bytes): a ring of JMB_ROWS_RING stages per warp filled with cp.async (LDGSTS)
//     RING - 1 row-loop trips ahead of their use, 16 bytes per copy by lane pairs (JMB_ROWS_PAIR);
//   * "early" data of a step - record header, state, draws of the iteration, visit rows of every outcome and
//     their per-draw linear predictors, about 1 kB per subject - for the warp's NEXT step (template parameter EARLY):
//     1: cp.async.bulk + mbarrier into a per-group stage; 3: the same stage filled by the lanes' 16-byte cp.async;
//     2: group slots + lane-private visit slots; 0: read in place behind an L2 prefetch.
// An L2 prefetch of whole records one step ahead (cp.async.bulk.prefetch.L2) was measured first and evicts itself
// (16 warps x 148 SMs x 37 KB = 87 MB in flight; DRAM reads 1.9x the compulsory bytes).
//
// Code size: the loop body of a step has to fit the 32 KB instruction cache of an SM.  Cold paths are out-of-line
// functions with SCALAR arguments (a pointer or reference to a local moves it to the stack for the whole kernel); the
// censoring code and the visit loops exist once.
//
// The sums over quadrature rows need the proposed-parameter hazard at the accepted b, which is only known after
// all rows have been visited: the row loop therefore accumulates it for both the old and the proposed b (three
// exponentials per row instead of two) and selects after the accept decision.
#pragma once
#include "jmb_step_static.cuh"

namespace jmb {

// per-subject addressing of the pipelined variant (EARLY = 2): everything the copies of a subject's early data and visit
// rows need, in one 32-byte read (the record head, where the visit counts live, would be a dependent load)
struct alignas(32) RowsMeta {
    long long d0, c0;        // offsets of the design / per-draw record (doubles)
    int v[3];                // visit counts of the outcomes (rows kernels exist for O <= 3)
    int ntr;                 // transition rows (multi-state fits; 1 otherwise)
};
static_assert(sizeof(RowsMeta) == 32, "RowsMeta is read as two 16-byte words");

struct RowsArgs {
    // warp steps (of 32 / G subjects) of every (CTA, warp) slot: order[slot * order_len + k], -1 = none.  The host deals the
    // steps to the slots by estimated cost (longest processing time first; jmb200.cu) - with an equal NUMBER of steps per warp
    // the slowest warp of the config-4 shard carries 16 % more row-loop trips than the average one
    const int* order; int order_len;
    const RowsMeta* meta;    // [n] (EARLY = 2)
    int vt;                  // EARLY = 2: visit-row trips of an outcome held in the lane-private visit slots (later trips are read in place)
    int prefetch;            // EARLY = 0 only: 1 = L2 bulk prefetch of the next step's early data
    int cap_dl, cap_cl;      // EARLY = 1: capacity (doubles) of the visit rows of the design / per-draw record of one subject
    // several ranks (peer_mbox != NULL): the last CTA to finish reduces the per-CTA partial sums and stores the rank's three sums
    // into every rank's mailbox itself, so that the NVLink flight overlaps the kernel boundary and the finalize kernel finds
    // the stamps waiting (fields as in FinalizeArgs)
    double* const* peer_mbox; int nranks, rank, slot; unsigned long long stamp; unsigned int* ticket;
};

#ifndef JMB_ROWS_THREADS
#define JMB_ROWS_THREADS 256
#endif
#ifndef JMB_ROWS_CTAS
#define JMB_ROWS_CTAS 2
#endif
#ifndef JMB_ROWS_RING
#define JMB_ROWS_RING 2
#endif
// pipelined variant: two group slots per group, so that the per-subject values that are only needed after the row loop (log u,
// the carried log-densities, counters, event) are re-read from the slot instead of being held in registers across the loop
#ifndef JMB_ROWS_REREAD
#define JMB_ROWS_REREAD 1
#endif
// ring copies by lane pairs: the lanes 2m / 2m + 1 of a warp walk neighbouring rows, so one 16-byte cp.async.cg of the even
// (odd) lane fetches the pair's two values of an even (odd) slot.  Half the copies, and none of them passes through the L1:
// the 8-byte copies (.ca is the only cache operator below 16 bytes) hold a 128-byte L1 line each while in flight, and 16 warps
// keep about 120 KB of lines in flight - more than the L1 that is left once the shared memory of two CTAs exceeds 132 KB
#ifndef JMB_ROWS_PAIR
#define JMB_ROWS_PAIR 1
#endif
constexpr int kRowsThreads = JMB_ROWS_THREADS, kRowsCtas = JMB_ROWS_CTAS, kRowsWarps = JMB_ROWS_THREADS / 32;
constexpr int kRowsHdr = 8;  // doubles of a record head that hold event, visit counts and the constant W2 row

__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
#if !defined(JMB_EMU)
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void cp_async8(double* dst, const double* src) {
#if defined(JMB_EMU)
    *dst = *src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async16(double* dst, const double* src) {
#if defined(JMB_EMU)
    dst[0] = src[0]; dst[1] = src[1];
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async4(int* dst, const int* src) {
#if defined(JMB_EMU)
    *dst = *src;
#else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if !defined(JMB_EMU)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#if !defined(JMB_EMU)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}
// generic-proxy reads of a shared-memory stage must be ordered before the async-proxy (bulk copy) writes that re-fill it
__device__ __forceinline__ void fence_proxy_async() {
#if !defined(JMB_EMU)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}

// slots of one hazard row in the ring: band offset | WB band values | Pw | W2 columns | per term: XXbeta, stored ZZ, stored U
template <class Spec>
struct RowSlots {
    static constexpr int nzz(int t) { return Spec::d.rt[t] - Spec::d.zz_ones0[t]; }
    static constexpr int nu(int t) { return Spec::d.u_ones[t] ? 0 : Spec::d.ut[t]; }
    static constexpr int term_base(int t) {
        int b = 2 + Spec::d.WB + (Spec::d.w2_const ? 0 : Spec::d.p);
        for (int k = 0; k < t; ++k) b += 1 + nzz(k) + nu(k);
        return b;
    }
    static constexpr int NV = term_base(Spec::d.T);
};

// branch-free censoring term for a warp whose groups hold subjects of different event types
// (src/fast_LAPRWM.cpp:640-650; same arithmetic per event type as log_ptb_static)
template <int S>
__device__ __forceinline__ double log_ptb_rows(double event, double lh, double H, double HL) {
    if constexpr (S == 2) {
        const double eH = exp(-H), eHL = exp(-HL);
        const double arg = (event == 2.0) ? (1.0 - eH) : ((event == 3.0) ? (eHL - eH) : 1.0);
        const double lg = log(arg);
        double r = 0.0;
        if (event == 1.0) r = lh - H;
        if (event == 0.0) r = 0.0 - H;
        if (event == 2.0 || event == 3.0) r = lg;
        return r;
    } else {
        return event * lh - H;
    }
}

// censoring terms evaluated lane-parallel inside the group (G >= 8).  (Evaluating the three terms side by side on every lane,
// and two visit rows per loop trip, were measured too: both lose more to code size than the extra independent chains gain,
// profiles/r2b_rows_kernel.md.)
#ifndef JMB_ROWS_PTBL
#define JMB_ROWS_PTBL 1
#endif

// GLMM log-density of this lane's visit rows (lin_predF + log_longF, src/fast_LAPRWM.cpp:67-78,168-202); lrec / lcrec point
// at the visit rows of the design / per-draw record, in global memory or in the warp's early stage (the call sites keep
// the two apart so that the loads compile to LDS / LDG rather than generic loads)
template <class Spec, int G>
__device__ __forceinline__ double visit_rows(const double* __restrict__ lrec, const double* __restrict__ lcrec, const int* vcount,
                                             const double (&bn)[Spec::d.q], const double (&isig)[Spec::d.O], const int l) {
    constexpr int O = Spec::d.O;
    double v_pyb = 0.0;
    sfor<O>([&](auto kc) {
        constexpr int k = kc;
        constexpr int ones0 = Spec::d.z_ones0[k], qk = Spec::d.qk[k];
        const int v = vcount[k];
        auto eta_of = [&](const int r) {
            double eta = lcrec[r];
            if constexpr (ones0 == 1) { constexpr int cz = Spec::d.re_inds[k][0]; eta += bn[cz]; }
            sfor<qk - ones0>([&](auto jc) {
                constexpr int j = jc;
                constexpr int cj = Spec::d.re_inds[k][j + ones0];
                eta += lrec[(1 + j) * v + r] * bn[cj];
            });
            return eta;
        };
        for (int r = l; r < v; r += G) v_pyb += logdens_static<Spec::d.fam[k], Spec::d.link[k]>(lrec[r], eta_of(r), isig[k]);
        constexpr int lcols = Spec::L.long_cols[k];
        lrec += lcols * v;
        lcrec += v;
    });
    return v_pyb;
}

// Pipelined variant (EARLY = 2): lane-private visit slots.  Trip t (rows l + t G) of outcome k occupies per(k) values - y, the
// per-draw linear predictor, the stored Z columns - at [(t VS + base(k)) * 32 + lane]; they are cp.async-copied by the lane that
// reads them, one row loop ahead of their use, so the visit loops read shared memory without bank conflicts and never wait for
// the L2.  Trips beyond ra.vt (segments longer than vt G rows) are read in place.
template <class Spec>
struct VisitSlots {
    static constexpr int per(int k) { return 2 + Spec::d.qk[k] - Spec::d.z_ones0[k]; }
    static constexpr int base(int k) { int b = 0; for (int i = 0; i < k; ++i) b += per(i); return b; }
    static constexpr int VS = base(Spec::d.O);
};

template <class Spec, int G>
__device__ __forceinline__ void issue_visit_rows(double* vs, const int vt, const double* __restrict__ lrec, const double* __restrict__ lcrec,
                                                 const int (&vcount)[Spec::d.O], const int l) {
    constexpr int O = Spec::d.O, VS = VisitSlots<Spec>::VS;
    sfor<O>([&](auto kc) {
        constexpr int k = kc;
        constexpr int ones0 = Spec::d.z_ones0[k], qk = Spec::d.qk[k], sb = VisitSlots<Spec>::base(k);
        const int v = vcount[k];
        int r = l;
        for (int t = 0; t < vt && r < v; ++t, r += G) {
            double* const dst = vs + (size_t)(t * VS + sb) * 32;
            cp_async8(dst, lrec + r);
            cp_async8(dst + 32, lcrec + r);
            sfor<qk - ones0>([&](auto jc) { constexpr int j = jc; cp_async8(dst + (2 + j) * 32, lrec + (1 + j) * v + r); });
        }
        constexpr int lcols = Spec::L.long_cols[k];
        lrec += lcols * v;
        lcrec += v;
    });
}

// same arithmetic and summation order as visit_rows; the first vt trips of every outcome come from the visit slots, later trips
// (segments longer than vt G rows) from the record - through ONE copy of the loop body: the addresses of a row's values are
// selected per trip and read with generic loads
template <class Spec, int G>
__device__ __forceinline__ double visit_rows_slots(const double* vs, const int vt, const double* __restrict__ lrec, const double* __restrict__ lcrec,
                                                   const int* vcount, const double (&bn)[Spec::d.q], const double (&isig)[Spec::d.O], const int l) {
    constexpr int O = Spec::d.O, VS = VisitSlots<Spec>::VS;
    double v_pyb = 0.0;
    sfor<O>([&](auto kc) {
        constexpr int k = kc;
        constexpr int ones0 = Spec::d.z_ones0[k], qk = Spec::d.qk[k], sb = VisitSlots<Spec>::base(k);
        const int v = vcount[k];
        int t = 0;
        for (int r = l; r < v; r += G, ++t) {
            const bool slot = t < vt;
            const double* const sl = vs + (size_t)(t * VS + sb) * 32;
            const double* const py = slot ? sl : lrec + r;
            const double* const pe = slot ? sl + 32 : lcrec + r;
            double eta = *pe;
            if constexpr (ones0 == 1) { constexpr int cz = Spec::d.re_inds[k][0]; eta += bn[cz]; }
            sfor<qk - ones0>([&](auto jc) {
                constexpr int j = jc;
                constexpr int cj = Spec::d.re_inds[k][j + ones0];
                const double* const pz = slot ? sl + (2 + j) * 32 : lrec + (1 + j) * v + r;
                eta += *pz * bn[cj];
            });
            v_pyb += logdens_static<Spec::d.fam[k], Spec::d.link[k]>(*py, eta, isig[k]);
        }
        constexpr int lcols = Spec::L.long_cols[k];
        lrec += lcols * v;
        lcrec += v;
    });
    return v_pyb;
}

// last CTA of a multi-rank launch: fixed-order sum of the per-CTA partials, then one store per rank (values, fence, stamp)
__device__ __noinline__ void rows_publish_sums(const double* partials, const int n_partials, const RowsArgs ra, double* s_pub /* [3 * JMB_ROWS_THREADS] */) {
    const int tid = threadIdx.x;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
    for (int i = tid; i < n_partials; i += JMB_ROWS_THREADS) { t0 += __ldcg(partials + i * 3); t1 += __ldcg(partials + i * 3 + 1); t2 += __ldcg(partials + i * 3 + 2); }
    s_pub[tid] = t0; s_pub[JMB_ROWS_THREADS + tid] = t1; s_pub[2 * JMB_ROWS_THREADS + tid] = t2;
    __syncthreads();
    if (tid < 3) {                                      // any block size: thread-order sum (this path runs once per launch, in one CTA)
        double t = 0.0;
        for (int i = 0; i < JMB_ROWS_THREADS; ++i) t += s_pub[tid * JMB_ROWS_THREADS + i];
        s_pub[tid * JMB_ROWS_THREADS] = t;
    }
    __syncthreads();
    if (tid < ra.nranks) {
        const long long base = ((long long)(ra.slot * 2 + (int)(ra.stamp & 1ull)) * ra.nranks);
        mbox_send(ra.peer_mbox[tid] + (base + ra.rank) * kMboxEntry, s_pub[0], s_pub[JMB_ROWS_THREADS], s_pub[2 * JMB_ROWS_THREADS], ra.stamp);
    }
    if (tid == 0) *ra.ticket = 0u;                    // ready for the next launch on this chain slot
}

template <class Spec, int G, int EARLY>
__global__ void __launch_bounds__(JMB_ROWS_THREADS, JMB_ROWS_CTAS)
jmb_step_rows(const int B, const __grid_constant__ ParamLayout pl, const __grid_constant__ ChainConst cc,
              const __grid_constant__ StepArgs a, const RowsArgs ra) {
    constexpr int q = Spec::d.q, p = Spec::d.p, A = Spec::d.A, O = Spec::d.O, T = Spec::d.T, K = Spec::d.K, S = Spec::d.S;
    constexpr int WB = Spec::d.WB, RP = Spec::L.RP, SS = Spec::L.state_stride;
    constexpr int oV = Spec::L.o_V, oLong = Spec::L.o_long, cLong = Spec::L.c_long, oW1off = Spec::L.o_W1off;
    constexpr int oW1v = Spec::L.o_W1v, oW2c = Spec::L.o_W2c, oW2 = Spec::L.o_W2, oPw = Spec::L.o_Pw;
    constexpr bool W2C = Spec::d.w2_const != 0;
    // multi-state fits (src/fast_LAPRWM.cpp:519-529,669,703): ntr >= 1 hazard-row blocks per subject (one per transition row,
    // HB doubles apart in the design record, cLong apart in the per-draw record); log p(T|b) sums the blocks' terms
    constexpr bool MS = Spec::d.MS != 0;
    constexpr int HB = Spec::L.HB;
    static_assert(!MS || (S == 1 && !W2C), "multi-state fits: no interval censoring, transition-specific W2");
    constexpr int THREADS = kRowsThreads, WARPS = THREADS / 32, SPW = 32 / G;
    constexpr int NR = 1 + K * S;                       // hazard rows of a subject
    constexpr int NJ = (NR + G - 1) / G;                // row-loop trips
    constexpr int NJ_H = (1 + K + G - 1) / G;           // trips that hold the event-time row and the K rows of H
    constexpr int RS = (q + 2) / 2 * 2;                 // rng stride: q increments + log u, padded to even
    constexpr int NV = RowSlots<Spec>::NV, RING = JMB_ROWS_RING, LEAD = RING > 0 ? RING - 1 : 0;
    constexpr bool PAIR = JMB_ROWS_PAIR != 0 && RING > 0;
    static_assert(!PAIR || (RP % 2 == 0 && G % 2 == 0 && oW1off % 2 == 0 && oW1v % 2 == 0 && oPw % 2 == 0 && (W2C || oW2 % 2 == 0) && HB % 2 == 0),
                  "paired ring copies move 16-byte-aligned pairs of rows");
    // early data of a step (head | state | draws, visit rows): 0 read in place (+ L2 prefetch); 1 bulk-copied per-group stage +
    // mbarrier; 2 cp.async group slots + lane-private visit slots; 3 cp.async (16-byte, by all lanes of the group) into the per-group
    // stage of variant 1 - a bulk copy takes its operands from uniform registers, so the four groups of a warp issue theirs one after
    // the other (9 instructions per copy and lane: 211 of the 2 490 instructions of a warp step), the lanes' own copies do not
    constexpr bool E1 = EARLY == 1, E3 = EARLY == 3, E2 = EARLY == 2 || E3;
    constexpr int EH = kRowsHdr + SS + RS;              // E2: head | state | draws of a subject, one slot per group
    constexpr int VS = VisitSlots<Spec>::VS;
    static_assert(EARLY >= 0 && EARLY <= 3, "EARLY");
    static_assert(!E2 || (RING == 2 && O <= 3), "the pipelined variant orders its cp.async groups around a two-stage ring");
    static_assert(32 % G == 0 && G >= 2 && G <= 16, "G lanes per subject");
    static_assert(LEAD <= NJ_H, "the trips copied ahead of the record head must be trips every subject needs");
    static_assert(SS % 2 == 0 && RS % 2 == 0 && oLong % 2 == 0 && cLong % 2 == 0, "16-byte pieces");
    static_assert(oPw <= kRowsHdr && oW2c + (W2C ? p : 0) <= kRowsHdr && oV * 2 + O <= 2 * kRowsHdr, "record head fits the staged header");
    const int P = B + p + A;
    extern __shared__ double smem[];
    // [P] (current, proposed) pairs of Bs_gammas | gammas | alphas: one 16-byte load serves both parameter sets
    double2* s_cp = reinterpret_cast<double2*>(smem);
    double* s_invD = smem + 2 * P;                       // [q*q]
    double* s_red = s_invD + (q * q + 1) / 2 * 2;        // [WARPS * SPW][3]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_red + (WARPS * SPW * 3 + 1) / 2 * 2);      // [WARPS] early-stage barriers
    double* const s_ring0 = reinterpret_cast<double*>(s_bar + (WARPS + 1) / 2 * 2);                        // [WARPS][RING][NV][32]
    // E1: early stage of a subject: head | state | draws | visit rows (design) | visit rows (per-draw)
    // E2: [WARPS][SPW][EH] group slots, then [WARPS][vt][VS][32] lane-private visit slots
    constexpr int ESLOTS = (E2 && JMB_ROWS_REREAD) ? 2 : 1;
    const int ES = E3 ? EH * ESLOTS + ra.cap_dl + ra.cap_cl : (E2 ? EH * ESLOTS : (kRowsHdr + SS + RS + ra.cap_dl + ra.cap_cl));
    double* const s_early0 = s_ring0 + (size_t)WARPS * RING * NV * 32;                       // [WARPS][SPW][ES]

    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    const int grp = lane32 / G, l = lane32 % G;
    double* const s_ring = s_ring0 + (size_t)warp * RING * NV * 32 + lane32;
    double* const my_early0 = s_early0 + ((size_t)warp * SPW + grp) * ES;
    double* my_early = my_early0;                        // E2 with two slots: slot of the current step
    double* const s_vis = E3 ? my_early0 + EH * ESLOTS       // E3: visit rows of the group's subject, design part then per-draw part
                             : s_early0 + (size_t)WARPS * SPW * ES + (size_t)warp * (E2 ? ra.vt : 0) * VS * 32 + lane32;
    uint64_t* const bar = s_bar + warp;
    // this warp's steps: the list the host dealt to the (CTA, warp) slot
    const long long n = a.n;
    const int* const my_order = ra.order + (size_t)(blockIdx.x * WARPS + warp) * ra.order_len;
    auto subject_of = [&](int step) { return min(n - 1, (long long)step * SPW + grp); };
    auto load_meta = [&](const long long sj, long long& e0, long long& f0, int (&mv)[O], int& mtr) {
        const RowsMeta m = ra.meta[sj];                  // 32-byte aligned: two 16-byte loads
        e0 = m.d0; f0 = m.c0;
        sfor<O>([&](auto kc) { mv[kc] = m.v[kc]; });
        mtr = m.ntr;
    };

    // offsets of this group's subject in the first step; loaded before the dependency wait (the offset tables never change)
    int ko = 0;                                           // position in the slot's list
    int t_step = ra.order_len > 0 ? my_order[0] : -1;
    int t_next = ra.order_len > 1 ? my_order[1] : -1;
    if (t_step < 0) t_next = -1;
    long long d0 = 0, d1 = 0, c0 = 0, c1 = 0;
    int ntr = 1;                                          // transition rows of this group's subject
    int mvc[O];                                           // E2: visit counts of the first subject (its copies start before the loop)
    sfor<O>([&](auto kc) { mvc[kc] = 0; });
    if (t_step >= 0) {
        const long long s = subject_of(t_step);
        if constexpr (E2) load_meta(s, d0, c0, mvc, ntr);
        else {
            d0 = a.doff[s]; d1 = a.doff[s + 1]; c0 = a.coff[s]; c1 = a.coff[s + 1];
            if constexpr (MS) ntr = a.trow[s + 1] - a.trow[s];
        }
    }
    if constexpr (E1) { if (lane32 == 0) mbar_init(bar, SPW); }
    pdl_wait();
    pdl_trigger();
    for (int j = tid; j < P; j += THREADS) s_cp[j] = make_double2(a.par[pl.cur + j], a.par[pl.prop + j]);
    for (int j = tid; j < q * q; j += THREADS) s_invD[j] = a.par[pl.invD + j];
#if !defined(JMB_EMU)
    if constexpr (E1) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
    const int i_blk = a.ctl->i;
    const bool last_acc = a.ctl->last_accept != 0;
    __syncthreads();
    double isig[O];
    sfor<O>([&](auto kc) { isig[kc] = 1.0 / a.par[pl.sigmas + kc]; });
    const double2* s_g = s_cp + B;
    const double2* s_al = s_cp + B + p;
    const double* rng_it = a.rng + (long long)(i_blk % a.rng_chunk) * n * RS;

    // bulk copies of one subject's early data into this group's stage slot (lane 0 of the group); `fits` = its visit rows
    // are within the stage capacity (else they are read in place)
    // visit rows of a record with nb hazard-row blocks start at oLong + (nb - 1) HB (design) / nb cLong (per-draw)
    auto fits_early = [&](long long e0, long long e1, long long f0, long long f1, int nb) {
        return (e1 - e0 - oLong - (long long)(nb - 1) * HB) <= ra.cap_dl && (f1 - f0 - (long long)nb * cLong) <= ra.cap_cl;
    };
    auto issue_early = [&](long long sn, long long e0, long long e1, long long f0, long long f1, int nb) {
        const bool fit = fits_early(e0, e1, f0, f1, nb);
        const long long lo = oLong + (long long)(nb - 1) * HB, lc = (long long)nb * cLong;
        const uint32_t dl = fit ? (uint32_t)(e1 - e0 - lo) * 8u : 0u, cl = fit ? (uint32_t)(f1 - f0 - lc) * 8u : 0u;
        mbar_expect_tx(bar, (uint32_t)(kRowsHdr + SS + RS) * 8u + dl + cl);
        bulk_g2s(my_early, a.drec + e0, kRowsHdr * 8u, bar);
        bulk_g2s(my_early + kRowsHdr, a.state + sn * SS, SS * 8u, bar);
        bulk_g2s(my_early + kRowsHdr + SS, rng_it + sn * RS, RS * 8u, bar);
        if (dl) bulk_g2s(my_early + kRowsHdr + SS + RS, a.drec + e0 + lo, dl, bar);
        if (cl) bulk_g2s(my_early + kRowsHdr + SS + RS + ra.cap_dl, a.crec + f0 + lc, cl, bar);
    };
    uint32_t phase = 0;
    if constexpr (E1) { if (t_step >= 0 && l == 0) issue_early(subject_of(t_step), d0, d1, c0, c1, ntr); }
    // E2: cp.async copies of one subject's early data (the group's lanes share the 16-byte pieces of head | state | draws into the
    // group slot) and of this lane's visit rows (into its visit slots); one commit group
    auto issue_ev = [&](double* const slot, const long long sj, const long long e0, const long long f0, const int (&mv)[O], const int mtr) {
        constexpr int H2 = kRowsHdr / 2, S2 = SS / 2, R2 = RS / 2;
#pragma unroll
        for (int i = 0; i < (H2 + S2 + R2 + G - 1) / G; ++i) {
            const int c = l + i * G;
            if (c < H2 + S2 + R2) {
                const double* src = (c < H2) ? (a.drec + e0 + 2 * c)
                                             : ((c < H2 + S2) ? (a.state + sj * SS + 2 * (c - H2)) : (rng_it + sj * RS + 2 * (c - H2 - S2)));
                cp_async16(slot + 2 * c, src);
            }
        }
        if constexpr (E3) {
            int dl = 0, cl = 0;                               // doubles of the visit rows in the design / per-draw record
            sfor<O>([&](auto kc) { constexpr int k = kc; constexpr int lc = Spec::L.long_cols[k]; dl += lc * mv[k]; cl += mv[k]; });
            if (dl <= ra.cap_dl && cl <= ra.cap_cl) {          // else: read in place
                const double* const gd = a.drec + e0 + oLong + (size_t)(mtr - 1) * HB;
                const double* const gc = a.crec + f0 + (size_t)mtr * cLong;
                for (int c = 2 * l; c < dl; c += 2 * G) cp_async16(s_vis + c, gd + c);
                for (int c = 2 * l; c < cl; c += 2 * G) cp_async16(s_vis + ra.cap_dl + c, gc + c);
            }
        } else {
            issue_visit_rows<Spec, G>(s_vis, ra.vt, a.drec + e0 + oLong + (size_t)(mtr - 1) * HB, a.crec + f0 + (size_t)mtr * cLong, mv, l);
        }
        cp_async_commit();
    };
    if constexpr (E2) { if (t_step >= 0) issue_ev(my_early0, subject_of(t_step), d0, c0, mvc, ntr); }

    double acc_cur = 0.0, acc_prop = 0.0, acc_lw = 0.0;
    for (; t_step >= 0; ++ko) {
        const bool valid = (long long)t_step * SPW + grp < n;
        const long long s = subject_of(t_step);
        const double* __restrict__ rec0 = a.drec + d0;
        const double* __restrict__ crec0 = a.crec + c0;
        const bool staged = E1 && fits_early(d0, d1, c0, c1, ntr);
        // ---- next step of this warp: offsets now (used after the visit loops) ----
        const bool has_next = t_next >= 0;
        const int t_next2 = (has_next && ko + 2 < ra.order_len) ? my_order[ko + 2] : -1;
        long long nd0 = 0, nd1 = 0, nc0 = 0, nc1 = 0;
        int n_ntr = 1;
        int nvc[O];                                       // E2: visit counts of the next subject (fresh registers: the load is consumed a visit loop later)
        sfor<O>([&](auto kc) { nvc[kc] = 0; });
        if (has_next) {
            const long long sn = subject_of(t_next);
            if constexpr (E2) load_meta(sn, nd0, nc0, nvc, n_ntr);
            else {
                nd0 = a.doff[sn]; nd1 = a.doff[sn + 1]; nc0 = a.coff[sn]; nc1 = a.coff[sn + 1];
                if constexpr (MS) n_ntr = a.trow[sn + 1] - a.trow[sn];
            }
        }
        if constexpr (E2) {
            // the copies of this step's early data and visit rows were issued one row loop ago; the group slot is shared
            cp_async_wait<0>();
            __syncwarp();
            my_early = my_early0 + (ESLOTS == 2 ? (ko & 1) * EH : 0);
        }
        // every group of the warp walks the same number of blocks (the shuffles inside the loop need the whole warp); a group
        // past its own last block re-evaluates that block and discards the result
        int ntr_w = 1;
        if constexpr (MS) {
            ntr_w = ntr;
#pragma unroll
            for (int o = 16; o >= G; o >>= 1) ntr_w = max(ntr_w, __shfl_xor_sync(0xFFFFFFFFu, ntr_w, o));
        }

        // copies of hazard-row trip jr (rows l + jr G of this lane's subject) into ring stage st
        auto issue_row = [&](const int f, const int st) {
            const int fb = MS ? min(f / NJ, ntr - 1) : 0, jr = MS ? f % NJ : f;
            const double* const rec = rec0 + (size_t)fb * HB;
            const double* const crec = crec0 + (size_t)fb * cLong;
            const int rq = PAIR ? (min(l + jr * G, RP - 1) & ~1) : min(l + jr * G, RP - 1);      // PAIR: the even row of the lane pair
            double* const dst = (PAIR ? s_ring - (lane32 & 1) : s_ring) + (size_t)st * NV * 32;
            [[maybe_unused]] const int par = lane32 & 1;
            // slot v of the stage <- field starting at src (row 0); PAIR: lanes of parity v & 1 copy the pair's two values
            auto put = [&](auto vc, const double* src) {
                constexpr int v = vc;
                if constexpr (PAIR) { if (par == (v & 1)) cp_async16(dst + v * 32, src + rq); }
                else cp_async8(dst + v * 32, src + rq);
            };
            if constexpr (PAIR) {     // band offsets: 32 ints packed at the head of slot 0
                if (par == 0) cp_async8(reinterpret_cast<double*>(reinterpret_cast<int*>(dst - (lane32 & ~1)) + (lane32 & ~1)),
                                        reinterpret_cast<const double*>(reinterpret_cast<const int*>(rec + oW1off) + rq));
            } else {
                cp_async4(reinterpret_cast<int*>(dst), reinterpret_cast<const int*>(rec + oW1off) + rq);
            }
            sfor<WB>([&](auto j) { constexpr int jv = j; put(std::integral_constant<int, 1 + jv>{}, rec + oW1v + jv * RP); });
            put(std::integral_constant<int, 1 + WB>{}, rec + oPw);
            if constexpr (!W2C) sfor<p>([&](auto j) { constexpr int jv = j; put(std::integral_constant<int, 2 + WB + jv>{}, rec + oW2 + jv * RP); });
            sfor<T>([&](auto tc) {
                constexpr int t = tc;
                constexpr int tb = RowSlots<Spec>::term_base(t), nz = RowSlots<Spec>::nzz(t), nuu = RowSlots<Spec>::nu(t);
                constexpr int oZZq = Spec::L.o_ZZ[t], oUq = Spec::L.o_U[t];
                put(std::integral_constant<int, tb>{}, crec + t * RP);
                sfor<nz>([&](auto jc) { constexpr int j = jc; put(std::integral_constant<int, tb + 1 + j>{}, rec + oZZq + j * RP); });
                sfor<nuu>([&](auto uc) { constexpr int u = uc; put(std::integral_constant<int, tb + 1 + nz + u>{}, rec + oUq + u * RP); });
            });
        };
        if constexpr (RING > 0) {
#pragma unroll
            for (int d = 0; d < LEAD; ++d) { if (d < NJ_H) issue_row(d, d); cp_async_commit(); }
        }

        // ---- head of the record, current state and this iteration's draws (uniform within the group, 16-byte loads) ----
        double stv[SS], rgv[RS], hdr[kRowsHdr];
        if constexpr (E1 || E2) {
            if constexpr (E1) mbar_wait(bar, phase);
            phase ^= 1u;
            const double2* e2 = reinterpret_cast<const double2*>(smem + (my_early - smem));      // shared-memory provenance: LDS
            sfor<kRowsHdr / 2>([&](auto j) { const double2 v = e2[j]; hdr[2 * j] = v.x; hdr[2 * j + 1] = v.y; });
            sfor<SS / 2>([&](auto j) { const double2 v = e2[kRowsHdr / 2 + j]; stv[2 * j] = v.x; stv[2 * j + 1] = v.y; });
            sfor<RS / 2>([&](auto j) { const double2 v = e2[(kRowsHdr + SS) / 2 + j]; rgv[2 * j] = v.x; rgv[2 * j + 1] = v.y; });
        } else {
            const double2* h2 = reinterpret_cast<const double2*>(rec0);
            sfor<kRowsHdr / 2>([&](auto j) { const double2 v = __ldg(h2 + j); hdr[2 * j] = v.x; hdr[2 * j + 1] = v.y; });
            const double2* st2 = reinterpret_cast<const double2*>(a.state + s * SS);
            sfor<SS / 2>([&](auto j) { const double2 v = st2[j]; stv[2 * j] = v.x; stv[2 * j + 1] = v.y; });
            const double2* rg2 = reinterpret_cast<const double2*>(rng_it + s * RS);
            sfor<RS / 2>([&](auto j) { const double2 v = __ldg(rg2 + j); rgv[2 * j] = v.x; rgv[2 * j + 1] = v.y; });
        }
        const double event = hdr[0];
        int vcount[O];
        sfor<O>([&](auto kc) {
            constexpr int k = kc;
            const double w = hdr[oV + k / 2];
            vcount[k] = (k % 2 == 0) ? __double2loint(w) : __double2hiint(w);
        });
        double bo[q], bn[q];
        sfor<q>([&](auto j) { bo[j] = stv[j]; bn[j] = stv[j] + rgv[j]; });     // :670
        const double log_u = rgv[q];
        const double pyb_old = stv[q + 1], pb_old = stv[q + 2];
        // log p(T | b_old) under the current survival parameters, carried from the previous iteration
        const double ptb_co = last_acc ? stv[q + 6] : stv[q + 5];

        // ---- log p(b) at the proposal (:673) ----
        double pb_new = 0.0;
        sfor<q>([&](auto jc) {
            constexpr int j = jc;
            double t = 0.0;
            sfor<q>([&](auto lc) { t += bn[lc] * s_invD[lc + j * q]; });
            pb_new += t * bn[j];
        });
        pb_new *= -0.5;

        // ---- GLMM log-density over the visit segment ----
        double v_pyb;
        if constexpr (E3) {
            int dl = 0, cl = 0;
            sfor<O>([&](auto kc) { constexpr int k = kc; constexpr int lc = Spec::L.long_cols[k]; dl += lc * vcount[k]; cl += vcount[k]; });
            // one copy of the visit loops for the staged rows and the rows read in place (generic loads through a selected pointer):
            // a second, shared-memory-only copy of the loops was measured and loses more to the instruction cache than LDS gains
            const bool st3 = dl <= ra.cap_dl && cl <= ra.cap_cl;
            const double* const gl = rec0 + oLong + (size_t)(ntr - 1) * HB;
            const double* const gc = crec0 + (size_t)ntr * cLong;
            v_pyb = visit_rows<Spec, G>(st3 ? s_vis : gl, st3 ? s_vis + ra.cap_dl : gc, vcount, bn, isig, l);
        } else if constexpr (E2) {
            v_pyb = visit_rows_slots<Spec, G>(s_vis, ra.vt, rec0 + oLong + (size_t)(ntr - 1) * HB, crec0 + (size_t)ntr * cLong, vcount, bn, isig, l);
        } else if constexpr (E1) {
            // one copy of the visit loops serves the staged rows and the rows read in place (segments beyond the stage capacity):
            // generic loads through a pointer that is selected here
            const double* const lr = my_early + kRowsHdr + SS + RS;
            const double* const gl = rec0 + oLong + (size_t)(ntr - 1) * HB;
            const double* const gc = crec0 + (size_t)ntr * cLong;
            v_pyb = visit_rows<Spec, G>(staged ? lr : gl, staged ? lr + ra.cap_dl : gc, vcount, bn, isig, l);
        } else {
            v_pyb = visit_rows<Spec, G>(rec0 + oLong + (size_t)(ntr - 1) * HB, crec0 + (size_t)ntr * cLong, vcount, bn, isig, l);
        }

        // ---- the early data of the warp's next step start to travel now; they are used a whole row loop later ----
        if constexpr (E1) {
            __syncwarp();                        // every lane has finished reading the stage
            if (has_next && l == 0) { fence_proxy_async(); issue_early(subject_of(t_next), nd0, nd1, nc0, nc1, n_ntr); }
        } else if (!E2 && ra.prefetch && has_next && l == 0) {
            const long long sn = subject_of(t_next);
            l2_prefetch_bulk(a.drec + nd0, 64u);
            const long long lo = oLong + (long long)(n_ntr - 1) * HB, lc = (long long)n_ntr * cLong;
            const uint32_t dl = (uint32_t)(nd1 - nd0 - lo) * 8u, cl = (uint32_t)(nc1 - nc0 - lc) * 8u;
            if (dl) l2_prefetch_bulk(a.drec + nd0 + lo, dl);
            if (cl) l2_prefetch_bulk(a.crec + nc0 + lc, cl);
            l2_prefetch_bulk(a.state + sn * SS, (uint32_t)SS * 8u);
            l2_prefetch_bulk(rng_it + sn * RS, (uint32_t)RS * 8u);
        }

        // ---- hazard rows l, l + G, ...: lin_pred_matF :80-109, log_h :684, H / HL :685-691, survival step :735-741 ----
        double s2c = 0.0, s2p = 0.0;
        if constexpr (W2C) sfor<p>([&](auto j) { constexpr int jv = j; const double wv = hdr[oW2c + jv]; const double2 g = s_g[jv]; s2c += wv * g.x; s2p += wv * g.y; });
        double H_cn = 0.0, HL_cn = 0.0, H_po = 0.0, HL_po = 0.0, H_pn = 0.0, HL_pn = 0.0;
        double lh_cn = 0.0, lh_po = 0.0, lh_pn = 0.0;
        // The lower-interval rows K+1..2K feed HL, which enters log p(T|b) of interval-censored subjects only (event == 3,
        // src/fast_LAPRWM.cpp:699): a warp none of whose subjects is interval-censored neither copies nor evaluates the
        // trips that hold only those rows (the reference evaluates them for everybody and discards the result).
        const int nj = MS ? ntr_w * NJ : ((S == 2 && NJ > NJ_H) ? (__any_sync(0xFFFFFFFFu, event == 3.0) ? NJ : NJ_H) : NJ);
        double ms_cn = 0.0, ms_po = 0.0, ms_pn = 0.0, ev_tb = 0.0;           // MS: sums over the subject's transition rows
#pragma unroll 1
        for (int jj = 0; jj < nj; ++jj) {
            const int blk = MS ? jj / NJ : 0, jt = MS ? jj % NJ : jj;        // block and trip inside the block
            const double* const rec = rec0 + (size_t)(MS ? min(blk, ntr - 1) : 0) * HB;
            const double* const crec = crec0 + (size_t)(MS ? min(blk, ntr - 1) : 0) * cLong;
            const int r = l + jt * G;
            const int rr = min(r, RP - 1);                                   // padded rows exist up to RP - 1
            // ring: start the copies of trip jj + LEAD into the stage consumed in the previous trip, then wait for trip jj
            const double* const rs = s_ring + (size_t)(RING > 0 ? jj % (RING > 0 ? RING : 1) : 0) * NV * 32;
            if constexpr (RING > 0) {
                if constexpr (PAIR) __syncwarp();        // the partner lane has read the stage that is filled next (trip jj - 1)
                if (jj + LEAD < nj) issue_row(jj + LEAD, (jj + LEAD) % RING);
                cp_async_commit();
                if constexpr (E2) {
                    // the next step's early data and visit rows start to travel behind trip 1: the groups younger than trip jj
                    // are then {trip 1, next step} at jj = 0, {next step, trip 2} at jj = 1 and {trip jj + 1} afterwards, so the
                    // copies have two trips (or, in a two-trip loop, until the top of the next step) to arrive
                    if (jj == 0) {
                        __syncwarp();            // every lane of the group has read the group slot (top of this step)
                        if (has_next) issue_ev(my_early0 + (ESLOTS == 2 ? ((ko + 1) & 1) * EH : 0), subject_of(t_next), nd0, nc0, nvc, n_ntr);
                        else cp_async_commit();
                    }
                    if (jj < 2) cp_async_wait<2>(); else cp_async_wait<1>();
                } else {
                    cp_async_wait<LEAD>();
                }
                if constexpr (PAIR) __syncwarp();        // half of this lane's slots were copied by its partner
            }
            // value v of this row: from the ring, or straight from the record
            auto rowld = [&](const int v, const double* g) { if constexpr (RING > 0) return rs[v * 32]; else return __ldg(g); };
            const int off = (RING > 0) ? (PAIR ? reinterpret_cast<const int*>(rs - lane32)[lane32] : *reinterpret_cast<const int*>(rs))
                                       : __ldg(reinterpret_cast<const int*>(rec + oW1off) + rr);
            double s1c = 0.0, s1p = 0.0;
            sfor<WB>([&](auto j) {
                constexpr int jv = j;
                const double wv = rowld(1 + jv, rec + oW1v + jv * RP + rr);
                const double2 cp = s_cp[off + jv];
                s1c += wv * cp.x;
                s1p += wv * cp.y;
            });
            double w2c_ = s2c, w2p_ = s2p;
            if constexpr (!W2C) sfor<p>([&](auto j) { constexpr int jv = j; const double wv = rowld(2 + WB + jv, rec + oW2 + jv * RP + rr); const double2 g = s_g[jv]; w2c_ += wv * g.x; w2p_ += wv * g.y; });
            double a_cn = 0.0, a_po = 0.0, a_pn = 0.0;
            sfor<T>([&](auto tc) {
                constexpr int t = tc;
                constexpr int ones0 = Spec::d.zz_ones0[t], rt = Spec::d.rt[t];
                constexpr int tb = RowSlots<Spec>::term_base(t), nz = RowSlots<Spec>::nzz(t);
                const double xxb = rowld(tb, crec + t * RP + rr);
                double zo = 0.0, zn = 0.0;
                if constexpr (ones0 == 1) { constexpr int cz = Spec::d.re2[t][0]; zo = bo[cz]; zn = bn[cz]; }
                constexpr int oZZ = Spec::L.o_ZZ[t], oU = Spec::L.o_U[t];
                sfor<rt - ones0>([&](auto jc) {
                    constexpr int j = jc;
                    constexpr int cj = Spec::d.re2[t][j + ones0];
                    const double zz = rowld(tb + 1 + j, rec + oZZ + j * RP + rr);
                    zo += zz * bo[cj];
                    zn += zz * bn[cj];
                });
                const double vo = trans_static<Spec::d.tf[t]>(xxb + zo), vn = trans_static<Spec::d.tf[t]>(xxb + zn);
                sfor<Spec::d.ut[t]>([&](auto uc) {
                    constexpr int u = uc;
                    constexpr int c = Spec::d.col[t][u];
                    double wo = vo, wn = vn;
                    if constexpr (Spec::d.u_ones[t] == 0) {
                        const double uu = rowld(tb + 1 + nz + u, rec + oU + u * RP + rr);
                        wo *= uu; wn *= uu;
                    }
                    const double2 al = s_al[c];
                    a_cn += wn * al.x;
                    a_po += wo * al.y; a_pn += wn * al.y;
                });
            });
            const double l_cn = (s1c + w2c_) + a_cn;
            const double l_po = (s1p + w2p_) + a_po, l_pn = (s1p + w2p_) + a_pn;
            const bool isH = (r >= 1 && r <= K), isHL = (S == 2 && r > K && r <= 2 * K);
            const double pw_raw = (MS || isH || isHL) ? rowld(1 + WB, rec + oPw + rr) : 0.0;
            const double pw = (isH || isHL) ? pw_raw : 0.0;
            if constexpr (MS) { if (r == 0) ev_tb = pw_raw; }                // the event-time row's Pw slot holds the transition's event indicator
            double e_cn, e_po, e_pn;
            exp3(l_cn, l_po, l_pn, e_cn, e_po, e_pn);
            e_cn *= pw; e_po *= pw; e_pn *= pw;
            if (r == 0) { lh_cn = l_cn; lh_po = l_po; lh_pn = l_pn; }
            if (isH) { H_cn += e_cn; H_po += e_po; H_pn += e_pn; }
            if constexpr (S == 2) { if (isHL) { HL_cn += e_cn; HL_po += e_po; HL_pn += e_pn; } }
            if constexpr (MS) {
                if (jt == NJ - 1) {          // end of a block: its term event * log h - H for the three (parameters, b) pairs
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) {
                        H_cn += __shfl_xor_sync(0xFFFFFFFFu, H_cn, o, G);
                        H_po += __shfl_xor_sync(0xFFFFFFFFu, H_po, o, G);
                        H_pn += __shfl_xor_sync(0xFFFFFFFFu, H_pn, o, G);
                    }
                    const double ev = __shfl_sync(0xFFFFFFFFu, ev_tb, 0, G);
                    const double h_cn = __shfl_sync(0xFFFFFFFFu, lh_cn, 0, G), h_po = __shfl_sync(0xFFFFFFFFu, lh_po, 0, G);
                    const double h_pn = __shfl_sync(0xFFFFFFFFu, lh_pn, 0, G);
                    if (blk < ntr) { ms_cn += ev * h_cn - H_cn; ms_po += ev * h_po - H_po; ms_pn += ev * h_pn - H_pn; }
                    H_cn = 0.0; H_po = 0.0; H_pn = 0.0;
                }
            }
        }

        // ---- group sums (xor butterflies inside the G lanes of a subject) ----
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            v_pyb += __shfl_xor_sync(0xFFFFFFFFu, v_pyb, o, G);
            if constexpr (!MS) {
                H_cn += __shfl_xor_sync(0xFFFFFFFFu, H_cn, o, G);
                H_po += __shfl_xor_sync(0xFFFFFFFFu, H_po, o, G);
                H_pn += __shfl_xor_sync(0xFFFFFFFFu, H_pn, o, G);
            }
            if constexpr (S == 2) {
                HL_cn += __shfl_xor_sync(0xFFFFFFFFu, HL_cn, o, G);
                HL_po += __shfl_xor_sync(0xFFFFFFFFu, HL_po, o, G);
                HL_pn += __shfl_xor_sync(0xFFFFFFFFu, HL_pn, o, G);
            }
        }
        lh_cn = __shfl_sync(0xFFFFFFFFu, lh_cn, 0, G);
        lh_po = __shfl_sync(0xFFFFFFFFu, lh_po, 0, G);
        lh_pn = __shfl_sync(0xFFFFFFFFu, lh_pn, 0, G);
        const double pyb_new = v_pyb;

        // per-subject values of the step head that are needed from here on: with two group slots they are read again from the
        // slot (nothing has overwritten it) instead of having occupied registers during the row loop
        constexpr bool REREAD = E2 && ESLOTS == 2;
        double event_t = event, log_u_t = log_u, pyb_old_t = pyb_old, pb_old_t = pb_old, ptb_co_t = ptb_co;
        double st_keep[SS - q];                                          // state slots q .. SS - 1 (sigma_b, counters, padding)
        if constexpr (REREAD) {
            const double* const es = smem + (my_early - smem);
            event_t = es[0];
            log_u_t = es[kRowsHdr + SS + q];
            sfor<SS - q>([&](auto j) { st_keep[j] = es[kRowsHdr + q + j]; });
            pyb_old_t = st_keep[1]; pb_old_t = st_keep[2];
            ptb_co_t = last_acc ? st_keep[6] : st_keep[5];
        } else {
            sfor<SS - q>([&](auto j) { st_keep[j] = stv[q + j]; });
        }
        {
        const double event = event_t, log_u = log_u_t, pyb_old = pyb_old_t, pb_old = pb_old_t, ptb_co = ptb_co_t;

        // The censoring term is evaluated twice - current parameters at the proposal, then (after the accept decision, :706)
        // proposed parameters at the accepted b (:735-752) - by ONE copy of its code (two exponentials and a logarithm of libm):
        // the kernel's hot loop has to fit the instruction cache.
        double ptb_cn = 0.0, ptb_p = 0.0;
        bool accept = false;
        const double cur_lp = pyb_old + ptb_co + pb_old;                 // :669
        constexpr bool PTBL = JMB_ROWS_PTBL != 0 && S == 2 && !MS && G >= 8;
        if constexpr (PTBL) {
            // The six exponentials exp(-H), exp(-HL) of the three (parameters, b) pairs are evaluated by six lanes of the group at
            // once (every lane holds all six sums after the butterflies), their three logarithms by three lanes: one exponential and
            // one logarithm in the instruction stream of a step instead of four and two, and the term under the proposed parameters
            // no longer waits for the accept decision.  Same values as log_ptb_rows.
            const int role = l & 7, pr_i = min(role >> 1, 2);            // pair 0 (cur, new b) | 1 (prop, old b) | 2 (prop, new b); odd lanes: HL
            const double Hs = pr_i == 0 ? H_cn : (pr_i == 1 ? H_po : H_pn), HLs = pr_i == 0 ? HL_cn : (pr_i == 1 ? HL_po : HL_pn);
            const double ex = exp1(-((role & 1) ? HLs : Hs));
            const double ex_o = __shfl_xor_sync(0xFFFFFFFFu, ex, 1);      // even lanes: exp(-HL) of their pair
            const double arg = (event == 2.0) ? (1.0 - ex) : ((event == 3.0) ? (ex_o - ex) : 1.0);
            const double lg = log(arg);
            const double lg_cn = __shfl_sync(0xFFFFFFFFu, lg, 0, G), lg_po = __shfl_sync(0xFFFFFFFFu, lg, 2, G), lg_pn = __shfl_sync(0xFFFFFFFFu, lg, 4, G);
            auto term = [&](const double lh, const double H, const double lgv) {
                double r = 0.0;
                if (event == 1.0) r = lh - H;
                if (event == 0.0) r = 0.0 - H;
                if (event == 2.0 || event == 3.0) r = lgv;
                return r;
            };
            ptb_cn = term(lh_cn, H_cn, lg_cn);
            accept = log_u < ((pyb_new + ptb_cn + pb_new) - cur_lp);     // :703, :706 (NaN rejects)
            ptb_p = accept ? term(lh_pn, H_pn, lg_pn) : term(lh_po, H_po, lg_po);
        } else if constexpr (MS || S != 2) {
            ptb_cn = MS ? ms_cn : log_ptb_rows<S>(event, lh_cn, H_cn, HL_cn);
            accept = log_u < ((pyb_new + ptb_cn + pb_new) - cur_lp);     // :703, :706 (NaN rejects)
            ptb_p = MS ? (accept ? ms_pn : ms_po) : log_ptb_rows<S>(event, accept ? lh_pn : lh_po, accept ? H_pn : H_po, 0.0);
        } else {
#pragma unroll 1
            for (int it = 0; it < 2; ++it) {
                const bool pn = accept;                                  // it == 1: which b the proposed parameters are evaluated at
                const double lh_i = it == 0 ? lh_cn : (pn ? lh_pn : lh_po), H_i = it == 0 ? H_cn : (pn ? H_pn : H_po);
                const double HL_i = it == 0 ? HL_cn : (pn ? HL_pn : HL_po);
                const double r = log_ptb_rows<S>(event, lh_i, H_i, HL_i);
                if (it == 0) { ptb_cn = r; accept = log_u < ((pyb_new + ptb_cn + pb_new) - cur_lp); }
                else ptb_p = r;
            }
        }
        const double ptb_c = accept ? ptb_cn : ptb_co;
        const double pyb_acc = accept ? pyb_new : pyb_old;
        const double pb_acc = accept ? pb_new : pb_old;
        if (valid) { acc_cur += ptb_c; acc_prop += ptb_p; acc_lw += pyb_acc + pb_acc; }

        // ---- state write-back by lane 0 of the group (:707-724); sigma_b (slot q) keeps the value read above ----
        if (valid && l == 0) {
            double outv[SS];
            sfor<SS - q>([&](auto j) { outv[q + j] = st_keep[j]; });
            sfor<q>([&](auto j) { outv[j] = accept ? bn[j] : bo[j]; });
            outv[q + 1] = pyb_acc; outv[q + 2] = pb_acc;
            outv[q + 3] = st_keep[3] + (accept ? 1.0 : 0.0);
            outv[q + 4] = st_keep[4] + (accept ? 1.0 : 0.0);
            outv[q + 5] = ptb_c; outv[q + 6] = ptb_p;
            double2* st2 = reinterpret_cast<double2*>(a.state + s * SS);
            sfor<SS / 2>([&](auto j) {
                constexpr int j0 = 2 * j;
                // pairs that hold only b are written when the proposal was accepted
                if (j0 + 1 < q) { if (accept) st2[j] = make_double2(outv[j0], outv[j0 + 1]); }
                else st2[j] = make_double2(outv[j0], outv[j0 + 1]);
            });
        }
        }
        d0 = nd0; d1 = nd1; c0 = nc0; c1 = nc1; ntr = n_ntr;
        t_step = t_next; t_next = t_next2;
    }
    // ---- deterministic per-CTA partial sums: one slot per (warp, group), summed in slot order ----
    if (l == 0) { const int slot = warp * SPW + grp; s_red[slot * 3] = acc_cur; s_red[slot * 3 + 1] = acc_prop; s_red[slot * 3 + 2] = acc_lw; }
    __syncthreads();
    if (tid < 3) {
        double t = 0.0;
        for (int g = 0; g < WARPS * SPW; ++g) t += s_red[g * 3 + tid];
        a.partials[blockIdx.x * 3 + tid] = t;
    }
    if (ra.peer_mbox != nullptr) {
        __shared__ unsigned int s_tk;
        __threadfence();
        __syncthreads();
        if (tid == 0) s_tk = atomicAdd(ra.ticket, 1u);
        __syncthreads();
        if (s_tk == gridDim.x - 1) rows_publish_sums(a.partials, (int)gridDim.x, ra, s_ring0);     // the ring is idle by now
    }
    (void)cc;
}

}  // namespace jmb
