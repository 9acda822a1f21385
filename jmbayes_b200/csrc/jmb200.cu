// jmb200.cu - host runtime + C ABI (include/jmb200.h) of the B200-native mvJointModelBayes
// MCMC hot path.  Product code: never includes or links anything under oracle/; every compute
// entry point fails loudly when no CUDA device is usable (no CPU fallback).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: the library itself is resolved with dlopen

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <queue>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "jmb_host.h"
#include "jmb_step_static.cuh"
#include "jmb_step_rows.cuh"

using namespace jmb;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& m) { g_err = m; return 1; }
int jmb_set_error(const std::string& m) { return fail(m); }     // other translation units of the library (jmb_host.h)
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

extern "C" const char* jmb_last_error(void) { return g_err.c_str(); }

extern "C" int jmb_device_count(int* count) {
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) { *count = 0; return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = c;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// NCCL through dlopen (reuses the copy torch already loaded when there is one)
// ---------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static int load_nccl() {
    if (g_nccl.lib) return 0;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(std::string("cannot load libnccl: ") + dlerror());
    g_nccl.lib = lib;
    *(void**)&g_nccl.GetUniqueId = dlsym(lib, "ncclGetUniqueId");
    *(void**)&g_nccl.CommInitRank = dlsym(lib, "ncclCommInitRank");
    *(void**)&g_nccl.AllReduce = dlsym(lib, "ncclAllReduce");
    *(void**)&g_nccl.CommDestroy = dlsym(lib, "ncclCommDestroy");
    *(void**)&g_nccl.GetErrorString = dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce) return fail("libnccl lacks required symbols");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
// everything that belongs to one chain (one mvglmer draw): several chains of the same fit can be in
// flight on their own streams, sharing the design records
struct ChainCtx {
    cudaStream_t stream = nullptr;
    ChainConst cc{};
    double* d_crec = nullptr;                       // per-draw records
    double* d_state = nullptr; double* d_par = nullptr; Ctl* d_ctl = nullptr; unsigned int* d_ticket = nullptr;
    double* d_partials = nullptr; double* d_tmp = nullptr;  // d_tmp: n*(q+1)
    double* d_rng = nullptr;                        // [rng_chunk][n][rng_stride] draws of the b-step
    size_t rng_capacity = 0; int rng_chunk = 1;
    double* h_crec = nullptr;                       // pinned staging for the per-draw records
    double* h_par = nullptr;                        // pinned staging for the parameter block
    double* h_tmp = nullptr;                        // pinned staging n*(q+1)
    bool draw_set = false, chain_open = false, fin_attr_set = false, wore = false, step_published = false;
    int iter_host = 0, total_iters = 0, keep_host = 0, check_host = 0, n_out = 0, n_td = 0;
    double *o_b = nullptr, *o_par = nullptr, *o_trace = nullptr;   // device outputs
    size_t cap_ob = 0, cap_opar = 0, cap_otr = 0;   // capacities (doubles) of the output buffers
    unsigned long long p2p_seq = 0;                 // stamp of the last fused all-reduce of this slot
    int slot_index = 0;
    FinalizeArgs fa{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evk0 = nullptr, evk1 = nullptr;
};

struct jmb_handle_s {
    int device = 0;
    ModelMeta mm{};
    ParamLayout pl{};
    int n = 0;
    // multi-state fits: nT transition rows in the survival part, trow[i] = first row of subject i
    bool ms = false; int nT = 0, nRisks = 1;
    std::vector<int> trow, kn_first, kn_last;
    int* d_trow = nullptr;
    long long subject_offset = 0;
    int n_block_cap = 1024;
    std::vector<long long> N;                       // rows per outcome
    std::vector<std::vector<int>> rowptr;           // host CSR per outcome
    std::vector<long long> doff, coff;              // host offsets
    std::vector<double> h_drec;                     // host copy of the design records (jmb_debug_pack_host handles only)
    bool host_only = false;
    // device data shared by all chains
    double* d_drec = nullptr; long long* d_doff = nullptr; long long* d_coff = nullptr;
    int* d_rowptr = nullptr;
    double* d_eval = nullptr;
    double* d_Rchol = nullptr;                      // [n][q(q+1)/2] proposal factors (Covs$b)
    int rng_stride = 2;
    // chains
    std::vector<ChainCtx*> slots;
    ChainCtx* c = nullptr;                          // current slot
    int grid_cap = 0;                               // partials per chain
    // survival-block Laplace step (logPosterior / gradient_logPosterior)
    double *d_Wlong = nullptr, *d_Wlongs = nullptr, *d_sp_partials = nullptr, *d_sp_out = nullptr, *d_theta = nullptr;
    std::vector<double> ecs;                        // event_colSums of [W1 | W2 | Wlong]
    std::vector<double> h_event;                    // host copy of Data$event
    std::vector<double> sp_last_theta, sp_last_out; // cache: optim() asks for fn and gr at the same point
    bool wlong_set = false;
    // launch configuration
    int grid = 0, subjects_per_cta = 0, smem_bytes = 0;
    long long launches = 0, launches_at_ev0 = 0, launches_timed = 0;
    // fixed-effects designs for Xbetas_calc on the device (jmb_set_xdesigns)
    double* d_xrec = nullptr; long long* d_xoff = nullptr; double* d_betas = nullptr;
    XbDesc xd{}; int n_betas = 0; bool xdesigns_set = false;
    // step kernel chosen for this model: a precompiled static specialisation or the generic one
    void (*static_kernel)(int, ParamLayout, ChainConst, StepArgs) = nullptr;
    void (*tma_kernel)(int, ParamLayout, ChainConst, StepArgs, TmaArgs) = nullptr;
    void (*rows_kernel)(int, ParamLayout, ChainConst, StepArgs, RowsArgs) = nullptr;
    const char* kernel_name = "generic";
    std::string kernel_label;
    int static_smem = 0, tma_smem = 0, tma_grid = 0, num_sms = 148;
    int rows_smem = 0, rows_grid = 0, rows_G = 0, rows_early = 0;
    int* d_rows_order = nullptr; RowsMeta* d_rows_meta = nullptr;    // step lists of the (CTA, warp) slots; per-subject addressing
    double rows_imbalance = 0.0;                                      // estimated cost of the slowest slot / the mean
    TmaArgs ta{};
    RowsArgs ra{};
    int kernel_mode = 0;            // 0 generic, 1 static direct, 2 static TMA-staged, 3 static row-strided groups (default)
    int select_kernel();
    int rows_schedule(int G, int total_steps);
    int add_slot();
    // multi-GPU
    ncclComm_t comm = nullptr; int nranks = 1, rank = 0;
    // peer-memory mailboxes for the fused all-reduce of the per-iteration sums
    static constexpr int kMboxSlots = 64;
    double* d_mbox = nullptr;                       // this rank's mailbox [kMboxSlots][2][nranks][kMboxEntry]
    std::vector<void*> peer_ptrs;                   // peers' mailboxes opened through CUDA IPC
    double** d_peer_mbox = nullptr;                 // device copy of the pointer table
    bool p2p = false;
};

template <class Spec, int EARLY>
static void (*rows_kernel_of(int G))(int, ParamLayout, ChainConst, StepArgs, RowsArgs) {
    switch (G) {
        case 2: return jmb_step_rows<Spec, 2, EARLY>;
        case 4: return jmb_step_rows<Spec, 4, EARLY>;
        case 8: return jmb_step_rows<Spec, 8, EARLY>;
        case 16: return jmb_step_rows<Spec, 16, EARLY>;
    }
    return nullptr;
}
template <class Spec>
static void (*rows_kernel_of_early(int G, int early))(int, ParamLayout, ChainConst, StepArgs, RowsArgs) {
    if constexpr (JMB_ROWS_RING == 2) {      // the pipelined variants are written for a two-stage ring
        if (early == 2) return rows_kernel_of<Spec, 2>(G);
        if (early == 3) return rows_kernel_of<Spec, 3>(G);
    }
    return early == 1 ? rows_kernel_of<Spec, 1>(G) : rows_kernel_of<Spec, 0>(G);
}
static void (*pick_rows_kernel(const ModelDesc& md, int G, int early))(int, ParamLayout, ChainConst, StepArgs, RowsArgs) {
    if (same_desc(md, SpecValueGaussian::d)) return rows_kernel_of_early<SpecValueGaussian>(G, early);
    if (same_desc(md, SpecThreeOutcomes::d)) return rows_kernel_of_early<SpecThreeOutcomes>(G, early);
    if (same_desc(md, SpecThreeOutcomesIC::d)) return rows_kernel_of_early<SpecThreeOutcomesIC>(G, early);
    if (same_desc(md, SpecMultiState::d)) return rows_kernel_of_early<SpecMultiState>(G, early);
    return nullptr;
}

// Steps of the row-strided kernel dealt to its (CTA, warp) slots.  A step's cost depends on the data - the row-loop trips (a warp
// with an interval-censored subject walks the lower-interval rows too; a multi-state warp walks max(ntr) blocks) and the visit-row
// trips of every outcome - so equal NUMBERS of steps per warp leave the slowest warp of the config-4 shard with 16 % more than the
// average (13.2 steps per warp).  The steps are sorted by estimated cost and handed, longest first, to the least-loaded slot
// (LPT); a slot walks its list in ascending step order.  The assignment is static, so the per-slot partial sums keep their fixed
// order.  JMB_ROWS_LPT=0 restores the CTA-contiguous, warp-interleaved assignment of equal counts.
int jmb_handle_s::rows_schedule(int G, int total_steps) {
    const ModelDesc& md = mm.d;
    const int spw = 32 / G, nslots = rows_grid * kRowsWarps;
    const int NR = 1 + md.K * md.S, NJ = (NR + G - 1) / G, NJ_H = (1 + md.K + G - 1) / G;
    double w_row = 0.135, w_vis = 0.05;              // cost of a step = 1 + w_row * row trips + w_vis * visit trips
    if (const char* e = getenv("JMB_LPT_ROW")) w_row = atof(e);
    if (const char* e = getenv("JMB_LPT_VIS")) w_vis = atof(e);
    bool lpt = true;
    if (const char* e = getenv("JMB_ROWS_LPT")) lpt = atoi(e) != 0;
    std::vector<float> cost(total_steps);
    for (int t = 0; t < total_steps; ++t) {
        int ntr_w = 1, vtr = 0; bool any3 = false;
        for (int k = 0; k < md.O; ++k) {
            int mv = 0;
            for (int g = 0; g < spw; ++g) { const int i = std::min(n - 1, t * spw + g); mv = std::max(mv, rowptr[k][i + 1] - rowptr[k][i]); }
            vtr += (mv + G - 1) / G;
        }
        for (int g = 0; g < spw; ++g) {
            const int i = std::min(n - 1, t * spw + g);
            if (ms) ntr_w = std::max(ntr_w, trow[i + 1] - trow[i]);
            else if (h_event[i] == 3.0) any3 = true;
        }
        const int nj = ms ? ntr_w * NJ : ((md.S == 2 && NJ > NJ_H) ? (any3 ? NJ : NJ_H) : NJ);
        cost[t] = (float)(1.0 + w_row * nj + w_vis * vtr);
    }
    std::vector<std::vector<int>> lists(nslots);
    std::vector<double> load(nslots, 0.0);
    if (lpt) {
        std::vector<int> idx(total_steps);
        for (int t = 0; t < total_steps; ++t) idx[t] = t;
        std::stable_sort(idx.begin(), idx.end(), [&](int x, int y) { return cost[x] > cost[y]; });
        typedef std::pair<double, int> LS;
        std::priority_queue<LS, std::vector<LS>, std::greater<LS>> pq;
        for (int sl = 0; sl < nslots; ++sl) pq.push(LS(0.0, sl));
        for (int t : idx) {
            LS top = pq.top(); pq.pop();
            lists[top.second].push_back(t);
            load[top.second] = top.first + cost[t];
            pq.push(LS(load[top.second], top.second));
        }
        for (auto& v : lists) std::sort(v.begin(), v.end());
    } else {
        const int spc = (total_steps + rows_grid - 1) / rows_grid;
        for (int c = 0; c < rows_grid; ++c)
            for (int w = 0; w < kRowsWarps; ++w)
                for (int t = c * spc + w; t < std::min(total_steps, (c + 1) * spc); t += kRowsWarps) { lists[c * kRowsWarps + w].push_back(t); load[c * kRowsWarps + w] += cost[t]; }
    }
    size_t len = 1; double tot = 0.0, mx = 0.0;
    for (int sl = 0; sl < nslots; ++sl) { len = std::max(len, lists[sl].size()); tot += load[sl]; mx = std::max(mx, load[sl]); }
    rows_imbalance = tot > 0.0 ? mx / (tot / nslots) : 1.0;
    std::vector<int> order((size_t)nslots * len, -1);
    for (int sl = 0; sl < nslots; ++sl) std::copy(lists[sl].begin(), lists[sl].end(), order.begin() + (size_t)sl * len);
    ra.order_len = (int)len;
    cudaFree(d_rows_order); d_rows_order = nullptr;
    cudaError_t e = cudaMalloc(&d_rows_order, sizeof(int) * order.size());
    if (e == cudaSuccess) e = cudaMemcpy(d_rows_order, order.data(), sizeof(int) * order.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { g_err = std::string("rows_schedule: ") + cudaGetErrorString(e); return 1; }
    ra.order = d_rows_order;
    ra.meta = nullptr;
    if (rows_early >= 2) {
        std::vector<RowsMeta> meta(n);
        for (int i = 0; i < n; ++i) {
            RowsMeta& m = meta[i];
            m.d0 = doff[i]; m.c0 = coff[i]; m.v[0] = m.v[1] = m.v[2] = 0;
            for (int k = 0; k < md.O && k < 3; ++k) m.v[k] = rowptr[k][i + 1] - rowptr[k][i];
            m.ntr = ms ? trow[i + 1] - trow[i] : 1;
        }
        cudaFree(d_rows_meta); d_rows_meta = nullptr;
        e = cudaMalloc(&d_rows_meta, sizeof(RowsMeta) * (size_t)std::max(n, 1));
        if (e == cudaSuccess) e = cudaMemcpy(d_rows_meta, meta.data(), sizeof(RowsMeta) * (size_t)n, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { g_err = std::string("rows_schedule(meta): ") + cudaGetErrorString(e); return 1; }
        ra.meta = d_rows_meta;
    }
    return 0;
}

int jmb_handle_s::select_kernel() {
    static_kernel = nullptr; tma_kernel = nullptr; rows_kernel = nullptr; kernel_name = "generic"; kernel_mode = 0;
    const ModelDesc& md = mm.d;
    if (same_desc(md, SpecValueGaussian::d)) {
        static_kernel = jmb_step_static<SpecValueGaussian>; tma_kernel = jmb_step_tma<SpecValueGaussian>; kernel_name = "value_gaussian";
    } else if (same_desc(md, SpecThreeOutcomes::d)) {
        static_kernel = jmb_step_static<SpecThreeOutcomes>; tma_kernel = jmb_step_tma<SpecThreeOutcomes>; kernel_name = "three_outcomes";
    } else if (same_desc(md, SpecThreeOutcomesIC::d)) {
        static_kernel = jmb_step_static<SpecThreeOutcomesIC>; tma_kernel = jmb_step_tma<SpecThreeOutcomesIC>; kernel_name = "three_outcomes_ic";
    }
    const int groups = 256 / md.G, spw = 32 / md.G, SS = mm.L.state_stride;
    static_smem = (int)sizeof(double) * (2 * mm.P + md.q * md.q + groups * 3);
    if (static_kernel) kernel_mode = 1;
    if (tma_kernel) {
        // per-warp stage capacity: the largest run of `spw` consecutive records
        long long max_d = 0, max_c = 0;
        for (int i = 0; i + spw <= n || i == 0; i += 1) {
            const int j = std::min(n, i + spw);
            max_d = std::max(max_d, doff[j] - doff[i]);
            max_c = std::max(max_c, coff[j] - coff[i]);
            if (j == n) break;
        }
        const int tgroups = kTmaThreads / md.G;
        const int head = (2 * mm.P + md.q * md.q + tgroups * 3 + kTmaWarps * 2 + 1) / 2 * 2;
        // kTmaCtas CTAs per SM share 226 KB; kTmaWarps warps x 2 stages per CTA
        const int RS = (md.q + 2) / 2 * 2;     // rng stride (q increments + log u)
        const long long budget = (((226LL / kTmaCtas) * 1024) / 8 - head) / (2 * kTmaWarps) - spw * (SS + RS);
        long long cap_d = (max_d + 1) / 2 * 2, cap_c = (max_c + 1) / 2 * 2;
        if (cap_d + cap_c > budget) {
            // keep the split proportional; oversize steps fall back to direct global loads
            const double f = (double)budget / (double)(cap_d + cap_c);
            cap_d = (long long)(cap_d * f) / 2 * 2;
            cap_c = (budget - cap_d) / 2 * 2;
        }
        ta.cap_d = (int)cap_d;
        ta.cap_c = (int)cap_c;
        // the typical step must fit, otherwise staging is pointless
        long long mean_d = (doff[n] / std::max(n, 1)) * spw, mean_c = (coff[n] / std::max(n, 1)) * spw;
        if (cap_d >= mean_d && cap_c >= mean_c && budget > 0) {
            const long long stage = (long long)ta.cap_d + ta.cap_c + spw * (SS + RS);
            tma_smem = (int)(8LL * (head + (long long)kTmaWarps * 2 * stage));
            const int unit = kTmaWarps * spw;
            int g = std::min((n + unit - 1) / unit, kTmaCtas * num_sms);
            int spc = ((n + g - 1) / g + unit - 1) / unit * unit;
            ta.subjects_per_cta = spc;
            tma_grid = (n + spc - 1) / spc;
            cudaError_t e = cudaFuncSetAttribute((const void*)tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tma_smem);
            if (e != cudaSuccess) { g_err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e); return 1; }
            kernel_mode = 2;
        }
    }
    // row-strided groups (jmb_step_rows.cuh): G lanes per subject, 32 / G subjects per warp step; the default
    if (same_desc(md, SpecMultiState::d)) kernel_name = "multistate_illness_death";
    {
        // group width and staging of a step's early data per model shape, from the sweeps in profiles/r2_rows_kernel.md and
        // profiles/r2b_rows_kernel.md: early = 0 read in place (+ L2 prefetch), 1 bulk-copied per-group stage, 2 cp.async group
        // slots + lane-private visit slots, 3 cp.async into the per-group stage
        int G = md.MS ? 16 : 8;
        int early = (md.MS || md.O == 1) ? 2 : 1;     // variants 1 and 3 are within 3 % of each other on configs 3 / 4; 1 was ahead on two of three boxes
        if (const char* g = getenv("JMB_ROWS_G")) G = atoi(g);
        if (G != 2 && G != 4 && G != 8 && G != 16) G = 8;
        if (const char* e = getenv("JMB_ROWS_EARLY")) early = std::max(0, std::min(3, atoi(e)));
        if (early >= 2 && md.O > 3) early = 0;
        if (JMB_ROWS_RING != 2 && early >= 2) early = 0;
        rows_kernel = pick_rows_kernel(md, G, early);
        rows_early = early;
        if (rows_kernel) {
            rows_G = G;
            const int spw = 32 / G;
            // shared memory: parameter pairs | invD | per-group partial sums | early-stage barriers | hazard-row ring | early stages
            int nv = 2 + md.WB + (md.w2_const ? 0 : md.p);
            for (int t = 0; t < md.T; ++t) nv += 1 + (md.rt[t] - md.zz_ones0[t]) + (md.u_ones[t] ? 0 : md.ut[t]);
            const long long fixed_b = 8LL * (2 * mm.P + (md.q * md.q + 1) / 2 * 2 + (kRowsWarps * spw * 3 + 1) / 2 * 2) + 8LL * ((kRowsWarps + 1) / 2 * 2) +
                                      8LL * kRowsWarps * JMB_ROWS_RING * nv * 32;
            const long long cta_budget = (227LL * 1024) / kRowsCtas - 1024 - 64;      // per-CTA share of an SM, minus the reserved KB and the static bytes
            const int RSs = (md.q + 2) / 2 * 2;
            ra.cap_dl = 0; ra.cap_cl = 0; ra.vt = 0;
            if (early == 1 || early == 3) {
                // early stage of a subject: head | state | draws | visit rows of the design and of the per-draw record; the
                // capacity covers the longest visit segment unless that would not leave room for kRowsCtas CTAs per SM
                // (subjects beyond the capacity read their visit rows in place)
                long long max_dl = 0, max_cl = 0;
                for (int i = 0; i < n; ++i) {
                    max_dl = std::max(max_dl, doff[i + 1] - doff[i] - mm.L.o_long);
                    max_cl = std::max(max_cl, coff[i + 1] - coff[i] - mm.L.c_long);
                }
                const int head = (kRowsHdr + mm.L.state_stride + RSs) * ((early == 3 && JMB_ROWS_REREAD) ? 2 : 1);
                const long long budget = (cta_budget - fixed_b) / 8 / ((long long)kRowsWarps * spw) - head;
                long long cd = (max_dl + 1) / 2 * 2, cc_ = (max_cl + 1) / 2 * 2;
                if (cd + cc_ > budget) {
                    const double f = (double)std::max<long long>(budget, 0) / (double)(cd + cc_);
                    cd = (long long)(cd * f) / 2 * 2; cc_ = (long long)(cc_ * f) / 2 * 2;
                }
                if (const char* e = getenv("JMB_ROWS_CAP")) {     // tests: a small stage, so that long visit segments are read in place
                    const long long lim = std::max(0, atoi(e)) / 2 * 2;
                    cd = std::min(cd, lim); cc_ = std::min(cc_, lim);
                }
                ra.cap_dl = (int)std::max<long long>(cd, 0); ra.cap_cl = (int)std::max<long long>(cc_, 0);
                const int RSd = head + ra.cap_dl + ra.cap_cl;
                rows_smem = (int)(fixed_b + 8LL * kRowsWarps * spw * RSd);
            } else if (early == 2) {
                // group slots (head | state | draws) + as many visit-row trips in the lane-private visit slots as the longest
                // segment needs and the CTA's share of the SM holds
                int vs = 0, maxv = 1;
                for (int k = 0; k < md.O; ++k) {
                    vs += 2 + md.qk[k] - md.z_ones0[k];
                    for (int i = 0; i < n; ++i) maxv = std::max(maxv, rowptr[k][i + 1] - rowptr[k][i]);
                }
                const long long slots_b = 8LL * kRowsWarps * spw * (kRowsHdr + mm.L.state_stride + RSs) * (JMB_ROWS_REREAD ? 2 : 1);
                const long long per_trip = 8LL * kRowsWarps * 32 * vs;
                int vt = std::min((maxv + G - 1) / G, md.MS ? 1 : 2);    // measured: more trips in the slots buy nothing (later trips are read in place)
                if (const char* v = getenv("JMB_ROWS_VT")) vt = std::max(0, atoi(v));
                vt = (int)std::min<long long>(vt, std::max<long long>(0, (cta_budget - fixed_b - slots_b) / per_trip));
                ra.vt = vt;
                rows_smem = (int)(fixed_b + slots_b + per_trip * vt);
            } else {
                rows_smem = (int)fixed_b;
            }
            {   // always: the kernel's static shared memory counts against the 48 KB default as well
                cudaError_t e = cudaFuncSetAttribute((const void*)rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rows_smem);
                if (e != cudaSuccess) { g_err = std::string("cudaFuncSetAttribute(rows): ") + cudaGetErrorString(e); return 1; }
            }
            if (const char* co = getenv("JMB_ROWS_CARVEOUT"))   // experiments: shared-memory carve-out in percent (the rest of the 256 KB is L1)
                cudaFuncSetAttribute((const void*)rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
            const long long total_steps = ((long long)n + spw - 1) / spw;
            int waves = 1;                                   // persistent: one resident wave of CTAs
            if (const char* w = getenv("JMB_ROWS_WAVES")) waves = std::max(1, atoi(w));
            const long long ctas = std::max<long long>(1, std::min<long long>((total_steps + kRowsWarps - 1) / kRowsWarps, (long long)kRowsCtas * num_sms * waves));
            rows_grid = (int)ctas;
            ra.prefetch = 1;
            if (const char* pf = getenv("JMB_ROWS_PREFETCH")) ra.prefetch = atoi(pf) ? 1 : 0;
            if (rows_schedule(G, (int)total_steps)) return 1;
            if (getenv("JMB_ROWS_DEBUG")) {
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)rows_kernel, kRowsThreads, (size_t)rows_smem);
                fprintf(stderr, "[jmb rows] G=%d early=%d smem=%d B vt=%d CTAs/SM=%d grid=%d steps=%lld list=%d est. slowest/mean slot=%.4f\n",
                        G, early, rows_smem, ra.vt, occ, rows_grid, total_steps, ra.order_len, rows_imbalance);
            }
            kernel_mode = 3;
        }
    }
    const char* env = getenv("JMB_KERNEL");   // generic | direct | tma | rows : force a variant (tests, A/B runs)
    if (env) {
        if (!strcmp(env, "generic")) kernel_mode = 0;
        else if (!strcmp(env, "direct") && static_kernel) kernel_mode = 1;
        else if (!strcmp(env, "tma") && tma_kernel && tma_smem > 0) kernel_mode = 2;
    }
    if (getenv("JMB_FORCE_GENERIC") && getenv("JMB_FORCE_GENERIC")[0] == '1') kernel_mode = 0;
    switch (kernel_mode) {
        case 0: kernel_label = "generic"; break;
        case 1: kernel_label = std::string("static:") + kernel_name; break;
        case 2: kernel_label = std::string("static-tma:") + kernel_name; break;
        default: kernel_label = std::string("static-rows") + std::to_string(rows_G) + (rows_early == 3 ? "q:" : (rows_early == 2 ? "p:" : (rows_early ? "e:" : ":"))) + kernel_name; break;
    }
    return 0;
}

template <typename F>
static void parallel_for(int n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
    if (n < 4096) nt = 1;
    std::vector<std::thread> th;
    int chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        int lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

static void build_param_layout(const ModelMeta& mm, int n_td, int n_block_cap, ParamLayout& pl) {
    int o = 0;
    auto take = [&](int len) { int r = o; o += std::max(len, 1); return r; };
    pl.cur = take(mm.P); pl.prop = take(mm.P);
    pl.invD = take(mm.d.q * mm.d.q); pl.sigmas = take(mm.d.O);
    pl.tau_Bs = take(JMB_MAX_RISKS); pl.tau_g = take(1); pl.tau_a = take(1); pl.xi_a = take(1);
    pl.phi_g = take(mm.d.p); pl.phi_a = take(mm.d.A); pl.nu_a = take(mm.d.A); pl.tau_td = take(n_td);
    pl.Tau_g = take(mm.d.p * mm.d.p); pl.Tau_a = take(mm.d.A * mm.d.A); pl.Tau_a_pen = take(mm.d.A * mm.d.A);
    pl.sig = take(3);
    pl.S_Bs = take(mm.B * mm.B); pl.S_g = take(mm.d.p * mm.d.p); pl.S_a = take(mm.d.A * mm.d.A);
    pl.H_Bs = take(mm.B * mm.B); pl.H_g = take(mm.d.p * mm.d.p); pl.H_a = take(mm.d.A * mm.d.A);
    pl.mean_Bs = take(mm.B); pl.Tau_Bs = take(mm.B * mm.B); pl.mean_g = take(mm.d.p); pl.mean_a = take(mm.d.A);
    pl.cur_lp = take(1);
    pl.block_par = take(mm.P * n_block_cap);
    pl.sums = take(4);
    pl.total = o;
}

// ---------------------------------------------------------------------------------------------
// jmb_create: validate, compact and repack the static designs into per-subject records
// ---------------------------------------------------------------------------------------------
static int create_impl(const jmb_data* d, int device, jmb_handle* out, bool host_only) {
    if (!d || !out) return fail("jmb_create: null argument");
    *out = nullptr;
    if (!host_only) {
        int ndev = 0;
        if (jmb_device_count(&ndev) || ndev == 0)
            return fail("jmb_create: no CUDA device available (this library has no CPU fallback): " + g_err);
        if (device < 0 || device >= ndev) return fail("jmb_create: bad device ordinal");
    }
    const bool ms = d->multi_state != 0;
    if (ms) {
        if (d->interval_cens) return fail("jmb_create: multiState cannot be combined with interval censoring (the reference's state copy "
                                          "assumes K rows per subject there, src/fast_LAPRWM.cpp:712-717)");
        if (!d->W1) return fail("jmb_create: multiState needs the dense (stratified) W1 / W1s; the knots alternative is single-stratum");
        if (d->nT < d->n || !d->idT) return fail("jmb_create: multiState needs nT >= n and idT");
        if (d->nRisks < 1 || d->nRisks > JMB_MAX_RISKS || !d->kn_strat_first || !d->kn_strat_last) return fail("jmb_create: nRisks out of range (1..16) or kn_strat_* missing");
        for (int k = 0; k < d->nRisks; ++k)
            if (d->kn_strat_first[k] < 0 || d->kn_strat_last[k] < d->kn_strat_first[k] || d->kn_strat_last[k] >= d->n_Bs)
                return fail("jmb_create: kn_strat_first / kn_strat_last out of range");
    }
    if (d->n <= 0 || d->n_outcomes <= 0 || d->n_outcomes > JMB_MAX_OUTCOMES) return fail("jmb_create: bad n / n_outcomes");
    if (d->q <= 0 || d->q > JMB_MAX_Q) return fail("jmb_create: q out of range");
    if (d->n_terms <= 0 || d->n_terms > JMB_MAX_TERMS) return fail("jmb_create: n_terms out of range");
    if (d->n_Bs <= 0 || d->n_Bs > JMB_MAX_BS || d->n_gammas < 0 || d->n_gammas > JMB_MAX_GAMMAS ||
        d->n_alphas <= 0 || d->n_alphas > JMB_MAX_ALPHAS) return fail("jmb_create: parameter block sizes out of range");
    const int S = d->interval_cens ? 2 : 1;
    if (d->K <= 0 || 1 + d->K * S > 32) return fail("jmb_create: 1 + K * (1 + interval_cens) must be <= 32");
    if (d->q + kStateExtra > 32) return fail("jmb_create: q + 7 must be <= 32");
    const bool from_knots = (d->W1 == nullptr);
    if (from_knots) {
        if (!d->knots || !d->Time || !d->st || (S == 2 && !d->st_int)) return fail("jmb_create: W1 is NULL, so knots / Time / st (/ st_int) are required");
        if (d->ord < 1 || d->ord > 8 || d->n_knots != d->n_Bs + d->ord) return fail("jmb_create: need 1 <= ord <= 8 and n_knots = n_Bs + ord");
        for (int j = 1; j < d->n_knots; ++j) if (d->knots[j] < d->knots[j - 1]) return fail("jmb_create: knots must be non-decreasing");
    } else if (!d->W1s || (S == 2 && !d->W1s_int)) return fail("jmb_create: W1s (/ W1s_int) required with a dense W1");
    if (S == 2 && (!d->Pw_int || !d->ZZs_int || !d->Us_int || (d->n_gammas && !d->W2s_int)))
        return fail("jmb_create: interval_cens needs the *_int designs");

    auto* h = new jmb_handle_s();
    h->device = device;
    h->n = d->n;
    h->subject_offset = d->subject_offset;
    const int n = d->n, O = d->n_outcomes, q = d->q, T = d->n_terms, B = d->n_Bs, p = d->n_gammas, K = d->K;
    // rows of the survival part: one per subject, or one per transition a subject is at risk for (multi-state,
    // src/fast_LAPRWM.cpp:519-529); trow[i] .. trow[i+1]-1 are subject i's rows (rows_wlong[[i]], idT_rsum)
    const int nT = ms ? d->nT : n;
    h->ms = ms; h->nT = nT; h->nRisks = ms ? d->nRisks : 1;
    h->trow.assign(n + 1, 0);
    if (ms) {
        int prev = 0;
        for (int r = 0; r < nT; ++r) {
            const int sbj = d->idT[r];
            if (sbj < 0 || sbj >= n || sbj < prev) { delete h; return fail("jmb_create: idT must be 0-based and grouped by subject in subject order"); }
            prev = sbj;
            h->trow[sbj + 1] += 1;
        }
        for (int i = 0; i < n; ++i) {
            if (h->trow[i + 1] == 0) { delete h; return fail("jmb_create: every subject needs at least one transition row"); }
            h->trow[i + 1] += h->trow[i];
        }
        h->kn_first.assign(d->kn_strat_first, d->kn_strat_first + d->nRisks);
        h->kn_last.assign(d->kn_strat_last, d->kn_strat_last + d->nRisks);
    } else {
        for (int i = 0; i <= n; ++i) h->trow[i] = i;
    }
    const long long ns = (long long)nT * K;
    ModelMeta& mm = h->mm;
    ModelDesc& md = mm.d;
    memset(&mm, 0, sizeof(mm));
    md.O = O; md.q = q; md.T = T; md.A = d->n_alphas; md.p = p; md.K = K; md.S = S; md.MS = ms ? 1 : 0;
    mm.B = B;
    mm.P = B + p + d->n_alphas;
    md.G = (1 + K * S <= 16) ? 16 : 32;
    if (q + kStateExtra > md.G) md.G = 32;
    for (int k = 0; k < O; ++k) {
        md.fam[k] = d->fams[k]; md.link[k] = d->links[k]; md.qk[k] = d->q_k[k];
        if (md.qk[k] < 0 || md.qk[k] > JMB_MAX_Q) { delete h; return fail("jmb_create: q_k out of range"); }
        for (int j = 0; j < md.qk[k]; ++j) {
            md.re_inds[k][j] = d->RE_inds[k][j];
            if (md.re_inds[k][j] < 0 || md.re_inds[k][j] >= q) { delete h; return fail("jmb_create: RE_inds out of range"); }
        }
    }
    for (int t = 0; t < T; ++t) {
        md.rt[t] = d->r_t[t]; md.ut[t] = d->u_t[t]; md.tf[t] = d->trans_Funs[t];
        if (md.rt[t] < 0 || md.rt[t] > JMB_MAX_RT || md.ut[t] <= 0 || md.ut[t] > JMB_MAX_UT) { delete h; return fail("jmb_create: term sizes out of range"); }
        for (int j = 0; j < md.rt[t]; ++j) {
            md.re2[t][j] = d->RE_inds2[t][j];
            if (md.re2[t][j] < 0 || md.re2[t][j] >= q) { delete h; return fail("jmb_create: RE_inds2 out of range"); }
        }
        for (int u = 0; u < md.ut[t]; ++u) {
            md.col[t][u] = d->col_inds[t][u];
            if (md.col[t][u] < 0 || md.col[t][u] >= d->n_alphas) { delete h; return fail("jmb_create: col_inds out of range"); }
        }
    }

    // ---- segment boundaries: rows grouped by subject, idL2 = last row, idGK_fast = (i+1)K-1 ----
    h->N.assign(d->N, d->N + O);
    h->rowptr.resize(O);
    for (int k = 0; k < O; ++k) {
        const long long Nk = d->N[k];
        if (Nk < 0 || Nk > 2147483000LL) { delete h; return fail("jmb_create: N_k out of range"); }
        std::vector<int>& rp = h->rowptr[k];
        rp.assign(n + 1, 0);
        const int32_t* id = d->idL[k];
        int prev = -1;
        for (long long r = 0; r < Nk; ++r) {
            int s = id[r];
            // grouped by subject; a subject without rows in an outcome is allowed (survfitJM new data:
            // log_longF_svft skips empty outcomes, src/survfitJM_mvJMbayes.cpp:109-110)
            if (s < 0 || s >= n || s < prev) { delete h; return fail("jmb_create: idL must be 0-based and grouped by subject"); }
            prev = s;
            rp[s + 1] += 1;
        }
        for (int i = 0; i < n; ++i) rp[i + 1] += rp[i];
        for (int i = 0; i < n; ++i)
            if (d->idL2[k][i] != rp[i + 1] - 1) { delete h; return fail("jmb_create: idL2 is not the last row of each subject"); }
    }
    if (d->idGK_fast)
        for (int i = 0; i < nT; ++i)
            if (d->idGK_fast[i] != (i + 1) * K - 1) { delete h; return fail("jmb_create: idGK_fast must be (i+1)K-1"); }

    // ---- band of the B-spline basis rows (splineDesign rows have `ord` contiguous non-zeros) ----
    auto band_of = [&](const double* W, long long rows, long long r, int& first, int& last) {
        first = B; last = -1;
        for (int j = 0; j < B; ++j)
            if (W[r + (long long)j * rows] != 0.0) { if (first == B) first = j; last = j; }
    };
    int WB = 1;
    if (from_knots) WB = std::min(d->ord, B);
    else {
        std::atomic<int> wmax{1};
        auto scan = [&](const double* W, long long rows) {
            parallel_for((int)std::min<long long>(rows, 2147483647LL), [&, W, rows](int lo, int hi) {
                int m = 1;
                for (long long r = lo; r < hi; ++r) { int f, l; band_of(W, rows, r, f, l); if (l >= f) m = std::max(m, l - f + 1); }
                int cur = wmax.load();
                while (m > cur && !wmax.compare_exchange_weak(cur, m)) {}
            });
        };
        scan(d->W1, nT); scan(d->W1s, ns);
        if (S == 2) scan(d->W1s_int, ns);
        WB = wmax.load();
        if (WB > 8) WB = B;                 // not banded: keep the dense rows
        WB = std::min(WB, B);
    }
    md.WB = WB;

    // ---- constant design columns that need not be stored (decided from the data) ----
    auto all_ones = [&](const double* col, long long rows) {
        std::atomic<int> ok{1};
        parallel_for((int)rows, [&, col](int lo, int hi) {
            for (long long r = lo; r < hi; ++r) if (col[r] != 1.0) { ok.store(0); return; }
        });
        return ok.load() == 1;
    };
    for (int k = 0; k < O; ++k) md.z_ones0[k] = (md.qk[k] >= 1 && all_ones(d->Z[k], d->N[k])) ? 1 : 0;
    for (int t = 0; t < T; ++t) {
        bool zo = md.rt[t] >= 1 && all_ones(d->ZZ[t], nT) && all_ones(d->ZZs[t], ns) && (S == 1 || all_ones(d->ZZs_int[t], ns));
        md.zz_ones0[t] = zo ? 1 : 0;
        bool uo = all_ones(d->U[t], (long long)nT * md.ut[t]) && all_ones(d->Us[t], ns * md.ut[t]) &&
                  (S == 1 || all_ones(d->Us_int[t], ns * md.ut[t]));
        md.u_ones[t] = uo ? 1 : 0;
    }
    {
        std::atomic<int> ok{(p > 0 && !ms) ? 1 : 0};   // multi-state: W2 rows differ between a subject's transitions
        if (p > 0 && !ms)
            parallel_for(n, [&](int lo, int hi) {
                for (int i = lo; i < hi; ++i)
                    for (int j = 0; j < p; ++j) {
                        const double w = d->W2[i + (long long)j * n];
                        for (int k = 0; k < K; ++k) {
                            if (d->W2s[(long long)i * K + k + (long long)j * ns] != w) { ok.store(0); return; }
                            if (S == 2 && d->W2s_int[(long long)i * K + k + (long long)j * ns] != w) { ok.store(0); return; }
                        }
                    }
            });
        md.w2_const = ok.load();
    }

    // ---- record layout ----
    mm.L = make_layout(md);
    const RecordLayout& L = mm.L;
    const int RP = L.RP;

    h->doff.assign(n + 1, 0);
    h->coff.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        const int ntr = h->trow[i + 1] - h->trow[i];
        long long dl = L.o_long + (long long)(ntr - 1) * L.HB, cl = (long long)L.c_long * ntr;
        for (int k = 0; k < O; ++k) {
            int v = h->rowptr[k][i + 1] - h->rowptr[k][i];
            dl += (long long)v * L.long_cols[k];
            cl += v;
        }
        h->doff[i + 1] = h->doff[i] + ((dl + 3) / 4) * 4;       // records start on 32-byte sectors
        h->coff[i + 1] = h->coff[i] + ((cl + 3) / 4) * 4;
    }

    // ---- pack the design records ----
    std::vector<double> drec((size_t)h->doff[n], 0.0);
    parallel_for(n, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            double* rec0 = drec.data() + h->doff[i];
            const int ntr = h->trow[i + 1] - h->trow[i];
            rec0[0] = ms ? (double)ntr : d->event[i];
            for (int k = 0; k < O; ++k) reinterpret_cast<int*>(rec0 + L.o_V)[k] = h->rowptr[k][i + 1] - h->rowptr[k][i];
            if (md.w2_const) for (int j = 0; j < p; ++j) rec0[L.o_W2c + j] = d->W2[i + (long long)j * n];
            for (int tb = 0; tb < ntr; ++tb) {
            double* rec = rec0 + (size_t)tb * L.HB;           // hazard-row block of transition row ti
            const long long ti = h->trow[i] + tb;
            if (ms) rec[L.o_Pw] = d->event[ti];               // lane 0's Pw slot: event indicator of the transition
            int* w1off = reinterpret_cast<int*>(rec + L.o_W1off);
            for (int r = 0; r < RP; ++r) {
                int cls = (r == 0) ? 0 : (r <= K ? 1 : ((S == 2 && r <= 2 * K) ? 2 : 3));
                w1off[r] = 0;
                if (cls == 3) continue;
                const long long rows = cls == 0 ? nT : ns;
                const long long src = cls == 0 ? ti : (cls == 1 ? ti * K + (r - 1) : ti * K + (r - 1 - K));
                const double* W1x = cls == 0 ? d->W1 : (cls == 1 ? d->W1s : d->W1s_int);
                const double* W2x = cls == 0 ? d->W2 : (cls == 1 ? d->W2s : d->W2s_int);
                if (cls == 1) rec[L.o_Pw + r] = d->Pw[src];
                if (cls == 2) rec[L.o_Pw + r] = d->Pw_int[src];
                if (from_knots) {
                    const double tx = cls == 0 ? d->Time[src] : (cls == 1 ? d->st[src] : d->st_int[src]);
                    int off; double bv[16];
                    spline_band_row(d->knots, d->n_knots, d->ord, tx, off, bv);
                    w1off[r] = off;
                    for (int j = 0; j < WB; ++j) rec[L.o_W1v + j * RP + r] = bv[j];
                } else {
                    int f, l;
                    band_of(W1x, rows, src, f, l);
                    int off = (l < f) ? 0 : std::min(f, B - WB);
                    if (WB == B) off = 0;
                    w1off[r] = off;
                    for (int j = 0; j < WB; ++j) rec[L.o_W1v + j * RP + r] = W1x[src + (long long)(off + j) * rows];
                }
                if (!md.w2_const) for (int j = 0; j < p; ++j) rec[L.o_W2 + j * RP + r] = W2x[src + (long long)j * rows];
                for (int t = 0; t < T; ++t) {
                    const double* ZZx = cls == 0 ? d->ZZ[t] : (cls == 1 ? d->ZZs[t] : d->ZZs_int[t]);
                    const double* Ux = cls == 0 ? d->U[t] : (cls == 1 ? d->Us[t] : d->Us_int[t]);
                    const int z0 = md.zz_ones0[t];
                    for (int j = z0; j < md.rt[t]; ++j) rec[L.o_ZZ[t] + (j - z0) * RP + r] = ZZx[src + (long long)j * rows];
                    if (!md.u_ones[t]) for (int u = 0; u < md.ut[t]; ++u) rec[L.o_U[t] + u * RP + r] = Ux[src + (long long)u * rows];
                }
            }
            }
            double* lrec = rec0 + L.o_long + (size_t)(ntr - 1) * L.HB;
            for (int k = 0; k < O; ++k) {
                const int r0 = h->rowptr[k][i], v = h->rowptr[k][i + 1] - r0;
                const long long Nk = d->N[k];
                const int z0 = md.z_ones0[k];
                for (int r = 0; r < v; ++r) lrec[r] = d->y[k][r0 + r];
                for (int j = z0; j < md.qk[k]; ++j)
                    for (int r = 0; r < v; ++r) lrec[(1 + j - z0) * v + r] = d->Z[k][r0 + r + (long long)j * Nk];
                lrec += L.long_cols[k] * v;
            }
        }
    });

    // proposal factors: upper Cholesky factors of Covs$b packed by column (R/mvJointModelBayes.R:728-729)
    const int rlen = q * (q + 1) / 2;
    std::vector<double> rchol((size_t)n * rlen);
    parallel_for(n, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            const double* Rm = d->Covs_b + (long long)i * q * q;
            for (int j = 0; j < q; ++j)
                for (int l = 0; l <= j; ++l) rchol[(size_t)i * rlen + j * (j + 1) / 2 + l] = Rm[l + j * q];
        }
    });
    h->rng_stride = (q + 2) / 2 * 2;

    // event_colSums of W1 and W2 (R/mvJointModelBayes.R:677-679) for gradient_logPosterior
    h->h_event.assign(d->event, d->event + nT);
    h->ecs.assign(B + p + d->n_alphas, 0.0);
    if (from_knots) {      // event_colSumsW1 from the band rows, subjects in order (the dense sum's order per column)
        for (int i = 0; i < n; ++i) {
            int off; double bv[16];
            spline_band_row(d->knots, d->n_knots, d->ord, d->Time[i], off, bv);
            for (int j = 0; j < std::min(d->ord, B); ++j) h->ecs[off + j] += d->event[i] * bv[j];
        }
    } else
    for (int j = 0; j < B; ++j) { double t = 0.0; for (int i = 0; i < nT; ++i) t += d->event[i] * d->W1[i + (long long)j * nT]; h->ecs[j] = t; }
    for (int j = 0; j < p; ++j) { double t = 0.0; for (int i = 0; i < nT; ++i) t += d->event[i] * d->W2[i + (long long)j * nT]; h->ecs[B + j] = t; }

    if (host_only) {       // layout inspection only (jmb_debug_pack_host): no device, no compute
        h->host_only = true;
        h->h_drec = std::move(drec);
        build_param_layout(mm, JMB_MAX_TERMS, h->n_block_cap, h->pl);
        *out = h;
        return 0;
    }

    // ---- device allocations ----
    cudaError_t e;
#define CUH(call) do { e = (call); if (e != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e); jmb_destroy(h); return fail(m_); } } while (0)
    CUH(cudaSetDevice(device));
    CUH(cudaMalloc(&h->d_drec, sizeof(double) * std::max<size_t>(drec.size(), 2)));
    CUH(cudaMalloc(&h->d_doff, sizeof(long long) * (n + 1)));
    CUH(cudaMalloc(&h->d_coff, sizeof(long long) * (n + 1)));
    CUH(cudaMalloc(&h->d_rowptr, sizeof(int) * (size_t)O * (n + 1)));
    CUH(cudaMalloc(&h->d_eval, sizeof(double) * (size_t)std::max(n, nT) * 6));
    CUH(cudaMalloc(&h->d_trow, sizeof(int) * (n + 1)));
    CUH(cudaMemcpy(h->d_trow, h->trow.data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));
    CUH(cudaMemcpy(h->d_drec, drec.data(), sizeof(double) * drec.size(), cudaMemcpyHostToDevice));
    CUH(cudaMalloc(&h->d_Rchol, sizeof(double) * rchol.size()));
    CUH(cudaMemcpy(h->d_Rchol, rchol.data(), sizeof(double) * rchol.size(), cudaMemcpyHostToDevice));
    CUH(cudaMemcpy(h->d_doff, h->doff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
    CUH(cudaMemcpy(h->d_coff, h->coff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
    for (int k = 0; k < O; ++k)
        CUH(cudaMemcpy(h->d_rowptr + (size_t)k * (n + 1), h->rowptr[k].data(), sizeof(int) * (n + 1), cudaMemcpyHostToDevice));

    // parameter block (n_td capacity = JMB_MAX_TERMS)
    build_param_layout(mm, JMB_MAX_TERMS, h->n_block_cap, h->pl);

    // launch configuration: G lanes per subject, 256 threads, a few subjects per group per CTA
    const int groups = 256 / md.G;
    h->subjects_per_cta = groups * 4;
    h->grid = (n + h->subjects_per_cta - 1) / h->subjects_per_cta;
    h->smem_bytes = (int)sizeof(double) * (2 * mm.P + q * q + O + groups * 3 * (q + kStateExtra) + groups * 3);
    {
        cudaDeviceProp prop;
        CUH(cudaGetDeviceProperties(&prop, device));
        h->num_sms = prop.multiProcessorCount;
        if (h->select_kernel()) { std::string m_ = g_err; jmb_destroy(h); return fail(m_); }
    }
    h->grid_cap = std::max(std::max(h->grid, h->tma_grid), h->rows_grid);
    if (h->add_slot()) { std::string m_ = g_err; jmb_destroy(h); return fail(m_); }   // chain slot 0
#undef CUH
    *out = h;
    return 0;
}

extern "C" int jmb_create(const jmb_data* d, int device, jmb_handle* out) { return create_impl(d, device, out, false); }

// one more chain slot (own stream, per-draw records, state, parameter block, staging)
int jmb_handle_s::add_slot() {
    ChainCtx* x = new ChainCtx();
    x->slot_index = (int)slots.size();
    slots.push_back(x);
    if (!c) c = x;
    const size_t crec_len = (size_t)std::max<long long>(coff[n], 2);
    const int q = mm.d.q;
    CU(cudaSetDevice(device));
    CU(cudaStreamCreateWithFlags(&x->stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&x->d_crec, sizeof(double) * crec_len));
    CU(cudaMalloc(&x->d_state, sizeof(double) * (size_t)n * mm.L.state_stride));
    CU(cudaMalloc(&x->d_tmp, sizeof(double) * (size_t)n * (q + 1)));
    CU(cudaMalloc(&x->d_ctl, sizeof(Ctl)));
    CU(cudaMalloc(&x->d_ticket, sizeof(unsigned int)));
    CU(cudaMemset(x->d_ticket, 0, sizeof(unsigned int)));
    CU(cudaMalloc(&x->d_par, sizeof(double) * pl.total));
    CU(cudaMalloc(&x->d_partials, sizeof(double) * 3 * (size_t)grid_cap));
    CU(cudaHostAlloc(&x->h_crec, sizeof(double) * crec_len, cudaHostAllocDefault));
    CU(cudaHostAlloc(&x->h_tmp, sizeof(double) * (size_t)n * (q + 1), cudaHostAllocDefault));
    CU(cudaHostAlloc(&x->h_par, sizeof(double) * pl.total, cudaHostAllocDefault));
    memset(x->h_par, 0, sizeof(double) * pl.total);
    CU(cudaMemset(x->d_state, 0, sizeof(double) * (size_t)n * mm.L.state_stride));
    CU(cudaMemset(x->d_ctl, 0, sizeof(Ctl)));
    CU(cudaMemset(x->d_par, 0, sizeof(double) * pl.total));
    CU(cudaEventCreate(&x->ev0)); CU(cudaEventCreate(&x->ev1));
    CU(cudaEventCreate(&x->evk0)); CU(cudaEventCreate(&x->evk1));
    return 0;
}

extern "C" int jmb_chain_slots(jmb_handle h, int count) {
    if (!h || count < 1 || count > 64) return fail("jmb_chain_slots: bad argument");
    while ((int)h->slots.size() < count)
        if (h->add_slot()) return 1;
    return 0;
}

extern "C" int jmb_select_chain(jmb_handle h, int slot) {
    if (!h || slot < 0 || slot >= (int)h->slots.size()) return fail("jmb_select_chain: no such chain slot (call jmb_chain_slots)");
    h->c = h->slots[slot];
    return 0;
}

extern "C" int jmb_destroy(jmb_handle h) {
    if (!h) return 0;
    if (h->host_only) { delete h; return 0; }
    cudaSetDevice(h->device);
    for (ChainCtx* x : h->slots) if (x->stream) cudaStreamSynchronize(x->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    cudaFree(h->d_drec); cudaFree(h->d_doff); cudaFree(h->d_coff); cudaFree(h->d_rowptr); cudaFree(h->d_eval);
    cudaFree(h->d_trow); cudaFree(h->d_rows_order); cudaFree(h->d_rows_meta);
    cudaFree(h->d_Rchol); cudaFree(h->d_xrec); cudaFree(h->d_xoff); cudaFree(h->d_betas);
    for (size_t r = 0; r < h->peer_ptrs.size(); ++r)
        if (h->peer_ptrs[r] && (int)r != h->rank) cudaIpcCloseMemHandle(h->peer_ptrs[r]);
    cudaFree(h->d_mbox); cudaFree(h->d_peer_mbox);
    cudaFree(h->d_Wlong); cudaFree(h->d_Wlongs); cudaFree(h->d_sp_partials); cudaFree(h->d_sp_out); cudaFree(h->d_theta);
    for (ChainCtx* x : h->slots) {
        cudaFree(x->d_crec); cudaFree(x->d_state); cudaFree(x->d_par); cudaFree(x->d_ctl); cudaFree(x->d_ticket); cudaFree(x->d_partials);
        cudaFree(x->d_tmp); cudaFree(x->o_b); cudaFree(x->o_par); cudaFree(x->o_trace); cudaFree(x->d_rng);
        if (x->h_crec) cudaFreeHost(x->h_crec);
        if (x->h_par) cudaFreeHost(x->h_par);
        if (x->h_tmp) cudaFreeHost(x->h_tmp);
        if (x->ev0) cudaEventDestroy(x->ev0);
        if (x->ev1) cudaEventDestroy(x->ev1);
        if (x->evk0) cudaEventDestroy(x->evk0);
        if (x->evk1) cudaEventDestroy(x->evk1);
        if (x->stream) cudaStreamDestroy(x->stream);
        delete x;
    }
    delete h;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-draw vectors
// ---------------------------------------------------------------------------------------------
// per-draw records of all subjects into `base` (coff[n] doubles)
static void pack_draw(jmb_handle h, const jmb_draw* w, double* base) {
    const ModelMeta& mm = h->mm;
    const int n = h->n, K = mm.d.K, RP = mm.L.RP, S = mm.d.S, T = mm.d.T, O = mm.d.O;
    parallel_for(n, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            double* rec = base + h->coff[i];
            const int ntr = h->trow[i + 1] - h->trow[i];
            for (int tb = 0; tb < ntr; ++tb) {              // one [T][RP] block per transition row (one unless multi-state)
                const long long ti = h->trow[i] + tb;
                for (int t = 0; t < T; ++t) {
                    double* x = rec + (size_t)tb * mm.L.c_long + t * RP;
                    x[0] = w->XXbetas[t][ti];
                    for (int r = 1; r <= K; ++r) x[r] = w->XXsbetas[t][ti * K + r - 1];
                    if (S == 2) for (int r = 1; r <= K; ++r) x[K + r] = w->XXsbetas_int[t][ti * K + r - 1];
                    for (int r = 1 + K * S; r < RP; ++r) x[r] = 0.0;
                }
            }
            double* l = rec + (size_t)mm.L.c_long * ntr;
            for (int k = 0; k < O; ++k) {
                const int r0 = h->rowptr[k][i], v = h->rowptr[k][i + 1] - r0;
                memcpy(l, w->Xbetas[k] + r0, sizeof(double) * v);
                l += v;
            }
        }
    });
}

extern "C" int jmb_set_draw(jmb_handle h, const jmb_draw* w) {
    if (!h || !w) return fail("jmb_set_draw: null argument");
    if (h->host_only) return fail("jmb_set_draw: this handle was made by jmb_debug_pack_host (no device)");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->c->stream));        // the staging buffer may still be in flight
    const ModelMeta& mm = h->mm;
    const int n = h->n, O = mm.d.O;
    if (mm.d.S == 2 && !w->XXsbetas_int) return fail("jmb_set_draw: XXsbetas_int required with interval censoring");
    pack_draw(h, w, h->c->h_crec);
    CU(cudaMemcpyAsync(h->c->d_crec, h->c->h_crec, sizeof(double) * h->coff[n], cudaMemcpyHostToDevice, h->c->stream));
    for (int j = 0; j < mm.d.q * mm.d.q; ++j) h->c->h_par[h->pl.invD + j] = w->invD[j];
    for (int k = 0; k < O; ++k) h->c->h_par[h->pl.sigmas + k] = (mm.d.fam[k] == JMB_FAM_GAUSSIAN) ? w->sigmas[k] : 1.0;
    CU(cudaMemcpyAsync(h->c->d_par + h->pl.invD, h->c->h_par + h->pl.invD, sizeof(double) * (mm.d.q * mm.d.q + O), cudaMemcpyHostToDevice, h->c->stream));
    h->c->draw_set = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Xbetas_calc on the device: fixed-effects designs once per fit, betas per draw
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_set_xdesigns(jmb_handle h, const jmb_xdesigns* x) {
    if (!h || !x || !x->p_k || !x->X || !x->outcome || !x->f_t || !x->indFixed || !x->XX || !x->XXs)
        return fail("jmb_set_xdesigns: null argument");
    const ModelMeta& mm = h->mm;
    const int n = h->n, K = mm.d.K, RP = mm.L.RP, S = mm.d.S, T = mm.d.T, O = mm.d.O;
    if (S == 2 && !x->XXs_int) return fail("jmb_set_xdesigns: XXs_int required with interval censoring");
    XbDesc& xd = h->xd;
    memset(&xd, 0, sizeof(xd));
    xd.T = T; xd.O = O; xd.RP = RP;
    int nb = 0;
    for (int k = 0; k < O; ++k) {
        if (x->p_k[k] <= 0 || x->p_k[k] > 64) return fail("jmb_set_xdesigns: p_k out of range");
        xd.pk[k] = x->p_k[k]; xd.boff[k] = nb; nb += x->p_k[k];
    }
    h->n_betas = nb;
    int tb = 0;
    for (int t = 0; t < T; ++t) {
        if (x->f_t[t] < 0 || x->f_t[t] > JMB_MAX_RT) return fail("jmb_set_xdesigns: f_t out of range (0..8)");
        if (x->outcome[t] < 0 || x->outcome[t] >= O) return fail("jmb_set_xdesigns: outcome[t] out of range");
        xd.ft[t] = x->f_t[t]; xd.tbase[t] = tb; tb += x->f_t[t]; xd.tboff[t] = xd.boff[x->outcome[t]];
        for (int f = 0; f < xd.ft[t]; ++f) {
            xd.indf[t][f] = x->indFixed[t][f];
            if (xd.indf[t][f] < 0 || xd.indf[t][f] >= xd.pk[x->outcome[t]]) return fail("jmb_set_xdesigns: indFixed out of range");
        }
    }
    const long long nT = h->ms ? h->nT : n, ns = nT * K;     // rows of XX / XXs (transition rows of a multi-state fit)
    std::vector<long long> xoff(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        long long len = (long long)tb * RP * (h->trow[i + 1] - h->trow[i]);
        for (int k = 0; k < O; ++k) len += (long long)(h->rowptr[k][i + 1] - h->rowptr[k][i]) * xd.pk[k];
        xoff[i + 1] = xoff[i] + len;
    }
    std::vector<double> xrec((size_t)std::max<long long>(xoff[n], 1), 0.0);
    parallel_for(n, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i) {
            double* xr = xrec.data() + xoff[i];
            const int ntr = h->trow[i + 1] - h->trow[i];
            for (int blk = 0; blk < ntr; ++blk) {           // one block of term columns per transition row
                const long long ti = h->trow[i] + blk;
                for (int t = 0; t < T; ++t)
                    for (int f = 0; f < xd.ft[t]; ++f) {
                        double* col = xr + ((size_t)blk * tb + xd.tbase[t] + f) * RP;
                        col[0] = x->XX[t][ti + (long long)f * nT];
                        for (int r = 1; r <= K; ++r) col[r] = x->XXs[t][ti * K + r - 1 + (long long)f * ns];
                        if (S == 2) for (int r = 1; r <= K; ++r) col[K + r] = x->XXs_int[t][ti * K + r - 1 + (long long)f * ns];
                    }
            }
            double* l = xr + (size_t)tb * RP * ntr;
            for (int k = 0; k < O; ++k) {
                const int r0 = h->rowptr[k][i], v = h->rowptr[k][i + 1] - r0;
                const long long Nk = h->N[k];
                for (int c = 0; c < xd.pk[k]; ++c)
                    for (int r = 0; r < v; ++r) l[(size_t)c * v + r] = x->X[k][r0 + r + (long long)c * Nk];
                l += (size_t)xd.pk[k] * v;
            }
        }
    });
    CU(cudaSetDevice(h->device));
    cudaFree(h->d_xrec); cudaFree(h->d_xoff); cudaFree(h->d_betas);
    h->d_xrec = nullptr; h->d_xoff = nullptr; h->d_betas = nullptr;
    CU(cudaMalloc(&h->d_xrec, sizeof(double) * xrec.size()));
    CU(cudaMalloc(&h->d_xoff, sizeof(long long) * (n + 1)));
    CU(cudaMalloc(&h->d_betas, sizeof(double) * std::max(nb, 1) * 64));       // one slice per chain slot
    CU(cudaMemcpy(h->d_xrec, xrec.data(), sizeof(double) * xrec.size(), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_xoff, xoff.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
    h->xdesigns_set = true;
    return 0;
}

extern "C" int jmb_set_draw_betas(jmb_handle h, const double* const* betas, const double* sigmas, const double* invD) {
    if (!h || !betas || !invD) return fail("jmb_set_draw_betas: null argument");
    if (!h->xdesigns_set) return fail("jmb_set_draw_betas: call jmb_set_xdesigns first");
    CU(cudaSetDevice(h->device));
    const ModelMeta& mm = h->mm;
    const int O = mm.d.O, n = h->n;
    ChainCtx* c = h->c;
    CU(cudaStreamSynchronize(c->stream));           // h_par staging may still be in flight
    double* hb = c->h_tmp;                          // pinned staging (n * (q + 1) doubles >= n_betas)
    if (h->n_betas > (long long)n * (mm.d.q + 1)) return fail("jmb_set_draw_betas: too many fixed effects for the staging buffer");
    for (int k = 0; k < O; ++k)
        for (int j = 0; j < h->xd.pk[k]; ++j) hb[h->xd.boff[k] + j] = betas[k][j];
    double* d_b = h->d_betas + (size_t)c->slot_index * std::max(h->n_betas, 1);
    CU(cudaMemcpyAsync(d_b, hb, sizeof(double) * h->n_betas, cudaMemcpyHostToDevice, c->stream));
    const unsigned grid = (unsigned)(((long long)n * 32 + 255) / 256);
    jmb_xbetas_kernel<<<grid, 256, 0, c->stream>>>(h->xd, h->d_xrec, h->d_xoff, h->d_coff, h->d_rowptr, d_b, n, c->d_crec,
                                                   h->ms ? h->d_trow : nullptr);
    h->launches += 1;
    CU(cudaGetLastError());
    for (int j = 0; j < mm.d.q * mm.d.q; ++j) c->h_par[h->pl.invD + j] = invD[j];
    for (int k = 0; k < O; ++k) c->h_par[h->pl.sigmas + k] = (mm.d.fam[k] == JMB_FAM_GAUSSIAN) ? sigmas[k] : 1.0;
    CU(cudaMemcpyAsync(c->d_par + h->pl.invD, c->h_par + h->pl.invD, sizeof(double) * (mm.d.q * mm.d.q + O), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));           // the staging buffers are reused by chain_begin
    c->draw_set = true;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// launches
// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch of the two per-iteration kernels (JMB_PDL=0 disables it for A/B runs): launch latency and
// CTA set-up of the step kernel overlap the previous finalize kernel, and the finalize kernel preloads its parameter block
// under the tail of the step kernel.  (Prefetching the first record stages before the wait as well costs the step kernel
// 180 B of extra spills and 35 % of its speed: profiles/r1_experiments.md.)
static bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("JMB_PDL"); on = (e && e[0] == '0') ? 0 : 1; }
    return on == 1;
}
// JMB_PUBLISH=1: the step kernel's last CTA stores the rank's sums into the mailboxes instead of the finalize kernel.  Off by
// default: measured on 2 GPUs it does not pay (0.1813 vs 0.1796 ms per iteration, profiles/r2_scaling.md) - the ticket and the
// fences in every CTA cost what the earlier NVLink flight saves.
static bool publish_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("JMB_PUBLISH"); on = (e && e[0] == '1') ? 1 : 0; }
    return on == 1;
}
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static int launch_step(jmb_handle h, int mode) {
    StepArgs a{};
    a.drec = h->d_drec; a.doff = h->d_doff; a.crec = h->c->d_crec; a.coff = h->d_coff; a.rowptr = h->d_rowptr;
    a.state = h->c->d_state; a.par = h->c->d_par; a.ctl = h->c->d_ctl; a.partials = h->c->d_partials; a.eval_out = h->d_eval;
    a.n = h->n; a.subject_offset = h->subject_offset; a.mode = mode; a.subjects_per_cta = h->subjects_per_cta;
    a.rng = h->c->d_rng; a.rng_chunk = h->c->rng_chunk; a.rng_stride = h->rng_stride;
    a.trow = h->d_trow; a.eval_stride = std::max(h->n, h->nT);
    if (mode == MODE_STEP && h->kernel_mode == 3) {
        RowsArgs ra = h->ra;
        h->c->step_published = false;
        if (h->nranks > 1 && h->p2p && !h->c->wore && publish_enabled()) {
            // this rank's sums travel to the mailboxes from the step kernel's last CTA (the finalize kernel of the same
            // iteration uses the same stamp and only waits)
            ra.peer_mbox = h->d_peer_mbox; ra.nranks = h->nranks; ra.rank = h->rank; ra.slot = h->c->slot_index;
            ra.stamp = h->c->p2p_seq + 1; ra.ticket = h->c->d_ticket;
            h->c->step_published = true;
        }
        if (pdl_enabled()) {
            CU(launch_pdl(h->rows_kernel, dim3(h->rows_grid), dim3(kRowsThreads), (size_t)h->rows_smem, h->c->stream, h->mm.B, h->pl, h->c->cc, a, ra));
        } else {
            h->rows_kernel<<<h->rows_grid, kRowsThreads, h->rows_smem, h->c->stream>>>(h->mm.B, h->pl, h->c->cc, a, ra);
        }
    }
    else if (mode == MODE_STEP && h->kernel_mode == 2) {
        if (pdl_enabled()) {
            CU(launch_pdl(h->tma_kernel, dim3(h->tma_grid), dim3(kTmaThreads), (size_t)h->tma_smem, h->c->stream, h->mm.B, h->pl, h->c->cc, a, h->ta));
        } else {
            h->tma_kernel<<<h->tma_grid, kTmaThreads, h->tma_smem, h->c->stream>>>(h->mm.B, h->pl, h->c->cc, a, h->ta);
        }
    }
    else if (mode == MODE_STEP && h->kernel_mode == 1)
        h->static_kernel<<<h->grid, 256, h->static_smem, h->c->stream>>>(h->mm.B, h->pl, h->c->cc, a);
    else if (h->ms && h->mm.d.G == 16)
        jmb_step_kernel<16, true><<<h->grid, 256, h->smem_bytes, h->c->stream>>>(h->mm, h->pl, h->c->cc, a);
    else if (h->ms)
        jmb_step_kernel<32, true><<<h->grid, 256, h->smem_bytes, h->c->stream>>>(h->mm, h->pl, h->c->cc, a);
    else if (h->mm.d.G == 16)
        jmb_step_kernel<16, false><<<h->grid, 256, h->smem_bytes, h->c->stream>>>(h->mm, h->pl, h->c->cc, a);
    else
        jmb_step_kernel<32, false><<<h->grid, 256, h->smem_bytes, h->c->stream>>>(h->mm, h->pl, h->c->cc, a);
    h->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

static int launch_finalize(jmb_handle h, int phase) {
    FinalizeArgs a = h->c->fa;
    a.par = h->c->d_par; a.ctl = h->c->d_ctl; a.partials = h->c->d_partials; a.n_partials = h->grid;
    a.use_reduced = (h->nranks > 1) ? (h->p2p ? 2 : 1) : 0; a.phase = phase;
    a.n_partials = (h->kernel_mode == 3) ? h->rows_grid : ((h->kernel_mode == 2) ? h->tma_grid : h->grid);
    a.wore = h->c->wore ? 1 : 0; a.sp_out = h->d_sp_out;
    if (a.wore) a.use_reduced = 3;                  // the data sums come from the survival-block kernels
    if (!a.wore && phase == 0 && h->nranks > 1 && h->p2p) {
        a.peer_mbox = h->d_peer_mbox; a.my_mbox = h->d_mbox; a.nranks = h->nranks; a.rank = h->rank;
        a.slot = h->c->slot_index; a.stamp = ++h->c->p2p_seq;
        a.published = h->c->step_published ? 1 : 0;
        h->c->step_published = false;
    }
    if (!a.wore && phase == 0 && h->nranks > 1 && !h->p2p) {
        jmb_reduce_partials_kernel<<<1, 128, 0, h->c->stream>>>(h->c->d_partials, a.n_partials, h->c->d_par + h->pl.sums);
        h->launches += 1;
        CU(cudaGetLastError());
        ncclResult_t rc = g_nccl.AllReduce(h->c->d_par + h->pl.sums, h->c->d_par + h->pl.sums, 3, ncclDouble, ncclSum, h->comm, h->c->stream);
        if (rc != ncclSuccess) return fail(std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    }
    const size_t fin_smem = sizeof(double) * (size_t)h->pl.block_par;
    a.use_smem = (fin_smem <= 200 * 1024) ? 1 : 0;
    if (a.use_smem && fin_smem > 48 * 1024 && !h->c->fin_attr_set) {
        CU(cudaFuncSetAttribute((const void*)jmb_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
        h->c->fin_attr_set = true;
    }
    if (pdl_enabled()) CU(launch_pdl(jmb_finalize_kernel, dim3(1), dim3(128), a.use_smem ? fin_smem : 0, h->c->stream, h->mm, h->pl, h->c->cc, a));
    else jmb_finalize_kernel<<<1, 128, a.use_smem ? fin_smem : 0, h->c->stream>>>(h->mm, h->pl, h->c->cc, a);
    h->launches += 1;
    CU(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// chain
// ---------------------------------------------------------------------------------------------
// parameter block, loop constants and output buffers of a chain (shared by lap_rwm_C and the woRE flavour)
static int chain_setup(jmb_handle h, const jmb_priors* pr, const jmb_control* ct, const jmb_chain_in* in, bool wore) {
    CU(cudaSetDevice(h->device));
    const ModelMeta& mm = h->mm;
    const ParamLayout& pl = h->pl;
    const int n = h->n, q = mm.d.q, B = mm.B, p = mm.d.p, A = mm.d.A;
    if (ct->n_block <= 0 || ct->n_block > h->n_block_cap) return fail("jmb_chain_begin: n_block out of range (1..1024)");
    if (ct->n_iter < 0 || ct->n_burnin < 0 || ct->n_thin <= 0) return fail("jmb_chain_begin: bad control");
    if (pr->n_td < 0 || pr->n_td > JMB_MAX_TERMS) return fail("jmb_chain_begin: too many td effects");
    ChainConst& cc = h->c->cc;
    memset(&cc, 0, sizeof(cc));
    cc.n_iter = ct->n_iter; cc.n_burnin = ct->n_burnin; cc.n_block = ct->n_block; cc.n_thin = ct->n_thin;
    const int total_it = ct->n_iter + ct->n_burnin;
    const int its = (int)std::ceil((double)total_it / ct->n_block);
    cc.n_out = (int)std::floor((double)ct->n_iter / ct->n_thin);
    cc.n_keep = 0;
    for (int v = ct->n_burnin + 1; v <= total_it; v += ct->n_thin) cc.n_keep += 1;
    cc.start_cov_update = (int)std::floor((double)ct->n_burnin / ct->n_block) - 1;
    cc.adaptCov = ct->adaptCov;
    cc.target_acc = ct->target_acc; cc.eps1 = ct->eps1; cc.eps2 = ct->eps2; cc.eps3 = ct->eps3;
    cc.c0 = ct->c0; cc.mc1 = -ct->c1; cc.seed = ct->seed; cc.chain_id = ct->chain_id;
    cc.B_tau_Bs = pr->B_tau_Bs_gammas; cc.Apost_tau_Bs = pr->A_tau_Bs_gammas + 0.5 * pr->rank_Tau_Bs_gammas;
    if (h->ms) {
        if (wore) return fail("lap_rwm_C_woRE is not defined for multi-state fits (the reference reads a scalar tau_Bs_gammas there)");
        cc.nRisks = h->nRisks;
        for (int k = 0; k < h->nRisks; ++k) { cc.kn_first[k] = h->kn_first[k]; cc.kn_last[k] = h->kn_last[k]; }
    }
    cc.shrink_gammas = pr->shrink_gammas; cc.B_tau_g = pr->B_tau_gammas;
    cc.Apost_tau_g = pr->A_tau_gammas + 0.5 * pr->rank_Tau_gammas;
    cc.B_phi_g = pr->B_phi_gammas; cc.Apost_phi_g = pr->A_phi_gammas + 0.5;
    cc.shrink_alphas = pr->shrink_alphas; cc.double_gamma_alphas = pr->double_gamma_alphas;
    cc.B_tau_a = pr->B_tau_alphas; cc.Apost_tau_a = pr->A_tau_alphas + 0.5 * pr->rank_Tau_alphas;
    cc.B_phi_a = pr->B_phi_alphas; cc.Apost_phi_a = pr->A_phi_alphas + 0.5;
    cc.B_xi_a = pr->B_xi_alphas; cc.Apost_xi_a = pr->A_xi_alphas + 0.5;
    cc.B_nu_a = pr->B_nu_alphas; cc.Apost_nu_a = pr->A_nu_alphas + 0.5;
    if (wore) { cc.double_gamma_alphas = 0; }      // lap_rwm_C_woRE has neither td-effects nor the double-gamma branch
    cc.n_td = wore ? 0 : pr->n_td; cc.B_tau_td = pr->B_tau_Bs_gammas;
    cc.Apost_tau_td = pr->A_tau_Bs_gammas + 0.5 * pr->rank_Tau_td_alphas;
    for (int k = 0; k < cc.n_td; ++k) {
        cc.td_len[k] = pr->td_len[k];
        if (cc.td_len[k] > JMB_MAX_UT) return fail("jmb_chain_begin: td block too wide");
        for (int j = 0; j < cc.td_len[k]; ++j) cc.td_cols[k][j] = pr->td_cols[k][j];
    }
    h->c->n_out = cc.n_out; h->c->n_td = cc.n_td;
    h->c->total_iters = its * ct->n_block;
    {   // the reference stores a sample whenever iter == keep[k] (src/fast_LAPRWM.cpp:836-850); when that
        // happens more than n_out times it indexes past its output arrays and fails ("index out of bounds")
        int hits = 0;
        for (int v = ct->n_burnin + 1; v <= total_it && v <= h->c->total_iters - 1; v += ct->n_thin) ++hits;
        if (hits > cc.n_out)
            return fail("jmb_chain_begin: control would store more than floor(n_iter / n_thin) samples "
                        "(the reference fails with 'index out of bounds' for this n_iter / n_thin / n_block)");
    }
    h->c->iter_host = 0; h->c->keep_host = 0; h->c->check_host = 0;

    CU(cudaStreamSynchronize(h->c->stream));
    double* hp = h->c->h_par;
    memcpy(hp + pl.cur, in->Bs_gammas, sizeof(double) * B);
    if (p) memcpy(hp + pl.cur + B, in->gammas, sizeof(double) * p);
    memcpy(hp + pl.cur + B + p, in->alphas, sizeof(double) * A);
    memcpy(hp + pl.prop, hp + pl.cur, sizeof(double) * mm.P);
    for (int k = 0; k < h->nRisks; ++k) hp[pl.tau_Bs + k] = in->tau_Bs_gammas[k];     // rep(., nRisks), R/mvJointModelBayes.R:723-724
    hp[pl.tau_g] = in->tau_gammas; hp[pl.tau_a] = in->tau_alphas;
    hp[pl.xi_a] = in->tau_alphas;                                        // src/fast_LAPRWM.cpp:518
    for (int j = 0; j < p; ++j) hp[pl.phi_g + j] = in->phi_gammas[j];
    for (int j = 0; j < A; ++j) { hp[pl.phi_a + j] = in->phi_alphas[j]; hp[pl.nu_a + j] = in->phi_alphas[j]; }   // :513
    for (int j = 0; j < cc.n_td; ++j) hp[pl.tau_td + j] = in->tau_td_alphas[j];
    if (p) memcpy(hp + pl.Tau_g, pr->Tau_gammas, sizeof(double) * p * p);
    memcpy(hp + pl.Tau_a, pr->Tau_alphas, sizeof(double) * A * A);
    memcpy(hp + pl.Tau_a_pen, pr->Tau_alphas, sizeof(double) * A * A);
    hp[pl.sig] = in->sigma_Bs_gammas; hp[pl.sig + 1] = in->sigma_gammas; hp[pl.sig + 2] = in->sigma_alphas;
    memcpy(hp + pl.S_Bs, in->Sigma_Bs_gammas, sizeof(double) * B * B);
    if (p) memcpy(hp + pl.S_g, in->Sigma_gammas, sizeof(double) * p * p);
    memcpy(hp + pl.S_a, in->Sigma_alphas, sizeof(double) * A * A);
    memcpy(hp + pl.mean_Bs, pr->mean_Bs_gammas, sizeof(double) * B);
    memcpy(hp + pl.Tau_Bs, pr->Tau_Bs_gammas, sizeof(double) * B * B);
    if (p) memcpy(hp + pl.mean_g, pr->mean_gammas, sizeof(double) * p);
    memcpy(hp + pl.mean_a, pr->mean_alphas, sizeof(double) * A);
    CU(cudaMemcpyAsync(h->c->d_par, hp, sizeof(double) * pl.block_par, cudaMemcpyHostToDevice, h->c->stream));
    CU(cudaMemsetAsync(h->c->d_ctl, 0, sizeof(Ctl), h->c->stream));

    // outputs on the device (buffers are kept across chains of the same shape: cudaMalloc / cudaFree
    // synchronise the device and are slow once peer mappings exist)
    const int no = std::max(cc.n_out, 1);
    const size_t par_out = (size_t)no * (B + 2 * p + 2 * A + 3 + JMB_MAX_RISKS + JMB_MAX_TERMS);
    const size_t need_b = wore ? 1 : (size_t)n * q * no, need_tr = 2 * (size_t)std::max(h->c->total_iters, 1);
    if (need_b > h->c->cap_ob) { cudaFree(h->c->o_b); h->c->o_b = nullptr; CU(cudaMalloc(&h->c->o_b, sizeof(double) * need_b)); h->c->cap_ob = need_b; }
    if (par_out > h->c->cap_opar) { cudaFree(h->c->o_par); h->c->o_par = nullptr; CU(cudaMalloc(&h->c->o_par, sizeof(double) * par_out)); h->c->cap_opar = par_out; }
    if (need_tr > h->c->cap_otr) { cudaFree(h->c->o_trace); h->c->o_trace = nullptr; CU(cudaMalloc(&h->c->o_trace, sizeof(double) * need_tr)); h->c->cap_otr = need_tr; }
    CU(cudaMemsetAsync(h->c->o_b, 0, sizeof(double) * need_b, h->c->stream));
    CU(cudaMemsetAsync(h->c->o_par, 0, sizeof(double) * par_out, h->c->stream));
    CU(cudaMemsetAsync(h->c->o_trace, 0, sizeof(double) * need_tr, h->c->stream));
    FinalizeArgs& fa = h->c->fa;
    memset(&fa, 0, sizeof(fa));
    double* o = h->c->o_par;
    fa.o_Bs = o; o += (size_t)no * B;
    fa.o_g = o; o += (size_t)no * p;
    fa.o_a = o; o += (size_t)no * A;
    fa.o_phi_g = o; o += (size_t)no * p;
    fa.o_phi_a = o; o += (size_t)no * A;
    fa.o_tau_Bs = o; o += (size_t)no * JMB_MAX_RISKS; fa.o_tau_g = o; o += no; fa.o_tau_a = o; o += no;
    fa.o_tau_td = o; o += (size_t)no * JMB_MAX_TERMS;
    fa.o_lw = o;
    fa.trace = h->c->o_trace;

    return 0;
}

extern "C" int jmb_chain_begin(jmb_handle h, const jmb_priors* pr, const jmb_control* ct, const jmb_chain_in* in) {
    if (!h || !pr || !ct || !in) return fail("jmb_chain_begin: null argument");
    if (!h->c->draw_set) return fail("jmb_chain_begin: call jmb_set_draw first");
    if (chain_setup(h, pr, ct, in, false)) return 1;
    const ModelMeta& mm = h->mm;
    const int n = h->n, q = mm.d.q;
    // state <- initials$b, scales$b
    memcpy(h->c->h_tmp, in->b, sizeof(double) * (size_t)n * q);
    memcpy(h->c->h_tmp + (size_t)n * q, in->sigma_b, sizeof(double) * n);
    CU(cudaMemcpyAsync(h->c->d_tmp, h->c->h_tmp, sizeof(double) * (size_t)n * (q + 1), cudaMemcpyHostToDevice, h->c->stream));
    {
        const long long tot = (long long)n * mm.L.state_stride;
        jmb_load_state_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->c->stream>>>(h->c->d_state, mm.L.state_stride, n, q, h->c->d_tmp, h->c->d_tmp + (size_t)n * q);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    {   // draws of the b-step: one buffer per chunk of iterations that divides n_block.  Short chunks (<= 10 iterations by
        // default, JMB_RNG_CHUNK overrides) keep the buffer small (80 MB per 125 000 subjects at q = 6) and spread the
        // once-per-chunk kernel evenly over the iterations; the draws are keyed by (iteration, subject), so the chunk
        // length never changes them.
        const size_t per_iter = (size_t)n * h->rng_stride * sizeof(double);
        int want = 10;
        if (const char* e = getenv("JMB_RNG_CHUNK")) want = std::max(1, atoi(e));
        int chunk = std::min(ct->n_block, want);
        while (chunk > 1 && ct->n_block % chunk != 0) --chunk;
        if ((size_t)chunk * per_iter > h->c->rng_capacity) {
            // growing the buffer is the only place the chain start touches the allocator (a warm chain start does not)
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            const size_t budget = std::max<size_t>(per_iter, std::min<size_t>(free_b / 4 + h->c->rng_capacity, (size_t)8 << 30));
            while (chunk > 1 && ((size_t)chunk * per_iter > budget || ct->n_block % chunk != 0)) --chunk;
            if ((size_t)chunk * per_iter > h->c->rng_capacity) {
                cudaFree(h->c->d_rng); h->c->d_rng = nullptr;
                CU(cudaMalloc(&h->c->d_rng, (size_t)chunk * per_iter));
                h->c->rng_capacity = (size_t)chunk * per_iter;
            }
        }
        h->c->rng_chunk = chunk;
    }
    if (launch_finalize(h, 1)) return 1;            // Cholesky factors + proposal of iteration 0
    if (launch_step(h, MODE_INIT)) return 1;        // log_pyb / log_pb at the initial b (:619-621)
    h->c->chain_open = true;
    return 0;
}

static bool is_keep_iter(const jmb_handle h, int iter) {
    const ChainConst& cc = h->c->cc;
    return h->c->check_host < cc.n_keep && h->c->keep_host < cc.n_out && iter == cc.n_burnin + 1 + h->c->check_host * cc.n_thin;
}

// enqueue one iteration of the current chain slot on its stream
static int enqueue_iteration(jmb_handle h, float* kernel_ms) {
    const ModelMeta& mm = h->mm;
    const int i_blk = h->c->iter_host % h->c->cc.n_block;
    if (i_blk % h->c->rng_chunk == 0) {            // draws for the next rng_chunk iterations (:658)
        PrepArgs pa{};
        pa.Rchol = h->d_Rchol; pa.state = h->c->d_state; pa.rng = h->c->d_rng; pa.n = h->n; pa.q = mm.d.q;
        pa.state_stride = mm.L.state_stride; pa.rng_stride = h->rng_stride; pa.iter0 = h->c->iter_host; pa.count = h->c->rng_chunk;
        pa.subject_offset = h->subject_offset;
        const long long tot = (long long)h->n * h->c->rng_chunk;
        const unsigned pg = (unsigned)((tot + 255) / 256);
        if (mm.d.q == 2) jmb_block_prep_kernel<2><<<pg, 256, 0, h->c->stream>>>(h->c->cc, pa);
        else if (mm.d.q == 4) jmb_block_prep_kernel<4><<<pg, 256, 0, h->c->stream>>>(h->c->cc, pa);
        else if (mm.d.q == 6) jmb_block_prep_kernel<6><<<pg, 256, 0, h->c->stream>>>(h->c->cc, pa);
        else jmb_block_prep_kernel<0><<<pg, 256, 0, h->c->stream>>>(h->c->cc, pa);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    if (kernel_ms) CU(cudaEventRecord(h->c->evk0, h->c->stream));
    if (launch_step(h, MODE_STEP)) return 1;
    if (kernel_ms) {
        CU(cudaEventRecord(h->c->evk1, h->c->stream));
        CU(cudaEventSynchronize(h->c->evk1));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, h->c->evk0, h->c->evk1));
        *kernel_ms += t;
    }
    if (launch_finalize(h, 0)) return 1;
    if (i_blk == h->c->cc.n_block - 1) {           // block end: per-subject scale adaptation (:853-859)
        jmb_adapt_sigma_kernel<<<(h->n + 255) / 256, 256, 0, h->c->stream>>>(h->c->cc, h->c->d_state, h->n, mm.d.q, mm.L.state_stride, h->c->d_ctl);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    if (is_keep_iter(h, h->c->iter_host)) {
        const long long tot = (long long)h->n * mm.d.q;
        jmb_extract_b_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->c->stream>>>(
            h->c->d_state, mm.L.state_stride, h->n, mm.d.q, 0, h->c->o_b + (size_t)h->c->keep_host * h->n * mm.d.q);
        h->launches += 1;
        CU(cudaGetLastError());
        h->c->keep_host += 1; h->c->check_host += 1;
    }
    h->c->iter_host += 1;
    return 0;
}

static int chain_iterate_impl(jmb_handle h, int lead, int iters, float* elapsed_ms, float* step_kernel_ms);
extern "C" int jmb_chain_iterate(jmb_handle h, int iters, float* elapsed_ms, float* step_kernel_ms) {
    return chain_iterate_impl(h, 0, iters, elapsed_ms, step_kernel_ms);
}
// `lead` untimed iterations enqueued in the same call right before the timed ones: with several ranks the start event then
// follows the all-reduce of the last lead iteration on the device, so every rank's clock starts within the latency of that
// exchange of the others' - a host-side barrier leaves the ranks hundreds of microseconds apart, which a short timed window
// would otherwise count as iteration time.
extern "C" int jmb_chain_iterate_after(jmb_handle h, int lead, int iters, float* elapsed_ms) {
    return chain_iterate_impl(h, lead, iters, elapsed_ms, nullptr);
}
static int chain_iterate_impl(jmb_handle h, int lead, int iters, float* elapsed_ms, float* step_kernel_ms) {
    if (!h || !h->c->chain_open) return fail("jmb_chain_iterate: no open chain");
    if (lead < 0 || iters < 0) return fail("jmb_chain_iterate: negative iteration count");
    CU(cudaSetDevice(h->device));
    float kms = 0.f;
    for (int k = 0; k < lead; ++k)
        if (enqueue_iteration(h, nullptr)) return 1;
    CU(cudaEventRecord(h->c->ev0, h->c->stream));
    h->launches_at_ev0 = h->launches;
    for (int k = 0; k < iters; ++k)
        if (enqueue_iteration(h, step_kernel_ms ? &kms : nullptr)) return 1;
    h->launches_timed = h->launches - h->launches_at_ev0;
    CU(cudaEventRecord(h->c->ev1, h->c->stream));
    CU(cudaEventSynchronize(h->c->ev1));
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, h->c->ev0, h->c->ev1));
    if (step_kernel_ms) *step_kernel_ms = kms;
    Ctl c;
    CU(cudaMemcpy(&c, h->c->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    if (c.error == 3) return fail("jmb_chain_iterate: a peer rank did not deliver its sums (fused all-reduce timed out)");
    if (c.error) return fail("jmb_chain_iterate: proposal covariance is not positive definite");
    return 0;
}

// Several chains of the same fit in flight: slots 0..nslots-1 (each begun with jmb_chain_begin after
// jmb_select_chain) advance `iters` iterations, enqueued round-robin on their own streams so that one
// chain's survival-block bookkeeping overlaps the other chains' subject kernels.  This is the
// reference's chain parallelism (foreach over draws, R/mvJointModelBayes.R:871-905) on one device.
extern "C" int jmb_chains_iterate(jmb_handle h, int nslots, int iters, float* elapsed_ms) {
    if (!h || nslots < 1 || nslots > (int)h->slots.size()) return fail("jmb_chains_iterate: bad slot count");
    CU(cudaSetDevice(h->device));
    ChainCtx* keep = h->c;
    for (int s = 0; s < nslots; ++s)
        if (!h->slots[s]->chain_open) return fail("jmb_chains_iterate: a slot has no open chain");
    ChainCtx* first = h->slots[0];
    // common start: every stream waits for slot 0's start event
    CU(cudaEventRecord(first->ev0, first->stream));
    for (int s = 1; s < nslots; ++s) CU(cudaStreamWaitEvent(h->slots[s]->stream, first->ev0, 0));
    for (int k = 0; k < iters; ++k)
        for (int s = 0; s < nslots; ++s) {
            h->c = h->slots[s];
            if (enqueue_iteration(h, nullptr)) { h->c = keep; return 1; }
        }
    // common end: slot 0's stream waits for all the others
    for (int s = 1; s < nslots; ++s) {
        CU(cudaEventRecord(h->slots[s]->ev1, h->slots[s]->stream));
        CU(cudaStreamWaitEvent(first->stream, h->slots[s]->ev1, 0));
    }
    CU(cudaEventRecord(first->ev1, first->stream));
    CU(cudaEventSynchronize(first->ev1));
    if (elapsed_ms) CU(cudaEventElapsedTime(elapsed_ms, first->ev0, first->ev1));
    for (int s = 0; s < nslots; ++s) {
        Ctl c;
        CU(cudaMemcpy(&c, h->slots[s]->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
        if (c.error == 3) { h->c = keep; return fail("jmb_chains_iterate: a peer rank did not deliver its sums (fused all-reduce timed out)"); }
        if (c.error) { h->c = keep; return fail("jmb_chains_iterate: proposal covariance is not positive definite"); }
    }
    h->c = keep;
    return 0;
}

extern "C" int jmb_chain_end(jmb_handle h, jmb_chain_out* out) {
    if (!h || !h->c->chain_open) return fail("jmb_chain_end: no open chain");
    CU(cudaSetDevice(h->device));
    const ModelMeta& mm = h->mm;
    const ParamLayout& pl = h->pl;
    const int n = h->n, q = mm.d.q, B = mm.B, p = mm.d.p, A = mm.d.A, no = h->c->n_out;
    CU(cudaStreamSynchronize(h->c->stream));
    h->c->chain_open = false;
    if (!out) return 0;
    const FinalizeArgs& fa = h->c->fa;
    auto d2h = [&](double* dst, const double* src, size_t len) -> cudaError_t {
        if (!dst || len == 0) return cudaSuccess;
        return cudaMemcpy(dst, src, sizeof(double) * len, cudaMemcpyDeviceToHost);
    };
    CU(d2h(out->b, h->c->o_b, (size_t)n * q * no));
    CU(d2h(out->Bs_gammas, fa.o_Bs, (size_t)B * no));
    CU(d2h(out->gammas, fa.o_g, (size_t)p * no));
    CU(d2h(out->alphas, fa.o_a, (size_t)A * no));
    CU(d2h(out->phi_gammas, fa.o_phi_g, (size_t)p * no));
    CU(d2h(out->phi_alphas, fa.o_phi_a, (size_t)A * no));
    CU(d2h(out->tau_Bs_gammas, fa.o_tau_Bs, (size_t)h->nRisks * no));
    CU(d2h(out->tau_gammas, fa.o_tau_g, no));
    CU(d2h(out->tau_alphas, fa.o_tau_a, no));
    CU(d2h(out->tau_td_alphas, fa.o_tau_td, (size_t)h->c->n_td * no));
    CU(d2h(out->logWeights, fa.o_lw, no));
    CU(d2h(out->sigma_surv, h->c->d_par + pl.sig, 3));
    CU(d2h(out->Sigma_Bs_gammas, h->c->d_par + pl.S_Bs, (size_t)B * B));
    CU(d2h(out->Sigma_gammas, h->c->d_par + pl.S_g, (size_t)p * p));
    CU(d2h(out->Sigma_alphas, h->c->d_par + pl.S_a, (size_t)A * A));
    CU(d2h(out->trace_surv, h->c->o_trace, 2 * (size_t)h->c->iter_host));
    if (out->sigma_b || out->accept_b) {
        // columns q (sigma_b) and q+4 (total accepts) of the state
        jmb_extract_b_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->c->stream>>>(h->c->d_state, mm.L.state_stride, n, 1, q, h->c->d_tmp);
        jmb_extract_b_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->c->stream>>>(h->c->d_state, mm.L.state_stride, n, 1, q + 4, h->c->d_tmp + n);
        h->launches += 2;
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->c->stream));
        CU(d2h(out->sigma_b, h->c->d_tmp, n));
        if (out->accept_b) {
            CU(d2h(out->accept_b, h->c->d_tmp + n, n));
            for (int i = 0; i < n; ++i) out->accept_b[i] /= std::max(h->c->iter_host, 1);
        }
    }
    if (out->accept_surv) {
        Ctl c;
        CU(cudaMemcpy(&c, h->c->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
        out->accept_surv[0] = (double)c.accept_s_total / std::max(h->c->iter_host, 1);
    }
    return 0;
}

extern "C" int jmb_lap_rwm(jmb_handle h, const jmb_priors* pr, const jmb_control* ct, const jmb_chain_in* in, jmb_chain_out* out) {
    if (jmb_chain_begin(h, pr, ct, in)) return 1;
    if (jmb_chain_iterate(h, h->c->total_iters, nullptr, nullptr)) return 1;
    return jmb_chain_end(h, out);
}

// ---------------------------------------------------------------------------------------------
// parity probe
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_eval_logpost_re(jmb_handle h, const double* b, const double* Bs, const double* gam,
                                   const double* alp, jmb_eval_out* out) {
    if (!h || !b || !out) return fail("jmb_eval_logpost_re: null argument");
    if (!h->c->draw_set) return fail("jmb_eval_logpost_re: call jmb_set_draw first");
    if (h->c->chain_open) return fail("jmb_eval_logpost_re: a chain is open on this handle");
    CU(cudaSetDevice(h->device));
    const ModelMeta& mm = h->mm;
    const ParamLayout& pl = h->pl;
    const int n = h->n, q = mm.d.q, B = mm.B, p = mm.d.p, A = mm.d.A;
    CU(cudaStreamSynchronize(h->c->stream));
    double* hp = h->c->h_par;
    memcpy(hp + pl.cur, Bs, sizeof(double) * B);
    if (p) memcpy(hp + pl.cur + B, gam, sizeof(double) * p);
    memcpy(hp + pl.cur + B + p, alp, sizeof(double) * A);
    memcpy(hp + pl.prop, hp + pl.cur, sizeof(double) * mm.P);
    CU(cudaMemcpyAsync(h->c->d_par + pl.cur, hp + pl.cur, sizeof(double) * 2 * mm.P, cudaMemcpyHostToDevice, h->c->stream));
    CU(cudaMemsetAsync(h->c->d_ctl, 0, sizeof(Ctl), h->c->stream));
    memcpy(h->c->h_tmp, b, sizeof(double) * (size_t)n * q);
    for (int i = 0; i < n; ++i) h->c->h_tmp[(size_t)n * q + i] = 1.0;
    CU(cudaMemcpyAsync(h->c->d_tmp, h->c->h_tmp, sizeof(double) * (size_t)n * (q + 1), cudaMemcpyHostToDevice, h->c->stream));
    const long long tot = (long long)n * mm.L.state_stride;
    jmb_load_state_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->c->stream>>>(h->c->d_state, mm.L.state_stride, n, q, h->c->d_tmp, h->c->d_tmp + (size_t)n * q);
    h->launches += 1;
    CU(cudaGetLastError());
    memset(&h->c->cc, 0, sizeof(h->c->cc));
    h->c->cc.n_block = 1;
    if (launch_step(h, MODE_EVAL)) return 1;
    CU(cudaStreamSynchronize(h->c->stream));
    double* dst[6] = {out->log_pyb, out->log_pb, out->log_h, out->H, out->HL, out->log_ptb};
    const size_t es = (size_t)std::max(n, h->nT);       // log_h / H / HL / log_ptb: one value per transition row
    for (int k = 0; k < 6; ++k)
        if (dst[k]) CU(cudaMemcpy(dst[k], h->d_eval + (size_t)k * es, sizeof(double) * (k < 2 ? n : h->nT), cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// survival-block Laplace step: logPosterior / gradient_logPosterior on fixed Wlong / Wlongs
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_set_wlong(jmb_handle h, const double* Wlong, const double* Wlongs) {
    if (!h || !Wlong || !Wlongs) return fail("jmb_set_wlong: null argument");
    CU(cudaSetDevice(h->device));
    // rows of Wlong: one per subject, or one per transition row of a multi-state fit (R/mvJointModelBayes.R:559-565)
    const int n = h->ms ? h->nT : h->n, A = h->mm.d.A, K = h->mm.d.K, B = h->mm.B, p = h->mm.d.p;
    const size_t ns = (size_t)n * K;
    if (!h->d_Wlong) {
        CU(cudaMalloc(&h->d_Wlong, sizeof(double) * (size_t)n * A));
        CU(cudaMalloc(&h->d_Wlongs, sizeof(double) * ns * A));
        CU(cudaMalloc(&h->d_sp_partials, sizeof(double) * (size_t)h->grid * (h->mm.P + 2)));
        CU(cudaMalloc(&h->d_sp_out, sizeof(double) * (h->mm.P + 2)));
        CU(cudaMalloc(&h->d_theta, sizeof(double) * h->mm.P));
    }
    CU(cudaMemcpy(h->d_Wlong, Wlong, sizeof(double) * (size_t)n * A, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_Wlongs, Wlongs, sizeof(double) * ns * A, cudaMemcpyHostToDevice));
    for (int c = 0; c < A; ++c) {      // event_colSumsWlong (R/mvJointModelBayes.R:681)
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += h->h_event[i] * Wlong[i + (size_t)c * n];
        h->ecs[B + p + c] = t;
    }
    h->wlong_set = true;
    h->sp_last_theta.clear();
    return 0;
}

// data sums [sum event*log_h, sum Pw exp(lin), score sums...] at theta; cached for the last theta
// enqueue the data sums of the survival-block log-posterior at a device-resident theta
static int launch_surv_post(jmb_handle h, const double* d_theta, int value_only) {
    const ModelMeta& mm = h->mm;
    const int NA = mm.P + 2;
    SurvPostArgs a{};
    a.drec = h->d_drec; a.doff = h->d_doff; a.Wlong = h->d_Wlong; a.Wlongs = h->d_Wlongs; a.theta = d_theta;
    a.partials = h->d_sp_partials; a.n = h->n; a.subjects_per_cta = h->subjects_per_cta; a.value_only = value_only;
    a.trow = h->d_trow; a.nT = h->nT; a.ms = h->ms ? 1 : 0;
    const int groups = 256 / mm.d.G;
    const int smem = (int)sizeof(double) * (mm.P + groups * NA);
    if (mm.d.G == 16) jmb_surv_post_kernel<16><<<h->grid, 256, smem, h->c->stream>>>(mm, a);
    else jmb_surv_post_kernel<32><<<h->grid, 256, smem, h->c->stream>>>(mm, a);
    jmb_colsum_kernel<<<value_only ? 2 : NA, 128, 0, h->c->stream>>>(h->d_sp_partials, h->grid, NA, h->d_sp_out);
    h->launches += 2;
    CU(cudaGetLastError());
    if (h->nranks > 1) {
        ncclResult_t rc = g_nccl.AllReduce(h->d_sp_out, h->d_sp_out, value_only ? 2 : NA, ncclDouble, ncclSum, h->comm, h->c->stream);
        if (rc != ncclSuccess) return fail("ncclAllReduce failed in the survival-block sums");
    }
    return 0;
}

static int surv_post_sums(jmb_handle h, const double* Bs, const double* gam, const double* alp, std::vector<double>& out) {
    if (!h->wlong_set) return fail("jmb_logPosterior: call jmb_set_wlong first");
    const ModelMeta& mm = h->mm;
    const int B = mm.B, p = mm.d.p, A = mm.d.A, P = mm.P, NA = P + 2;
    std::vector<double> theta(P);
    memcpy(theta.data(), Bs, sizeof(double) * B);
    if (p) memcpy(theta.data() + B, gam, sizeof(double) * p);
    memcpy(theta.data() + B + p, alp, sizeof(double) * A);
    if (h->sp_last_theta.size() == theta.size() && !memcmp(h->sp_last_theta.data(), theta.data(), sizeof(double) * P)) {
        out = h->sp_last_out;
        return 0;
    }
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->d_theta, theta.data(), sizeof(double) * P, cudaMemcpyHostToDevice, h->c->stream));
    if (launch_surv_post(h, h->d_theta, 0)) return 1;
    out.assign(NA, 0.0);
    CU(cudaMemcpyAsync(out.data(), h->d_sp_out, sizeof(double) * NA, cudaMemcpyDeviceToHost, h->c->stream));
    CU(cudaStreamSynchronize(h->c->stream));
    h->sp_last_theta = theta;
    h->sp_last_out = out;
    return 0;
}

static double host_quad(const double* x, const double* mean, const double* Tau, int m) {
    double s = 0.0;
    for (int j = 0; j < m; ++j) {
        double t = 0.0;
        for (int l = 0; l < m; ++l) t += (x[l] - mean[l]) * Tau[l + j * m];
        s += t * (x[j] - mean[j]);
    }
    return s;
}

extern "C" int jmb_logPosterior(jmb_handle h, double temp, const double* Bs, const double* gam, const double* alp,
                                const jmb_priors* pr, double tauBs, double A_tauBs, double B_tauBs, double* out) {
    if (!h || !Bs || !alp || !pr || !out) return fail("jmb_logPosterior: null argument");
    std::vector<double> s;
    if (surv_post_sums(h, Bs, gam, alp, s)) return 1;
    const int B = h->mm.B, p = h->mm.d.p, A = h->mm.d.A;
    double v = temp * (s[0] - s[1]) - 0.5 * tauBs * host_quad(Bs, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, B);
    if (p) v -= 0.5 * host_quad(gam, pr->mean_gammas, pr->Tau_gammas, p);
    v -= 0.5 * host_quad(alp, pr->mean_alphas, pr->Tau_alphas, A);
    v += (A_tauBs - 1.0) * std::log(tauBs) - B_tauBs * tauBs;
    *out = v;
    return 0;
}

extern "C" int jmb_gradient_logPosterior(jmb_handle h, double temp, const double* Bs, const double* gam, const double* alp,
                                         const jmb_priors* pr, double tauBs, double A_tauBs, double B_tauBs, double* out) {
    if (!h || !Bs || !alp || !pr || !out) return fail("jmb_gradient_logPosterior: null argument");
    std::vector<double> s;
    if (surv_post_sums(h, Bs, gam, alp, s)) return 1;
    const int B = h->mm.B, p = h->mm.d.p, A = h->mm.d.A;
    std::vector<double> ecs = h->ecs;
    if (h->nranks > 1) {   // event_colSums over the whole cohort
        CU(cudaMemcpyAsync(h->d_sp_out, ecs.data(), sizeof(double) * (B + p + A), cudaMemcpyHostToDevice, h->c->stream));
        if (g_nccl.AllReduce(h->d_sp_out, h->d_sp_out, B + p + A, ncclDouble, ncclSum, h->comm, h->c->stream) != ncclSuccess)
            return fail("ncclAllReduce failed in jmb_gradient_logPosterior");
        CU(cudaMemcpyAsync(ecs.data(), h->d_sp_out, sizeof(double) * (B + p + A), cudaMemcpyDeviceToHost, h->c->stream));
        CU(cudaStreamSynchronize(h->c->stream));
        h->sp_last_theta.clear();
    }
    const double* par[3] = {Bs, gam, alp};
    const double* mean[3] = {pr->mean_Bs_gammas, pr->mean_gammas, pr->mean_alphas};
    const double* Tau[3] = {pr->Tau_Bs_gammas, pr->Tau_gammas, pr->Tau_alphas};
    const int dims[3] = {B, p, A};
    const double mult[3] = {tauBs, 1.0, 1.0};
    int o = 0;
    for (int blk = 0; blk < 3; ++blk) {
        const int m = dims[blk];
        for (int i = 0; i < m; ++i) {
            double pen = 0.0;
            for (int l = 0; l < m; ++l) pen += Tau[blk][i + l * m] * (par[blk][l] - mean[blk][l]);
            out[o + i] = temp * (ecs[o + i] - s[2 + o + i]) - mult[blk] * pen;     // src/fast_score.cpp:21-41
        }
        o += m;
    }
    out[o] = tauBs * (-0.5 * host_quad(Bs, pr->mean_Bs_gammas, pr->Tau_Bs_gammas, B) + (A_tauBs - 1.0) / tauBs - B_tauBs);  // :44
    return 0;
}

// lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (src/fast_LAPRWM.cpp:934-1297): the survival block alone on
// the fixed association designs of jmb_set_wlong (control$update_RE = FALSE)
extern "C" int jmb_lap_rwm_woRE(jmb_handle h, const jmb_priors* pr, const jmb_control* ct, const jmb_chain_in* in,
                                jmb_chain_out* out, double* logPost) {
    if (!h || !pr || !ct || !in || !out) return fail("jmb_lap_rwm_woRE: null argument");
    if (!h->wlong_set) return fail("jmb_lap_rwm_woRE: call jmb_set_wlong first");
    if (h->c->chain_open) return fail("jmb_lap_rwm_woRE: a chain is open on this slot");
    if (chain_setup(h, pr, ct, in, true)) return 1;
    const ModelMeta& mm = h->mm;
    const ParamLayout& pl = h->pl;
    const int B = mm.B, p = mm.d.p, A = mm.d.A, no = h->c->n_out;
    h->c->wore = true;
    // current_logpost at the initial values (:1024-1028), Cholesky factors, first proposal
    if (launch_surv_post(h, h->c->d_par + pl.cur, 1)) { h->c->wore = false; return 1; }
    if (launch_finalize(h, 1)) { h->c->wore = false; return 1; }
    for (int k = 0; k < h->c->total_iters; ++k) {
        if (launch_surv_post(h, h->c->d_par + pl.prop, 1) || launch_finalize(h, 0)) { h->c->wore = false; return 1; }
        h->c->iter_host += 1;
    }
    h->c->wore = false;
    CU(cudaStreamSynchronize(h->c->stream));
    Ctl c;
    CU(cudaMemcpy(&c, h->c->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost));
    if (c.error) return fail("jmb_lap_rwm_woRE: proposal covariance is not positive definite");
    const FinalizeArgs& fa = h->c->fa;
    auto d2h = [&](double* dst, const double* src, size_t len) -> cudaError_t {
        if (!dst || len == 0) return cudaSuccess;
        return cudaMemcpy(dst, src, sizeof(double) * len, cudaMemcpyDeviceToHost);
    };
    CU(d2h(out->Bs_gammas, fa.o_Bs, (size_t)B * no)); CU(d2h(out->gammas, fa.o_g, (size_t)p * no));
    CU(d2h(out->alphas, fa.o_a, (size_t)A * no)); CU(d2h(out->phi_gammas, fa.o_phi_g, (size_t)p * no));
    CU(d2h(out->phi_alphas, fa.o_phi_a, (size_t)A * no)); CU(d2h(out->tau_Bs_gammas, fa.o_tau_Bs, no));
    CU(d2h(out->tau_gammas, fa.o_tau_g, no)); CU(d2h(out->tau_alphas, fa.o_tau_a, no));
    CU(d2h(logPost, fa.o_lw, no));
    CU(d2h(out->sigma_surv, h->c->d_par + pl.sig, 3));
    CU(d2h(out->Sigma_Bs_gammas, h->c->d_par + pl.S_Bs, (size_t)B * B));
    CU(d2h(out->Sigma_gammas, h->c->d_par + pl.S_g, (size_t)p * p));
    CU(d2h(out->Sigma_alphas, h->c->d_par + pl.S_a, (size_t)A * A));
    CU(d2h(out->trace_surv, h->c->o_trace, 2 * (size_t)h->c->iter_host));
    if (out->accept_surv) out->accept_surv[0] = (double)c.accept_s_total / std::max(h->c->iter_host, 1);
    return 0;
}

// log_post_RE_svft (src/survfitJM_mvJMbayes.cpp:190-238; core :140-166): log p(y_i|b) + log p(b) - H_i per
// subject, and survPred_svft_2 (:241-271; core :168-187): -H_i, on the handle's quadrature designs
static int svft_eval(jmb_handle h, const double* b, const double* Bs, const double* gam, const double* alp,
                     double* log_post, double* surv_pred) {
    const int n = h->n;
    if (h->ms) return fail("survfitJM evaluators are not available on a multi-state handle");
    std::vector<double> pyb(n), pb(n), H(n);
    jmb_eval_out eo{};
    eo.log_pyb = pyb.data(); eo.log_pb = pb.data(); eo.H = H.data();
    if (jmb_eval_logpost_re(h, b, Bs, gam, alp, &eo)) return 1;
    for (int i = 0; i < n; ++i) {
        if (log_post) log_post[i] = (pyb[i] + (0.0 - H[i])) + pb[i];      // log_pyb + log_ptb + log_pb, :163-165
        if (surv_pred) surv_pred[i] = 0.0 - H[i];
    }
    return 0;
}

extern "C" int jmb_log_post_RE_svft(jmb_handle h, const double* b, const double* Bs, const double* gam, const double* alp, double* out) {
    if (!h || !b || !out) return fail("jmb_log_post_RE_svft: null argument");
    return svft_eval(h, b, Bs, gam, alp, out, nullptr);
}

extern "C" int jmb_survPred_svft_2(jmb_handle h, const double* b, const double* Bs, const double* gam, const double* alp, double* out) {
    if (!h || !b || !out) return fail("jmb_survPred_svft_2: null argument");
    return svft_eval(h, b, Bs, gam, alp, nullptr, out);
}

// ---------------------------------------------------------------------------------------------
// layout inspection without a device (CPU tests of the packer; never part of a compute path)
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_debug_pack_host(const jmb_data* d, jmb_handle* out) { return create_impl(d, 0, out, true); }

extern "C" int jmb_debug_layout(jmb_handle h, int32_t* ints /* [64] */, int64_t* doff /* [n+1] */, int64_t* coff /* [n+1] */) {
    if (!h || !ints) return fail("jmb_debug_layout: null argument");
    const RecordLayout& L = h->mm.L; const ModelDesc& md = h->mm.d;
    int k = 0;
    ints[k++] = L.RP; ints[k++] = L.o_V; ints[k++] = L.o_W2c; ints[k++] = L.o_Pw; ints[k++] = L.o_W1off; ints[k++] = L.o_W1v;
    ints[k++] = L.o_W2; ints[k++] = L.o_long; ints[k++] = L.c_long; ints[k++] = L.HB; ints[k++] = L.state_stride;
    ints[k++] = md.WB; ints[k++] = md.G; ints[k++] = md.w2_const; ints[k++] = md.MS; ints[k++] = md.S;       // k == 16
    for (int t = 0; t < kMaxT; ++t) { ints[16 + t] = L.o_ZZ[t]; ints[32 + t] = L.o_U[t]; }
    for (int t = 0; t < kMaxT; ++t) ints[48 + t] = (t < md.T) ? (md.zz_ones0[t] | (md.u_ones[t] << 1)) : 0;
    for (int o = 0; o < md.O && o < 8; ++o) ints[48 + 8 + o] |= (md.z_ones0[o] << 8) | (L.long_cols[o] << 16);
    if (doff) for (int i = 0; i <= h->n; ++i) doff[i] = h->doff[i];
    if (coff) for (int i = 0; i <= h->n; ++i) coff[i] = h->coff[i];
    return 0;
}

extern "C" int jmb_debug_records(jmb_handle h, const jmb_draw* w, double* drec /* [doff[n]] */, double* crec /* [coff[n]] */) {
    if (!h || !h->host_only) return fail("jmb_debug_records: needs a handle made by jmb_debug_pack_host");
    if (drec) memcpy(drec, h->h_drec.data(), sizeof(double) * h->h_drec.size());
    if (crec && w) pack_draw(h, w, crec);
    return 0;
}

extern "C" int jmb_get_row_ptr(jmb_handle h, int outcome, int64_t* row_ptr) {
    if (!h || outcome < 0 || outcome >= h->mm.d.O) return fail("jmb_get_row_ptr: bad argument");
    std::vector<int> rp(h->n + 1);
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpy(rp.data(), h->d_rowptr + (size_t)outcome * (h->n + 1), sizeof(int) * (h->n + 1), cudaMemcpyDeviceToHost));
    for (int i = 0; i <= h->n; ++i) row_ptr[i] = rp[i];
    return 0;
}

extern "C" int jmb_layout_info(jmb_handle h, double* shared_bytes, double* chain_bytes, int32_t* w1_band, int32_t* group_lanes) {
    if (!h) return fail("jmb_layout_info: null handle");
    const double n = h->n;
    if (shared_bytes) *shared_bytes = (8.0 * h->doff[h->n] + 8.0 * 2 * (h->n + 1) + 4.0 * h->mm.d.O * (h->n + 1)) / n;
    // per-draw record + state read/write + b-step draws (written once per block, read once)
    if (chain_bytes) *chain_bytes = 8.0 * h->coff[h->n] / n + 2.0 * 8.0 * h->mm.L.state_stride + 2.0 * 8.0 * h->rng_stride;
    if (w1_band) *w1_band = h->mm.d.WB;
    if (group_lanes) *group_lanes = h->mm.d.G;
    return 0;
}

extern "C" int jmb_param_dims(jmb_handle h, int32_t* n_Bs, int32_t* n_gammas, int32_t* n_alphas) {
    if (!h) return fail("jmb_param_dims: null handle");
    if (n_Bs) *n_Bs = h->mm.B;
    if (n_gammas) *n_gammas = h->mm.d.p;
    if (n_alphas) *n_alphas = h->mm.d.A;
    return 0;
}

extern "C" int jmb_launch_count_timed(jmb_handle h, int64_t* launches) {
    if (!h || !launches) return fail("jmb_launch_count_timed: null argument");
    *launches = h->launches_timed;
    return 0;
}

extern "C" int jmb_launch_count(jmb_handle h, int64_t* launches) {
    if (!h) return fail("jmb_launch_count: null handle");
    *launches = h->launches;
    return 0;
}

extern "C" const char* jmb_kernel_name(jmb_handle h) {
    if (!h) return "";
    return h->kernel_label.c_str();
}

extern "C" void* jmb_stream(jmb_handle h) { return h ? (void*)h->c->stream : nullptr; }

// ---------------------------------------------------------------------------------------------
// multi-GPU
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_comm_unique_id(char id[128]) {
    if (load_nccl()) return 1;
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&uid) != ncclSuccess) return fail("ncclGetUniqueId failed");
    memcpy(id, &uid, 128);
    return 0;
}

extern "C" int jmb_comm_init(jmb_handle h, int nranks, int rank, const char id[128]) {
    if (!h) return fail("jmb_comm_init: null handle");
    if (nranks <= 1) { h->nranks = 1; h->rank = 0; return 0; }
    if (load_nccl()) return 1;
    CU(cudaSetDevice(h->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    ncclResult_t rc = g_nccl.CommInitRank(&h->comm, nranks, uid, rank);
    if (rc != ncclSuccess) return fail(std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "error"));
    h->nranks = nranks; h->rank = rank;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Fused all-reduce over NVLink peer memory.  jmb_p2p_export allocates this rank's mailbox and
// returns its CUDA IPC handle (64 bytes); the host language gathers the handles of all ranks
// (torch.distributed.all_gather) and passes them to jmb_p2p_attach, after which the finalize kernel
// exchanges the per-iteration sums itself (stores into every peer's mailbox, waits for the peers'
// stores) and NCCL is no longer on the per-iteration path.
// ---------------------------------------------------------------------------------------------
extern "C" int jmb_p2p_export(jmb_handle h, int nranks, char handle[64]) {
    if (!h || nranks < 2 || nranks > 64) return fail("jmb_p2p_export: bad argument");
    CU(cudaSetDevice(h->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!h->d_mbox) {
        const size_t bytes = sizeof(double) * jmb_handle_s::kMboxSlots * 2 * (size_t)nranks * kMboxEntry;
        CU(cudaMalloc(&h->d_mbox, bytes));
        CU(cudaMemset(h->d_mbox, 0, bytes));
        CU(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t ih;
    CU(cudaIpcGetMemHandle(&ih, h->d_mbox));
    memcpy(handle, &ih, 64);
    return 0;
}

extern "C" int jmb_p2p_attach(jmb_handle h, int nranks, int rank, const char* handles) {
    if (!h || !handles || nranks < 2 || rank < 0 || rank >= nranks) return fail("jmb_p2p_attach: bad argument");
    if (!h->d_mbox) return fail("jmb_p2p_attach: call jmb_p2p_export first");
    if ((int)h->slots.size() > jmb_handle_s::kMboxSlots) return fail("jmb_p2p_attach: too many chain slots");
    CU(cudaSetDevice(h->device));
    h->peer_ptrs.assign(nranks, nullptr);
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { h->peer_ptrs[r] = h->d_mbox; continue; }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, handles + (size_t)r * 64, 64);
        CU(cudaIpcOpenMemHandle(&h->peer_ptrs[r], ih, cudaIpcMemLazyEnablePeerAccess));
    }
    CU(cudaMalloc(&h->d_peer_mbox, sizeof(double*) * nranks));
    CU(cudaMemcpy(h->d_peer_mbox, h->peer_ptrs.data(), sizeof(double*) * nranks, cudaMemcpyHostToDevice));
    h->nranks = nranks; h->rank = rank; h->p2p = true;
    const char* env = getenv("JMB_ALLREDUCE");
    if (env && !strcmp(env, "nccl")) h->p2p = false;     // A/B: keep NCCL on the per-iteration path
    return 0;
}
