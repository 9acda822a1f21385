// jmb_survfit.cu - host runtime + C ABI (include/jmb200_survfit.h) of the survfitJM.mvJMbayes() loop.
// Product code: never includes or links anything under oracle/; fails loudly without a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jmb_host.h"
#include "jmb_survfit.cuh"

using namespace jmb;

static int fail(const std::string& m) { return jmb_set_error(m); }
#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

template <typename F>
static void parallel_for(long long n, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    int nt = (int)std::max(1u, std::min(hw ? hw : 1u, 32u));
    if (n < 4096) nt = 1;
    std::vector<std::thread> th;
    long long chunk = (n + nt - 1) / nt;
    for (int t = 0; t < nt; ++t) {
        long long lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& x : th) x.join();
}

struct jmb_svft_handle_s {
    int device = 0, n = 0, B = 0, L = 0, ps_len = 0, num_sms = 148;
    long long subject_offset = 0;
    SvMeta mm{};
    std::vector<long long> off;                     // host record offsets
    double* d_rec = nullptr; long long* d_off = nullptr;
    double *d_modes = nullptr, *d_hess = nullptr, *d_S = nullptr, *d_acc = nullptr, *d_blast = nullptr, *d_sum = nullptr;
    double *d_par_mean = nullptr, *d_par_draws = nullptr, *d_cbuf = nullptr;
    size_t cap_cbuf = 0;
    bool split_horizons = false;                    // linear model with a precompiled shape: jmb_svft_horizons does the intervals
    int *d_conv = nullptr, *d_counts = nullptr, *d_work = nullptr;
    size_t cap_S = 0, cap_draws = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    float last_ms = 0.f;
    long long launches = 0;
    int kernel = 0;                                 // 0 generic, 1 three outcomes, 2 value gaussian
    int grid[2] = {0, 0}, smem = 0;                 // launch shape of the mode-search and of the sampling kernel
};

extern "C" int jmb_svft_destroy(jmb_svft_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->d_rec); cudaFree(h->d_off); cudaFree(h->d_modes); cudaFree(h->d_hess); cudaFree(h->d_S);
    cudaFree(h->d_acc); cudaFree(h->d_blast); cudaFree(h->d_sum); cudaFree(h->d_par_mean); cudaFree(h->d_par_draws); cudaFree(h->d_cbuf);
    cudaFree(h->d_conv); cudaFree(h->d_counts); cudaFree(h->d_work);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// jmb_svft_create: validate, decide the compaction flags from the data, pack one record per subject
// ---------------------------------------------------------------------------------------------------
extern "C" int jmb_svft_create(const jmb_svft_data* d, int device, jmb_svft_handle* out) {
    if (!d || !out) return fail("jmb_svft_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("jmb_svft_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail("jmb_svft_create: bad device ordinal");
    const int n = d->n, O = d->n_outcomes, q = d->q, T = d->n_terms, B = d->n_Bs, p = d->n_gammas, K = d->K, L = d->L, A = d->n_alphas;
    if (n <= 0 || O <= 0 || O > JMB_MAX_OUTCOMES) return fail("jmb_svft_create: bad n / n_outcomes");
    if (q <= 0 || q > JMB_SVFT_MAX_Q) return fail("jmb_svft_create: q must be in 1..8");
    if (T <= 0 || T > JMB_MAX_TERMS) return fail("jmb_svft_create: n_terms out of range");
    if (B <= 0 || B > JMB_MAX_BS || p < 0 || p > JMB_MAX_GAMMAS || A <= 0 || A > JMB_MAX_ALPHAS)
        return fail("jmb_svft_create: parameter block sizes out of range");
    if (K <= 0 || K > kSvG) return fail("jmb_svft_create: K must be in 1..16");
    if (L < 0 || L > 4096) return fail("jmb_svft_create: L out of range");

    auto* h = new jmb_svft_handle_s();
    h->device = device; h->n = n; h->B = B; h->L = L; h->subject_offset = d->subject_offset;
    SvDesc& md = h->mm.d;
    memset(&h->mm, 0, sizeof(h->mm));
    md.O = O; md.q = q; md.T = T; md.A = A; md.p = p; md.K = K;
#define BAD(msg) do { delete h; return fail(msg); } while (0)
    int nb = 0;
    for (int k = 0; k < O; ++k) {
        md.fam[k] = d->fams[k]; md.link[k] = d->links[k]; md.qk[k] = d->q_k[k]; md.pk[k] = d->p_k[k];
        if (md.qk[k] < 0 || md.qk[k] > JMB_SVFT_MAX_Q) BAD("jmb_svft_create: q_k out of range");
        if (md.pk[k] <= 0 || md.pk[k] > JMB_SVFT_MAX_PK) BAD("jmb_svft_create: p_k out of range (1..8)");
        md.boff[k] = nb; nb += md.pk[k];
        for (int j = 0; j < md.qk[k]; ++j) {
            md.re_inds[k][j] = d->RE_inds[k][j];
            if (md.re_inds[k][j] < 0 || md.re_inds[k][j] >= q) BAD("jmb_svft_create: RE_inds out of range");
        }
    }
    md.NB = nb;
    for (int t = 0; t < T; ++t) {
        md.rt[t] = d->r_t[t]; md.ut[t] = d->u_t[t]; md.tf[t] = d->trans_Funs[t]; md.ft[t] = d->f_t[t]; md.outc[t] = d->outcome[t];
        if (md.rt[t] < 0 || md.rt[t] > JMB_MAX_RT || md.ut[t] <= 0 || md.ut[t] > JMB_MAX_UT || md.ft[t] < 0 || md.ft[t] > JMB_SVFT_MAX_FT)
            BAD("jmb_svft_create: term sizes out of range");
        if (md.outc[t] < 0 || md.outc[t] >= O) BAD("jmb_svft_create: outcome[t] out of range");
        for (int j = 0; j < md.rt[t]; ++j) {
            md.re2[t][j] = d->RE_inds2[t][j];
            if (md.re2[t][j] < 0 || md.re2[t][j] >= q) BAD("jmb_svft_create: RE_inds2 out of range");
        }
        for (int u = 0; u < md.ut[t]; ++u) {
            md.col[t][u] = d->col_inds[t][u];
            if (md.col[t][u] < 0 || md.col[t][u] >= A) BAD("jmb_svft_create: col_inds out of range");
        }
        for (int f = 0; f < md.ft[t]; ++f) {
            md.indf[t][f] = d->indFixed[t][f];
            if (md.indf[t][f] < 0 || md.indf[t][f] >= md.pk[md.outc[t]]) BAD("jmb_svft_create: indFixed out of range");
        }
    }
    const long long nK = (long long)n * K, nH = (long long)n * L * K;
    const bool from_knots = (d->W1s == nullptr);
    if (from_knots) {
        if (!d->knots || !d->st || (nH && !d->st_h)) BAD("jmb_svft_create: W1s is NULL, so knots / st / st_h are required");
        if (d->ord < 1 || d->ord > 8 || d->n_knots != B + d->ord) BAD("jmb_svft_create: need 1 <= ord <= 8 and n_knots = n_Bs + ord");
        for (int j = 1; j < d->n_knots; ++j) if (d->knots[j] < d->knots[j - 1]) BAD("jmb_svft_create: knots must be non-decreasing");
    } else if (nH && !d->W1s_h) BAD("jmb_svft_create: W1s_h is required with a dense W1s");

    // ---- segment boundaries of the longitudinal outcomes (rows grouped by subject; empty segments allowed) ----
    std::vector<std::vector<int>> rowptr(O);
    for (int k = 0; k < O; ++k) {
        const long long Nk = d->N[k];
        if (Nk < 0 || Nk > 2147483000LL) BAD("jmb_svft_create: N_k out of range");
        rowptr[k].assign(n + 1, 0);
        int prev = -1;
        for (long long r = 0; r < Nk; ++r) {
            const int s = d->idL[k][r];
            if (s < 0 || s >= n || s < prev) BAD("jmb_svft_create: idL must be 0-based and grouped by subject");
            prev = s;
            rowptr[k][s + 1] += 1;
        }
        for (int i = 0; i < n; ++i) rowptr[k][i + 1] += rowptr[k][i];
    }
    for (int i = 0; i < n; ++i)
        if (d->n_horizons[i] < 0 || d->n_horizons[i] > L) BAD("jmb_svft_create: n_horizons out of range");

    // rows of the horizon designs that are actually read (l < n_horizons[i])
    auto live_h = [&](long long r) { const long long il = r / K; return (int)(il % std::max(L, 1)) < d->n_horizons[il / std::max(L, 1)]; };

    // ---- band of the B-spline rows ----
    auto band_of = [&](const double* W, long long rows, long long r, int& first, int& last) {
        first = B; last = -1;
        for (int j = 0; j < B; ++j)
            if (W[r + (long long)j * rows] != 0.0) { if (first == B) first = j; last = j; }
    };
    int WB = 1;
    if (from_knots) WB = std::min(d->ord, B);
    else {
        std::atomic<int> wmax{1};
        auto scan = [&](const double* W, long long rows, bool horizon) {
            parallel_for(rows, [&, W, rows, horizon](long long lo, long long hi) {
                int m = 1;
                for (long long r = lo; r < hi; ++r) {
                    if (horizon && !live_h(r)) continue;
                    int f, l; band_of(W, rows, r, f, l);
                    if (l >= f) m = std::max(m, l - f + 1);
                }
                int cur = wmax.load();
                while (m > cur && !wmax.compare_exchange_weak(cur, m)) {}
            });
        };
        scan(d->W1s, nK, false);
        if (nH) scan(d->W1s_h, nH, true);
        WB = wmax.load();
        if (WB > 8) WB = B;
        WB = std::min(WB, B);
    }
    md.WB = WB;

    // ---- constant columns that need not be stored ----
    auto all_ones = [&](const double* col, long long rows, bool horizon) {
        std::atomic<int> ok{1};
        parallel_for(rows, [&, col, horizon](long long lo, long long hi) {
            for (long long r = lo; r < hi; ++r) {
                if (horizon && !live_h(r)) continue;
                if (col[r] != 1.0) { ok.store(0); return; }
            }
        });
        return ok.load() == 1;
    };
    for (int k = 0; k < O; ++k) {
        md.z_ones0[k] = (md.qk[k] >= 1 && all_ones(d->Z[k], d->N[k], false)) ? 1 : 0;
        md.x_ones0[k] = all_ones(d->X[k], d->N[k], false) ? 1 : 0;
    }
    // XXs / ZZs columns: all ones -> not stored; identical to an earlier column over every row -> stored once
    struct ColRef { const double* q; const double* h; };
    std::vector<ColRef> pool;
    auto same_col = [&](const double* a, const double* b, long long rows, bool horizon) {
        if (a == b) return true;
        std::atomic<int> ok{1};
        parallel_for(rows, [&, a, b, horizon](long long lo, long long hi) {
            for (long long r = lo; r < hi; ++r) {
                if (horizon && !live_h(r)) continue;
                if (a[r] != b[r]) { ok.store(0); return; }
            }
        });
        return ok.load() == 1;
    };
    auto place = [&](const double* cq, const double* ch) -> int {
        if (all_ones(cq, nK, false) && (!nH || all_ones(ch, nH, true))) return -1;
        for (size_t c = 0; c < pool.size(); ++c)
            if (same_col(cq, pool[c].q, nK, false) && (!nH || same_col(ch, pool[c].h, nH, true))) return (int)c;
        pool.push_back({cq, ch});
        return (int)pool.size() - 1;
    };
    for (int t = 0; t < T; ++t) {
        for (int f = 0; f < md.ft[t]; ++f)
            md.xx_col[t][f] = place(d->XXs[t] + (long long)f * nK, nH ? d->XXs_h[t] + (long long)f * nH : nullptr);
        for (int j = 0; j < md.rt[t]; ++j)
            md.zz_col[t][j] = place(d->ZZs[t] + (long long)j * nK, nH ? d->ZZs_h[t] + (long long)j * nH : nullptr);
        bool uo = true;
        for (int u = 0; u < md.ut[t] && uo; ++u)
            uo = all_ones(d->Us[t] + (long long)u * nK, nK, false) && (!nH || all_ones(d->Us_h[t] + (long long)u * nH, nH, true));
        md.u_ones[t] = uo ? 1 : 0;
    }
    md.NC = (int)pool.size();
    {
        std::atomic<int> ok{p > 0 ? 1 : 0};
        if (p > 0)
            parallel_for(n, [&](long long lo, long long hi) {
                for (long long i = lo; i < hi; ++i)
                    for (int j = 0; j < p; ++j) {
                        const double w = d->W2s[i * K + (long long)j * nK];
                        for (int k = 0; k < K; ++k) if (d->W2s[i * K + k + (long long)j * nK] != w) { ok.store(0); return; }
                        for (int l = 0; l < d->n_horizons[i]; ++l)
                            for (int k = 0; k < K; ++k) if (d->W2s_h[(i * L + l) * K + k + (long long)j * nH] != w) { ok.store(0); return; }
                    }
            });
        md.w2_const = ok.load();
    }
    h->mm.L = sv_make_layout(md);
    const SvLayout& LY = h->mm.L;
    h->ps_len = LY.ps_Bs + B;

    // ---- record offsets and packing ----
    h->off.assign(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        long long len = LY.o_blk0 + (long long)(d->n_horizons[i] + 1) * LY.blk;
        for (int k = 0; k < O; ++k) len += (long long)(rowptr[k][i + 1] - rowptr[k][i]) * LY.long_cols[k];
        h->off[i + 1] = h->off[i] + (len + 1) / 2 * 2;
    }
    std::vector<double> rec((size_t)h->off[n], 0.0);
    parallel_for(n, [&](long long lo, long long hi) {
        for (long long i = lo; i < hi; ++i) {
            double* r0 = rec.data() + h->off[i];
            int* hdr = reinterpret_cast<int*>(r0 + LY.o_V);
            for (int k = 0; k < O; ++k) hdr[k] = rowptr[k][i + 1] - rowptr[k][i];
            const int nh = d->n_horizons[i];
            hdr[O] = nh;
            if (md.w2_const) for (int j = 0; j < p; ++j) r0[LY.o_W2c + j] = d->W2s[i * K + (long long)j * nK];
            for (int b = 0; b <= nh; ++b) {
                double* blk = r0 + LY.o_blk0 + (long long)b * LY.blk;
                const bool hz = b > 0;
                const long long rows = hz ? nH : nK;
                const long long base = hz ? (i * L + (b - 1)) * K : i * K;
                const double* W1x = hz ? d->W1s_h : d->W1s;
                const double* W2x = hz ? d->W2s_h : d->W2s;
                const double* Pwx = hz ? d->Pw_h : d->Pw;
                int* w1off = reinterpret_cast<int*>(blk + LY.b_W1off);
                for (int k = 0; k < K; ++k) {
                    const long long src = base + k;
                    blk[LY.b_Pw + k] = Pwx[src];
                    if (from_knots) {
                        int off; double bv[16];
                        spline_band_row(d->knots, d->n_knots, d->ord, (hz ? d->st_h : d->st)[src], off, bv);
                        // keep the window inside the basis when ord > n_Bs - off cannot happen: off <= n_Bs - ord
                        w1off[k] = off;
                        for (int j = 0; j < WB; ++j) blk[LY.b_W1v + j * kSvG + k] = bv[j];
                    } else {
                        int f, l;
                        band_of(W1x, rows, src, f, l);
                        int off = (l < f) ? 0 : std::min(f, B - WB);
                        if (WB == B) off = 0;
                        w1off[k] = off;
                        for (int j = 0; j < WB; ++j) blk[LY.b_W1v + j * kSvG + k] = W1x[src + (long long)(off + j) * rows];
                    }
                    if (!md.w2_const) for (int j = 0; j < p; ++j) blk[LY.b_W2 + j * kSvG + k] = W2x[src + (long long)j * rows];
                    for (int c = 0; c < md.NC; ++c) blk[LY.b_col + c * kSvG + k] = (hz ? pool[c].h : pool[c].q)[src];
                    for (int t = 0; t < T; ++t) {
                        const double* Ux = hz ? d->Us_h[t] : d->Us[t];
                        if (!md.u_ones[t]) for (int u = 0; u < md.ut[t]; ++u) blk[LY.b_U[t] + u * kSvG + k] = Ux[src + (long long)u * rows];
                    }
                }
            }
            double* lrec = r0 + LY.o_blk0 + (long long)(nh + 1) * LY.blk;
            for (int k = 0; k < O; ++k) {
                const int s0 = rowptr[k][i], v = rowptr[k][i + 1] - s0;
                const long long Nk = d->N[k];
                for (int r = 0; r < v; ++r) lrec[r] = d->y[k][s0 + r];
                int c = 1;
                for (int j = md.x_ones0[k]; j < md.pk[k]; ++j, ++c)
                    for (int r = 0; r < v; ++r) lrec[(long long)c * v + r] = d->X[k][s0 + r + (long long)j * Nk];
                for (int j = md.z_ones0[k]; j < md.qk[k]; ++j, ++c)
                    for (int r = 0; r < v; ++r) lrec[(long long)c * v + r] = d->Z[k][s0 + r + (long long)j * Nk];
                lrec += (long long)LY.long_cols[k] * v;
            }
        }
    });

    cudaError_t e;
#define CUH(call) do { e = (call); if (e != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e); jmb_svft_destroy(h); return fail(m_); } } while (0)
    CUH(cudaSetDevice(device));
    CUH(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CUH(cudaEventCreate(&h->ev0)); CUH(cudaEventCreate(&h->ev1));
    CUH(cudaMalloc(&h->d_rec, sizeof(double) * std::max<size_t>(rec.size(), 2)));
    CUH(cudaMemcpy(h->d_rec, rec.data(), sizeof(double) * rec.size(), cudaMemcpyHostToDevice));
    CUH(cudaMalloc(&h->d_off, sizeof(long long) * (n + 1)));
    CUH(cudaMemcpy(h->d_off, h->off.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
    CUH(cudaMalloc(&h->d_modes, sizeof(double) * (size_t)n * q));
    CUH(cudaMalloc(&h->d_hess, sizeof(double) * (size_t)n * q * q));
    CUH(cudaMalloc(&h->d_acc, sizeof(double) * n));
    CUH(cudaMalloc(&h->d_blast, sizeof(double) * (size_t)n * q));
    CUH(cudaMalloc(&h->d_sum, sizeof(double) * (size_t)n * 4 * std::max(L, 1)));
    CUH(cudaMalloc(&h->d_conv, sizeof(int) * n));
    CUH(cudaMalloc(&h->d_counts, sizeof(int) * 2 * (size_t)n));
    CUH(cudaMalloc(&h->d_work, sizeof(int)));
    CUH(cudaMalloc(&h->d_par_mean, sizeof(double) * h->ps_len));
    cudaDeviceProp prop;
    CUH(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;

    // kernel choice and launch shape: persistent CTAs, groups pull subjects from an atomic counter
    h->kernel = 0;
    if (sv_same_desc(md, SvSpecThree::d)) h->kernel = 1;
    else if (sv_same_desc(md, SvSpecValue::d)) h->kernel = 2;
    const char* env = getenv("JMB_SVFT_KERNEL");
    if (env && !strcmp(env, "generic")) h->kernel = 0;
    h->split_horizons = (h->kernel == 1 && sv_is_linear<SvSpecThree>()) || (h->kernel == 2 && sv_is_linear<SvSpecValue>());
    const int groups = JMB_SVFT_THREADS / kSvG;
    const int gd = (sv_group_doubles(q, h->ps_len) + 1) / 2 * 2;
    h->smem = (int)sizeof(double) * groups * gd;
    if (h->smem > 220 * 1024) { jmb_svft_destroy(h); return fail("jmb_svft_create: model too large for the shared-memory state"); }
    const void* fns[2] = {
        h->kernel == 1 ? (const void*)jmb_svft_modes_static<SvSpecThree> : (h->kernel == 2 ? (const void*)jmb_svft_modes_static<SvSpecValue> : (const void*)jmb_svft_modes_generic),
        h->kernel == 1 ? (const void*)jmb_svft_sample_static<SvSpecThree> : (h->kernel == 2 ? (const void*)jmb_svft_sample_static<SvSpecValue> : (const void*)jmb_svft_sample_generic)};
    for (int f = 0; f < 2; ++f) {
        CUH(cudaFuncSetAttribute(fns[f], cudaFuncAttributeMaxDynamicSharedMemorySize, h->smem));
        int per_sm = 1;
        CUH(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fns[f], JMB_SVFT_THREADS, h->smem));
        per_sm = std::max(per_sm, 1);
        h->grid[f] = std::max(1, std::min(h->num_sms * per_sm, (n + groups - 1) / groups));
    }
#undef CUH
#undef BAD
    *out = h;
    return 0;
}

// mcmc matrices (draws in rows, leading dimension M) -> [M][ps_len] parameter sets
static int pack_params(jmb_svft_handle h, const jmb_svft_params* P, std::vector<double>& out) {
    const SvDesc& md = h->mm.d;
    const SvLayout& LY = h->mm.L;
    const int M = P->M, q = md.q, B = h->B;
    if (M <= 0) return fail("jmb_svft: M must be positive");
    if (!P->betas || !P->invD || !P->Bs_gammas || !P->alphas || (md.p && !P->gammas)) return fail("jmb_svft: null parameter array");
    out.assign((size_t)M * h->ps_len, 0.0);
    for (int m = 0; m < M; ++m) {
        double* o = out.data() + (size_t)m * h->ps_len;
        for (int k = 0; k < md.O; ++k) {
            for (int j = 0; j < md.pk[k]; ++j) o[LY.ps_betas + md.boff[k] + j] = P->betas[k][m + (size_t)j * M];
            double sg = 1.0;
            if (md.fam[k] == JMB_FAM_GAUSSIAN) {
                if (!P->sigmas || !P->sigmas[k]) return fail("jmb_svft: sigmas[k] is required for a gaussian outcome");
                sg = P->sigmas[k][m];
            }
            o[LY.ps_isig + k] = 1.0 / sg;
        }
        for (int j = 0; j < md.p; ++j) o[LY.ps_gam + j] = P->gammas[m + (size_t)j * M];
        for (int j = 0; j < md.A; ++j) o[LY.ps_alp + j] = P->alphas[m + (size_t)j * M];
        for (int j = 0; j < q * q; ++j) o[LY.ps_invD + j] = P->invD[m + (size_t)j * M];
        for (int j = 0; j < B; ++j) o[LY.ps_Bs + j] = P->Bs_gammas[m + (size_t)j * M];
    }
    return 0;
}

static int run(jmb_svft_handle h, const jmb_svft_params* pm, const jmb_svft_params* pd, const jmb_svft_control* ct,
               jmb_svft_out* out, bool do_modes, bool do_sample) {
    if (!h || !ct || !out) return fail("jmb_svft: null argument");
    if (do_modes && (!pm || pm->M < 1)) return fail("jmb_svft: posterior means required");
    if (do_sample && !pd) return fail("jmb_svft: posterior draws required");
    if (!out->modes || (!out->hessians && !(do_modes && do_sample)))
        return fail("jmb_svft: out->modes is required; out->hessians may be NULL only for jmb_survfitJM (it is internal there)");
    if (do_sample && !out->summaries) return fail("jmb_svft: out->summaries is required");
    if (!(ct->parscale > 0.0) || !(ct->scale > 0.0) || !(ct->df > 0.0) || ct->maxit <= 0 || !(ct->cd_eps > 0.0) || !(ct->ndeps > 0.0))
        return fail("jmb_svft: bad control");
    CU(cudaSetDevice(h->device));
    const SvDesc& md = h->mm.d;
    const int n = h->n, q = md.q, L = h->L, M = do_sample ? pd->M : 0;
    std::vector<double> pk;
    if (do_modes) {
        if (pack_params(h, pm, pk)) return 1;
        CU(cudaMemcpyAsync(h->d_par_mean, pk.data(), sizeof(double) * h->ps_len, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    std::vector<double> pdk;
    if (do_sample) {
        if (pack_params(h, pd, pdk)) return 1;
        if (pdk.size() > h->cap_draws) {
            cudaFree(h->d_par_draws); h->d_par_draws = nullptr;
            CU(cudaMalloc(&h->d_par_draws, sizeof(double) * pdk.size()));
            h->cap_draws = pdk.size();
        }
        CU(cudaMemcpyAsync(h->d_par_draws, pdk.data(), sizeof(double) * pdk.size(), cudaMemcpyHostToDevice, h->stream));
        if (h->split_horizons && L > 0) {
            const size_t needC = (size_t)n * M * (md.NC + 1);
            if (needC > h->cap_cbuf) {
                cudaFree(h->d_cbuf); h->d_cbuf = nullptr;
                CU(cudaMalloc(&h->d_cbuf, sizeof(double) * needC));
                h->cap_cbuf = needC;
            }
        }
        const size_t needS = (size_t)n * M * std::max(L, 1);
        if (needS > h->cap_S) {
            cudaFree(h->d_S); h->d_S = nullptr;
            CU(cudaMalloc(&h->d_S, sizeof(double) * needS));
            h->cap_S = needS;
        }
    }
    std::vector<double> tmp;
    if (!do_modes) {        // modes / Hessians are inputs: n x q column-major and q x q x n -> subject-major
        tmp.resize((size_t)n * q);
        for (int i = 0; i < n; ++i) for (int j = 0; j < q; ++j) tmp[(size_t)i * q + j] = out->modes[i + (size_t)j * n];
        CU(cudaMemcpyAsync(h->d_modes, tmp.data(), sizeof(double) * tmp.size(), cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->d_hess, out->hessians, sizeof(double) * (size_t)n * q * q, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(cudaMemsetAsync(h->d_work, 0, sizeof(int), h->stream));
    CU(cudaMemsetAsync(h->d_conv, 0, sizeof(int) * n, h->stream));
    CU(cudaMemsetAsync(h->d_counts, 0, sizeof(int) * 2 * (size_t)n, h->stream));

    SvArgs a{};
    a.rec = h->d_rec; a.off = h->d_off; a.par_mean = do_modes ? h->d_par_mean : nullptr; a.par_draws = h->d_par_draws;
    a.n = n; a.M = M; a.L = L; a.B = h->B; a.ps_len = h->ps_len; a.subject_offset = h->subject_offset;
    a.do_modes = do_modes ? 1 : 0; a.do_sample = do_sample ? 1 : 0;
    a.c.scale = ct->scale; a.c.df = ct->df; a.c.ci_lo = ct->CI_levels[0]; a.c.ci_hi = ct->CI_levels[1];
    a.c.parscale = ct->parscale; a.c.reltol = ct->reltol; a.c.ndeps = ct->ndeps; a.c.cd_eps = ct->cd_eps;
    a.c.log = ct->log; a.c.maxit = ct->maxit; a.c.seed = ct->seed;
    a.Cbuf = (do_sample && h->split_horizons && L > 0) ? h->d_cbuf : nullptr; a.ncoef = md.NC + 1;
    a.modes = h->d_modes; a.hess = h->d_hess; a.conv = h->d_conv; a.counts = h->d_counts; a.S = h->d_S;
    a.acc_rate = h->d_acc; a.b_last = h->d_blast; a.work = h->d_work;

    CU(cudaEventRecord(h->ev0, h->stream));
    if (do_modes) {
        if (h->kernel == 1) jmb_svft_modes_static<SvSpecThree><<<h->grid[0], JMB_SVFT_THREADS, h->smem, h->stream>>>(a);
        else if (h->kernel == 2) jmb_svft_modes_static<SvSpecValue><<<h->grid[0], JMB_SVFT_THREADS, h->smem, h->stream>>>(a);
        else jmb_svft_modes_generic<<<h->grid[0], JMB_SVFT_THREADS, h->smem, h->stream>>>(h->mm, a);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    if (do_sample) {
        if (h->kernel == 1) jmb_svft_sample_static<SvSpecThree><<<h->grid[1], JMB_SVFT_THREADS, h->smem, h->stream>>>(a);
        else if (h->kernel == 2) jmb_svft_sample_static<SvSpecValue><<<h->grid[1], JMB_SVFT_THREADS, h->smem, h->stream>>>(a);
        else jmb_svft_sample_generic<<<h->grid[1], JMB_SVFT_THREADS, h->smem, h->stream>>>(h->mm, a);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    if (a.Cbuf) {                                   // horizon integrals: one thread per draw
        const int MC = 256;
        const size_t smem = sizeof(double) * ((size_t)MC * h->B + (size_t)md.K * (2 + md.WB + md.NC));
        const int grid = std::max(1, std::min(n, h->num_sms * 4));
        if (h->kernel == 1) {
            CU(cudaFuncSetAttribute((const void*)jmb_svft_horizons<SvSpecThree>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            jmb_svft_horizons<SvSpecThree><<<grid, MC, smem, h->stream>>>(a);
        } else {
            CU(cudaFuncSetAttribute((const void*)jmb_svft_horizons<SvSpecValue>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            jmb_svft_horizons<SvSpecValue><<<grid, MC, smem, h->stream>>>(a);
        }
        h->launches += 1;
        CU(cudaGetLastError());
    }
    if (do_sample && L > 0) {
        int Mp = 1;
        while (Mp < M) Mp <<= 1;
        if (Mp > 4096) return fail("jmb_svft: M too large for the summary kernel (max 4096)");
        const long long warps = (long long)n * L;
        const unsigned wg = (unsigned)((warps * 32 + 127) / 128);
#define SUMMARY_WARP(E) jmb_svft_summary_warp<E><<<wg, 128, 0, h->stream>>>(h->d_S, h->d_off, h->d_rec, md.O, n, M, L, ct->CI_levels[0], ct->CI_levels[1], h->d_sum)
        if (Mp <= 32) SUMMARY_WARP(1);
        else if (Mp <= 64) SUMMARY_WARP(2);
        else if (Mp <= 128) SUMMARY_WARP(4);
        else if (Mp <= 256) SUMMARY_WARP(8);
        else jmb_svft_summary<<<n, 128, sizeof(double) * Mp, h->stream>>>(h->d_S, h->d_off, h->d_rec, md.O, n, M, L, Mp,
                                                                          ct->CI_levels[0], ct->CI_levels[1], h->d_sum);
#undef SUMMARY_WARP
        h->launches += 1;
        CU(cudaGetLastError());
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));

    // ---- results ----
    if (do_modes) {
        tmp.resize((size_t)n * q);
        CU(cudaMemcpy(tmp.data(), h->d_modes, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) for (int j = 0; j < q; ++j) out->modes[i + (size_t)j * n] = tmp[(size_t)i * q + j];
        if (out->hessians) CU(cudaMemcpy(out->hessians, h->d_hess, sizeof(double) * (size_t)n * q * q, cudaMemcpyDeviceToHost));
        if (out->counts) CU(cudaMemcpy(out->counts, h->d_counts, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost));
    }
    if (out->convergence) {
        std::vector<int> cv(n);
        CU(cudaMemcpy(cv.data(), h->d_conv, sizeof(int) * n, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; ++i) if (do_modes || cv[i] == 20) out->convergence[i] = cv[i];
    }
    if (do_sample) {
        if (L > 0) CU(cudaMemcpy(out->summaries, h->d_sum, sizeof(double) * (size_t)n * 4 * L, cudaMemcpyDeviceToHost));
        if (out->full && L > 0) CU(cudaMemcpy(out->full, h->d_S, sizeof(double) * (size_t)n * M * L, cudaMemcpyDeviceToHost));
        if (out->accept_rate) CU(cudaMemcpy(out->accept_rate, h->d_acc, sizeof(double) * n, cudaMemcpyDeviceToHost));
        if (out->b_last) {
            tmp.resize((size_t)n * q);
            CU(cudaMemcpy(tmp.data(), h->d_blast, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
            for (int i = 0; i < n; ++i) for (int j = 0; j < q; ++j) out->b_last[i + (size_t)j * n] = tmp[(size_t)i * q + j];
        }
    }
    return 0;
}

extern "C" int jmb_svft_modes(jmb_svft_handle h, const jmb_svft_params* pm, const jmb_svft_control* ct, jmb_svft_out* out) {
    return run(h, pm, nullptr, ct, out, true, false);
}
extern "C" int jmb_svft_sample(jmb_svft_handle h, const jmb_svft_params* pd, const jmb_svft_control* ct, jmb_svft_out* out) {
    return run(h, nullptr, pd, ct, out, false, true);
}
extern "C" int jmb_survfitJM(jmb_svft_handle h, const jmb_svft_params* pm, const jmb_svft_params* pd,
                             const jmb_svft_control* ct, jmb_svft_out* out) {
    return run(h, pm, pd, ct, out, true, true);
}

extern "C" int jmb_svft_info(jmb_svft_handle h, float* last_kernel_ms, int64_t* launches, double* record_bytes) {
    if (!h) return fail("jmb_svft_info: null handle");
    if (last_kernel_ms) *last_kernel_ms = h->last_ms;
    if (launches) *launches = h->launches;
    if (record_bytes) *record_bytes = 8.0 * (double)h->off[h->n] / std::max(h->n, 1);
    return 0;
}
