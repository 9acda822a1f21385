// jmb_accuracy.cu - pair / subject sums of aucJM.mvJMbayes and prederrJM.mvJMbayes (include/jmb200_accuracy.h).
// Product code: nothing under oracle/ is used; every entry point fails loudly without a CUDA device.
#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <vector>

#include "../../include/jmb200_accuracy.h"
#include "jmb_host.h"

#define CUA(call)                                                                                  \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return jmb_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

namespace {

constexpr int kTile = 256;      // subjects per tile = threads per CTA
constexpr int kSplit = 16;      // CTAs that share the column tiles of one row tile

// Per-subject factors of the pair weight (R/aucJM.mvJMbayes.R:74-77,98,114,139): a_i for the first member of a pair,
// b_j for the second.  0 = the subject cannot take that role.
__global__ void auc_factors_kernel(int n, const double* __restrict__ Time, const double* __restrict__ event, double Thoriz,
                                   const double* __restrict__ pi2, const double* __restrict__ pi3,
                                   double* __restrict__ a, double* __restrict__ b) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double T = Time[i], d = event[i];
    const bool early = T <= Thoriz;
    const bool ev = (d == 1.0 || d == 3.0), cens = (d == 0.0 || d == 2.0);
    a[i] = (early && ev) ? 1.0 : ((early && cens) ? pi2[i] : 0.0);
    b[i] = (T > Thoriz) ? 1.0 : ((early && cens) ? pi3[i] : 0.0);
}

// sum over the pairs i < j of ind[i, j] and of (pi_i < pi_j) * ind[i, j], ind = a_i b_j (:141).  CTA (ti, y) owns the rows
// of tile ti and the column tiles ti + y, ti + y + kSplit, ...: one thread per row, the column tile staged in shared
// memory.  Every thread adds its pairs in a fixed order and the CTA combines its threads in a fixed order, so the result
// does not depend on scheduling.
__global__ void __launch_bounds__(kTile)
auc_pairs_kernel(int n, const double* __restrict__ pi, const double* __restrict__ a, const double* __restrict__ b,
                 double* __restrict__ partials /* [gridDim.x][kSplit][2] */) {
    __shared__ double s_pi[kTile], s_b[kTile];
    __shared__ double s_red[2][kTile];
    const int ti = blockIdx.x, y = blockIdx.y, tid = threadIdx.x;
    const int tiles = (n + kTile - 1) / kTile;
    const int i = ti * kTile + tid;
    const double ai = (i < n) ? a[i] : 0.0, pii = (i < n) ? pi[i] : 0.0;
    const bool row_live = (i < n) && !(ai == 0.0);          // NaN a_i stays live: its pairs are NA unless b_j = 0
    double num = 0.0, den = 0.0;
    for (int tj = ti + y; tj < tiles; tj += kSplit) {
        const int j0 = tj * kTile;
        __syncthreads();
        s_pi[tid] = (j0 + tid < n) ? pi[j0 + tid] : 0.0;
        s_b[tid] = (j0 + tid < n) ? b[j0 + tid] : 0.0;
        __syncthreads();
        if (!row_live) continue;
        const int jbeg = (tj == ti) ? tid + 1 : 0;          // i < j
        const int jend = min(kTile, n - j0);
        for (int jj = jbeg; jj < jend; ++jj) {
            const double bj = s_b[jj];
            if (bj == 0.0) continue;                        // not a comparable pair: ind = FALSE
            const double w = ai * bj;
            if (w != w) continue;                           // NA weight: dropped from both sums (na.rm = TRUE)
            den += w;
            const double pj = s_pi[jj];
            if (pii != pii || pj != pj) continue;           // NA probability: dropped from the numerator
            if (pii < pj) num += w;
        }
    }
    s_red[0][tid] = num; s_red[1][tid] = den;
    __syncthreads();
    for (int wdt = kTile / 2; wdt > 0; wdt >>= 1) {
        if (tid < wdt) { s_red[0][tid] += s_red[0][tid + wdt]; s_red[1][tid] += s_red[1][tid + wdt]; }
        __syncthreads();
    }
    if (tid == 0) {
        partials[((size_t)ti * kSplit + y) * 2] = s_red[0][0];
        partials[((size_t)ti * kSplit + y) * 2 + 1] = s_red[1][0];
    }
}

// fixed-order sum of `cols` interleaved columns of a [rows][cols] matrix, one CTA per column
__global__ void __launch_bounds__(256) colsum_kernel(const double* __restrict__ x, long long rows, int cols, double* __restrict__ out) {
    __shared__ double s[256];
    const int c = blockIdx.x;
    double t = 0.0;
    for (long long r = threadIdx.x; r < rows; r += 256) t += x[r * cols + c];
    s[threadIdx.x] = t;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) { if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) out[c] = s[0];
}

// prederrJM: per-subject loss terms of the three groups (R/prederrJM.mvJMbayes.R:104-108); NaN terms count as 0 (na.rm)
__global__ void prederr_terms_kernel(int loss, int n_alive, const double* __restrict__ Sa, int n_dead, const double* __restrict__ Sd,
                                     int n_cens, const double* __restrict__ Sc, const double* __restrict__ wc,
                                     double* __restrict__ terms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = n_alive + n_dead + n_cens;
    if (i >= n) return;
    auto L = [loss](double x) { return loss == 0 ? fabs(x) : x * x; };
    double t;
    if (i < n_alive) t = L(1.0 - Sa[i]);
    else if (i < n_alive + n_dead) t = L(0.0 - Sd[i - n_alive]);
    else { const int k = i - n_alive - n_dead; t = wc[k] * L(1.0 - Sc[k]) + (1.0 - wc[k]) * L(0.0 - Sc[k]); }
    terms[i] = (t != t) ? 0.0 : t;
}

// rocJM: thread k of a CTA owns threshold k and walks the CTA's chunk of subjects (staged in shared memory) in order;
// partials [CTA][n_thr][3] (+ the chunk's sum(ind), sum(1 - ind) in slots [CTA][n_thr][0..1]) are combined in CTA order by colsum_kernel
constexpr int kRocThreads = 128, kRocChunk = 1024;
__global__ void __launch_bounds__(kRocThreads)
roc_sums_kernel(int n, const double* __restrict__ Time, const double* __restrict__ event, double Thoriz, const double* __restrict__ pi,
                const double* __restrict__ pi2, int n_thr, const double* __restrict__ thrs, double* __restrict__ partials) {
    __shared__ double s_pi[kRocChunk], s_ind[kRocChunk];
    const int base = blockIdx.x * kRocChunk, cnt = min(kRocChunk, n - base);
    for (int e = threadIdx.x; e < cnt; e += kRocThreads) {
        const int i = base + e;
        const double T = Time[i], d = event[i];
        const bool before = T < Thoriz;                                             // strict, R/rocJM.mvJMbayes.R:69,71
        s_ind[e] = (before && (d == 1.0 || d == 3.0)) ? 1.0 : ((before && (d == 0.0 || d == 2.0)) ? pi2[i] : 0.0);
        s_pi[e] = pi[i];
    }
    __syncthreads();
    const size_t row = (size_t)blockIdx.x * (n_thr + 1) * 3;
    for (int k = threadIdx.x; k <= n_thr; k += kRocThreads) {
        double tp = 0.0, fp = 0.0, q = 0.0;
        if (k < n_thr) {
            const double th = thrs[k];
            for (int e = 0; e < cnt; ++e) {
                const double lt = (s_pi[e] != s_pi[e]) ? s_pi[e] : ((s_pi[e] < th) ? 1.0 : 0.0);   // NA < x is NA
                tp += lt * s_ind[e]; fp += lt * (1.0 - s_ind[e]); q += lt;
            }
        } else {
            // sum(ind) and sum(1 - ind) of the chunk, summed in the order of nTP / nFP so that nFN = sum(ind) - nTP and
            // nTN = sum(1 - ind) - nFP (R/rocJM.mvJMbayes.R:90,93) cancel exactly where every probability is below the threshold
            for (int e = 0; e < cnt; ++e) { tp += s_ind[e]; fp += 1.0 - s_ind[e]; }
        }
        partials[row + (size_t)k * 3] = tp; partials[row + (size_t)k * 3 + 1] = fp; partials[row + (size_t)k * 3 + 2] = q;
    }
}

struct DevBuf {
    double* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

int pick_device(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return jmb_set_error(std::string("no CUDA device available (this library has no CPU fallback)") + (e != cudaSuccess ? std::string(": ") + cudaGetErrorString(e) : ""));
    if (device < 0 || device >= ndev) return jmb_set_error("bad device ordinal");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return jmb_set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace

extern "C" int jmb_aucJM_pairs(int device, int32_t n, const double* Time, const double* event, double Thoriz, const double* pi_u_t,
                               const double* pi2, const double* pi3, double out[3], float* kernel_ms) {
    if (!Time || !event || !pi_u_t || !pi2 || !pi3 || !out) return jmb_set_error("jmb_aucJM_pairs: null argument");
    if (n < 0) return jmb_set_error("jmb_aucJM_pairs: n < 0");
    if (pick_device(device)) return 1;
    if (kernel_ms) *kernel_ms = 0.f;
    if (n < 2) { out[0] = std::nan(""); out[1] = 0.0; out[2] = 0.0; return 0; }      // R/aucJM.mvJMbayes.R:66,142-144
    const int tiles = (n + kTile - 1) / kTile;
    DevBuf in, ab, part, res;
    CUA(cudaMalloc(&in.p, sizeof(double) * 5 * (size_t)n));
    CUA(cudaMalloc(&ab.p, sizeof(double) * 2 * (size_t)n));
    CUA(cudaMalloc(&part.p, sizeof(double) * 2 * (size_t)tiles * kSplit));
    CUA(cudaMalloc(&res.p, sizeof(double) * 2));
    const double* src[5] = {Time, event, pi_u_t, pi2, pi3};
    for (int k = 0; k < 5; ++k) CUA(cudaMemcpy(in.p + (size_t)k * n, src[k], sizeof(double) * n, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CUA(cudaEventCreate(&e0)); CUA(cudaEventCreate(&e1));
    CUA(cudaEventRecord(e0, 0));
    auc_factors_kernel<<<(n + 255) / 256, 256>>>(n, in.p, in.p + n, Thoriz, in.p + 3 * (size_t)n, in.p + 4 * (size_t)n, ab.p, ab.p + n);
    auc_pairs_kernel<<<dim3(tiles, kSplit), kTile>>>(n, in.p + 2 * (size_t)n, ab.p, ab.p + n, part.p);
    colsum_kernel<<<2, 256>>>(part.p, (long long)tiles * kSplit, 2, res.p);
    CUA(cudaEventRecord(e1, 0));
    CUA(cudaGetLastError());
    double h[2];
    CUA(cudaMemcpy(h, res.p, sizeof(h), cudaMemcpyDeviceToHost));
    if (kernel_ms) CUA(cudaEventElapsedTime(kernel_ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    out[1] = h[0]; out[2] = h[1];
    out[0] = h[0] / h[1];                                    // 0 / 0 = NaN like R
    return 0;
}

extern "C" int jmb_prederrJM_sum(int device, int32_t loss, int32_t nr, int32_t n_alive, const double* S_alive, int32_t n_dead,
                                 const double* S_dead, int32_t n_cens, const double* S_cens, const double* w_cens, double* out) {
    if (!out || (n_alive > 0 && !S_alive) || (n_dead > 0 && !S_dead) || (n_cens > 0 && (!S_cens || !w_cens)))
        return jmb_set_error("jmb_prederrJM_sum: null argument");
    if (n_alive < 0 || n_dead < 0 || n_cens < 0 || nr <= 0 || (loss != 0 && loss != 1)) return jmb_set_error("jmb_prederrJM_sum: bad argument");
    if (pick_device(device)) return 1;
    const int n = n_alive + n_dead + n_cens;
    if (n == 0) { *out = 0.0; return 0; }
    DevBuf in, terms, res;
    CUA(cudaMalloc(&in.p, sizeof(double) * ((size_t)n + n_cens + 1)));
    CUA(cudaMalloc(&terms.p, sizeof(double) * (size_t)n));
    CUA(cudaMalloc(&res.p, sizeof(double)));
    double* dSa = in.p; double* dSd = dSa + n_alive; double* dSc = dSd + n_dead; double* dwc = dSc + n_cens;
    if (n_alive) CUA(cudaMemcpy(dSa, S_alive, sizeof(double) * n_alive, cudaMemcpyHostToDevice));
    if (n_dead) CUA(cudaMemcpy(dSd, S_dead, sizeof(double) * n_dead, cudaMemcpyHostToDevice));
    if (n_cens) {
        CUA(cudaMemcpy(dSc, S_cens, sizeof(double) * n_cens, cudaMemcpyHostToDevice));
        CUA(cudaMemcpy(dwc, w_cens, sizeof(double) * n_cens, cudaMemcpyHostToDevice));
    }
    prederr_terms_kernel<<<(n + 255) / 256, 256>>>(loss, n_alive, dSa, n_dead, dSd, n_cens, dSc, dwc, terms.p);
    colsum_kernel<<<1, 256>>>(terms.p, n, 1, res.p);
    CUA(cudaGetLastError());
    double h = 0.0;
    CUA(cudaMemcpy(&h, res.p, sizeof(h), cudaMemcpyDeviceToHost));
    *out = (1.0 / nr) * h;
    return 0;
}

extern "C" int jmb_rocJM_sums(int device, int32_t n, const double* Time, const double* event, double Thoriz, const double* pi_u_t,
                              const double* pi2, int32_t n_thr, const double* thrs, double* nTP, double* nFP, double* Q, double* sum_ind) {
    if (!Time || !event || !pi_u_t || !pi2 || !thrs || !nTP || !nFP || !Q || !sum_ind) return jmb_set_error("jmb_rocJM_sums: null argument");
    if (n <= 0 || n_thr <= 0 || n_thr > 4096) return jmb_set_error("jmb_rocJM_sums: bad n / n_thr");
    if (pick_device(device)) return 1;
    const int ctas = (n + kRocChunk - 1) / kRocChunk, cols = (n_thr + 1) * 3;
    DevBuf in, part, res;
    CUA(cudaMalloc(&in.p, sizeof(double) * (4 * (size_t)n + n_thr)));
    CUA(cudaMalloc(&part.p, sizeof(double) * (size_t)ctas * cols));
    CUA(cudaMalloc(&res.p, sizeof(double) * cols));
    const double* src[4] = {Time, event, pi_u_t, pi2};
    for (int k = 0; k < 4; ++k) CUA(cudaMemcpy(in.p + (size_t)k * n, src[k], sizeof(double) * n, cudaMemcpyHostToDevice));
    CUA(cudaMemcpy(in.p + 4 * (size_t)n, thrs, sizeof(double) * n_thr, cudaMemcpyHostToDevice));
    roc_sums_kernel<<<ctas, kRocThreads>>>(n, in.p, in.p + n, Thoriz, in.p + 2 * (size_t)n, in.p + 3 * (size_t)n, n_thr, in.p + 4 * (size_t)n, part.p);
    colsum_kernel<<<cols, 256>>>(part.p, ctas, cols, res.p);
    CUA(cudaGetLastError());
    std::vector<double> h(cols);
    CUA(cudaMemcpy(h.data(), res.p, sizeof(double) * cols, cudaMemcpyDeviceToHost));
    for (int k = 0; k < n_thr; ++k) { nTP[k] = h[(size_t)k * 3]; nFP[k] = h[(size_t)k * 3 + 1]; Q[k] = h[(size_t)k * 3 + 2] / n; }
    sum_ind[0] = h[(size_t)n_thr * 3]; sum_ind[1] = h[(size_t)n_thr * 3 + 1];
    return 0;
}
