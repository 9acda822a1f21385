// jmb_designs.cu - design construction on the device (include/jmb200_designs.h): right_rows /
// right_rows_mstate (R/supportFuns_mvJointModelBayes.R:16-40) and splineDesign in band form
// (R/mvJointModelBayes.R:382-383,389-392).  One thread per quadrature point; integer results are bit-exact.
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/jmb200_designs.h"
#include "jmb_host.h"

namespace {

#define CUD(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return jmb_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

// findInterval(q, times_i) with the defaults (R: number of times_i[j] <= q), then ind[ind < 1] <- 1, as a 0-based global row
__global__ void right_rows_kernel(const long long n_rows, const int K, const long long* __restrict__ row_ptr, const double* __restrict__ times,
                                  const int* __restrict__ idT, const double* __restrict__ Q, long long* __restrict__ out) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_rows * K) return;
    const long long r = e / K;
    const long long i = idT ? idT[r] : r;
    const long long lo0 = row_ptr[i], hi0 = row_ptr[i + 1];
    const double q = Q[e];
    long long lo = lo0, hi = hi0;                 // first position whose time is > q (upper bound)
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (times[mid] <= q) lo = mid + 1; else hi = mid;
    }
    long long cnt = lo - lo0;                     // findInterval
    if (!(q == q)) cnt = 0;                       // findInterval(NA, .) is NA; ind[ind < 1] <- 1 leaves NA: first row is the least wrong
    if (cnt < 1) cnt = 1;
    out[e] = lo0 + cnt - 1;
}

__global__ void spline_band_kernel(const double* __restrict__ t, const int nk, const int ord, const double* __restrict__ x, const long long n,
                                   int* __restrict__ off, double* __restrict__ vals, double* __restrict__ dense) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int nb = nk - ord;
    double Bl[16], dl[16], dr[16];
    for (int r = 0; r < ord; ++r) Bl[r] = 0.0;
    int o = 0;
    const double xv = x[i];
    if (xv >= t[ord - 1] && xv <= t[nb]) {
        int lo = 0, hi = nk;                       // upper bound: first knot > x
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (t[mid] <= xv) lo = mid + 1; else hi = mid; }
        int j = lo - 1;
        j = min(max(j, ord - 1), nb - 1);
        Bl[0] = 1.0;
        for (int d = 1; d < ord; ++d) {
            dr[d - 1] = t[j + d] - xv;
            dl[d - 1] = xv - t[j + 1 - d];
            double saved = 0.0;
            for (int r = 0; r < d; ++r) {
                const double den = dr[r] + dl[d - 1 - r];
                const double term = den != 0.0 ? Bl[r] / den : 0.0;
                Bl[r] = saved + dr[r] * term;
                saved = dl[d - 1 - r] * term;
            }
            Bl[d] = saved;
        }
        o = j - (ord - 1);
    }
    off[i] = o;
    for (int r = 0; r < ord; ++r) vals[i * ord + r] = Bl[r];
    if (dense) for (int r = 0; r < ord; ++r) if (Bl[r] != 0.0 || (xv >= t[ord - 1] && xv <= t[nb])) dense[i + (long long)(o + r) * n] = Bl[r];
}

// data[c(ind), ] with the time column replaced: one thread per selected row, column-major in and out (writes coalesced over the rows,
// reads nearly so: consecutive quadrature points of a subject select the same or neighbouring visits)
__global__ void data_rows_kernel(const double* __restrict__ X, const long long N, const int p, const long long* __restrict__ rows, const long long M,
                                 const int time_col, const double* __restrict__ new_time, double* __restrict__ out) {
    const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const long long r = rows[m];
    for (int c = 0; c < p; ++c) out[m + (long long)c * M] = (c == time_col) ? new_time[m] : X[r + (long long)c * N];
}

struct Dev { void* p = nullptr; ~Dev() { if (p) cudaFree(p); } };

int pick(int device) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return jmb_set_error("no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return jmb_set_error("bad device ordinal");
    CUD(cudaSetDevice(device));
    return 0;
}

int right_rows_impl(int device, int32_t n, long long n_rows, int32_t K, const int64_t* row_ptr, const double* times, const int32_t* idT,
                    const double* Q, int64_t* out) {
    if (!row_ptr || !times || !Q || !out) return jmb_set_error("jmb_right_rows: null argument");
    if (n <= 0 || n_rows <= 0 || K <= 0) return jmb_set_error("jmb_right_rows: bad n / K");
    for (int i = 0; i < n; ++i)
        if (row_ptr[i + 1] <= row_ptr[i]) return jmb_set_error("jmb_right_rows: subject " + std::to_string(i) + " has no rows in the long data");
    if (idT) for (long long r = 0; r < n_rows; ++r) if (idT[r] < 0 || idT[r] >= n) return jmb_set_error("jmb_right_rows_mstate: idT out of range");
    if (pick(device)) return 1;
    const long long N = row_ptr[n], E = n_rows * K;
    Dev d_rp, d_t, d_q, d_o, d_id;
    CUD(cudaMalloc(&d_rp.p, sizeof(long long) * (n + 1))); CUD(cudaMalloc(&d_t.p, sizeof(double) * N));
    CUD(cudaMalloc(&d_q.p, sizeof(double) * E)); CUD(cudaMalloc(&d_o.p, sizeof(long long) * E));
    CUD(cudaMemcpy(d_rp.p, row_ptr, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(d_t.p, times, sizeof(double) * N, cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(d_q.p, Q, sizeof(double) * E, cudaMemcpyHostToDevice));
    if (idT) { CUD(cudaMalloc(&d_id.p, sizeof(int) * n_rows)); CUD(cudaMemcpy(d_id.p, idT, sizeof(int) * n_rows, cudaMemcpyHostToDevice)); }
    right_rows_kernel<<<(unsigned)((E + 255) / 256), 256>>>(n_rows, K, (const long long*)d_rp.p, (const double*)d_t.p, (const int*)d_id.p,
                                                              (const double*)d_q.p, (long long*)d_o.p);
    CUD(cudaGetLastError());
    CUD(cudaMemcpy(out, d_o.p, sizeof(long long) * E, cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace

extern "C" int jmb_right_rows(int device, int32_t n, int32_t K, const int64_t* row_ptr, const double* times, const double* Q, int64_t* out) {
    return right_rows_impl(device, n, n, K, row_ptr, times, nullptr, Q, out);
}

extern "C" int jmb_right_rows_mstate(int device, int32_t n, int32_t nT, int32_t K, const int64_t* row_ptr, const double* times,
                                     const int32_t* idT, const double* Q, int64_t* out) {
    if (!idT) return jmb_set_error("jmb_right_rows_mstate: idT required");
    return right_rows_impl(device, n, nT, K, row_ptr, times, idT, Q, out);
}

extern "C" int jmb_spline_design(int device, const double* knots, int32_t n_knots, int32_t ord, const double* x, int64_t n,
                                 int32_t* off, double* vals, double* dense) {
    if (!knots || !x || !off || !vals) return jmb_set_error("jmb_spline_design: null argument");
    if (ord < 1 || ord > 8 || n_knots < 2 * ord || n <= 0) return jmb_set_error("jmb_spline_design: need 1 <= ord <= 8, n_knots >= 2 ord, n > 0");
    for (int j = 1; j < n_knots; ++j) if (knots[j] < knots[j - 1]) return jmb_set_error("jmb_spline_design: knots must be non-decreasing");
    if (pick(device)) return 1;
    const int nb = n_knots - ord;
    Dev d_k, d_x, d_off, d_v, d_d;
    CUD(cudaMalloc(&d_k.p, sizeof(double) * n_knots)); CUD(cudaMalloc(&d_x.p, sizeof(double) * n));
    CUD(cudaMalloc(&d_off.p, sizeof(int) * n)); CUD(cudaMalloc(&d_v.p, sizeof(double) * n * ord));
    CUD(cudaMemcpy(d_k.p, knots, sizeof(double) * n_knots, cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(d_x.p, x, sizeof(double) * n, cudaMemcpyHostToDevice));
    if (dense) { CUD(cudaMalloc(&d_d.p, sizeof(double) * n * nb)); CUD(cudaMemset(d_d.p, 0, sizeof(double) * n * nb)); }
    spline_band_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const double*)d_k.p, n_knots, ord, (const double*)d_x.p, n, (int*)d_off.p,
                                                              (double*)d_v.p, (double*)d_d.p);
    CUD(cudaGetLastError());
    CUD(cudaMemcpy(off, d_off.p, sizeof(int) * n, cudaMemcpyDeviceToHost));
    CUD(cudaMemcpy(vals, d_v.p, sizeof(double) * n * ord, cudaMemcpyDeviceToHost));
    if (dense) CUD(cudaMemcpy(dense, d_d.p, sizeof(double) * n * nb, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int jmb_data_rows(int device, const double* X, int64_t N, int32_t p, const int64_t* rows, int64_t M, int32_t time_col,
                             const double* new_time, double* out) {
    if (!X || !rows || !out) return jmb_set_error("jmb_data_rows: null argument");
    if (N <= 0 || p <= 0 || M <= 0) return jmb_set_error("jmb_data_rows: need N, p, M > 0");
    if (time_col >= p || time_col < -1) return jmb_set_error("jmb_data_rows: time_col out of range");
    if (time_col >= 0 && !new_time) return jmb_set_error("jmb_data_rows: new_time required with a time column");
    for (int64_t m = 0; m < M; ++m) if (rows[m] < 0 || rows[m] >= N) return jmb_set_error("jmb_data_rows: row index " + std::to_string(m) + " out of range");
    if (pick(device)) return 1;
    Dev d_x, d_r, d_t, d_o;
    CUD(cudaMalloc(&d_x.p, sizeof(double) * N * p)); CUD(cudaMalloc(&d_r.p, sizeof(long long) * M));
    CUD(cudaMalloc(&d_o.p, sizeof(double) * M * p));
    CUD(cudaMemcpy(d_x.p, X, sizeof(double) * N * p, cudaMemcpyHostToDevice));
    CUD(cudaMemcpy(d_r.p, rows, sizeof(long long) * M, cudaMemcpyHostToDevice));
    if (time_col >= 0) { CUD(cudaMalloc(&d_t.p, sizeof(double) * M)); CUD(cudaMemcpy(d_t.p, new_time, sizeof(double) * M, cudaMemcpyHostToDevice)); }
    data_rows_kernel<<<(unsigned)((M + 255) / 256), 256>>>((const double*)d_x.p, (long long)N, p, (const long long*)d_r.p, (long long)M, time_col,
                                                           (const double*)d_t.p, (double*)d_o.p);
    CUD(cudaGetLastError());
    CUD(cudaMemcpy(out, d_o.p, sizeof(double) * M * p, cudaMemcpyDeviceToHost));
    return 0;
}
