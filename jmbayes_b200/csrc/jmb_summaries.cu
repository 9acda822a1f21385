// jmb_summaries.cu - posterior summaries of the kept draws (include/jmb200_summaries.h).
//
// One warp owns one series (the M kept draws of one scalar of the mcmc list) in shared memory and computes the ten entries
// of `statistics` (R/mvJointModelBayes.R:959-971) for it: moments, weighted moments, the Yule-Walker spectrum at zero
// (R/effectiveSize.R), the sort, type-7 and Hmisc "i/(n+1)" quantiles, and the binned kernel density of R/modes.R whose
// FFT convolution is evaluated here as the direct sum over the occupied cells (the same numbers: the kernel's wrap-around
// terms fall on empty cells).  A CTA loads the series of its warps together so that global reads are coalesced for both
// layouts (draws contiguous, or series contiguous as in the n x q x M array of random effects).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/jmb200_summaries.h"
#include "jmb_host.h"

namespace jmb {

constexpr int kStN = 1024;          // density.default: n = 1000 -> 2^ceiling(log2(n)) cells between lo and up
constexpr int kStUser = 1000;       // modes(): density(..., n = 1000)
constexpr unsigned kFull = 0xFFFFFFFFu;

struct StArgs {
    const double* x; long long S; int M, P, warps; long long sstride, dstride;
    const double* w;                // weights [M]
    double sumw, sumwt2, p_lo, p_hi, m_pow; int om;
    double *mean, *wmean, *mode, *ess, *sd, *wsd, *se, *ci, *wci, *pval;
    double* scratch;                // kStScratch doubles per resident warp
};

__device__ __forceinline__ double st_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// quantile(x, p, type = 7) on the sorted keys (stats:::quantile.default)
__device__ __forceinline__ double st_q7(const double* k, int M, double p) {
    const double index = (double)(M - 1) * p;
    const int lo = (int)floor(index), hi = (int)ceil(index);
    const double h = index - (double)lo;
    double qs = k[lo + (lo >> 5)];
    if (index > (double)lo && k[hi + (hi >> 5)] != qs) qs = (1.0 - h) * qs + h * k[hi + (hi >> 5)];
    return qs;
}

// Shared memory per warp (doubles): R1 [max(P + P / 32, 528)]: the series in draw order, then the sorted keys, then the density on the
// evaluated cells | R2 [512 + 1024]: cumulative weights, then the occupied BinDist cells and the kernel table | idx (uint16).
// Sorted keys and cumulative weights are stored skewed, element e at e + e / 32, so that both access patterns - lane l
// touching e = l * (P / 32) + r (the sort's registers, the per-lane chunks) and e = l + 32 i - are at most 2-way conflicted.
constexpr int kStYFast = 512, kStKFast = 1024;      // cells / kernel ordinates the shared-memory density path holds
constexpr int kStScratch = 4608;                    // doubles of global scratch per warp for series with longer tails
__host__ __device__ inline int st_half(int P) { return P + P / 32; }
__host__ __device__ inline int st_r1(int P) { return st_half(P) > 528 ? st_half(P) : 528; }
__host__ __device__ inline int st_warp_doubles(int P) { return st_r1(P) + 512 + 1024 + (P + 3) / 4; }
#define SKEW(e) ((e) + ((e) >> 5))

// Bitonic sort of the P = 32 EPL keys of a warp in registers: lane l holds positions l EPL .. l EPL + EPL - 1, so the
// exchanges at distance < EPL stay inside a lane and the others are one shuffle per element.  Writes the sorted keys
// (skewed) and the draw index of each to shared memory.
template <int EPL>
__device__ __forceinline__ void st_sort(const double* A, double* Bk, unsigned short* idx, int M, int lane) {
    constexpr int P = 32 * EPL;
    double key[EPL];
    int id[EPL];
#pragma unroll
    for (int r = 0; r < EPL; ++r) {                     // any initial placement will do: read without bank conflicts
        const int src = r * 32 + lane;
        key[r] = src < M ? A[src] : __longlong_as_double(0x7ff0000000000000LL);
        id[r] = src;
    }
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j >= EPL; j >>= 1) {
            // The partner of every key of this lane lives in lane ^ lm; the lane keeps the smaller key where it holds the
            // lower position of an ascending pair (or the upper of a descending one), else the larger: compared with the
            // sign bits flipped in the second case.  Bit masks, not branches - the compiler turns selects of three
            // registers into divergent branches.
            const int lm = j / EPL;
            const bool keep_min = (((lane * EPL) & k) == 0) == ((lane & lm) == 0);
            const int flip = keep_min ? 0 : (int)0x80000000u;
#pragma unroll
            for (int r = 0; r < EPL; ++r) {
                int mh = __double2hiint(key[r]), ml = __double2loint(key[r]);
                const int oh = __shfl_xor_sync(kFull, mh, lm), ol = __shfl_xor_sync(kFull, ml, lm);
                const int oi = __shfl_xor_sync(kFull, id[r], lm);
                const int tk = (__hiloint2double(oh ^ flip, ol) < __hiloint2double(mh ^ flip, ml)) ? -1 : 0;
                mh = (oh & tk) | (mh & ~tk); ml = (ol & tk) | (ml & ~tk);
                id[r] = (oi & tk) | (id[r] & ~tk);
                key[r] = __hiloint2double(mh, ml);
            }
        }
#pragma unroll
        for (int jj = EPL >> 1; jj > 0; jj >>= 1) {
            if (jj <= (k >> 1)) {
#pragma unroll
                for (int r = 0; r < EPL; ++r) {
                    if ((r & jj) == 0) {
                        const int p = r | jj;
                        const bool up = ((lane * EPL + r) & k) == 0;
                        const double x0 = key[r], x1 = key[p];
                        const bool ex = (x0 > x1) == up;          // selects, not branches: the warp must stay converged
                        const int i0 = id[r], i1 = id[p];
                        key[r] = ex ? x1 : x0; key[p] = ex ? x0 : x1;
                        id[r] = ex ? i1 : i0; id[p] = ex ? i0 : i1;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < EPL; ++r) { const int e = lane * EPL + r; Bk[SKEW(e)] = key[r]; idx[e] = (unsigned short)id[r]; }
}

// Autocovariance sums of the lags 0..om of the centred, zero-padded series xc (P + 32 entries): every lane keeps its P / 32
// elements in registers, one pass per lag; lane l returns the sum of lag l.
template <int EPL>
__device__ __forceinline__ double st_acf(const double* xc, int om, int lane) {
    double xr[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) xr[i] = xc[lane + 32 * i];
    double mine = 0.0;
    for (int l = 0; l <= om; ++l) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < EPL; ++i) s += xr[i] * xc[lane + 32 * i + l];       // the zero padding ends the sum at M - l
        s = st_sum(s);
        mine = (lane == l) ? s : mine;
    }
    return mine;
}

// The binned Gaussian kernel estimate of density.default on the cells jlo..jhi (into Ro) from the sorted keys, and the argmax
// over the grid points t_lo..t_hi of density()'s 1000-point grid.  Yb holds the occupied cells klo..khi, Kb the two-sided
// table of dnorm(|d| dk, sd = bw) over the cell distances d; the three buffers are shared memory or, for long-tailed
// series, global scratch.
struct StGrid { double from, to, lo, xdelta, by, dk, bw; int klo, khi, jlo, jhi, span, t_lo, t_hi; };

// Padding of the two tables: four zero cells in front of and eight behind the occupied cells (the sum runs over whole
// groups of four cells), and the kernel table reaches kStKPad ordinates further than the widest cell distance on both sides
// (the last block of 128 output cells may hang over jhi; those sums are discarded, their reads stay inside the table).
constexpr int kStYPad = 12, kStKPad = 136;
__host__ __device__ inline int st_ktable_quarter(int span) { return (2 * span + 2 * kStKPad + 1 + 3) / 4; }

__device__ __forceinline__ int st_density_argmax(const StGrid& g, const double* Bk, int M, int C, double* Yb, double* Kb, double* Ro, int lane) {
    const int ncell = g.khi - g.klo + 1;
    double* Yz = Yb + 4;                                     // Yz[c] = cell klo + c
    for (int k = lane; k < ncell + kStYPad; k += 32) Yb[k] = 0.0;
    // Kernel table, entry e = d + off for the cell distance d, stored permuted - entry e at (e & 3) Q + (e >> 2) - so that
    // lanes whose entries are four apart read neighbouring words.
    const int off = g.span + kStKPad, Q = st_ktable_quarter(g.span);
    const double cn = 0.3989422804014327 / g.bw;
    for (int d = lane; off + d < 4 * Q; d += 32) {           // one exp per distance, stored on both sides
        const double z = (double)d * g.dk / g.bw;
        const double kv = d <= g.span ? cn * exp(-0.5 * z * z) : 0.0;
        const int e1 = off + d, e0 = off - d;
        Kb[(e1 & 3) * Q + (e1 >> 2)] = kv;
        if (e0 >= 0) Kb[(e0 & 3) * Q + (e0 >> 2)] = kv;
    }
    __syncwarp();
    const double wi = 1.0 / (double)M;
    for (int r = 0; r < C; ++r) {                            // BinDist: the data lie 7 bw inside (lo, up), no edge cases;
        const int k = lane * C + r;                          // a lane's chunk of the sorted keys is far from its neighbours'
        if (k < M) {
            const double xpos = (Bk[SKEW(k)] - g.lo) / g.xdelta;
            const int ix = (int)floor(xpos);
            const double fx = xpos - (double)ix;
            atomicAdd(&Yz[ix - g.klo], wi * (1.0 - fx));
            atomicAdd(&Yz[ix + 1 - g.klo], wi * fx);
        }
    }
    __syncwarp();                                            // the sorted keys are no longer read: Ro may overlay them
    // Lane l sums the four neighbouring cells jb + 4 l + i: stepping to the next occupied cell k shifts the four kernel
    // ordinates by one, so each step costs one new ordinate and one (broadcast) cell: 2 LDS per 4 DFMA.
    for (int jb = g.jlo; jb <= g.jhi; jb += 128) {
        const int kstart = g.klo - ((g.klo - jb + off) & 3);             // (kstart - jb + off) is a multiple of four
        const int c0 = ((kstart - jb + off) >> 2) - lane;                // entry of (kstart + 4 t + u, i = 0) is 4 (c0 + t) + u
        const double* y = Yz + (kstart - g.klo);
        const double *p0 = Kb + c0, *p1 = Kb + Q + c0, *p2 = Kb + 2 * Q + c0, *p3 = Kb + 3 * Q + c0;
        // ordinates of i = 1, 2, 3 at kstart: entries 4 c0 - 1, - 2, - 3
        double w1 = p3[-1], w2 = p2[-1], w3 = p1[-1];
        double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
        const int nt = (g.khi - kstart + 4) >> 2;
#pragma unroll 2
        for (int t = 0; t < nt; ++t) {
            const double y0 = y[4 * t], y1 = y[4 * t + 1], y2 = y[4 * t + 2], y3 = y[4 * t + 3];
            const double a0 = p0[t], a1 = p1[t], a2 = p2[t], a3 = p3[t];
            r0 += y0 * a0; r1 += y0 * w1; r2 += y0 * w2; r3 += y0 * w3;
            r0 += y1 * a1; r1 += y1 * a0; r2 += y1 * w1; r3 += y1 * w2;
            r0 += y2 * a2; r1 += y2 * a1; r2 += y2 * a0; r3 += y2 * w1;
            r0 += y3 * a3; r1 += y3 * a2; r2 += y3 * a1; r3 += y3 * a0;
            w1 = a3; w2 = a2; w3 = a1;
        }
        const int j = jb + 4 * lane;
        if (j <= g.jhi) Ro[j - g.jlo] = fmax(0.0, r0);
        if (j + 1 <= g.jhi) Ro[j + 1 - g.jlo] = fmax(0.0, r1);
        if (j + 2 <= g.jhi) Ro[j + 2 - g.jlo] = fmax(0.0, r2);
        if (j + 3 <= g.jhi) Ro[j + 3 - g.jlo] = fmax(0.0, r3);
    }
    __syncwarp();
    double bv = -1.0; int bt = kStUser;
    for (int t = g.t_lo + lane; t <= g.t_hi; t += 32) {      // approx(xords, kords, x) and which.max
        const double xt = (t == kStUser - 1) ? g.to : g.from + (double)t * g.by;
        int jj = min(max((int)floor((xt - g.lo) / g.xdelta), g.jlo), g.jhi - 1);
        if (jj > g.jlo && g.lo + (double)jj * g.xdelta > xt) --jj;
        else if (jj < g.jhi - 1 && g.lo + (double)(jj + 1) * g.xdelta <= xt) ++jj;
        const double xa = g.lo + (double)jj * g.xdelta, xb = g.lo + (double)(jj + 1) * g.xdelta;
        const double ya = Ro[jj - g.jlo], yb = Ro[jj + 1 - g.jlo];
        const double val = ya + (yb - ya) * ((xt - xa) / (xb - xa));
        if (val > bv) { bv = val; bt = t; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(kFull, bv, o);
        const int ot = __shfl_xor_sync(kFull, bt, o);
        if (ov > bv || (ov == bv && ot < bt)) { bv = ov; bt = ot; }
    }
    return bt;
}

__global__ void __launch_bounds__(192, 2) jmb_stats_kernel(const StArgs a) {
    extern __shared__ double st_smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int M = a.M, P = a.P, W = a.warps;
    double* sw = st_smem;                                       // weights [M] (padded to even)
    const int wd = st_warp_doubles(P);
    double* base = st_smem + (M + 1) / 2 * 2 + (size_t)wid * wd;
    double* A = base;                                           // R1: the series in draw order ...
    double* Bk = base;                                          // ... the sorted keys (skewed) ...
    double* Rb = base;                                          // ... the density on the cells jlo..jhi
    double* cum = base + st_r1(P);                              // R2: cumulative weights in sorted order (skewed) ...
    double* Y = base + st_r1(P);                                // ... the occupied BinDist cells ...
    double* Kn = Y + kStYFast;                                  // ... and the kernel table
    unsigned short* idx = reinterpret_cast<unsigned short*>(Kn + kStKFast);
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    for (int j = tid; j < M; j += blockDim.x) sw[j] = a.w[j];

    for (long long tile = blockIdx.x; tile * W < a.S; tile += gridDim.x) {
        const long long s0 = tile * W;
        __syncthreads();                                         // the previous tile's series are no longer read
        if (a.dstride == 1) {                                    // draws contiguous: lanes run over the draws of one series
            for (int e = tid; e < W * M; e += blockDim.x) {
                const int sl = e / M, m = e - sl * M;
                if (s0 + sl < a.S) st_smem[(M + 1) / 2 * 2 + (size_t)sl * wd + m] = a.x[(s0 + sl) * a.sstride + m];
            }
        } else {                                                 // series contiguous: W neighbours of one draw per read
            for (int e = tid; e < W * M; e += blockDim.x) {
                const int m = e / W, sl = e - m * W;
                if (s0 + sl < a.S) st_smem[(M + 1) / 2 * 2 + (size_t)sl * wd + m] = a.x[(s0 + sl) * a.sstride + (long long)m * a.dstride];
            }
        }
        __syncthreads();
        const long long s = s0 + wid;
        if (s >= a.S) continue;

        // ---- moments: mean (two passes, as R's mean), var, weighted mean / variance (cov.wt), computeP -------------------
        double t0 = 0.0;
        for (int m = lane; m < M; m += 32) t0 += A[m];
        double mean = st_sum(t0) / (double)M;
        t0 = 0.0;
        for (int m = lane; m < M; m += 32) t0 += A[m] - mean;
        mean += st_sum(t0) / (double)M;
        double ss = 0.0, swx = 0.0, cnt = 0.0, sty = 0.0;
        const double tb = 0.5 * (double)(M + 1);
        for (int m = lane; m < M; m += 32) {
            const double v = A[m], d = v - mean;
            ss += d * d; swx += v * sw[m]; cnt += (v >= 0.0) ? 1.0 : 0.0; sty += ((double)(m + 1) - tb) * d;
        }
        const double var = st_sum(ss) / (double)(M - 1);
        const double wmean = st_sum(swx);
        cnt = st_sum(cnt);
        const double center = wmean / a.sumw;
        double sw2 = 0.0;
        for (int m = lane; m < M; m += 32) { const double d = A[m] - center; sw2 += sw[m] * d * d; }
        const double wvar = st_sum(sw2) / a.sumw / (1.0 - a.sumwt2);
        const double above = cnt / (double)M, below = ((double)M - cnt) / (double)M;

        // ---- effectiveSize: residuals of the linear trend, autocovariances, Levinson recursion, AIC order ----------------
        const double beta = st_sum(sty) / ((double)M * ((double)M * (double)M - 1.0) / 12.0);
        double rs = 0.0;
        for (int m = lane; m < M; m += 32) { const double r = (A[m] - mean) - beta * ((double)(m + 1) - tb); rs += r * r; }
        const double sd_res = sqrt(st_sum(rs) / (double)(M - 1));
        double spec = 0.0;
        if (sd_res > 1.5e-8) {                                   // identical(all.equal(sd(res), 0), TRUE) is FALSE
            double* xc = cum;                                    // the centred series (R2 is free until the sort is done)
            for (int m = lane; m < P + 32; m += 32) xc[m] = m < M ? A[m] - mean : 0.0;
            __syncwarp();
            double r;                                            // lane l: autocovariance at lag l (divisor n)
            switch (P) {
                case 128: r = st_acf<4>(xc, a.om, lane); break;
                case 256: r = st_acf<8>(xc, a.om, lane); break;
                case 512: r = st_acf<16>(xc, a.om, lane); break;
                default: r = st_acf<32>(xc, a.om, lane); break;
            }
            r /= (double)M;
            double phi = 0.0;                                    // lane j: coefficient j of the current order
            double v = __shfl_sync(kFull, r, 0);
            double v_own = v, sp_own = 0.0;                      // lane l keeps var.pred and sum(ar) of order l
            for (int l = 1; l <= a.om; ++l) {
                const bool in = lane >= 1 && lane < l;
                const double rj = __shfl_sync(kFull, r, (l - lane) & 31);
                const double acc = __shfl_sync(kFull, r, l) - st_sum(in ? phi * rj : 0.0);
                const double k = acc / v;
                const double prev = __shfl_sync(kFull, phi, (l - lane) & 31);
                phi = in ? phi - k * prev : (lane == l ? k : phi);
                v *= 1.0 - k * k;
                const double sp = st_sum((lane >= 1 && lane <= l) ? phi : 0.0);
                if (lane == l) { v_own = v; sp_own = sp; }
            }
            // xaic = n.used * log(var.pred) + 2 * order + 2; the first minimum over the orders 0..order.max
            double aic = (lane <= a.om) ? (double)M * log(v_own) + 2.0 * (double)lane + 2.0 : inf;
            if (!(aic == aic)) aic = inf;
            int order = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double oa = __shfl_xor_sync(kFull, aic, o);
                const int oo = __shfl_xor_sync(kFull, order, o);
                if (oa < aic || (oa == aic && oo < order)) { aic = oa; order = oo; }
            }
            const double vbest = __shfl_sync(kFull, v_own, order), sumphi = __shfl_sync(kFull, sp_own, order);
            const double var_pred = vbest * (double)M / (double)(M - (order + 1));
            spec = var_pred / ((1.0 - sumphi) * (1.0 - sumphi));
        }
        const double ess = (spec == 0.0) ? 0.0 : (double)M * var / spec;

        // ---- sort (bitonic in registers, keys with their draw index); the sorted keys replace the series ------------------
        __syncwarp();
        switch (P) {
            case 128: st_sort<4>(A, Bk, idx, M, lane); break;
            case 256: st_sort<8>(A, Bk, idx, M, lane); break;
            case 512: st_sort<16>(A, Bk, idx, M, lane); break;
            default: st_sort<32>(A, Bk, idx, M, lane); break;
        }
        __syncwarp();
        const double q_lo = st_q7(Bk, M, a.p_lo), q_hi = st_q7(Bk, M, a.p_hi);
        const double iqr = st_q7(Bk, M, 0.75) - st_q7(Bk, M, 0.25);

        // ---- Hmisc::wtd.quantile(type = "i/(n+1)"): cumulative weights in sorted order, knots at the last of tied keys ----
        const int C = P / 32;                                    // contiguous elements per lane
        {
            double part = 0.0;
            for (int k = lane * C; k < (lane + 1) * C; ++k) if (k < M) part += sw[idx[k]];
            double incl = part;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double t = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += t; }
            double run = incl - part;
            for (int k = lane * C; k < (lane + 1) * C; ++k) if (k < M) { run += sw[idx[k]]; cum[SKEW(k)] = run; }
        }
        __syncwarp();
        const double wn1 = cum[SKEW(M - 1)] + 1.0;               // n + 1, n = sum of the weights
        double wq[2];
        {
            int first_k = P, last_k = -1;                        // first / last knot of this lane's chunk
            int jge[2] = {P, P}, ilt[2] = {-1, -1};
            const double pr[2] = {a.p_lo, a.p_hi};
            for (int k = lane * C; k < (lane + 1) * C; ++k) {
                if (k >= M || sw[idx[k]] == 0.0) continue;       // zero-weight observations are dropped
                int k2 = k + 1;
                while (k2 < M && sw[idx[k2]] == 0.0) ++k2;
                if (k2 < M && Bk[SKEW(k2)] == Bk[SKEW(k)]) continue;         // wtd.table: tied values are one point
                first_k = min(first_k, k); last_k = k;
                const double cdf = cum[SKEW(k)] / wn1;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (cdf >= pr[e]) jge[e] = min(jge[e], k); else ilt[e] = k;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                first_k = min(first_k, __shfl_xor_sync(kFull, first_k, o));
                last_k = max(last_k, __shfl_xor_sync(kFull, last_k, o));
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    jge[e] = min(jge[e], __shfl_xor_sync(kFull, jge[e], o));
                    ilt[e] = max(ilt[e], __shfl_xor_sync(kFull, ilt[e], o));
                }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                if (last_k < 0) wq[e] = nan;                                   // no positive weight
                else if (jge[e] >= P) wq[e] = Bk[SKEW(last_k)];                // approx(rule = 2) beyond the last point
                else if (ilt[e] < 0) wq[e] = Bk[SKEW(first_k)];                // between the prepended (0, x[1]) and x[1]
                else {
                    const double ci = cum[SKEW(ilt[e])] / wn1, cj = cum[SKEW(jge[e])] / wn1, yi = Bk[SKEW(ilt[e])], yj = Bk[SKEW(jge[e])];
                    wq[e] = yi + (yj - yi) * ((pr[e] - ci) / (cj - ci));
                }
            }
        }

        // ---- modes(): density(y, bw = "nrd", adjust = 3, n = 1000), x of the maximum -------------------------------------
        double mode = nan;
        const double bw = 3.0 * (1.06 * fmin(sqrt(var), iqr / 1.34) * a.m_pow);
        const double xmin = Bk[0], xmax = Bk[SKEW(M - 1)];
        __syncwarp();                                            // the cumulative weights are no longer read
        if (bw > 0.0 && bw < inf) {
            StGrid g;
            g.bw = bw;
            g.from = xmin - 3.0 * bw; g.to = xmax + 3.0 * bw; g.lo = g.from - 4.0 * bw;
            const double up = g.to + 4.0 * bw;
            g.xdelta = (up - g.lo) / (double)(kStN - 1);
            g.by = (g.to - g.from) / (double)(kStUser - 1);
            g.dk = 2.0 * (up - g.lo) / (double)(2 * kStN - 1);
            g.klo = (int)floor((xmin - g.lo) / g.xdelta); g.khi = min(kStN - 1, (int)floor((xmax - g.lo) / g.xdelta) + 1);
            // Left of the first and right of the last occupied cell the estimate is monotone (every kernel term is), so the
            // maximum over density()'s 1000-point grid is taken between the draws: only those grid points (and a margin of
            // four, a cell being at most 2.3 grid steps wide) and the cells under them are evaluated.
            g.t_lo = max(0, (int)floor((xmin - g.from) / g.by) - 4); g.t_hi = min(kStUser - 1, (int)ceil((xmax - g.from) / g.by) + 4);
            const double xt_hi = (g.t_hi == kStUser - 1) ? g.to : g.from + (double)g.t_hi * g.by;
            g.jlo = max(0, (int)floor((g.from + (double)g.t_lo * g.by - g.lo) / g.xdelta) - 1);
            g.jhi = min(kStN - 1, (int)floor((xt_hi - g.lo) / g.xdelta) + 2);
            g.span = max(g.khi - g.jlo, g.jhi - g.klo);
            int bt;
            if (g.khi - g.klo + 1 + kStYPad <= kStYFast && 4 * st_ktable_quarter(g.span) <= kStKFast && g.jhi - g.jlo + 1 <= st_r1(P)) {
                bt = st_density_argmax(g, Bk, M, C, Y, Kn, Rb, lane);
            } else {                                             // long tails: the same arithmetic on global scratch
                double* sc = a.scratch + ((size_t)blockIdx.x * W + wid) * kStScratch;
                bt = st_density_argmax(g, Bk, M, C, sc, sc + 1056, sc + 3456, lane);   // cells | kernel table (<= 2320) | density
            }
            mode = (bt == kStUser - 1) ? g.to : g.from + (double)bt * g.by;
        }

        if (lane == 0) {
            if (a.mean) a.mean[s] = mean;
            if (a.wmean) a.wmean[s] = wmean;
            if (a.mode) a.mode[s] = mode;
            if (a.ess) a.ess[s] = ess;
            if (a.sd) a.sd[s] = sqrt(var);
            if (a.wsd) a.wsd[s] = sqrt(wvar);
            if (a.se) a.se[s] = sqrt(var / ess);
            if (a.ci) { a.ci[2 * s] = q_lo; a.ci[2 * s + 1] = q_hi; }
            if (a.wci) { a.wci[2 * s] = wq[0]; a.wci[2 * s + 1] = wq[1]; }
            if (a.pval) a.pval[s] = 2.0 * fmin(above, below);
        }
    }
}

}  // namespace jmb

using namespace jmb;

namespace {
int fail(const std::string& m) { return jmb_set_error(m); }

// Device buffers and streams of jmb_mcmc_statistics, kept per device between calls (grow-only; jmb_stats_release frees
// them): allocating and freeing ~0.7 GB per call costs more than the kernel at small S.
struct StWorkspace {
    cudaStream_t st[2] = {nullptr, nullptr};
    void* buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};     // weights, scratch, staging 0, staging 1, outputs
    size_t cap[5] = {0, 0, 0, 0, 0};
};
constexpr int kStMaxDevices = 64;
std::mutex g_st_mu;
StWorkspace g_st_ws[kStMaxDevices];

cudaError_t st_reserve(StWorkspace& w, int slot, size_t bytes) {
    if (w.cap[slot] >= bytes) return cudaSuccess;
    if (w.buf[slot]) cudaFree(w.buf[slot]);
    w.buf[slot] = nullptr; w.cap[slot] = 0;
    const cudaError_t e = cudaMalloc(&w.buf[slot], bytes);
    if (e == cudaSuccess) w.cap[slot] = bytes;
    return e;
}
}  // namespace

extern "C" int jmb_stats_release(void) {
    std::lock_guard<std::mutex> lock(g_st_mu);
    for (int d = 0; d < kStMaxDevices; ++d) {
        StWorkspace& w = g_st_ws[d];
        bool used = w.st[0] || w.st[1];
        for (int k = 0; k < 5; ++k) used = used || w.buf[k];
        if (!used) continue;
        if (cudaSetDevice(d) != cudaSuccess) continue;
        for (int k = 0; k < 5; ++k) { if (w.buf[k]) cudaFree(w.buf[k]); w.buf[k] = nullptr; w.cap[k] = 0; }
        for (auto& s : w.st) { if (s) cudaStreamDestroy(s); s = nullptr; }
    }
    return 0;
}

extern "C" int jmb_stand_weights(const double* LogLiks, int32_t M, double* weights) {
    if (!LogLiks || !weights || M <= 0) return fail("jmb_stand_weights: NULL argument or M <= 0");
    double mx = -INFINITY;
    for (int m = 0; m < M; ++m) if (LogLiks[m] > mx) mx = LogLiks[m];             // max(x, na.rm = TRUE)
    const double upp = mx + std::log((double)M);
    double s = 0.0;
    for (int m = 0; m < M; ++m) { weights[m] = std::exp(LogLiks[m] - upp); s += weights[m]; }
    for (int m = 0; m < M; ++m) weights[m] /= s;
    return 0;
}

extern "C" int jmb_mcmc_statistics(int device, const double* x, int64_t S, int32_t M, int64_t series_stride,
                                   int64_t draw_stride, const double* weights, const double probs[2], int on_device,
                                   jmb_stats_out* out, float* kernel_ms) {
    if (!x || !weights || !probs || !out) return fail("jmb_mcmc_statistics: NULL argument");
    if (S <= 0) return fail("jmb_mcmc_statistics: S <= 0");
    if (M < 4 || M > JMB_STATS_MAX_M) return fail("jmb_mcmc_statistics: M outside 4.." + std::to_string(JMB_STATS_MAX_M));
    if (!(series_stride == 1 && draw_stride >= S) && !(draw_stride == 1 && series_stride >= M))
        return fail("jmb_mcmc_statistics: layout must be series-contiguous (series_stride = 1, draw_stride >= S) or "
                    "draw-contiguous (draw_stride = 1, series_stride >= M)");
    for (int e = 0; e < 2; ++e) if (!(probs[e] >= 0.0 && probs[e] <= 1.0)) return fail("jmb_mcmc_statistics: probs outside [0, 1]");
    double sumw = 0.0, sumwt2 = 0.0;
    for (int m = 0; m < M; ++m) {
        if (!(weights[m] >= 0.0)) return fail("jmb_mcmc_statistics: weights must be non-negative and not NA");
        sumw += weights[m];
    }
    if (!(sumw > 0.0)) return fail("jmb_mcmc_statistics: weights sum to zero");
    for (int m = 0; m < M; ++m) { const double t = weights[m] / sumw; sumwt2 += t * t; }

    if (device < 0 || device >= kStMaxDevices) return fail("jmb_mcmc_statistics: device out of range");
    std::lock_guard<std::mutex> lock(g_st_mu);
    StWorkspace& ws = g_st_ws[device];
    cudaError_t e;
    std::vector<cudaEvent_t> evs;
    auto cleanup = [&]() { for (auto ev : evs) cudaEventDestroy(ev); };
#define CUS(call) do { e = (call); if (e != cudaSuccess) { std::string m_ = std::string(#call) + ": " + cudaGetErrorString(e); cleanup(); return fail(m_); } } while (0)
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail("jmb_mcmc_statistics: no CUDA device (there is no CPU fallback)");
    CUS(cudaSetDevice(device));
    int num_sms = 0;
    CUS(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
    for (auto& sx : ws.st) if (!sx) CUS(cudaStreamCreateWithFlags(&sx, cudaStreamNonBlocking));
    cudaStream_t* st = ws.st;

    StArgs a{};
    a.M = M; a.P = 4; while (a.P < M) a.P <<= 1;
    a.P = std::max(a.P, 128);
    a.warps = 4;                                  // three CTAs per SM at M <= 512 (74 kB each), two at M <= 1024
    if (const char* ev = getenv("JMB_STATS_WARPS")) { const int wv = atoi(ev); if (wv >= 1 && wv <= 6) a.warps = wv; }   // experiments
    a.sumw = sumw; a.sumwt2 = sumwt2; a.p_lo = probs[0]; a.p_hi = probs[1];
    a.m_pow = std::pow((double)M, -0.2);
    a.om = (int)std::min<double>(M - 1, std::floor(10.0 * std::log10((double)M)));
    const size_t smem = sizeof(double) * ((size_t)(M + 1) / 2 * 2 + (size_t)a.warps * st_warp_doubles(a.P));
    CUS(cudaFuncSetAttribute(jmb_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUS(st_reserve(ws, 0, sizeof(double) * JMB_STATS_MAX_M));
    double* d_w = static_cast<double*>(ws.buf[0]);
    int per_sm = 1;
    CUS(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, jmb_stats_kernel, 32 * a.warps, smem));
    const int max_grid = std::max(per_sm, 1) * num_sms;
    const size_t scratch_doubles = (size_t)kStScratch * max_grid * a.warps;       // one set per stream
    CUS(st_reserve(ws, 1, sizeof(double) * 2 * scratch_doubles));
    double* d_scratch = static_cast<double*>(ws.buf[1]);
    a.scratch = d_scratch;
    CUS(cudaMemcpy(d_w, weights, sizeof(double) * M, cudaMemcpyHostToDevice));
    a.w = d_w;

    double* host_out[10] = {out->postMeans, out->postwMeans, out->postModes, out->EffectiveSize, out->StDev, out->wStDev,
                            out->StErr, out->CIs, out->wCIs, out->Pvalues};
    const int width[10] = {1, 1, 1, 1, 1, 1, 1, 2, 2, 1};
    double* dev_out[10] = {nullptr};
    if (!on_device) CUS(st_reserve(ws, 4, sizeof(double) * 12 * (size_t)S));
    for (int k = 0, col = 0; k < 10; col += width[k], ++k) {
        if (!host_out[k]) continue;
        dev_out[k] = on_device ? host_out[k] : static_cast<double*>(ws.buf[4]) + (size_t)col * S;
    }
    auto launch = [&](const double* xd, long long s0, long long count, long long sstride, long long dstride, cudaStream_t stream) {
        StArgs b = a;
        b.scratch = d_scratch + (stream == st[1] ? scratch_doubles : 0);
        b.x = xd; b.S = count; b.sstride = sstride; b.dstride = dstride;
        b.mean = dev_out[0] ? dev_out[0] + s0 : nullptr; b.wmean = dev_out[1] ? dev_out[1] + s0 : nullptr;
        b.mode = dev_out[2] ? dev_out[2] + s0 : nullptr; b.ess = dev_out[3] ? dev_out[3] + s0 : nullptr;
        b.sd = dev_out[4] ? dev_out[4] + s0 : nullptr; b.wsd = dev_out[5] ? dev_out[5] + s0 : nullptr;
        b.se = dev_out[6] ? dev_out[6] + s0 : nullptr; b.ci = dev_out[7] ? dev_out[7] + 2 * s0 : nullptr;
        b.wci = dev_out[8] ? dev_out[8] + 2 * s0 : nullptr; b.pval = dev_out[9] ? dev_out[9] + s0 : nullptr;
        const long long tiles = (count + a.warps - 1) / a.warps;
        const int grid = (int)std::max<long long>(1, std::min<long long>(tiles, max_grid));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1); evs.push_back(e0); evs.push_back(e1);
        cudaEventRecord(e0, stream);
        jmb_stats_kernel<<<grid, 32 * a.warps, smem, stream>>>(b);
        cudaEventRecord(e1, stream);
    };

    if (on_device) {
        launch(x, 0, S, series_stride, draw_stride, st[0]);
        CUS(cudaGetLastError());
    } else {
        // stream the series through two device buffers: copy of chunk c + 1 overlaps the kernel of chunk c
        const long long chunk = std::max<long long>(a.warps, std::min<long long>(S, (256LL << 20) / (8LL * M)) / a.warps * a.warps);
        double* d_x[2] = {nullptr, nullptr};
        for (int c2 = 0; c2 < 2; ++c2) { CUS(st_reserve(ws, 2 + c2, sizeof(double) * chunk * M)); d_x[c2] = static_cast<double*>(ws.buf[2 + c2]); }
        int c = 0;
        for (long long s0 = 0; s0 < S; s0 += chunk, c ^= 1) {
            const long long cnt = std::min<long long>(chunk, S - s0);
            if (series_stride == 1) {                            // rows = draws, cnt series wide
                CUS(cudaMemcpy2DAsync(d_x[c], sizeof(double) * cnt, x + s0, sizeof(double) * draw_stride, sizeof(double) * cnt, M,
                                      cudaMemcpyHostToDevice, st[c]));
                launch(d_x[c], s0, cnt, 1, cnt, st[c]);
            } else {                                             // rows = series, M draws wide
                CUS(cudaMemcpy2DAsync(d_x[c], sizeof(double) * M, x + s0 * series_stride, sizeof(double) * series_stride,
                                      sizeof(double) * M, cnt, cudaMemcpyHostToDevice, st[c]));
                launch(d_x[c], s0, cnt, M, 1, st[c]);
            }
            CUS(cudaGetLastError());
        }
    }
    for (int c2 = 0; c2 < 2; ++c2) CUS(cudaStreamSynchronize(st[c2]));
    if (kernel_ms) {
        *kernel_ms = 0.f;
        for (size_t k = 0; k + 1 < evs.size(); k += 2) { float ms = 0.f; CUS(cudaEventElapsedTime(&ms, evs[k], evs[k + 1])); *kernel_ms += ms; }
    }
    if (!on_device)
        for (int k = 0; k < 10; ++k)
            if (host_out[k]) CUS(cudaMemcpy(host_out[k], dev_out[k], sizeof(double) * width[k] * S, cudaMemcpyDeviceToHost));
#undef CUS
    cleanup();
    return 0;
}
