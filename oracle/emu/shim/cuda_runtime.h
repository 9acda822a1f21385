// cuda_runtime.h - stand-in for the CUDA runtime that executes the library's own kernels on CPU threads.
//
// TEST INFRASTRUCTURE ONLY (oracle/emu): libjmb200_emu.so is the product's jmb200.cu / jmb_kernels.cuh / jmb_device.cuh,
// passed through the textual substitutions of make_emu.py (kernel<<<...>>>(...) launches, __shared__ -> static, the two PDL
// asm statements) and compiled by g++ against this header, in the same way oracle/_ref compiles the reference's sources
// against a stand-in Armadillo/Rcpp header.  It lets the CPU test suite run the generic step kernel, the finalize
// kernel, the survival-posterior kernel and the Xbetas kernel - the code a B200 runs - on small cohorts when no GPU is
// available.  Nothing in the product loads it; it is not a fallback and is never timed.
//
// Execution model: a launch runs its CTAs one after the other; the threads of a CTA are OS threads that really run
// concurrently, __syncthreads() / __syncwarp(mask) / __shfl_*_sync(mask, ...) are barriers over exactly the threads they
// name, so divergence between the half-warp groups of the step kernel behaves as on the device.  Streams are synchronous.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __constant__ static const
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
using std::isfinite;
using std::isnan;
using std::isinf;

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorEmu = 1 };
struct emuStream { int dummy; };
struct emuEvent { std::chrono::steady_clock::time_point t; };
typedef emuStream* cudaStream_t;
typedef emuEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaHostAllocDefault = 0, cudaStreamNonBlocking = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp { int multiProcessorCount; };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 6 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes = 0; cudaStream_t stream = nullptr; cudaLaunchAttribute* attrs = nullptr; unsigned numAttrs = 0; };

namespace emu {
struct Barrier {
    std::mutex m; std::condition_variable cv; int count = 0; unsigned long long gen = 0;
    void wait(int n) {
        std::unique_lock<std::mutex> l(m);
        const unsigned long long g = gen;
        if (++count == n) { count = 0; ++gen; cv.notify_all(); }
        else cv.wait(l, [&] { return g != gen; });
    }
};
struct Cta {
    Barrier all;
    std::mutex wm[32];
    std::map<unsigned, Barrier> warp[32];   // [warp]: one barrier per participation mask
    unsigned long long slot[32][32];        // shuffle / vote exchange
};
extern Cta g_cta;
extern thread_local unsigned t_tid;
void mbar_init(const void* bar, int count);       // emulated mbarrier objects, keyed by their shared-memory address
void mbar_arrive_expect_tx(const void* bar, unsigned bytes);
void mbar_complete_tx(const void* bar, unsigned bytes);
void mbar_wait(const void* bar, unsigned parity);
double* dyn_smem();                         // the CTA's dynamic shared memory (every `extern __shared__` array aliases it)
inline void warp_barrier(unsigned mask) {
    const unsigned w = t_tid >> 5;
    Barrier* b;
    { std::lock_guard<std::mutex> l(g_cta.wm[w]); b = &g_cta.warp[w][mask]; }
    b->wait(__builtin_popcount(mask));
}
template <class T> inline T shfl(unsigned mask, T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle of more than 8 bytes");
    const unsigned w = t_tid >> 5, lane = t_tid & 31;
    unsigned long long raw = 0; std::memcpy(&raw, &v, sizeof(T));
    g_cta.slot[w][lane] = raw;
    warp_barrier(mask);
    const unsigned long long got = g_cta.slot[w][src_lane & 31];
    warp_barrier(mask);
    T r; std::memcpy(&r, &got, sizeof(T));
    return r;
}
void launch(dim3 grid, dim3 block, size_t smem, cudaStream_t st, const std::function<void()>& fn);
}  // namespace emu

extern thread_local uint3 threadIdx;
extern uint3 blockIdx;
extern dim3 blockDim, gridDim;

inline void __syncthreads() { emu::g_cta.all.wait((int)(blockDim.x * blockDim.y * blockDim.z)); }
inline void __syncwarp(unsigned mask = 0xFFFFFFFFu) { emu::warp_barrier(mask); }
inline void __threadfence_system() {}
template <class T> inline T __shfl_xor_sync(unsigned mask, T v, int off, int width = 32) { (void)width; return emu::shfl(mask, v, (int)(emu::t_tid & 31) ^ off); }
template <class T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = (int)(emu::t_tid & 31);
    return emu::shfl(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
template <class T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)(emu::t_tid & 31), src = lane - (int)delta;
    const T got = emu::shfl(mask, v, src < (lane & ~(width - 1)) ? lane : src);
    return got;
}
template <class T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = (int)(emu::t_tid & 31), src = lane + (int)delta;
    return emu::shfl(mask, v, src > (lane | (width - 1)) ? lane : src);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    const unsigned w = emu::t_tid >> 5, lane = emu::t_tid & 31;
    emu::g_cta.slot[w][lane] = pred ? 1ull : 0ull;
    emu::warp_barrier(mask);
    unsigned r = 0;
    for (unsigned l = 0; l < 32; ++l) if ((mask >> l) & 1u) r |= (unsigned)(emu::g_cta.slot[w][l] & 1ull) << l;
    emu::warp_barrier(mask);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline double __longlong_as_double(long long x) { double d; std::memcpy(&d, &x, 8); return d; }
inline long long __double_as_longlong(double d) { long long x; std::memcpy(&x, &d, 8); return x; }
inline double __hiloint2double(int hi, int lo) { const unsigned long long u = ((unsigned long long)(unsigned)hi << 32) | (unsigned)lo; double d; std::memcpy(&d, &u, 8); return d; }
inline int __double2loint(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return (int)(unsigned)(u & 0xFFFFFFFFull); }
inline int __double2hiint(double d) { unsigned long long u; std::memcpy(&u, &d, 8); return (int)(unsigned)(u >> 32); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline double atomicAdd(double* p, double v) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = __atomic_load_n(q, __ATOMIC_RELAXED), neu;
    double d;
    do { std::memcpy(&d, &old, 8); d += v; std::memcpy(&neu, &d, 8); } while (!__atomic_compare_exchange_n(q, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
    std::memcpy(&d, &old, 8);
    return d;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline void sincospi(double x, double* s, double* c) {      // exact reduction to a multiple of 1/2, then sin / cos of pi r
    const double q = std::rint(2.0 * x), r = x - 0.5 * q;
    const double sr = std::sin(M_PI * r), cr = std::cos(M_PI * r);
    switch (((long long)q) & 3) {
        case 0: *s = sr; *c = cr; break;
        case 1: *s = cr; *c = -sr; break;
        case 2: *s = -sr; *c = -cr; break;
        default: *s = -cr; *c = sr; break;
    }
}
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
struct double2 { double x, y; };
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }

// ---- runtime -----------------------------------------------------------------------------------------------------------
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA runtime: unsupported call"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* c) { *c = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
// one SM: few, large CTAs, so that every warp of the TMA-staged kernel runs several steps of its pipeline
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { p->multiProcessorCount = 1; return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* f, size_t* t) { *f = (size_t)1 << 30; *t = (size_t)2 << 30; return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t bytes) { *p = (T*)std::malloc(bytes ? bytes : 8); return *p ? cudaSuccess : cudaErrorEmu; }
template <class T> inline cudaError_t cudaHostAlloc(T** p, size_t bytes, unsigned) { *p = (T*)std::malloc(bytes ? bytes : 8); return *p ? cudaSuccess : cudaErrorEmu; }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emuStream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emuEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = std::chrono::steady_clock::now(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count(); return cudaSuccess; }
inline cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F*, cudaFuncAttribute, int) { return cudaSuccess; }
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 1; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < h; ++r) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorEmu; }
inline cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned) { return cudaErrorEmu; }
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorEmu; }
template <typename... KA, typename... A>
inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* c, void (*k)(KA...), A... a) {
    emu::launch(c->gridDim, c->blockDim, c->dynamicSmemBytes, c->stream, [&] { k(a...); });
    return cudaSuccess;
}
