// emu_runtime.cpp - CTA execution for the stand-in CUDA runtime (oracle/emu/shim/cuda_runtime.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_runtime.h"

thread_local uint3 threadIdx;
uint3 blockIdx;
dim3 blockDim, gridDim;
namespace emu {
Cta g_cta;
thread_local unsigned t_tid;
double* dyn_smem() { alignas(128) static double buf[1 << 16]; return buf; }    // 512 KB >= any CTA's dynamic shared memory

namespace {
struct Mbar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
std::mutex g_mb_m;
std::condition_variable g_mb_cv;
std::map<const void*, Mbar> g_mb;
void mbar_maybe_complete(Mbar& b) {
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; g_mb_cv.notify_all(); }
}
}  // namespace
void mbar_init(const void* bar, int count) { std::lock_guard<std::mutex> l(g_mb_m); g_mb[bar] = Mbar{count, count, 0, 0}; }
void mbar_arrive_expect_tx(const void* bar, unsigned bytes) {
    std::lock_guard<std::mutex> l(g_mb_m);
    Mbar& b = g_mb[bar]; b.tx += bytes; b.pending -= 1; mbar_maybe_complete(b);
}
void mbar_complete_tx(const void* bar, unsigned bytes) {
    std::lock_guard<std::mutex> l(g_mb_m);
    Mbar& b = g_mb[bar]; b.tx -= bytes; mbar_maybe_complete(b);
}
void mbar_wait(const void* bar, unsigned parity) {
    std::unique_lock<std::mutex> l(g_mb_m);
    g_mb_cv.wait(l, [&] { return (g_mb[bar].phase & 1u) != (parity & 1u); });
}

namespace {
// One OS thread per CUDA thread of the running CTA, kept between launches (a chain is thousands of small launches).
struct Pool {
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::vector<std::thread> th;
    unsigned long long gen = 0;
    unsigned active = 0, remaining = 0;
    const std::function<void()>* fn = nullptr;
    dim3 block;

    void worker(unsigned id) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> l(m);
            cv_job.wait(l, [&] { return gen != seen; });
            seen = gen;
            if (id >= active) continue;
            const std::function<void()>* f = fn;
            const dim3 b = block;
            l.unlock();
            t_tid = id;
            ::threadIdx = uint3{id % b.x, (id / b.x) % b.y, id / (b.x * b.y)};
            (*f)();
            l.lock();
            if (--remaining == 0) cv_done.notify_one();
        }
    }
    void run_cta(unsigned n, dim3 b, const std::function<void()>& f) {
        while (th.size() < n) {
            const unsigned id = (unsigned)th.size();
            th.emplace_back([this, id] { worker(id); });
            th.back().detach();
        }
        std::unique_lock<std::mutex> l(m);
        active = n; remaining = n; fn = &f; block = b; ++gen;
        cv_job.notify_all();
        cv_done.wait(l, [&] { return remaining == 0; });
    }
};
Pool& pool() { static Pool* p = new Pool(); return *p; }      // never destroyed: its threads are detached
std::mutex g_launch;                                            // one launch at a time (host threads of the library's packers do not launch)
}  // namespace

void launch(dim3 grid, dim3 block, size_t, cudaStream_t, const std::function<void()>& fn) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > 1024) { std::fprintf(stderr, "emu: bad block size\n"); std::abort(); }
    std::lock_guard<std::mutex> g(g_launch);
    ::gridDim = grid; ::blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                ::blockIdx = uint3{bx, by, bz};
                pool().run_cta(nthreads, block, fn);
            }
}
}  // namespace emu
