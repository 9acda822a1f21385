// emu_runtime.cpp - CTA execution for the stand-in CUDA runtime (oracle/emu/shim/cuda_runtime.h).  TEST INFRASTRUCTURE ONLY.
#include "cuda_runtime.h"

thread_local uint3 threadIdx;
uint3 blockIdx;
dim3 blockDim, gridDim;
namespace emu {
Cta g_cta;
thread_local unsigned t_tid;

void launch(dim3 grid, dim3 block, size_t, cudaStream_t, const std::function<void()>& fn) {
    const unsigned nthreads = block.x * block.y * block.z;
    if (nthreads > 1024) { std::fprintf(stderr, "emu: block too large\n"); std::abort(); }
    ::gridDim = grid; ::blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                ::blockIdx = uint3{bx, by, bz};
                std::vector<std::thread> th;
                th.reserve(nthreads);
                for (unsigned t = 0; t < nthreads; ++t)
                    th.emplace_back([&, t] {
                        t_tid = t;
                        ::threadIdx = uint3{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        fn();
                    });
                for (auto& x : th) x.join();
            }
}
}  // namespace emu
