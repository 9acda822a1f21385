// jmb_step_static.cuh (oracle/emu stand-in): the compile-time specialised / TMA-staged step kernels are sm_100a code (inline
// PTX: mbarrier, cp.async.bulk) and cannot run on CPU threads; the emulated library always uses the generic kernel
// (JMB_KERNEL=generic), so these are declarations that are never launched.
#pragma once
#include "jmb_kernels.cuh"
namespace jmb {
constexpr int kTmaThreads = 256, kTmaCtas = 1, kTmaWarps = kTmaThreads / 32;
struct TmaArgs { int cap_d, cap_c, subjects_per_cta; };
template <class Spec> void jmb_step_static(int, ParamLayout, ChainConst, StepArgs) { std::abort(); }
template <class Spec> void jmb_step_tma(int, ParamLayout, ChainConst, StepArgs, TmaArgs) { std::abort(); }
}  // namespace jmb
