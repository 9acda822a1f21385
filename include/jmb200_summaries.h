/*
 * jmb200_summaries.h - C ABI of the posterior summaries of mvJointModelBayes().
 *
 * Replaces the ten `summary_fun(FUN)` sweeps of R/mvJointModelBayes.R:925-971 (paths relative to the JMbayes source tree),
 * which call an R closure once per series and statistic - 10 x n x q calls for the random effects `mcmc$b` (n x q x M), i.e.
 * 6e7 R-level calls for the 1M-subject fit.  A series is the vector of the M kept draws of one scalar: a column of an
 * M x p matrix (Bs_gammas, gammas, alphas, betas ...), a cell of an M x q x q array (D, inv_D) or a cell (i, j) of `b`
 * (`apply(x, dd, fun)` with dd = 2 / c(2, 3) / c(1, 2), :929-933).  One launch computes, per series:
 *   postMeans      mean(x, na.rm = TRUE)                                                             :960
 *   postwMeans     wmean: sum(x * weights)                                                           :948,961
 *   postModes      modes (R/modes.R:1-5): density(y, bw = "nrd", adjust = 3, n = 1000), x at the maximum   :962
 *   EffectiveSize  effectiveSize (R/effectiveSize.R:1-24): n var(x) / spectrum0.ar(x), ar(x, aic = TRUE)   :963
 *   StDev          sd(x, na.rm = TRUE)                                                               :964
 *   wStDev         wsd: sqrt(cov.wt(cbind(x), wt = weights)$cov)                                     :949,965
 *   StErr          stdErr (R/stdErr.R:1-7): sqrt(var / effectiveSize)                                :966
 *   CIs            quantile(x, probs = c(0.025, 0.975), na.rm = TRUE) (type 7)                       :967
 *   wCIs           Hmisc::wtd.quantile(x, weights, probs, type = "i/(n+1)", na.rm = TRUE)            :968-970
 *   Pvalues        computeP (R/computeP.R:1-6): 2 min(mean(x >= 0), mean(x < 0))                     :971
 * density() is the algorithm of R <= 4.3, the releases the reference was written against (from R 4.4.0 on:
 * density(old.coords = TRUE); the new default grid changes the estimate by a factor of about 0.999 and can move a mode by
 * one cell of the 1000-point grid).
 * and jmb_stand_weights is stand() (:940-945).  Draws are assumed free of NA (the chains never produce one); a series that
 * holds a NaN yields NaN statistics.  Conventions as in jmb200.h: plain pointers and sizes, caller-allocated outputs, 0 on
 * success / jmb_last_error() otherwise, no CPU fallback.
 */
#ifndef JMB200_SUMMARIES_H
#define JMB200_SUMMARIES_H

#include "jmb200.h"

#ifdef __cplusplus
extern "C" {
#endif

#define JMB_STATS_MAX_M 1024   /* kept draws per series */

/* Outputs, each [S] (CIs / wCIs: 2 x S column-major, i.e. lower and upper of series s at [2 s], [2 s + 1]); any may be
 * NULL.  Host pointers unless the call says the buffers live on the device. */
typedef struct jmb_stats_out {
    double* postMeans;
    double* postwMeans;
    double* postModes;
    double* EffectiveSize;
    double* StDev;
    double* wStDev;
    double* StErr;
    double* CIs;
    double* wCIs;
    double* Pvalues;
} jmb_stats_out;

/* weights <- stand(LogLiks) (:940-947); host arithmetic on M numbers. */
int jmb_stand_weights(const double* LogLiks, int32_t M, double* weights);

/*
 * Element m of series s is x[s * series_stride + m * draw_stride]:
 *   mcmc$b (n x q x M)            S = n q, series_stride = 1,  draw_stride = n q
 *   an M x p matrix of draws      S = p,   series_stride = M,  draw_stride = 1
 *   an M x q x q array            S = q q, series_stride = M,  draw_stride = 1
 * weights: [M] (stand(LogLiks)); probs: the two quantile levels (0.025, 0.975).
 * on_device = 0: x and the outputs are host buffers; the library streams x through the device in chunks of series
 * (copies overlapped with the kernel).  on_device = 1: x and every non-NULL output are device pointers on `device`.
 * kernel_ms (may be NULL): device time of the kernel launches of this call.
 */
int jmb_mcmc_statistics(int device, const double* x, int64_t S, int32_t M, int64_t series_stride, int64_t draw_stride,
                        const double* weights, const double probs[2], int on_device, jmb_stats_out* out, float* kernel_ms);

/* jmb_mcmc_statistics keeps its device buffers (staging, scratch, outputs) and streams per device between calls; this frees
 * them (additive: the reference has no such state). */
int jmb_stats_release(void);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_SUMMARIES_H */
