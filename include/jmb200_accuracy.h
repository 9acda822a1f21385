/*
 * jmb200_accuracy.h - C ABI of the pair / subject sums of the predictive-accuracy measures built on survfitJM
 * ("the aucJM / rocJM / prederrJM loops over survfitJM").
 *
 * aucJM.mvJMbayes (R/aucJM.mvJMbayes.R; paths relative to the JMbayes source tree) obtains the conditional survival
 * probabilities from survfitJM (include/jmb200_survfit.h replaces that loop) and then enumerates ALL pairs of subjects
 * with combn() (:67), builds four indicator vectors over the pairs (:74-77), weights three of them with further survfitJM
 * probabilities (:80-140) and returns sum((pi_i < pi_j) * ind) / sum(ind) (:141) - n (n - 1) / 2 R-level vector
 * elements, 5e11 for the 1M-subject fit.  jmb_aucJM_pairs computes the two sums on the device from the per-subject
 * vectors; the pair weight factorises (the four classes are disjoint):
 *     ind[i, j] = a_i * b_j,   a_i = 1        if T_i <= Thoriz and d_i in {1, 3}      (ind1, ind3)
 *                                  = pi2_i    if T_i <= Thoriz and d_i in {0, 2}      (ind2, ind4; pi2 = 1 - S_i(Thoriz | T_i))
 *                              b_j = 1        if T_j >  Thoriz                         (ind1, ind2)
 *                                  = pi3_j    if T_j <= Thoriz and d_j in {0, 2}      (ind3, ind4; pi3 = S_j(Thoriz | T_j))
 * for i before j in the order of the rows (newdata2 is sorted by Time, :35-38).  NA handling is sum(, na.rm = TRUE)'s: a pair
 * whose weight is NA is dropped from both sums, a pair whose weight is valid but one of whose probabilities is NA is
 * dropped from the numerator only.
 *
 * prederrJM.mvJMbayes (R/prederrJM.mvJMbayes.R:104-108) is one sum over three groups of subjects:
 *     (1 / nr) * sum(L(1 - S_alive), L(0 - S_dead), w * L(1 - S_cens) + (1 - w) * L(0 - S_cens), na.rm = TRUE)
 * with L = abs or square (:12-16).  Conventions as in jmb200.h; no CPU fallback.
 */
#ifndef JMB200_ACCURACY_H
#define JMB200_ACCURACY_H

#include "jmb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Time, event, pi_u_t, pi2, pi3: [n] host vectors in the order of newdata2's subjects (pi2 is read only where
 * T_i <= Thoriz and d_i in {0, 2}, pi3 likewise; NaN = NA).  out[0] = auc (NaN when n < 2, :142-144), out[1] = numerator,
 * out[2] = denominator.  kernel_ms may be NULL. */
int jmb_aucJM_pairs(int device, int32_t n, const double* Time, const double* event, double Thoriz, const double* pi_u_t,
                    const double* pi2, const double* pi3, double out[3], float* kernel_ms);

/* loss: 0 = absolute, 1 = square (:12-16).  S_cens / w_cens may be NULL when n_cens = 0.  out[0] = prederr. */
int jmb_prederrJM_sum(int device, int32_t loss, int32_t nr, int32_t n_alive, const double* S_alive, int32_t n_dead,
                      const double* S_dead, int32_t n_cens, const double* S_cens, const double* w_cens, double* out);

/* rocJM.mvJMbayes (R/rocJM.mvJMbayes.R:68-95): ind_i = 1 for subjects who died before Thoriz (T_i < Thoriz, d_i in {1, 3}),
 * pi2_i = 1 - S_i(Thoriz | T_i) for those censored before it (d_i in {0, 2}), 0 otherwise; then for every threshold
 *     nTP[k] = sum_i (pi_i < thrs[k]) ind_i,  nFP[k] = sum_i (pi_i < thrs[k]) (1 - ind_i),  Q[k] = mean_i (pi_i < thrs[k])
 * - the column sums of the n x 101 matrix outer(pi.u.t, thrs, "<") (:89,92,95) - and sum_ind = { sum(ind), sum(1 - ind) } (:90-93), each
 * accumulated in the order of nTP / nFP so that nFN and nTN cancel exactly as they do in R.  No
 * na.rm in the reference: NaN inputs propagate.  Everything after these sums (:90-105) is arithmetic on 101 numbers and
 * stays with the caller.  nTP, nFP, Q: [n_thr]; sum_ind: [2]. */
int jmb_rocJM_sums(int device, int32_t n, const double* Time, const double* event, double Thoriz, const double* pi_u_t,
                   const double* pi2, int32_t n_thr, const double* thrs, double* nTP, double* nFP, double* Q, double* sum_ind);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_ACCURACY_H */
