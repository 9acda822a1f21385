/*
 * jmb200_designs.h - C ABI of the design construction that precedes the MCMC path.
 *
 * mvJointModelBayes() builds, in R, the n x K grid of Gauss-Kronrod times st (R/mvJointModelBayes.R:117-134) and then
 *   (1) right_rows(dataL, times, ids, st) (R/supportFuns_mvJointModelBayes.R:16-25; called at R/mvJointModelBayes.R:263,267):
 *       for every subject i and quadrature point k, the row of the long data whose visit is the last one at or before
 *       st[i, k] - mapply(findInterval, ...) over the subjects, ind[ind < 1] <- 1, then data[c(ind), ] - so that
 *       time-varying covariates of the longitudinal submodels are taken at the right visit when XXs / ZZs / Us are built;
 *   (2) splines::splineDesign(knots, x, ord = 4, outer.ok = TRUE) at Time, st and st_int (:382-383,389-392) for W1 / W1s /
 *       W1s_int.
 * Both are per-point, index / byte level work (n K = 1.5e7 points at 1M subjects, one R-level findInterval call per subject):
 * here one thread per point.  Results are bit-exact for (1) and equal to the host path of jmb_create (csrc/jmb_host.h) for (2).
 * Host pointers in and out; `device` is the CUDA ordinal.  No CPU fallback: without a device both fail with a message.
 */
#ifndef JMB200_DESIGNS_H
#define JMB200_DESIGNS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* right_rows: row_ptr [n + 1] CSR pointers of the subjects' visits in the long data (rows grouped by subject, visit times
 * non-decreasing within a subject, as dataL is ordered at R/mvJointModelBayes.R:243), times [row_ptr[n]], Q [n x K]
 * row-major (Q[i * K + k] = st[i, k], i.e. c(t(st))).  out [n * K]: 0-based row of the long data for every (i, k), in the
 * order of c(ind) = subject-major - the row index vector that data[c(ind), ] takes, minus 1.  A subject without visits is an
 * error (the reference indexes NA row names there). */
int jmb_right_rows(int device, int32_t n, int32_t K, const int64_t* row_ptr, const double* times, const double* Q, int64_t* out);

/* right_rows_mstate (R/supportFuns_mvJointModelBayes.R:27-40): the same per transition row; idT [nT] = 0-based subject of
 * each transition row, Q [nT x K]; out [nT * K]. */
int jmb_right_rows_mstate(int device, int32_t n, int32_t nT, int32_t K, const int64_t* row_ptr, const double* times,
                          const int32_t* idT, const double* Q, int64_t* out);

/* splineDesign(knots, x, ord, outer.ok = TRUE) in band form: for every x[i], off[i] = index of the first basis function
 * that can be non-zero and vals[i * ord + r] = value of basis function off[i] + r (r < ord); rows outside
 * [knots[ord - 1], knots[n_knots - ord]] are all zero (outer.ok), the basis is right-closed at the last knot.  With dense != NULL
 * the dense n x (n_knots - ord) column-major matrix is written as well (what W1 / W1s hold in R). */
int jmb_spline_design(int device, const double* knots, int32_t n_knots, int32_t ord, const double* x, int64_t n,
                      int32_t* off, double* vals, double* dense);

/* The data frame the quadrature designs are built from (R/mvJointModelBayes.R:256-268): `dataL.id2 <- right_rows(dataL, ...)`
 * is `data[c(ind), ]` (R/supportFuns_mvJointModelBayes.R:24), followed by `dataL.id2[[timeVar]] <- c(t(st))` (and the same for
 * dataL.id / dataL_int.id2 with Time / TimeL / c(t(st_int))).  X [N x p] column-major = the numeric columns of the long data,
 * rows [M] = the 0-based row indices jmb_right_rows returned, time_col = column that holds timeVar (-1: none), new_time [M] =
 * c(t(st)).  out [M x p] column-major: out[m, c] = X[rows[m], c], except out[m, time_col] = new_time[m].  Bit-exact (a gather).
 * model.matrix() on this frame stays in R: it is formula semantics, not arithmetic. */
int jmb_data_rows(int device, const double* X, int64_t N, int32_t p, const int64_t* rows, int64_t M, int32_t time_col,
                  const double* new_time, double* out);

#ifdef __cplusplus
}
#endif
#endif /* JMB200_DESIGNS_H */
