/*
 * jmb_oracle_svft.h - CPU restatement of survfitJM.mvJMbayes()'s per-subject loops (see jmb_oracle_svft.c).
 * TEST INFRASTRUCTURE ONLY.  The input / output structs are the public ones of include/jmb200_survfit.h
 * (the oracle reads the reference's dense designs straight from them).
 */
#ifndef JMB_ORACLE_SVFT_H
#define JMB_ORACLE_SVFT_H
#include "../include/jmb200_survfit.h"
#ifdef __cplusplus
extern "C" {
#endif
int orc_svft_modes(const jmb_svft_data* d, const jmb_svft_params* post_means, const jmb_svft_control* ct, jmb_svft_out* out);
int orc_svft_sample(const jmb_svft_data* d, const jmb_svft_params* draws, const jmb_svft_control* ct, jmb_svft_out* out);
int orc_survfitJM(const jmb_svft_data* d, const jmb_svft_params* post_means, const jmb_svft_params* draws,
                  const jmb_svft_control* ct, jmb_svft_out* out);
int orc_svft_modes_range(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_control* ct, jmb_svft_out* out,
                         int i_begin, int i_end);
int orc_svft_sample_range(const jmb_svft_data* d, const jmb_svft_params* pd, const jmb_svft_control* ct, jmb_svft_out* out,
                          int i_begin, int i_end);
int orc_survfitJM_range(const jmb_svft_data* d, const jmb_svft_params* pm, const jmb_svft_params* pd,
                        const jmb_svft_control* ct, jmb_svft_out* out, int i_begin, int i_end);
double orc_svft_log_post(const jmb_svft_data* d, const jmb_svft_params* P, int m, int i, const double* b);
double orc_svft_surv_pred(const jmb_svft_data* d, const jmb_svft_params* P, int m, int i, int l, const double* b);
int orc_sv_inverse(const double* A, int n, double* inv);
void orc_sv_eigen_sym(const double* S, int n, double* val, double* vec);
void orc_sv_summary(const double* x, int M, const double* probs, double* out4);
#ifdef __cplusplus
}
#endif
#endif
