// ref_driver.cpp - C entry points around the reference's own functions (compiled from
// /root/reference/src/*.cpp against oracle/ref_shim/RcppArmadillo.h).  TEST INFRASTRUCTURE ONLY.
//
// The driver (1) turns the flat structs of jmb_oracle.h back into the R lists the reference
// reads by name, with R's 1-based indices where the reference subtracts 1
// (List2Field_uvec, src/fast_LAPRWM.cpp:45-56), and (2) provides the random numbers: the
// reference draws from R's RNG in its own order (all accept uniforms up front :617-618, proposal
// normals per block :658-661, Rf_rgamma per iteration :781-832); the provider below hands out,
// in that order, exactly the values the DESIGN.md counter streams assign to (chain, iteration,
// subject, component), so that the reference's chain can be compared draw for draw with the
// restated oracle and with the CUDA path.
#include <RcppArmadillo.h>

#include <cstdint>
#include <string>

#include "jmb_oracle.h"

using namespace Rcpp;

// ---- the reference's exported functions (defined in the reference sources) -------------------
List lap_rwm_C(List initials, List Data, List priors, List scales, List Covs, List control,
               bool interval_cens, bool multiState);
List lap_rwm_C_woRE(List initials, List Data, List priors, List scales, List Covs, List control);
List lap_rwm_C_woRE_nogammas(List initials, List Data, List priors, List scales, List Covs, List control);
arma::vec log_post_RE_svft(arma::vec b, List Data);      // src/survfitJM_mvJMbayes.cpp:190
arma::vec survPred_svft_2(arma::vec b, List Data);       // src/survfitJM_mvJMbayes.cpp:241
double logPosterior(double temp, arma::vec event, arma::uvec idGK_fast, arma::mat W1, arma::mat W1s, arma::vec Bs_gammas,
                    arma::mat W2, arma::mat W2s, arma::vec gammas, arma::mat Wlong, arma::mat Wlongs, arma::vec alphas,
                    arma::vec Pw, arma::vec mean_Bs_gammas, arma::mat Tau_Bs_gammas, arma::vec mean_gammas,
                    arma::mat Tau_gammas, arma::vec mean_alphas, arma::mat Tau_alphas, double tauBs, double A_tauBs,
                    double B_tauBs);
arma::vec gradient_logPosterior(double temp, arma::vec event, arma::uvec idGK_fast, arma::vec event_colSumsW1,
                                arma::mat W1s, arma::vec Bs_gammas, arma::vec event_colSumsW2, arma::mat W2s,
                                arma::vec gammas, arma::vec event_colSumsWlong, arma::mat Wlongs, arma::vec alphas,
                                arma::vec Pw, arma::vec mean_Bs_gammas, arma::mat Tau_Bs_gammas, arma::vec mean_gammas,
                                arma::mat Tau_gammas, arma::vec mean_alphas, arma::mat Tau_alphas, double tauBs,
                                double A_tauBs, double B_tauBs);

// ---- RNG provider -----------------------------------------------------------------------------
jmbref_fill_fn jmbref_fill = nullptr;
jmbref_rgamma_fn jmbref_rgamma = nullptr;

namespace {
struct Provider {
    uint32_t seed = 0, chain = 0;
    long long offset = 0;
    int n = 0, q = 0, B = 0, p = 0, A = 0, n_block = 1, draws_per_iter = 1;
    long long randu_calls = 0, randn_calls = 0, gamma_calls = 0;
} g;

enum { ST_B_PROP = 0, ST_B_ACC = 1, ST_SURV = 2 };

double normal_at(uint32_t c0, uint32_t iter, int comp, uint32_t stream) {
    double z[2];
    orc_rnorm_pair(g.seed, g.chain, c0, iter, (uint32_t)(comp >> 1), stream, z);
    return z[comp & 1];
}

void provider_fill(int kind, unsigned long long rows, unsigned long long cols, double* out) {
    if (kind == 0) {
        uint32_t r[4];
        if (g.randu_calls == 0) {            // log_us: one uniform per iteration (:617)
            for (unsigned long long it = 0; it < rows * cols; ++it) {
                orc_philox4x32(g.seed, g.chain, 0u, (uint32_t)it, 0xFFFFFFFFu, ST_SURV, r);
                out[it] = orc_u01(r[0], r[1]);
            }
        } else {                             // log_us_RE: n x iterations (:618)
            for (unsigned long long it = 0; it < cols; ++it)
                for (unsigned long long m = 0; m < rows; ++m) {
                    orc_philox4x32(g.seed, g.chain, (uint32_t)(g.offset + (long long)m), (uint32_t)it, 0u, ST_B_ACC, r);
                    out[m + it * rows] = orc_u01(r[0], r[1]);
                }
        }
        g.randu_calls += 1;
        return;
    }
    // randn: per outer block n calls of (n_block x q) for mvrnorm_cube (:123-136,658), then
    // (n_block x B), (n_block x p), (n_block x A) for the survival proposals (:659-661)
    // g.n == -1: lap_rwm_C_woRE_nogammas draws two proposal blocks (Bs_gammas, alphas) per outer block
    const bool nogam = g.n < 0;
    const long long nsub = nogam ? 0 : g.n;
    const long long per_block = nsub + (nogam ? 2 : 3);
    const long long blk = g.randn_calls / per_block;
    long long pos = g.randn_calls % per_block;
    if (nogam && pos == 1) pos = 2;          // second call = alphas
    for (unsigned long long c = 0; c < cols; ++c)
        for (unsigned long long r2 = 0; r2 < rows; ++r2) {
            const uint32_t iter = (uint32_t)(blk * g.n_block + (long long)r2);
            double v;
            if (pos < nsub) v = normal_at((uint32_t)(g.offset + pos), iter, (int)c, ST_B_PROP);
            else if (pos == nsub) v = normal_at(0u, iter, (int)c, ST_SURV);
            else if (pos == nsub + 1) v = normal_at(0u, iter, g.B + (int)c, ST_SURV);
            else v = normal_at(0u, iter, g.B + g.p + (int)c, ST_SURV);
            out[r2 + c * rows] = v;
        }
    g.randn_calls += 1;
}

double provider_rgamma(double shape, double scale) {
    const long long iter = g.gamma_calls / g.draws_per_iter, draw = g.gamma_calls % g.draws_per_iter;
    g.gamma_calls += 1;
    return orc_rgamma(g.seed, g.chain, (uint32_t)draw, (uint32_t)iter, shape, scale);
}

// ---- R-list builders ----------------------------------------------------------------------------
ValuePtr num_vec(const double* x, size_t len) { auto v = std::make_shared<Value>(); v->kind = Value::NUM; if (len) v->num.assign(x, x + len); return v; }
ValuePtr num_mat(const double* x, arma::uword r, arma::uword c) { auto v = num_vec(x, (size_t)(r * c)); v->dim = {r, c}; return v; }
ValuePtr num_arr3(const double* x, arma::uword a, arma::uword b, arma::uword c) { auto v = num_vec(x, (size_t)(a * b * c)); v->dim = {a, b, c}; return v; }
ValuePtr scalar(double x) { return num_vec(&x, 1); }
ValuePtr int_vec(const int32_t* x, size_t len, int add) { auto v = std::make_shared<Value>(); v->kind = Value::INT; for (size_t i = 0; i < len; ++i) v->ints.push_back((long long)x[i] + add); return v; }
ValuePtr seq_vec(long long from, long long len, long long each = 1) { auto v = std::make_shared<Value>(); v->kind = Value::INT; for (long long i = 0; i < len; ++i) for (long long e = 0; e < each; ++e) v->ints.push_back(from + i); return v; }
ValuePtr str_vec(const std::vector<std::string>& s) { auto v = std::make_shared<Value>(); v->kind = Value::STR; v->str = s; return v; }
ValuePtr lgl_vec(const std::vector<int>& s) { auto v = std::make_shared<Value>(); v->kind = Value::LGL; v->lgl = s; return v; }
ValuePtr list_of(const std::vector<ValuePtr>& xs) { List l; for (auto& x : xs) l.push("", x); return wrap(l); }

const char* FAM[] = {"gaussian", "binomial", "poisson"};
const char* LINK[] = {"identity", "logit", "probit", "cloglog", "log"};
const char* TF[] = {"identity", "expit", "exp", "log", "log2", "log10", "sqrt"};

List build_data(const orc_data* d, const orc_draw* w) {
    const int n = d->n, O = d->n_outcomes, T = d->n_terms, K = d->K, B = d->n_Bs, p = d->n_gammas;
    const bool ms = d->multi_state != 0;
    const int nT = ms ? d->nT : n;            // rows of the survival part (R/mvJointModelBayes.R:76-101)
    const arma::uword ns = (arma::uword)nT * K;
    // Data$idT2 / idT2s (1-based subject of each transition / quadrature row, R/mvJointModelBayes.R:559-565)
    auto idT_vec = [&](int each) {
        auto v = std::make_shared<Value>(); v->kind = Value::INT;
        for (int r = 0; r < nT; ++r) for (int e = 0; e < each; ++e) v->ints.push_back((long long)(ms ? d->idT[r] : r) + 1);
        return v;
    };
    List D;
    std::vector<ValuePtr> y, Xb, Z, REi, idL, idL2, sig;
    std::vector<std::string> fams, links, tfs;
    for (int k = 0; k < O; ++k) {
        const size_t N = (size_t)d->N[k];
        y.push_back(num_vec(d->y[k], N)); Xb.push_back(num_vec(w->Xbetas[k], N));
        Z.push_back(num_mat(d->Z[k], N, d->q_k[k]));
        REi.push_back(int_vec(d->RE_inds[k], d->q_k[k], 1));          // 1-based in R
        idL.push_back(int_vec(d->idL[k], N, 1));
        idL2.push_back(int_vec(d->idL2[k], n, 0));                     // already 0-based (:449)
        sig.push_back(scalar(w->sigmas[k]));
        fams.push_back(FAM[d->fams[k]]); links.push_back(LINK[d->links[k]]);
    }
    std::vector<ValuePtr> RE2, cols, XXb, XXsb, XXsbi, ZZ, ZZs, ZZsi, U, Us, Usi, idT, idTs, rowsw, rowsws;
    for (int t = 0; t < T; ++t) {
        RE2.push_back(int_vec(d->RE_inds2[t], d->r_t[t], 1)); cols.push_back(int_vec(d->col_inds[t], d->u_t[t], 1));
        XXb.push_back(num_vec(w->XXbetas[t], nT)); XXsb.push_back(num_vec(w->XXsbetas[t], ns));
        ZZ.push_back(num_mat(d->ZZ[t], nT, d->r_t[t])); ZZs.push_back(num_mat(d->ZZs[t], ns, d->r_t[t]));
        U.push_back(num_mat(d->U[t], nT, d->u_t[t])); Us.push_back(num_mat(d->Us[t], ns, d->u_t[t]));
        idT.push_back(idT_vec(1)); idTs.push_back(idT_vec(K));
        tfs.push_back(TF[d->trans_Funs[t]]);
        if (d->interval_cens) {
            XXsbi.push_back(num_vec(w->XXsbetas_int[t], ns)); ZZsi.push_back(num_mat(d->ZZs_int[t], ns, d->r_t[t]));
            Usi.push_back(num_mat(d->Us_int[t], ns, d->u_t[t]));
        }
    }
    // rows_wlong / rows_wlongs = tapply(seq_along(idT), idT, c) (:569-570,580-581); idT_rsum = last row of each subject
    std::vector<long long> rsum;
    for (int r = 0, i = 0; r < nT; ++i) {
        int r1 = r;
        while (r1 + 1 < nT && (ms ? d->idT[r1 + 1] == d->idT[r] : false)) ++r1;
        rowsw.push_back(seq_vec(r + 1, r1 - r + 1)); rowsws.push_back(seq_vec((long long)r * K + 1, (long long)(r1 - r + 1) * K));
        rsum.push_back(r1);
        r = r1 + 1;
    }
    auto rsum_v = std::make_shared<Value>(); rsum_v->kind = Value::INT; rsum_v->ints = rsum;
    D.push("y", list_of(y)); D.push("Xbetas", list_of(Xb)); D.push("Z", list_of(Z)); D.push("RE_inds", list_of(REi));
    D.push("RE_inds2", list_of(RE2)); D.push("sigmas", list_of(sig)); D.push("invD", num_mat(w->invD, d->q, d->q));
    D.push("idL", list_of(idL)); D.push("idL2", list_of(idL2)); D.push("fams", str_vec(fams)); D.push("links", str_vec(links));
    D.push("trans_Funs", str_vec(tfs)); D.push("event", num_vec(d->event, nT));
    D.push("idGK_fast", int_vec(d->idGK_fast, nT, 0)); D.push("idT_rsum", rsum_v);
    D.push("W1", num_mat(d->W1, nT, B)); D.push("W1s", num_mat(d->W1s, ns, B));
    D.push("W2", num_mat(d->W2, nT, p)); D.push("W2s", num_mat(d->W2s, ns, p));
    D.push("U", list_of(U)); D.push("Us", list_of(Us)); D.push("row_inds_U", seq_vec(1, nT)); D.push("row_inds_Us", seq_vec(1, (long long)ns));
    D.push("col_inds", list_of(cols)); D.push("XXbetas", list_of(XXb)); D.push("XXsbetas", list_of(XXsb));
    D.push("ZZ", list_of(ZZ)); D.push("ZZs", list_of(ZZs)); D.push("idT", list_of(idT)); D.push("idTs", list_of(idTs));
    D.push("idT2", list_of(idT)); D.push("idT2s", list_of(idTs)); D.push("Pw", num_vec(d->Pw, ns));
    D.push("rows_wlong", list_of(rowsw)); D.push("rows_wlongs", list_of(rowsws));
    if (ms) {                              // R/mvJointModelBayes.R:667-669
        std::vector<double> kf(d->nRisks), kl(d->nRisks);
        for (int j = 0; j < d->nRisks; ++j) { kf[j] = d->kn_strat_first[j]; kl[j] = d->kn_strat_last[j]; }
        D.push("nRisks", scalar((double)d->nRisks));
        D.push("kn_strat_last", num_vec(kl.data(), kl.size())); D.push("kn_strat_first", num_vec(kf.data(), kf.size()));
    } else {
        D.push("nRisks", scalar(1.0));
        double kl = B - 1, kf = 0;
        D.push("kn_strat_last", num_vec(&kl, 1)); D.push("kn_strat_first", num_vec(&kf, 1));
    }
    if (d->interval_cens) {
        std::vector<int> l1(n), l01(n), l2(n), l3(n);
        for (int i = 0; i < n; ++i) {      // R/mvJointModelBayes.R:694-696
            const double e = d->event[i];
            l1[i] = e == 1; l01[i] = (e == 1 || e == 0); l2[i] = e == 2; l3[i] = e == 3;
        }
        D.push("Levent1", lgl_vec(l1)); D.push("Levent01", lgl_vec(l01)); D.push("Levent2", lgl_vec(l2)); D.push("Levent3", lgl_vec(l3));
        D.push("W1s_int", num_mat(d->W1s_int, ns, B)); D.push("W2s_int", num_mat(d->W2s_int, ns, p));
        D.push("Us_int", list_of(Usi)); D.push("XXsbetas_int", list_of(XXsbi)); D.push("ZZs_int", list_of(ZZsi));
        D.push("Pw_int", num_vec(d->Pw_int, ns));
    } else {                               // R/mvJointModelBayes.R:703-713
        D.push("Levent1", lgl_vec({})); D.push("Levent01", lgl_vec({})); D.push("Levent2", lgl_vec({})); D.push("Levent3", lgl_vec({}));
        D.push("W1s_int", num_mat(nullptr, 0, 0)); D.push("W2s_int", num_mat(nullptr, 0, 0));
        D.push("Us_int", list_of({num_mat(nullptr, 0, 0)})); D.push("XXsbetas_int", list_of({num_vec(nullptr, 0)}));
        D.push("ZZs_int", list_of({num_mat(nullptr, 0, 0)})); D.push("Pw_int", num_vec(nullptr, 0));
    }
    return D;
}

void copy_out(const ValuePtr& v, double* dst) { if (dst && v && !v->num.empty()) std::memcpy(dst, v->num.data(), sizeof(double) * v->num.size()); }
thread_local std::string g_error;
}  // namespace

extern "C" const char* jmbref_last_error() { return g_error.c_str(); }

extern "C" int jmbref_lap_rwm(const orc_data* d, const orc_draw* w, const orc_priors* pr, const orc_control* ct,
                              const orc_chain_in* in, orc_chain_out* out) {
    try {
        const int n = d->n, q = d->q, B = d->n_Bs, p = d->n_gammas, A = d->n_alphas;
        List Data = build_data(d, w);
        List initials, priors, scales, Covs, control;
        initials.push("b", num_mat(in->b, n, q)); initials.push("Bs_gammas", num_vec(in->Bs_gammas, B));
        initials.push("gammas", num_vec(in->gammas, p)); initials.push("alphas", num_vec(in->alphas, A));
        initials.push("phi_gammas", num_vec(in->phi_gammas, p)); initials.push("phi_alphas", num_vec(in->phi_alphas, A));
        initials.push("tau_td_alphas", num_vec(in->tau_td_alphas, pr->n_td)); initials.push("tau_Bs_gammas", num_vec(in->tau_Bs_gammas, d->multi_state ? d->nRisks : 1));
        initials.push("tau_gammas", scalar(in->tau_gammas)); initials.push("tau_alphas", scalar(in->tau_alphas));
        priors.push("mean_Bs_gammas", num_vec(pr->mean_Bs_gammas, B)); priors.push("Tau_Bs_gammas", num_mat(pr->Tau_Bs_gammas, B, B));
        priors.push("mean_gammas", num_vec(pr->mean_gammas, p)); priors.push("Tau_gammas", num_mat(pr->Tau_gammas, p, p));
        priors.push("mean_alphas", num_vec(pr->mean_alphas, A)); priors.push("Tau_alphas", num_mat(pr->Tau_alphas, A, A));
#define SC(name) priors.push(#name, scalar((double)pr->name))
        SC(A_tau_Bs_gammas); SC(B_tau_Bs_gammas); SC(rank_Tau_Bs_gammas); SC(shrink_gammas); SC(A_tau_gammas); SC(B_tau_gammas);
        SC(rank_Tau_gammas); SC(A_phi_gammas); SC(B_phi_gammas); SC(shrink_alphas); SC(double_gamma_alphas); SC(A_tau_alphas);
        SC(B_tau_alphas); SC(rank_Tau_alphas); SC(A_phi_alphas); SC(B_phi_alphas); SC(A_xi_alphas); SC(B_xi_alphas);
        SC(A_nu_alphas); SC(B_nu_alphas); SC(rank_Tau_td_alphas);
#undef SC
        std::vector<ValuePtr> td;
        for (int k = 0; k < pr->n_td; ++k) td.push_back(int_vec(pr->td_cols[k], pr->td_len[k], 1));
        priors.push("td_cols", list_of(td));
        scales.push("b", num_vec(in->sigma_b, n)); scales.push("Bs_gammas", scalar(in->sigma_Bs_gammas));
        scales.push("gammas", scalar(in->sigma_gammas)); scales.push("alphas", scalar(in->sigma_alphas));
        Covs.push("b", num_arr3(d->Covs_b, q, q, n)); Covs.push("Bs_gammas", num_mat(in->Sigma_Bs_gammas, B, B));
        Covs.push("gammas", num_mat(in->Sigma_gammas, p, p)); Covs.push("alphas", num_mat(in->Sigma_alphas, A, A));
#define CT(name) control.push(#name, scalar((double)ct->name))
        CT(n_iter); CT(n_burnin); CT(n_block); CT(n_thin); CT(target_acc); CT(eps1); CT(eps2); CT(eps3); CT(c0); CT(c1); CT(adaptCov);
        CT(seed); CT(chain_id);      // control$seed exists in R; chain_id = index of the draw (read by the CUDA glue only)
#undef CT
        g = Provider();
        g.seed = ct->seed; g.chain = ct->chain_id; g.offset = d->subject_offset;
        g.n = n; g.q = q; g.B = B; g.p = p; g.A = A; g.n_block = ct->n_block;
        g.draws_per_iter = (d->multi_state ? d->nRisks : 1) + pr->n_td + (pr->shrink_gammas ? p + 1 : 0) +
                           (pr->shrink_alphas ? A + 1 + (pr->double_gamma_alphas ? A + 1 : 0) : 0);
        jmbref_fill = provider_fill; jmbref_rgamma = provider_rgamma;
        List res = lap_rwm_C(initials, Data, priors, scales, Covs, control, d->interval_cens != 0, d->multi_state != 0);
        List mc = as<List>(res["mcmc"]);
        copy_out(mc["b"].v, out->b); copy_out(mc["Bs_gammas"].v, out->Bs_gammas); copy_out(mc["gammas"].v, out->gammas);
        copy_out(mc["alphas"].v, out->alphas); copy_out(mc["phi_gammas"].v, out->phi_gammas); copy_out(mc["phi_alphas"].v, out->phi_alphas);
        copy_out(mc["tau_Bs_gammas"].v, out->tau_Bs_gammas); copy_out(mc["tau_gammas"].v, out->tau_gammas);
        copy_out(mc["tau_alphas"].v, out->tau_alphas); copy_out(mc["tau_td_alphas"].v, out->tau_td_alphas);
        copy_out(res["logWeights"].v, out->logWeights);
        List sc = as<List>(res["scales"]); List sg = as<List>(sc["sigma"]); List Sg = as<List>(sc["Sigma"]);
        copy_out(sg["b"].v, out->sigma_b);
        if (out->sigma_surv) { out->sigma_surv[0] = as<double>(sg["Bs_gammas"]); out->sigma_surv[1] = as<double>(sg["gammas"]); out->sigma_surv[2] = as<double>(sg["alphas"]); }
        copy_out(Sg["Bs_gammas"].v, out->Sigma_Bs_gammas); copy_out(Sg["gammas"].v, out->Sigma_gammas); copy_out(Sg["alphas"].v, out->Sigma_alphas);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

static arma::vec V(const double* x, size_t n) { arma::vec v(n); for (size_t i = 0; i < n; ++i) v.mem[i] = x[i]; return v; }
static arma::mat M(const double* x, size_t r, size_t c) { arma::mat m(r, c); for (size_t i = 0; i < r * c; ++i) m.mem[i] = x[i]; return m; }

extern "C" double jmbref_logPosterior(double temp, int32_t n, int64_t ns, int32_t B, int32_t p, int32_t A, const double* event,
                                      const double* W1, const double* W1s, const double* Bs, const double* W2, const double* W2s,
                                      const double* gam, const double* Wlong, const double* Wlongs, const double* alp,
                                      const double* Pw, const double* mean_Bs, const double* Tau_Bs, const double* mean_g,
                                      const double* Tau_g, const double* mean_a, const double* Tau_a, double tauBs,
                                      double A_tauBs, double B_tauBs) {
    return logPosterior(temp, V(event, n), arma::uvec(), M(W1, n, B), M(W1s, ns, B), V(Bs, B), M(W2, n, p), M(W2s, ns, p), V(gam, p),
                        M(Wlong, n, A), M(Wlongs, ns, A), V(alp, A), V(Pw, ns), V(mean_Bs, B), M(Tau_Bs, B, B), V(mean_g, p),
                        M(Tau_g, p, p), V(mean_a, A), M(Tau_a, A, A), tauBs, A_tauBs, B_tauBs);
}

extern "C" void jmbref_gradient_logPosterior(double temp, int64_t ns, int32_t B, int32_t p, int32_t A, const double* ecs_W1,
                                             const double* W1s, const double* Bs, const double* ecs_W2, const double* W2s,
                                             const double* gam, const double* ecs_Wlong, const double* Wlongs, const double* alp,
                                             const double* Pw, const double* mean_Bs, const double* Tau_Bs, const double* mean_g,
                                             const double* Tau_g, const double* mean_a, const double* Tau_a, double tauBs,
                                             double A_tauBs, double B_tauBs, double* out) {
    arma::vec r = gradient_logPosterior(temp, arma::vec(), arma::uvec(), V(ecs_W1, B), M(W1s, ns, B), V(Bs, B), V(ecs_W2, p),
                                        M(W2s, ns, p), V(gam, p), V(ecs_Wlong, A), M(Wlongs, ns, A), V(alp, A), V(Pw, ns),
                                        V(mean_Bs, B), M(Tau_Bs, B, B), V(mean_g, p), M(Tau_g, p, p), V(mean_a, A), M(Tau_a, A, A),
                                        tauBs, A_tauBs, B_tauBs);
    for (arma::uword i = 0; i < r.n_elem; ++i) out[i] = r.mem[i];
}

// lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (src/fast_LAPRWM.cpp:934-1297)
extern "C" int jmbref_lap_rwm_woRE(const orc_data* d, const double* Wlong, const double* Wlongs, const orc_priors* pr,
                                   const orc_control* ct, const orc_chain_in* in, orc_chain_out* out, double* logPost) {
    try {
        const int n = d->n, B = d->n_Bs, p = d->n_gammas, A = d->n_alphas, K = d->K;
        const arma::uword ns = (arma::uword)n * K;
        List Data, initials, priors, scales, Covs, control;
        Data.push("event", num_vec(d->event, n)); Data.push("idGK_fast", int_vec(d->idGK_fast, n, 0));
        Data.push("W1", num_mat(d->W1, n, B)); Data.push("W1s", num_mat(d->W1s, ns, B));
        Data.push("W2", num_mat(d->W2, n, p)); Data.push("W2s", num_mat(d->W2s, ns, p));
        Data.push("Wlong", num_mat(Wlong, n, A)); Data.push("Wlongs", num_mat(Wlongs, ns, A)); Data.push("Pw", num_vec(d->Pw, ns));
        initials.push("Bs_gammas", num_vec(in->Bs_gammas, B)); initials.push("gammas", num_vec(in->gammas, p));
        initials.push("alphas", num_vec(in->alphas, A)); initials.push("phi_gammas", num_vec(in->phi_gammas, p));
        initials.push("phi_alphas", num_vec(in->phi_alphas, A)); initials.push("tau_Bs_gammas", scalar(in->tau_Bs_gammas[0]));
        initials.push("tau_gammas", scalar(in->tau_gammas)); initials.push("tau_alphas", scalar(in->tau_alphas));
        priors.push("mean_Bs_gammas", num_vec(pr->mean_Bs_gammas, B)); priors.push("Tau_Bs_gammas", num_mat(pr->Tau_Bs_gammas, B, B));
        priors.push("mean_gammas", num_vec(pr->mean_gammas, p)); priors.push("Tau_gammas", num_mat(pr->Tau_gammas, p, p));
        priors.push("mean_alphas", num_vec(pr->mean_alphas, A)); priors.push("Tau_alphas", num_mat(pr->Tau_alphas, A, A));
#define SC(name) priors.push(#name, scalar((double)pr->name))
        SC(A_tau_Bs_gammas); SC(B_tau_Bs_gammas); SC(rank_Tau_Bs_gammas); SC(shrink_gammas); SC(A_tau_gammas); SC(B_tau_gammas);
        SC(rank_Tau_gammas); SC(A_phi_gammas); SC(B_phi_gammas); SC(shrink_alphas); SC(A_tau_alphas);
        SC(B_tau_alphas); SC(rank_Tau_alphas); SC(A_phi_alphas); SC(B_phi_alphas);
#undef SC
        scales.push("Bs_gammas", scalar(in->sigma_Bs_gammas)); scales.push("gammas", scalar(in->sigma_gammas));
        scales.push("alphas", scalar(in->sigma_alphas));
        Covs.push("Bs_gammas", num_mat(in->Sigma_Bs_gammas, B, B)); Covs.push("gammas", num_mat(in->Sigma_gammas, p, p));
        Covs.push("alphas", num_mat(in->Sigma_alphas, A, A));
#define CT(name) control.push(#name, scalar((double)ct->name))
        CT(n_iter); CT(n_burnin); CT(n_block); CT(n_thin); CT(target_acc); CT(eps1); CT(eps2); CT(eps3); CT(c0); CT(c1); CT(adaptCov);
        CT(seed); CT(chain_id);      // control$seed exists in R; chain_id = index of the draw (read by the CUDA glue only)
#undef CT
        g = Provider();
        g.seed = ct->seed; g.chain = ct->chain_id; g.n = 0; g.B = B; g.p = p; g.A = A; g.n_block = ct->n_block;
        g.draws_per_iter = 1 + (pr->shrink_gammas ? p + 1 : 0) + (pr->shrink_alphas ? A + 1 : 0);
        jmbref_fill = provider_fill; jmbref_rgamma = provider_rgamma;
        List res;
        if (p > 0) res = lap_rwm_C_woRE(initials, Data, priors, scales, Covs, control);
        else {
            // the _nogammas flavour draws only two proposal blocks per outer block
            g.n = -1;
            res = lap_rwm_C_woRE_nogammas(initials, Data, priors, scales, Covs, control);
        }
        List mc = as<List>(res["mcmc"]);
        copy_out(mc["Bs_gammas"].v, out->Bs_gammas); copy_out(mc["alphas"].v, out->alphas);
        copy_out(mc["phi_alphas"].v, out->phi_alphas); copy_out(mc["tau_Bs_gammas"].v, out->tau_Bs_gammas);
        copy_out(mc["tau_alphas"].v, out->tau_alphas);
        if (p > 0) { copy_out(mc["gammas"].v, out->gammas); copy_out(mc["phi_gammas"].v, out->phi_gammas); copy_out(mc["tau_gammas"].v, out->tau_gammas); }
        copy_out(res["logPost"].v, logPost);
        List sc = as<List>(res["scales"]); List sg = as<List>(sc["sigma"]); List Sg = as<List>(sc["Sigma"]);
        if (out->sigma_surv) { out->sigma_surv[0] = as<double>(sg["Bs_gammas"]); out->sigma_surv[1] = p > 0 ? as<double>(sg["gammas"]) : in->sigma_gammas; out->sigma_surv[2] = as<double>(sg["alphas"]); }
        copy_out(Sg["Bs_gammas"].v, out->Sigma_Bs_gammas); copy_out(Sg["alphas"].v, out->Sigma_alphas);
        if (p > 0) copy_out(Sg["gammas"].v, out->Sigma_gammas);
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// log_post_RE_svft / survPred_svft_2 (src/survfitJM_mvJMbayes.cpp:190-271), called one subject at a time
// with the per-subject Data list survfitJM.mvJMbayes builds (R/survfitJM.mvJMbayes.R:281-294,381-393)
static List svft_subject(const orc_data* d, const orc_draw* w, int i, const double* Bs, const double* gam, const double* alp) {
    const int O = d->n_outcomes, T = d->n_terms, K = d->K, B = d->n_Bs, p = d->n_gammas, A = d->n_alphas;
    const arma::uword ns = (arma::uword)d->n * K;
    List D;
    std::vector<ValuePtr> y, Xb, Z, REi, idL, idL2, sig;
    std::vector<std::string> fams, links, tfs;
    for (int k = 0; k < O; ++k) {
        const long long Nk = d->N[k];
        const long long r1 = d->idL2[k][i], r0 = i == 0 ? 0 : (long long)d->idL2[k][i - 1] + 1, v = r1 - r0 + 1;
        y.push_back(num_vec(d->y[k] + r0, (size_t)v)); Xb.push_back(num_vec(w->Xbetas[k] + r0, (size_t)v));
        std::vector<double> zk((size_t)v * d->q_k[k]);
        for (int j = 0; j < d->q_k[k]; ++j) for (long long r = 0; r < v; ++r) zk[r + j * v] = d->Z[k][r0 + r + (long long)j * Nk];
        Z.push_back(num_mat(zk.data(), (arma::uword)v, d->q_k[k]));
        REi.push_back(int_vec(d->RE_inds[k], d->q_k[k], 1));
        idL.push_back(seq_vec(1, 1, v));                                   // rep(1, v)
        idL2.push_back(seq_vec(v, 1));                                     // lapply(idL, length), R :203
        sig.push_back(scalar(w->sigmas[k]));
        fams.push_back(FAM[d->fams[k]]); links.push_back(LINK[d->links[k]]);
    }
    std::vector<ValuePtr> RE2, cols, XXsb, ZZs, Us, idTs;
    auto rows_of = [&](const double* M, int ncol) {
        std::vector<double> m((size_t)K * ncol);
        for (int j = 0; j < ncol; ++j) for (int r = 0; r < K; ++r) m[r + j * K] = M[(arma::uword)i * K + r + (arma::uword)j * ns];
        return num_mat(m.data(), K, ncol);
    };
    for (int t = 0; t < T; ++t) {
        RE2.push_back(int_vec(d->RE_inds2[t], d->r_t[t], 1)); cols.push_back(int_vec(d->col_inds[t], d->u_t[t], 1));
        XXsb.push_back(num_vec(w->XXsbetas[t] + (arma::uword)i * K, K));
        ZZs.push_back(rows_of(d->ZZs[t], d->r_t[t])); Us.push_back(rows_of(d->Us[t], d->u_t[t]));
        idTs.push_back(seq_vec(1, 1, K));
        tfs.push_back(TF[d->trans_Funs[t]]);
    }
    D.push("Bs_gammas", num_vec(Bs, B)); D.push("gammas", num_vec(gam, p)); D.push("alphas", num_vec(alp, A));
    D.push("y", list_of(y)); D.push("Xbetas", list_of(Xb)); D.push("Z", list_of(Z)); D.push("RE_inds", list_of(REi));
    D.push("RE_inds2", list_of(RE2)); D.push("idL", list_of(idL)); D.push("idL2", list_of(idL2));
    D.push("fams", str_vec(fams)); D.push("links", str_vec(links)); D.push("sigmas", list_of(sig));
    D.push("invD", num_mat(w->invD, d->q, d->q));
    D.push("W1s", rows_of(d->W1s, B)); D.push("W2s", rows_of(d->W2s, p));
    D.push("XXsbetas", list_of(XXsb)); D.push("ZZs", list_of(ZZs)); D.push("Us", list_of(Us));
    D.push("Pw", num_vec(d->Pw + (arma::uword)i * K, K));
    long long last = K - 1;
    auto gk = std::make_shared<Value>(); gk->kind = Value::INT; gk->ints = {last};   // which(idGK_fast) - 1, R :282
    D.push("idGK", gk);
    D.push("idTs", list_of(idTs)); D.push("col_inds", list_of(cols)); D.push("row_inds_Us", seq_vec(1, K));
    D.push("trans_Funs", str_vec(tfs));
    return D;
}

extern "C" int jmbref_svft(const orc_data* d, const orc_draw* w, const double* b, const double* Bs, const double* gam,
                           const double* alp, double* log_post, double* surv_pred) {
    try {
        for (int i = 0; i < d->n; ++i) {
            List D = svft_subject(d, w, i, Bs, gam, alp);
            arma::vec bi(d->q);
            for (int j = 0; j < d->q; ++j) bi.mem[j] = b[i + (size_t)j * d->n];
            if (log_post) log_post[i] = log_post_RE_svft(bi, D).mem[0];
            if (surv_pred) surv_pred[i] = survPred_svft_2(bi, D).mem[0];
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}
