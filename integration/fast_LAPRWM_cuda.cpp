// fast_LAPRWM_cuda.cpp - Rcpp glue that replaces the BODIES of JMbayes' exported C++ functions by calls into
// libjmb200.so (include/jmb200.h).  Drop-in for src/fast_LAPRWM.cpp, src/fast_logPost.cpp, src/fast_score.cpp and
// the two evaluators of src/survfitJM_mvJMbayes.cpp: same exported names, same argument lists, same returned
// lists (src/RcppExports.cpp:9-228), so R/RcppExports.R and every R caller stay as they are.
//
//   lap_rwm_C                      src/fast_LAPRWM.cpp:431-896     -> jmb_create / jmb_set_draw / jmb_lap_rwm
//   lap_rwm_C_woRE[_nogammas]      src/fast_LAPRWM.cpp:934-1297    -> jmb_set_wlong / jmb_lap_rwm_woRE
//   logPosterior[_nogammas]        src/fast_logPost.cpp:8-46       -> jmb_set_wlong / jmb_logPosterior
//   gradient_logPosterior[_nog.]   src/fast_score.cpp:8-84         -> jmb_set_wlong / jmb_gradient_logPosterior
//   log_post_RE_svft               src/survfitJM_mvJMbayes.cpp:190 -> jmb_log_post_RE_svft   (one-subject handle)
//   survPred_svft_2                src/survfitJM_mvJMbayes.cpp:241 -> jmb_survPred_svft_2    (one-subject handle)
//   dmvnorm2                       src/fast_dmvnorm.cpp            stays as it is (host, once per draw)
//
// Only the part of the Rcpp API that the reference's own sources use appears here (List lookup by name, as<arma::...>,
// List::create / Named), so the file compiles against RcppArmadillo in an R session (src/Makevars: PKG_CPPFLAGS =
// -I<repo>/include, PKG_LIBS = -L<repo>/jmbayes_b200 -ljmb200) and, in this repository, against the stand-in header the
// oracle uses for the reference sources (oracle/ref_shim/RcppArmadillo.h): `make -C oracle shim` links it with
// oracle/ref_driver.cpp into oracle/_ref/libjmb_shim.so, and tests/test_gpu_rcpp_shim.py feeds the SAME R lists to the
// reference's own function bodies (oracle/_ref/libjmb_ref.so) and to these, and compares what they return.
//
// Index conventions follow List2Field_uvec (src/fast_LAPRWM.cpp:45-56): RE_inds, RE_inds2, idL, col_inds, td_cols
// arrive 1-based and lose 1 here; idL2, idGK_fast, idT_rsum arrive 0-based.  Errors: jmb_last_error() -> Rcpp::stop
// (a C++ exception that BEGIN_RCPP / END_RCPP turn into an R error, src/RcppExports.cpp:12,25).  There is no CPU
// fallback: without the library or a device every entry point raises.
//
// Random numbers: the reference consumes R's global stream; the library draws from Philox streams keyed by
// (control$seed, control$chain_id [optional, default 0], iteration, global subject) - DESIGN.md section 4.
#include <RcppArmadillo.h>

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "jmb200.h"

using namespace Rcpp;

namespace {

[[noreturn]] void fail(const std::string& where) { throw std::runtime_error(where + ": " + jmb_last_error()); }

int fam_code(const std::string& f) {
    if (f == "gaussian") return JMB_FAM_GAUSSIAN;
    if (f == "binomial") return JMB_FAM_BINOMIAL;
    if (f == "poisson") return JMB_FAM_POISSON;
    throw std::runtime_error("unknown family '" + f + "'");
}
int link_code(const std::string& l) {
    if (l == "identity") return JMB_LINK_IDENTITY;
    if (l == "logit") return JMB_LINK_LOGIT;
    if (l == "probit") return JMB_LINK_PROBIT;
    if (l == "cloglog") return JMB_LINK_CLOGLOG;
    if (l == "log") return JMB_LINK_LOG;
    throw std::runtime_error("unknown link '" + l + "'");
}
int tf_code(const std::string& t) {
    static const char* nm[] = {"identity", "expit", "exp", "log", "log2", "log10", "sqrt"};
    for (int i = 0; i < 7; ++i) if (t == nm[i]) return i;
    throw std::runtime_error("unknown transformation '" + t + "'");
}

// a list of matrices / vectors, kept alive for the call, with the pointer table the C ABI takes
struct MatList {
    std::vector<arma::mat> m; std::vector<const double*> p;
    explicit MatList(const List& l) {
        m.reserve(l.size());
        for (int i = 0; i < l.size(); ++i) m.push_back(as<arma::mat>(l[i]));
        for (auto& x : m) p.push_back(x.n_elem ? x.memptr() : nullptr);
    }
    MatList() {}
};
struct IntList {
    std::vector<std::vector<int32_t>> v; std::vector<const int32_t*> p; std::vector<int32_t> len;
    IntList(const List& l, int minus) {
        v.reserve(l.size());
        for (int i = 0; i < l.size(); ++i) {
            const arma::vec x = as<arma::vec>(l[i]);                  // integer vectors convert exactly
            std::vector<int32_t> c(x.n_elem);
            for (arma::uword k = 0; k < x.n_elem; ++k) c[k] = (int32_t)x[k] - minus;
            v.push_back(std::move(c));
        }
        for (auto& x : v) { p.push_back(x.data()); len.push_back((int32_t)x.size()); }
    }
    IntList() {}
};
std::vector<int32_t> ints_of(const arma::vec& x, int minus) {
    std::vector<int32_t> c(x.n_elem);
    for (arma::uword k = 0; k < x.n_elem; ++k) c[k] = (int32_t)x[k] - minus;
    return c;
}

// 64-bit fingerprint of the static designs of a fit: the handle (device records) is rebuilt only when it changes.
// R callers may instead put a scalar `jmb_key` into Data (one value per fit), which skips the hashing.
struct Hash {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    void add(const void* p, size_t bytes) {
        const unsigned char* c = static_cast<const unsigned char*>(p);
        size_t i = 0;
        for (; i + 8 <= bytes; i += 8) { uint64_t w; std::memcpy(&w, c + i, 8); h = (h ^ w) * 0x100000001B3ull; h ^= h >> 29; }
        for (; i < bytes; ++i) h = (h ^ c[i]) * 0x100000001B3ull;
    }
    void add(const arma::mat& m) { const uint64_t d[2] = {m.n_rows, m.n_cols}; add(d, sizeof(d)); if (m.n_elem) add(m.memptr(), sizeof(double) * m.n_elem); }
};

// everything jmb_create needs, flattened from Data (table D of SURVEY.md 8(a); built at R/mvJointModelBayes.R:673-714)
struct FlatData {
    jmb_data d{};
    MatList y, Z, ZZ, ZZs, ZZs_int, U, Us, Us_int;
    IntList RE_inds, RE_inds2, idL, idL2, col_inds;
    arma::vec event, Pw, Pw_int;
    arma::mat W1, W1s, W1s_int, W2, W2s, W2s_int;
    std::vector<double> Covs_b;          // q x q x n, contiguous (copied slice by slice)
    std::vector<int64_t> N; std::vector<int32_t> fam, lnk, qk, rt, ut, tf, idGK, idT, knf, knl;
    uint64_t key = 0;

    FlatData(const List& Data, const arma::cube& Covs_b_, int n_subjects, bool interval_cens, bool multiState)
        : y(as<List>(Data["y"])), Z(as<List>(Data["Z"])), ZZ(as<List>(Data["ZZ"])), ZZs(as<List>(Data["ZZs"])),
          U(as<List>(Data["U"])), Us(as<List>(Data["Us"])),
          RE_inds(as<List>(Data["RE_inds"]), 1), RE_inds2(as<List>(Data["RE_inds2"]), 1), idL(as<List>(Data["idL"]), 1),
          idL2(as<List>(Data["idL2"]), 0), col_inds(as<List>(Data["col_inds"]), 1),
          event(as<arma::vec>(Data["event"])), Pw(as<arma::vec>(Data["Pw"])),
          W1(as<arma::mat>(Data["W1"])), W1s(as<arma::mat>(Data["W1s"])), W2(as<arma::mat>(Data["W2"])), W2s(as<arma::mat>(Data["W2s"])),
          Covs_b((size_t)Covs_b_.n_rows * Covs_b_.n_cols * Covs_b_.n_slices) {
        const size_t qq = (size_t)Covs_b_.n_rows * Covs_b_.n_cols;
        for (arma::uword i = 0; i < Covs_b_.n_slices; ++i) std::memcpy(Covs_b.data() + i * qq, Covs_b_.slice(i).memptr(), sizeof(double) * qq);
        const int O = (int)y.m.size(), T = (int)ZZ.m.size();
        const int nT = (int)event.n_elem;
        d.n = multiState ? n_subjects : nT; d.n_outcomes = O; d.n_terms = T;
        d.q = (int)Covs_b_.n_rows; d.n_Bs = (int)W1.n_cols; d.n_gammas = (int)W2.n_cols;
        d.K = (int)(Pw.n_elem / (arma::uword)nT);
        d.interval_cens = interval_cens; d.multi_state = multiState; d.subject_offset = 0;
        const CharacterVector fams = as<CharacterVector>(Data["fams"]), links = as<CharacterVector>(Data["links"]);
        const CharacterVector tfs = as<CharacterVector>(Data["trans_Funs"]);
        for (int k = 0; k < O; ++k) {
            N.push_back((int64_t)y.m[k].n_elem); fam.push_back(fam_code(std::string(fams[k]))); lnk.push_back(link_code(std::string(links[k])));
            qk.push_back((int32_t)Z.m[k].n_cols);
        }
        int A = 0;
        for (int t = 0; t < T; ++t) { rt.push_back(RE_inds2.len[t]); ut.push_back(col_inds.len[t]); A += col_inds.len[t]; tf.push_back(tf_code(std::string(tfs[t]))); }
        d.n_alphas = A;
        d.N = N.data(); d.fams = fam.data(); d.links = lnk.data(); d.q_k = qk.data();
        d.RE_inds = RE_inds.p.data(); d.y = y.p.data(); d.Z = Z.p.data(); d.idL = idL.p.data(); d.idL2 = idL2.p.data();
        d.event = event.memptr(); d.W1 = W1.memptr(); d.W1s = W1s.memptr();
        d.W2 = W2.n_elem ? W2.memptr() : nullptr; d.W2s = W2s.n_elem ? W2s.memptr() : nullptr; d.Pw = Pw.memptr();
        idGK = ints_of(as<arma::vec>(Data["idGK_fast"]), 0); d.idGK_fast = idGK.data();
        d.r_t = rt.data(); d.RE_inds2 = RE_inds2.p.data(); d.u_t = ut.data(); d.col_inds = col_inds.p.data(); d.trans_Funs = tf.data();
        d.ZZ = ZZ.p.data(); d.ZZs = ZZs.p.data(); d.U = U.p.data(); d.Us = Us.p.data();
        if (interval_cens) {
            W1s_int = as<arma::mat>(Data["W1s_int"]); W2s_int = as<arma::mat>(Data["W2s_int"]); Pw_int = as<arma::vec>(Data["Pw_int"]);
            ZZs_int = MatList(as<List>(Data["ZZs_int"])); Us_int = MatList(as<List>(Data["Us_int"]));
            d.W1s_int = W1s_int.memptr(); d.W2s_int = W2s_int.n_elem ? W2s_int.memptr() : nullptr; d.Pw_int = Pw_int.memptr();
            d.ZZs_int = ZZs_int.p.data(); d.Us_int = Us_int.p.data();
        }
        d.Covs_b = Covs_b.empty() ? nullptr : Covs_b.data();
        if (multiState) {                      // src/fast_LAPRWM.cpp:519-529; R/mvJointModelBayes.R:559-570,667-669
            d.nT = nT; d.nRisks = as<int>(Data["nRisks"]);
            const List idT2 = as<List>(Data["idT2"]);
            idT = ints_of(as<arma::vec>(idT2[0]), 1);
            knf = ints_of(as<arma::vec>(Data["kn_strat_first"]), 0); knl = ints_of(as<arma::vec>(Data["kn_strat_last"]), 0);
            d.idT = idT.data(); d.kn_strat_first = knf.data(); d.kn_strat_last = knl.data();
        }
        if (Data.containsElementNamed("jmb_key")) key = 0xA5A5A5A500000000ull ^ (uint64_t)as<double>(Data["jmb_key"]);
        else {
            Hash h;
            const int dims[6] = {d.n, d.q, d.K, (int)interval_cens, (int)multiState, d.n_alphas};
            h.add(dims, sizeof(dims));
            h.add(event); h.add(Pw); h.add(W1); h.add(W1s); h.add(W2s);
            for (auto& m : y.m) h.add(m); for (auto& m : Z.m) h.add(m); for (auto& m : ZZs.m) h.add(m); for (auto& m : Us.m) h.add(m);
            if (!Covs_b.empty()) h.add(Covs_b.data(), sizeof(double) * Covs_b.size());
            key = h.h;
        }
    }
};

// one resident fit: the device copy of the designs outlives the call and serves every chain (draw) of the fit
struct Resident { jmb_handle h = nullptr; uint64_t key = 0; bool valid = false; };
Resident g_fit;      // lap_rwm_C
Resident g_surv;     // logPosterior / gradient / woRE (survival-only designs)

jmb_handle resident_handle(Resident& r, const jmb_data& d, uint64_t key) {
    if (r.valid && r.key == key) return r.h;
    if (r.valid) { jmb_destroy(r.h); r.valid = false; }
    if (jmb_create(&d, /*device*/ 0, &r.h)) fail("jmb_create");
    r.key = key; r.valid = true;
    return r.h;
}

struct PriorsFlat {
    jmb_priors pr{};
    arma::vec mean_Bs, mean_g, mean_a; arma::mat Tau_Bs, Tau_g, Tau_a;
    IntList td;
    PriorsFlat(const List& priors, bool full) {
        mean_Bs = as<arma::vec>(priors["mean_Bs_gammas"]); Tau_Bs = as<arma::mat>(priors["Tau_Bs_gammas"]);
        mean_g = as<arma::vec>(priors["mean_gammas"]); Tau_g = as<arma::mat>(priors["Tau_gammas"]);
        mean_a = as<arma::vec>(priors["mean_alphas"]); Tau_a = as<arma::mat>(priors["Tau_alphas"]);
        pr.mean_Bs_gammas = mean_Bs.memptr(); pr.Tau_Bs_gammas = Tau_Bs.memptr();
        pr.mean_gammas = mean_g.n_elem ? mean_g.memptr() : nullptr; pr.Tau_gammas = Tau_g.n_elem ? Tau_g.memptr() : nullptr;
        pr.mean_alphas = mean_a.memptr(); pr.Tau_alphas = Tau_a.memptr();
        pr.A_tau_Bs_gammas = as<double>(priors["A_tau_Bs_gammas"]); pr.B_tau_Bs_gammas = as<double>(priors["B_tau_Bs_gammas"]);
        pr.rank_Tau_Bs_gammas = as<double>(priors["rank_Tau_Bs_gammas"]);
        pr.shrink_gammas = as<bool>(priors["shrink_gammas"]);
        pr.A_tau_gammas = as<double>(priors["A_tau_gammas"]); pr.B_tau_gammas = as<double>(priors["B_tau_gammas"]);
        pr.rank_Tau_gammas = as<double>(priors["rank_Tau_gammas"]);
        pr.A_phi_gammas = as<double>(priors["A_phi_gammas"]); pr.B_phi_gammas = as<double>(priors["B_phi_gammas"]);
        pr.shrink_alphas = as<bool>(priors["shrink_alphas"]);
        pr.A_tau_alphas = as<double>(priors["A_tau_alphas"]); pr.B_tau_alphas = as<double>(priors["B_tau_alphas"]);
        pr.rank_Tau_alphas = as<double>(priors["rank_Tau_alphas"]);
        pr.A_phi_alphas = as<double>(priors["A_phi_alphas"]); pr.B_phi_alphas = as<double>(priors["B_phi_alphas"]);
        if (full) {                            // read by lap_rwm_C only (src/fast_LAPRWM.cpp:551-569)
            pr.double_gamma_alphas = as<bool>(priors["double_gamma_alphas"]);
            pr.A_xi_alphas = as<double>(priors["A_xi_alphas"]); pr.B_xi_alphas = as<double>(priors["B_xi_alphas"]);
            pr.A_nu_alphas = as<double>(priors["A_nu_alphas"]); pr.B_nu_alphas = as<double>(priors["B_nu_alphas"]);
            pr.rank_Tau_td_alphas = as<double>(priors["rank_Tau_td_alphas"]);
            td = IntList(as<List>(priors["td_cols"]), 1);
            pr.n_td = (int32_t)td.v.size(); pr.td_len = td.len.data(); pr.td_cols = td.p.data();
        }
    }
};

jmb_control control_of(const List& control) {
    jmb_control ct{};
    ct.n_iter = as<int>(control["n_iter"]); ct.n_burnin = as<int>(control["n_burnin"]);
    ct.n_block = as<int>(control["n_block"]); ct.n_thin = as<int>(control["n_thin"]);
    ct.target_acc = as<double>(control["target_acc"]); ct.eps1 = as<double>(control["eps1"]);
    ct.eps2 = as<double>(control["eps2"]); ct.eps3 = as<double>(control["eps3"]);
    ct.c0 = as<double>(control["c0"]); ct.c1 = as<double>(control["c1"]); ct.adaptCov = as<bool>(control["adaptCov"]);
    ct.seed = control.containsElementNamed("seed") ? (uint32_t)as<double>(control["seed"]) : 1u;
    ct.chain_id = control.containsElementNamed("chain_id") ? (uint32_t)as<double>(control["chain_id"]) : 0u;
    return ct;
}

// survival-only handle for the Laplace-step functions: they see W1/W1s/W2/W2s/Pw/event and the association designs
// Wlong / Wlongs, never the longitudinal part - one empty outcome and one placeholder term complete the descriptor
struct SurvOnly {
    jmb_data d{};
    arma::vec event, Pw; arma::mat W1, W1s, W2, W2s, zz_n, zz_ns, u_n, u_ns, covb;
    std::vector<int32_t> idL2, idGK, one_i{0}, fam{JMB_FAM_GAUSSIAN}, lnk{JMB_LINK_IDENTITY}, qk{1}, rt{1}, ut{1}, tf{JMB_TF_IDENTITY};
    std::vector<int64_t> N{0};
    const int32_t* re_p[1]; const int32_t* idl_p[1]; const int32_t* idl2_p[1]; const double* y_p[1]; const double* z_p[1];
    const double* zz_p[1]; const double* zzs_p[1]; const double* u_p[1]; const double* us_p[1];
    uint64_t key = 0;
    SurvOnly(const arma::vec& event_, const arma::mat& W1_, const arma::mat& W1s_, const arma::mat& W2_, const arma::mat& W2s_,
             const arma::vec& Pw_, int A) : event(event_), Pw(Pw_), W1(W1_), W1s(W1s_), W2(W2_), W2s(W2s_) {
        const int n = (int)event.n_elem; const arma::uword ns = Pw.n_elem;
        d.n = n; d.n_outcomes = 1; d.q = 1; d.n_terms = 1; d.n_alphas = A; d.n_Bs = (int)W1.n_cols; d.n_gammas = (int)W2.n_cols;
        d.K = (int)(ns / (arma::uword)n);
        idL2.assign(n, -1);                     // no longitudinal rows: "last row" of every subject is -1
        idGK.resize(n); for (int i = 0; i < n; ++i) idGK[i] = (i + 1) * d.K - 1;
        zz_n = arma::mat((arma::uword)n, 1, arma::fill::zeros); zz_ns = arma::mat(ns, 1, arma::fill::zeros);
        u_n = arma::mat((arma::uword)n, 1, arma::fill::ones); u_ns = arma::mat(ns, 1, arma::fill::ones); covb = arma::mat((arma::uword)n, 1, arma::fill::ones);
        re_p[0] = one_i.data(); idl_p[0] = nullptr; idl2_p[0] = idL2.data(); y_p[0] = nullptr; z_p[0] = nullptr;
        zz_p[0] = zz_n.memptr(); zzs_p[0] = zz_ns.memptr(); u_p[0] = u_n.memptr(); us_p[0] = u_ns.memptr();
        d.N = N.data(); d.fams = fam.data(); d.links = lnk.data(); d.q_k = qk.data(); d.RE_inds = re_p; d.y = y_p; d.Z = z_p;
        d.idL = idl_p; d.idL2 = idl2_p;
        d.event = event.memptr(); d.W1 = W1.memptr(); d.W1s = W1s.memptr(); d.W2 = W2.n_elem ? W2.memptr() : nullptr;
        d.W2s = W2s.n_elem ? W2s.memptr() : nullptr; d.Pw = Pw.memptr(); d.idGK_fast = idGK.data();
        d.r_t = rt.data(); d.RE_inds2 = re_p; d.u_t = ut.data(); d.col_inds = re_p; d.trans_Funs = tf.data();
        d.ZZ = zz_p; d.ZZs = zzs_p; d.U = u_p; d.Us = us_p; d.Covs_b = covb.memptr();
        Hash h; const int dims[3] = {n, d.K, A}; h.add(dims, sizeof(dims));
        h.add(event); h.add(Pw); h.add(W1); h.add(W1s); h.add(W2); h.add(W2s);
        key = h.h;
    }
};

uint64_t g_wlong_key = 0;
jmb_handle surv_handle(const arma::vec& event, const arma::mat& W1, const arma::mat& W1s, const arma::mat& W2, const arma::mat& W2s,
                       const arma::vec& Pw, const arma::mat& Wlong, const arma::mat& Wlongs) {
    SurvOnly so(event, W1, W1s, W2, W2s, Pw, (int)Wlongs.n_cols);
    const bool fresh = !(g_surv.valid && g_surv.key == so.key);
    jmb_handle h = resident_handle(g_surv, so.d, so.key);
    Hash hw; hw.add(Wlong); hw.add(Wlongs);
    if (fresh || hw.h != g_wlong_key) {
        if (jmb_set_wlong(h, Wlong.memptr(), Wlongs.memptr())) fail("jmb_set_wlong");
        g_wlong_key = hw.h;
    }
    return h;
}

jmb_priors laplace_priors(const arma::vec& mean_Bs, const arma::mat& Tau_Bs, const arma::vec& mean_g, const arma::mat& Tau_g,
                          const arma::vec& mean_a, const arma::mat& Tau_a) {
    jmb_priors pr{};
    pr.mean_Bs_gammas = mean_Bs.memptr(); pr.Tau_Bs_gammas = Tau_Bs.memptr();
    pr.mean_gammas = mean_g.n_elem ? mean_g.memptr() : nullptr; pr.Tau_gammas = Tau_g.n_elem ? Tau_g.memptr() : nullptr;
    pr.mean_alphas = mean_a.memptr(); pr.Tau_alphas = Tau_a.memptr();
    return pr;
}

arma::mat col_out(int rows, int n_out) { return arma::mat((arma::uword)rows, (arma::uword)n_out, arma::fill::zeros); }

}  // namespace

// ---------------------------------------------------------------------------------------------------------------------
// lap_rwm_C (src/fast_LAPRWM.cpp:431-896)
// ---------------------------------------------------------------------------------------------------------------------
// [[Rcpp::export]]
List lap_rwm_C(List initials, List Data, List priors, List scales, List Covs, List control, bool interval_cens, bool multiState) {
    const arma::mat b0 = as<arma::mat>(initials["b"]);
    const int n = (int)b0.n_rows, q = (int)b0.n_cols;
    FlatData fd(Data, as<arma::cube>(Covs["b"]), n, interval_cens, multiState);
    jmb_handle h = resident_handle(g_fit, fd.d, fd.key);

    // per-draw fields (R/mvJointModelBayes.R:760-770)
    MatList Xb(as<List>(Data["Xbetas"])), XXb(as<List>(Data["XXbetas"])), XXsb(as<List>(Data["XXsbetas"])), XXsbi;
    if (interval_cens) XXsbi = MatList(as<List>(Data["XXsbetas_int"]));
    const List sig = as<List>(Data["sigmas"]);
    std::vector<double> sg((size_t)sig.size());
    for (int k = 0; k < sig.size(); ++k) sg[(size_t)k] = as<double>(sig[k]);
    const arma::mat invD = as<arma::mat>(Data["invD"]);
    jmb_draw w{};
    w.Xbetas = Xb.p.data(); w.XXbetas = XXb.p.data(); w.XXsbetas = XXsb.p.data(); w.XXsbetas_int = interval_cens ? XXsbi.p.data() : nullptr;
    w.sigmas = sg.data(); w.invD = invD.memptr();
    if (jmb_set_draw(h, &w)) fail("jmb_set_draw");

    PriorsFlat pf(priors, true);
    const jmb_control ct = control_of(control);
    const arma::vec Bs = as<arma::vec>(initials["Bs_gammas"]), gam = as<arma::vec>(initials["gammas"]), alp = as<arma::vec>(initials["alphas"]);
    const arma::vec tauBs = as<arma::vec>(initials["tau_Bs_gammas"]), phi_g = as<arma::vec>(initials["phi_gammas"]);
    const arma::vec phi_a = as<arma::vec>(initials["phi_alphas"]), tau_td = as<arma::vec>(initials["tau_td_alphas"]);
    const arma::vec sigma_b = as<arma::vec>(scales["b"]);
    const arma::mat S_Bs = as<arma::mat>(Covs["Bs_gammas"]), S_g = as<arma::mat>(Covs["gammas"]), S_a = as<arma::mat>(Covs["alphas"]);
    jmb_chain_in in{};
    in.b = b0.memptr(); in.Bs_gammas = Bs.memptr(); in.gammas = gam.n_elem ? gam.memptr() : nullptr; in.alphas = alp.memptr();
    in.tau_Bs_gammas = tauBs.memptr(); in.tau_gammas = as<double>(initials["tau_gammas"]); in.tau_alphas = as<double>(initials["tau_alphas"]);
    in.phi_gammas = phi_g.n_elem ? phi_g.memptr() : nullptr; in.phi_alphas = phi_a.memptr(); in.tau_td_alphas = tau_td.n_elem ? tau_td.memptr() : nullptr;
    in.sigma_b = sigma_b.memptr(); in.sigma_Bs_gammas = as<double>(scales["Bs_gammas"]); in.sigma_gammas = as<double>(scales["gammas"]);
    in.sigma_alphas = as<double>(scales["alphas"]);
    in.Sigma_Bs_gammas = S_Bs.memptr(); in.Sigma_gammas = S_g.n_elem ? S_g.memptr() : nullptr; in.Sigma_alphas = S_a.memptr();

    const int B = fd.d.n_Bs, p = fd.d.n_gammas, A = fd.d.n_alphas, nR = multiState ? fd.d.nRisks : 1, n_td = pf.pr.n_td;
    const int n_out = ct.n_iter / ct.n_thin;                  // floor(n_iter / n_thin), :585
    arma::cube out_b((arma::uword)n, (arma::uword)q, (arma::uword)n_out, arma::fill::zeros);
    arma::mat out_Bs = col_out(B, n_out), out_g = col_out(p, n_out), out_a = col_out(A, n_out), out_phi_g = col_out(p, n_out);
    arma::mat out_phi_a = col_out(A, n_out), out_tau_Bs = col_out(nR, n_out), out_tau_td = col_out(n_td, n_out);
    arma::vec out_tau_g((arma::uword)n_out, arma::fill::zeros), out_tau_a((arma::uword)n_out, arma::fill::zeros);
    arma::vec out_lw((arma::uword)n_out, arma::fill::zeros), o_sigma_b((arma::uword)n, arma::fill::zeros), o_sig3(3, arma::fill::zeros);
    arma::mat o_S_Bs((arma::uword)B, (arma::uword)B, arma::fill::zeros), o_S_g((arma::uword)p, (arma::uword)p, arma::fill::zeros);
    arma::mat o_S_a((arma::uword)A, (arma::uword)A, arma::fill::zeros);
    // the cube's slices are contiguous in R and in Armadillo; the stand-in keeps one matrix per slice, so go through a flat buffer
    std::vector<double> flat_b((size_t)n * q * (size_t)std::max(n_out, 1), 0.0);
    jmb_chain_out out{};
    out.b = flat_b.data(); out.Bs_gammas = out_Bs.memptr(); out.gammas = p ? out_g.memptr() : nullptr; out.alphas = out_a.memptr();
    out.phi_gammas = p ? out_phi_g.memptr() : nullptr; out.phi_alphas = out_phi_a.memptr(); out.tau_Bs_gammas = out_tau_Bs.memptr();
    out.tau_gammas = out_tau_g.memptr(); out.tau_alphas = out_tau_a.memptr(); out.tau_td_alphas = n_td ? out_tau_td.memptr() : nullptr;
    out.logWeights = out_lw.memptr(); out.sigma_b = o_sigma_b.memptr(); out.sigma_surv = o_sig3.memptr();
    out.Sigma_Bs_gammas = o_S_Bs.memptr(); out.Sigma_gammas = p ? o_S_g.memptr() : nullptr; out.Sigma_alphas = o_S_a.memptr();
    if (jmb_lap_rwm(h, &pf.pr, &ct, &in, &out)) fail("jmb_lap_rwm");
    for (int s = 0; s < n_out; ++s) std::memcpy(out_b.slice((arma::uword)s).memptr(), flat_b.data() + (size_t)s * n * q, sizeof(double) * (size_t)n * q);

    return List::create(Named("mcmc") = List::create(
        Named("b") = out_b,
        Named("Bs_gammas") = out_Bs,
        Named("gammas") = out_g,
        Named("alphas") = out_a,
        Named("phi_gammas") = out_phi_g,
        Named("phi_alphas") = out_phi_a,
        Named("tau_Bs_gammas") = out_tau_Bs,
        Named("tau_gammas") = out_tau_g,
        Named("tau_alphas") = out_tau_a,
        Named("tau_td_alphas") = out_tau_td
    ),
    Named("logWeights") = out_lw,
    Named("scales") = List::create(
        Named("sigma") = List::create(
            Named("b") = o_sigma_b,
            Named("Bs_gammas") = o_sig3[0],
            Named("gammas") = o_sig3[1],
            Named("alphas") = o_sig3[2]
        ),
        Named("Sigma") = List::create(
            Named("Bs_gammas") = o_S_Bs,
            Named("gammas") = o_S_g,
            Named("alphas") = o_S_a
        )
    ));
}

// ---------------------------------------------------------------------------------------------------------------------
// lap_rwm_C_woRE / lap_rwm_C_woRE_nogammas (src/fast_LAPRWM.cpp:934-1139,1142-1297): control$update_RE = FALSE
// ---------------------------------------------------------------------------------------------------------------------
static List woRE_impl(const List& initials, const List& Data, const List& priors, const List& scales, const List& Covs,
                      const List& control, bool nogammas) {
    const arma::vec event = as<arma::vec>(Data["event"]), Pw = as<arma::vec>(Data["Pw"]);
    const arma::mat W1 = as<arma::mat>(Data["W1"]), W1s = as<arma::mat>(Data["W1s"]);
    arma::mat W2, W2s;
    if (!nogammas) { W2 = as<arma::mat>(Data["W2"]); W2s = as<arma::mat>(Data["W2s"]); }
    const arma::mat Wlong = as<arma::mat>(Data["Wlong"]), Wlongs = as<arma::mat>(Data["Wlongs"]);
    jmb_handle h = surv_handle(event, W1, W1s, W2, W2s, Pw, Wlong, Wlongs);
    const int B = (int)W1.n_cols, p = (int)W2.n_cols, A = (int)Wlongs.n_cols;

    // priors: the woRE flavours read no double-gamma / td fields (:960-985)
    jmb_priors pr{};
    const arma::vec mean_Bs = as<arma::vec>(priors["mean_Bs_gammas"]), mean_a = as<arma::vec>(priors["mean_alphas"]);
    const arma::mat Tau_Bs = as<arma::mat>(priors["Tau_Bs_gammas"]), Tau_a = as<arma::mat>(priors["Tau_alphas"]);
    arma::vec mean_g; arma::mat Tau_g;
    pr.mean_Bs_gammas = mean_Bs.memptr(); pr.Tau_Bs_gammas = Tau_Bs.memptr(); pr.mean_alphas = mean_a.memptr(); pr.Tau_alphas = Tau_a.memptr();
    pr.A_tau_Bs_gammas = as<double>(priors["A_tau_Bs_gammas"]); pr.B_tau_Bs_gammas = as<double>(priors["B_tau_Bs_gammas"]);
    pr.rank_Tau_Bs_gammas = as<double>(priors["rank_Tau_Bs_gammas"]);
    pr.shrink_alphas = as<bool>(priors["shrink_alphas"]);
    pr.A_tau_alphas = as<double>(priors["A_tau_alphas"]); pr.B_tau_alphas = as<double>(priors["B_tau_alphas"]);
    pr.rank_Tau_alphas = as<double>(priors["rank_Tau_alphas"]);
    pr.A_phi_alphas = as<double>(priors["A_phi_alphas"]); pr.B_phi_alphas = as<double>(priors["B_phi_alphas"]);
    if (!nogammas) {
        mean_g = as<arma::vec>(priors["mean_gammas"]); Tau_g = as<arma::mat>(priors["Tau_gammas"]);
        pr.mean_gammas = mean_g.memptr(); pr.Tau_gammas = Tau_g.memptr();
        pr.shrink_gammas = as<bool>(priors["shrink_gammas"]);
        pr.A_tau_gammas = as<double>(priors["A_tau_gammas"]); pr.B_tau_gammas = as<double>(priors["B_tau_gammas"]);
        pr.rank_Tau_gammas = as<double>(priors["rank_Tau_gammas"]);
        pr.A_phi_gammas = as<double>(priors["A_phi_gammas"]); pr.B_phi_gammas = as<double>(priors["B_phi_gammas"]);
    }
    const jmb_control ct = control_of(control);
    const arma::vec Bs = as<arma::vec>(initials["Bs_gammas"]), alp = as<arma::vec>(initials["alphas"]), phi_a = as<arma::vec>(initials["phi_alphas"]);
    arma::vec gam, phi_g; arma::mat S_g;
    const double tauBs0 = as<double>(initials["tau_Bs_gammas"]);
    const arma::mat S_Bs = as<arma::mat>(Covs["Bs_gammas"]), S_a = as<arma::mat>(Covs["alphas"]);
    jmb_chain_in in{};
    in.Bs_gammas = Bs.memptr(); in.alphas = alp.memptr(); in.tau_Bs_gammas = &tauBs0; in.tau_alphas = as<double>(initials["tau_alphas"]);
    in.phi_alphas = phi_a.memptr(); in.sigma_Bs_gammas = as<double>(scales["Bs_gammas"]); in.sigma_alphas = as<double>(scales["alphas"]);
    in.Sigma_Bs_gammas = S_Bs.memptr(); in.Sigma_alphas = S_a.memptr();
    if (!nogammas) {
        gam = as<arma::vec>(initials["gammas"]); phi_g = as<arma::vec>(initials["phi_gammas"]); S_g = as<arma::mat>(Covs["gammas"]);
        in.gammas = gam.memptr(); in.phi_gammas = phi_g.memptr(); in.tau_gammas = as<double>(initials["tau_gammas"]);
        in.sigma_gammas = as<double>(scales["gammas"]); in.Sigma_gammas = S_g.memptr();
    }
    const int n_out = ct.n_iter / ct.n_thin;
    arma::mat out_Bs = col_out(B, n_out), out_g = col_out(p, n_out), out_a = col_out(A, n_out), out_phi_g = col_out(p, n_out), out_phi_a = col_out(A, n_out);
    arma::vec out_tau_Bs((arma::uword)n_out, arma::fill::zeros), out_tau_g((arma::uword)n_out, arma::fill::zeros), out_tau_a((arma::uword)n_out, arma::fill::zeros);
    arma::vec out_lp((arma::uword)n_out, arma::fill::zeros), o_sig3(3, arma::fill::zeros);
    arma::mat o_S_Bs((arma::uword)B, (arma::uword)B, arma::fill::zeros), o_S_g((arma::uword)p, (arma::uword)p, arma::fill::zeros), o_S_a((arma::uword)A, (arma::uword)A, arma::fill::zeros);
    jmb_chain_out out{};
    out.Bs_gammas = out_Bs.memptr(); out.gammas = p ? out_g.memptr() : nullptr; out.alphas = out_a.memptr();
    out.phi_gammas = p ? out_phi_g.memptr() : nullptr; out.phi_alphas = out_phi_a.memptr(); out.tau_Bs_gammas = out_tau_Bs.memptr();
    out.tau_gammas = out_tau_g.memptr(); out.tau_alphas = out_tau_a.memptr(); out.sigma_surv = o_sig3.memptr();
    out.Sigma_Bs_gammas = o_S_Bs.memptr(); out.Sigma_gammas = p ? o_S_g.memptr() : nullptr; out.Sigma_alphas = o_S_a.memptr();
    if (jmb_lap_rwm_woRE(h, &pr, &ct, &in, &out, out_lp.memptr())) fail("jmb_lap_rwm_woRE");
    if (nogammas)
        return List::create(Named("mcmc") = List::create(
            Named("Bs_gammas") = out_Bs, Named("alphas") = out_a, Named("phi_alphas") = out_phi_a,
            Named("tau_Bs_gammas") = out_tau_Bs, Named("tau_alphas") = out_tau_a),
            Named("logPost") = out_lp,
            Named("scales") = List::create(
                Named("sigma") = List::create(Named("Bs_gammas") = o_sig3[0], Named("alphas") = o_sig3[2]),
                Named("Sigma") = List::create(Named("Bs_gammas") = o_S_Bs, Named("alphas") = o_S_a)));
    return List::create(Named("mcmc") = List::create(
        Named("Bs_gammas") = out_Bs, Named("gammas") = out_g, Named("alphas") = out_a, Named("phi_gammas") = out_phi_g,
        Named("phi_alphas") = out_phi_a, Named("tau_Bs_gammas") = out_tau_Bs, Named("tau_gammas") = out_tau_g, Named("tau_alphas") = out_tau_a),
        Named("logPost") = out_lp,
        Named("scales") = List::create(
            Named("sigma") = List::create(Named("Bs_gammas") = o_sig3[0], Named("gammas") = o_sig3[1], Named("alphas") = o_sig3[2]),
            Named("Sigma") = List::create(Named("Bs_gammas") = o_S_Bs, Named("gammas") = o_S_g, Named("alphas") = o_S_a)));
}

// [[Rcpp::export]]
List lap_rwm_C_woRE(List initials, List Data, List priors, List scales, List Covs, List control) {
    return woRE_impl(initials, Data, priors, scales, Covs, control, false);
}
// [[Rcpp::export]]
List lap_rwm_C_woRE_nogammas(List initials, List Data, List priors, List scales, List Covs, List control) {
    return woRE_impl(initials, Data, priors, scales, Covs, control, true);
}

// ---------------------------------------------------------------------------------------------------------------------
// logPosterior[_nogammas] (src/fast_logPost.cpp:8-46), gradient_logPosterior[_nogammas] (src/fast_score.cpp:8-84)
// ---------------------------------------------------------------------------------------------------------------------
// [[Rcpp::export]]
double logPosterior(double temp, arma::vec event, arma::uvec idGK_fast, arma::mat W1, arma::mat W1s, arma::vec Bs_gammas,
                    arma::mat W2, arma::mat W2s, arma::vec gammas, arma::mat Wlong, arma::mat Wlongs, arma::vec alphas,
                    arma::vec Pw, arma::vec mean_Bs_gammas, arma::mat Tau_Bs_gammas, arma::vec mean_gammas,
                    arma::mat Tau_gammas, arma::vec mean_alphas, arma::mat Tau_alphas, double tauBs, double A_tauBs, double B_tauBs) {
    (void)idGK_fast;                              // accepted and unused by the reference as well (:12)
    jmb_handle h = surv_handle(event, W1, W1s, W2, W2s, Pw, Wlong, Wlongs);
    const jmb_priors pr = laplace_priors(mean_Bs_gammas, Tau_Bs_gammas, mean_gammas, Tau_gammas, mean_alphas, Tau_alphas);
    double out = 0.0;
    if (jmb_logPosterior(h, temp, Bs_gammas.memptr(), gammas.n_elem ? gammas.memptr() : nullptr, alphas.memptr(), &pr, tauBs, A_tauBs, B_tauBs, &out))
        fail("jmb_logPosterior");
    return out;
}
// [[Rcpp::export]]
double logPosterior_nogammas(double temp, arma::vec event, arma::uvec idGK_fast, arma::mat W1, arma::mat W1s, arma::vec Bs_gammas,
                             arma::mat Wlong, arma::mat Wlongs, arma::vec alphas, arma::vec Pw, arma::vec mean_Bs_gammas,
                             arma::mat Tau_Bs_gammas, arma::vec mean_alphas, arma::mat Tau_alphas, double tauBs, double A_tauBs, double B_tauBs) {
    return logPosterior(temp, event, idGK_fast, W1, W1s, Bs_gammas, arma::mat(), arma::mat(), arma::vec(), Wlong, Wlongs, alphas, Pw,
                        mean_Bs_gammas, Tau_Bs_gammas, arma::vec(), arma::mat(), mean_alphas, Tau_alphas, tauBs, A_tauBs, B_tauBs);
}

// The score needs the data behind the event_colSums* arguments (the library derives the sums from its own event / W1 / W2 /
// Wlong, R/mvJointModelBayes.R:677-681), so the shim keeps the handle of the last logPosterior call: marglogLik2 always
// evaluates fn before gr at a point (R/margLogLik2.R:42-65; optim() does the same).
// [[Rcpp::export]]
arma::vec gradient_logPosterior(double temp, arma::vec event, arma::uvec idGK_fast, arma::vec event_colSumsW1, arma::mat W1s,
                                arma::vec Bs_gammas, arma::vec event_colSumsW2, arma::mat W2s, arma::vec gammas,
                                arma::vec event_colSumsWlong, arma::mat Wlongs, arma::vec alphas, arma::vec Pw,
                                arma::vec mean_Bs_gammas, arma::mat Tau_Bs_gammas, arma::vec mean_gammas, arma::mat Tau_gammas,
                                arma::vec mean_alphas, arma::mat Tau_alphas, double tauBs, double A_tauBs, double B_tauBs) {
    (void)event; (void)idGK_fast; (void)event_colSumsW1; (void)event_colSumsW2; (void)event_colSumsWlong; (void)W1s; (void)W2s; (void)Wlongs; (void)Pw;
    if (!g_surv.valid) throw std::runtime_error("gradient_logPosterior: call logPosterior on the same Data first (the designs live on the device)");
    const jmb_priors pr = laplace_priors(mean_Bs_gammas, Tau_Bs_gammas, mean_gammas, Tau_gammas, mean_alphas, Tau_alphas);
    arma::vec out(Bs_gammas.n_elem + gammas.n_elem + alphas.n_elem + 1, arma::fill::zeros);
    if (jmb_gradient_logPosterior(g_surv.h, temp, Bs_gammas.memptr(), gammas.n_elem ? gammas.memptr() : nullptr, alphas.memptr(), &pr, tauBs, A_tauBs,
                                  B_tauBs, out.memptr())) fail("jmb_gradient_logPosterior");
    return out;
}
// [[Rcpp::export]]
arma::vec gradient_logPosterior_nogammas(double temp, arma::vec event, arma::uvec idGK_fast, arma::vec event_colSumsW1, arma::mat W1s,
                                         arma::vec Bs_gammas, arma::vec event_colSumsWlong, arma::mat Wlongs, arma::vec alphas, arma::vec Pw,
                                         arma::vec mean_Bs_gammas, arma::mat Tau_Bs_gammas, arma::vec mean_alphas, arma::mat Tau_alphas,
                                         double tauBs, double A_tauBs, double B_tauBs) {
    return gradient_logPosterior(temp, event, idGK_fast, event_colSumsW1, W1s, Bs_gammas, arma::vec(), arma::mat(), arma::vec(), event_colSumsWlong,
                                 Wlongs, alphas, Pw, mean_Bs_gammas, Tau_Bs_gammas, arma::vec(), arma::mat(), mean_alphas, Tau_alphas, tauBs, A_tauBs, B_tauBs);
}

// ---------------------------------------------------------------------------------------------------------------------
// log_post_RE_svft / survPred_svft_2 (src/survfitJM_mvJMbayes.cpp:190-271): one subject per call, as survfitJM.mvJMbayes
// calls them (R/survfitJM.mvJMbayes.R:297,372-373,393).  A one-subject handle per call keeps the two names working; the
// fast path for this loop is the batched jmb_survfitJM (include/jmb200_survfit.h), which replaces the R loop around them.
// ---------------------------------------------------------------------------------------------------------------------
static double svft_eval(const arma::vec& b, const List& Data, bool with_longitudinal) {
    const arma::vec Bs = as<arma::vec>(Data["Bs_gammas"]), gam = as<arma::vec>(Data["gammas"]), alp = as<arma::vec>(Data["alphas"]);
    const arma::mat W1s = as<arma::mat>(Data["W1s"]), W2s = as<arma::mat>(Data["W2s"]);
    const arma::vec Pw = as<arma::vec>(Data["Pw"]);
    const int K = (int)Pw.n_elem, q = (int)b.n_elem, B = (int)W1s.n_cols, p = (int)W2s.n_cols;
    MatList ZZs(as<List>(Data["ZZs"])), Us(as<List>(Data["Us"])), XXsb(as<List>(Data["XXsbetas"]));
    IntList RE2(as<List>(Data["RE_inds2"]), 1), cols(as<List>(Data["col_inds"]), 1);
    const CharacterVector tfs = as<CharacterVector>(Data["trans_Funs"]);
    const int T = (int)ZZs.m.size();
    jmb_data d{};
    d.n = 1; d.q = q; d.n_terms = T; d.n_Bs = B; d.n_gammas = p; d.K = K; d.n_alphas = (int)alp.n_elem;
    std::vector<int32_t> rt, ut, tf;
    std::vector<arma::mat> zz1, u1; std::vector<const double*> zz1p, u1p, xxb1p; std::vector<double> xxb1((size_t)T, 0.0);
    for (int t = 0; t < T; ++t) {
        rt.push_back(RE2.len[t]); ut.push_back(cols.len[t]); tf.push_back(tf_code(std::string(tfs[t])));
        zz1.push_back(arma::mat(1, (arma::uword)RE2.len[t], arma::fill::zeros)); u1.push_back(arma::mat(1, (arma::uword)cols.len[t], arma::fill::zeros));
    }
    for (int t = 0; t < T; ++t) { zz1p.push_back(zz1[t].memptr()); u1p.push_back(u1[t].memptr()); xxb1p.push_back(&xxb1[(size_t)t]); }
    // longitudinal part (log_post_RE_svft only; survPred_svft_2 reads none)
    MatList y, Z, Xb; IntList REi; std::vector<int64_t> N; std::vector<int32_t> fam, lnk, qk; std::vector<std::vector<int32_t>> idL, idL2;
    std::vector<const int32_t*> idLp, idL2p; std::vector<double> sg; arma::mat invD;
    int32_t zero_i = 0; const int32_t* zero_p[1] = {&zero_i}; const double* null_d[1] = {nullptr};
    std::vector<int32_t> m1{-1}; const int32_t* m1p[1] = {m1.data()}; const int32_t* null_i[1] = {nullptr};
    int64_t N0 = 0; int32_t fam0 = JMB_FAM_GAUSSIAN, lnk0 = JMB_LINK_IDENTITY, qk0 = 1;
    if (with_longitudinal) {
        y = MatList(as<List>(Data["y"])); Z = MatList(as<List>(Data["Z"])); Xb = MatList(as<List>(Data["Xbetas"]));
        REi = IntList(as<List>(Data["RE_inds"]), 1);
        const CharacterVector fams = as<CharacterVector>(Data["fams"]), links = as<CharacterVector>(Data["links"]);
        const List sig = as<List>(Data["sigmas"]);
        invD = as<arma::mat>(Data["invD"]);
        const int O = (int)y.m.size();
        for (int k = 0; k < O; ++k) {
            const int v = (int)y.m[k].n_elem;
            N.push_back(v); fam.push_back(fam_code(std::string(fams[k]))); lnk.push_back(link_code(std::string(links[k]))); qk.push_back((int32_t)Z.m[k].n_cols);
            idL.push_back(std::vector<int32_t>((size_t)v, 0)); idL2.push_back(std::vector<int32_t>{v - 1});
            sg.push_back(as<double>(sig[k]));
        }
        for (int k = 0; k < O; ++k) { idLp.push_back(idL[k].data()); idL2p.push_back(idL2[k].data()); }
        d.n_outcomes = O; d.N = N.data(); d.fams = fam.data(); d.links = lnk.data(); d.q_k = qk.data(); d.RE_inds = REi.p.data();
        d.y = y.p.data(); d.Z = Z.p.data(); d.idL = idLp.data(); d.idL2 = idL2p.data();
    } else {
        d.n_outcomes = 1; d.N = &N0; d.fams = &fam0; d.links = &lnk0; d.q_k = &qk0; d.RE_inds = zero_p; d.y = null_d; d.Z = null_d;
        d.idL = null_i; d.idL2 = m1p;
        invD = arma::mat((arma::uword)q, (arma::uword)q, arma::fill::zeros);
        sg.push_back(1.0);
    }
    const double ev0 = 0.0; const int32_t gk = K - 1;
    arma::mat W1_0(1, (arma::uword)B, arma::fill::zeros), W2_0(1, (arma::uword)p, arma::fill::zeros), covb((arma::uword)q, (arma::uword)q, arma::fill::zeros);
    for (int j = 0; j < q; ++j) covb(j, j) = 1.0;
    d.event = &ev0; d.W1 = W1_0.memptr(); d.W1s = W1s.memptr(); d.W2 = p ? W2_0.memptr() : nullptr; d.W2s = p ? W2s.memptr() : nullptr;
    d.Pw = Pw.memptr(); d.idGK_fast = &gk; d.r_t = rt.data(); d.RE_inds2 = RE2.p.data(); d.u_t = ut.data(); d.col_inds = cols.p.data();
    d.trans_Funs = tf.data(); d.ZZ = zz1p.data(); d.ZZs = ZZs.p.data(); d.U = u1p.data(); d.Us = Us.p.data(); d.Covs_b = covb.memptr();
    jmb_handle h = nullptr;
    if (jmb_create(&d, 0, &h)) fail("jmb_create (svft)");
    std::vector<const double*> xbp = with_longitudinal ? Xb.p : std::vector<const double*>{nullptr};
    jmb_draw w{};
    w.Xbetas = xbp.data(); w.XXbetas = xxb1p.data(); w.XXsbetas = XXsb.p.data(); w.sigmas = sg.data(); w.invD = invD.memptr();
    double out = 0.0;
    int rc = jmb_set_draw(h, &w);
    if (!rc) rc = with_longitudinal ? jmb_log_post_RE_svft(h, b.memptr(), Bs.memptr(), p ? gam.memptr() : nullptr, alp.memptr(), &out)
                                    : jmb_survPred_svft_2(h, b.memptr(), Bs.memptr(), p ? gam.memptr() : nullptr, alp.memptr(), &out);
    const std::string msg = rc ? jmb_last_error() : "";
    jmb_destroy(h);
    if (rc) throw std::runtime_error("svft evaluator: " + msg);
    return out;
}

// [[Rcpp::export]]
arma::vec log_post_RE_svft(arma::vec b, List Data) { arma::vec o(1); o[0] = svft_eval(b, Data, true); return o; }
// [[Rcpp::export]]
arma::vec survPred_svft_2(arma::vec b, List Data) { arma::vec o(1); o[0] = svft_eval(b, Data, false); return o; }
