// RcppArmadillo.h - STAND-IN for the Rcpp + Armadillo headers, which are not in this image.
//
// TEST INFRASTRUCTURE ONLY.  Purpose: compile the reference's own, unmodified C++ sources
// (/root/reference/src/fast_LAPRWM.cpp, fast_logPost.cpp, fast_score.cpp, survfitJM_mvJMbayes.cpp)
// where they lie, into oracle/_ref/, so that the restated oracle can be checked against the
// reference's actual control flow and arithmetic.  This header implements, eagerly and without
// expression templates, exactly the subset of the two APIs those files use, with the documented
// Armadillo / Rcpp semantics (column-major dense storage, sum(X,1) = row sums, chol = upper
// factor, cov = N-1 normalisation with rows as observations, ...).  It is written from the public
// API documentation; it contains no Armadillo or Rcpp code.  Random numbers and Rf_rgamma are
// routed to a provider installed by the driver (oracle/ref_driver.cpp).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

// ------------------------------------------------------------------------------------------------
// RNG / Rmath provider hooks (set by the driver)
// ------------------------------------------------------------------------------------------------
extern "C" {
typedef void (*jmbref_fill_fn)(int kind /*0 randu, 1 randn*/, unsigned long long rows, unsigned long long cols, double* out);
typedef double (*jmbref_rgamma_fn)(double shape, double scale);
extern jmbref_fill_fn jmbref_fill;
extern jmbref_rgamma_fn jmbref_rgamma;
}

inline double Rf_pnorm5(double x, double mu, double sigma, int lower, int logp) {
    double z = (x - mu) / sigma;
    double p = 0.5 * std::erfc(-z * 0.70710678118654752440);
    if (!lower) p = 1.0 - p;
    return logp ? std::log(p) : p;
}
inline double Rf_rgamma(double shape, double scale) {
    if (!jmbref_rgamma) throw std::runtime_error("Rf_rgamma: no provider");
    return jmbref_rgamma(shape, scale);
}
namespace R {
inline double rnorm(double, double) { throw std::runtime_error("R::rnorm is not available in the stand-in"); }
inline double rchisq(double) { throw std::runtime_error("R::rchisq is not available in the stand-in"); }
}  // namespace R

// ------------------------------------------------------------------------------------------------
// arma
// ------------------------------------------------------------------------------------------------
namespace arma {
typedef unsigned long long uword;
namespace fill {
struct zeros_t {};
struct ones_t {};
static const zeros_t zeros{};
static const ones_t ones{};
}  // namespace fill

template <class eT> class Mat;
template <class eT> class Col;
template <class eT> class Row;

// generic sub-matrix proxy: rows ridx x columns cidx of a parent matrix
template <class eT>
class subview {
 public:
    Mat<eT>& m;
    std::vector<uword> ridx, cidx;
    subview(Mat<eT>& m_, std::vector<uword> r, std::vector<uword> c) : m(m_), ridx(std::move(r)), cidx(std::move(c)) {}
    operator Mat<eT>() const;
    subview& operator=(const Mat<eT>& x);
    subview& operator=(const subview& x) { Mat<eT> t = x; return (*this = t); }
};

// linear-index element proxy (x.elem(indices))
template <class eT>
class subview_elem {
 public:
    Mat<eT>& m;
    std::vector<uword> idx;
    subview_elem(Mat<eT>& m_, std::vector<uword> i) : m(m_), idx(std::move(i)) {}
    operator Mat<eT>() const;
    subview_elem& operator=(const Mat<eT>& x);
    subview_elem& operator=(const subview_elem& x) { Mat<eT> t = x; return (*this = t); }
};

template <class eT>
class diagview {
 public:
    Mat<eT>& m;
    explicit diagview(Mat<eT>& m_) : m(m_) {}
    diagview& operator=(const Mat<eT>& x);
    operator Mat<eT>() const;
};

template <class eT>
class each_col_proxy {
 public:
    const Mat<eT>& m;
    explicit each_col_proxy(const Mat<eT>& m_) : m(m_) {}
    Mat<eT> operator%(const Mat<eT>& v) const;
};

template <class eT>
class Mat {
 public:
    uword n_rows = 0, n_cols = 0, n_elem = 0;
    std::vector<eT> mem;
    Mat() {}
    Mat(uword r, uword c) { set_size(r, c); }
    Mat(uword r, uword c, fill::zeros_t) { set_size(r, c); }
    Mat(uword r, uword c, fill::ones_t) { set_size(r, c); std::fill(mem.begin(), mem.end(), eT(1)); }
    void set_size(uword r, uword c) { n_rows = r; n_cols = c; n_elem = r * c; mem.assign(n_elem, eT(0)); }
    uword size() const { return n_elem; }
    eT* memptr() { return mem.data(); }
    const eT* memptr() const { return mem.data(); }
    eT& at(uword i) { return mem[i]; }
    const eT& at(uword i) const { return mem[i]; }
    eT& at(uword r, uword c) { return mem[r + c * n_rows]; }
    const eT& at(uword r, uword c) const { return mem[r + c * n_rows]; }
    eT& operator()(uword i) { return mem.at(i); }
    const eT& operator()(uword i) const { return mem.at(i); }
    eT& operator()(uword r, uword c) { return mem.at(r + c * n_rows); }
    const eT& operator()(uword r, uword c) const { return mem.at(r + c * n_rows); }
    // unchecked like Armadillo's operator[]; a read past the end (the reference does one after its
    // last kept iteration, src/fast_LAPRWM.cpp:836) yields a value that never matches an index
    eT& operator[](uword i) { static thread_local eT past_end; if (i >= n_elem) { past_end = eT(-1); return past_end; } return mem[i]; }
    const eT& operator[](uword i) const { static thread_local eT past_end; if (i >= n_elem) { past_end = eT(-1); return past_end; } return mem[i]; }
    Mat& operator+=(const Mat& x) { if (x.n_elem != n_elem) throw std::runtime_error("+=: size mismatch"); for (uword i = 0; i < n_elem; ++i) mem[i] += x.mem[i]; return *this; }
    Mat& operator+=(eT v) { for (uword i = 0; i < n_elem; ++i) mem[i] += v; return *this; }   // scalar
    Mat& operator-=(const Mat& x) { if (x.n_elem != n_elem) throw std::runtime_error("-=: size mismatch"); for (uword i = 0; i < n_elem; ++i) mem[i] -= x.mem[i]; return *this; }

    static std::vector<uword> range(uword a, uword b) { std::vector<uword> v; for (uword i = a; i <= b && b != (uword)-1; ++i) v.push_back(i); return v; }
    std::vector<uword> all_rows() const { return n_rows ? range(0, n_rows - 1) : std::vector<uword>(); }
    std::vector<uword> all_cols() const { return n_cols ? range(0, n_cols - 1) : std::vector<uword>(); }
    static std::vector<uword> from(const Mat<uword>& u);

    // non-const: proxies that can be assigned to; const: gathered copies
    subview<eT> rows(const Mat<uword>& r) { return subview<eT>(*this, from(r), all_cols()); }
    Mat<eT> rows(const Mat<uword>& r) const { return subview<eT>(const_cast<Mat&>(*this), from(r), all_cols()); }
    subview<eT> rows(uword a, uword b) { return subview<eT>(*this, range(a, b), all_cols()); }
    Mat<eT> rows(uword a, uword b) const { return subview<eT>(const_cast<Mat&>(*this), range(a, b), all_cols()); }
    subview<eT> cols(const Mat<uword>& c) { return subview<eT>(*this, all_rows(), from(c)); }
    Mat<eT> cols(const Mat<uword>& c) const { return subview<eT>(const_cast<Mat&>(*this), all_rows(), from(c)); }
    subview<eT> cols(uword a, uword b) { return subview<eT>(*this, all_rows(), range(a, b)); }
    Mat<eT> cols(uword a, uword b) const { return subview<eT>(const_cast<Mat&>(*this), all_rows(), range(a, b)); }
    subview<eT> col(uword c) { return subview<eT>(*this, all_rows(), range(c, c)); }
    Mat<eT> col(uword c) const { return subview<eT>(const_cast<Mat&>(*this), all_rows(), range(c, c)); }
    subview<eT> row(uword r) { return subview<eT>(*this, range(r, r), all_cols()); }
    Mat<eT> row(uword r) const { return subview<eT>(const_cast<Mat&>(*this), range(r, r), all_cols()); }
    subview<eT> submat(const Mat<uword>& r, const Mat<uword>& c) { return subview<eT>(*this, from(r), from(c)); }
    Mat<eT> submat(const Mat<uword>& r, const Mat<uword>& c) const { return subview<eT>(const_cast<Mat&>(*this), from(r), from(c)); }
    subview<eT> submat(uword r1, uword c1, uword r2, uword c2) { return subview<eT>(*this, range(r1, r2), range(c1, c2)); }
    Mat<eT> submat(uword r1, uword c1, uword r2, uword c2) const { return subview<eT>(const_cast<Mat&>(*this), range(r1, r2), range(c1, c2)); }
    subview<eT> subvec(uword a, uword b) { return n_cols == 1 ? rows(a, b) : cols(a, b); }
    Mat<eT> subvec(uword a, uword b) const { return n_cols == 1 ? rows(a, b) : cols(a, b); }
    subview_elem<eT> elem(const Mat<uword>& i) { return subview_elem<eT>(*this, from(i)); }
    Mat<eT> elem(const Mat<uword>& i) const { return subview_elem<eT>(const_cast<Mat&>(*this), from(i)); }
    diagview<eT> diag() { return diagview<eT>(*this); }
    Mat<eT> diag() const { return diagview<eT>(const_cast<Mat&>(*this)); }
    each_col_proxy<eT> each_col() const { return each_col_proxy<eT>(*this); }
    Mat<eT> t() const {
        Mat<eT> o(n_cols, n_rows);
        for (uword c = 0; c < n_cols; ++c) for (uword r = 0; r < n_rows; ++r) o.at(c, r) = at(r, c);
        return o;
    }
    // insert `count` zero rows before row `pos`
    void insert_rows(uword pos, uword count) {
        Mat<eT> o(n_rows + count, n_cols);
        for (uword c = 0; c < n_cols; ++c) {
            for (uword r = 0; r < pos; ++r) o.at(r, c) = at(r, c);
            for (uword r = pos; r < n_rows; ++r) o.at(r + count, c) = at(r, c);
        }
        *this = o;
    }
};

template <class eT>
std::vector<uword> Mat<eT>::from(const Mat<uword>& u) { return std::vector<uword>(u.mem.begin(), u.mem.end()); }

template <class eT>
class Col : public Mat<eT> {
 public:
    Col() { this->n_cols = 1; }
    explicit Col(uword n) : Mat<eT>(n, 1) {}
    Col(uword n, fill::zeros_t z) : Mat<eT>(n, 1, z) {}
    Col(uword n, fill::ones_t o) : Mat<eT>(n, 1, o) {}
    Col(const Mat<eT>& m) : Mat<eT>(m) { if (this->n_cols != 1) { this->n_rows = this->n_elem; this->n_cols = 1; } }
    Col(const subview<eT>& s) : Col(Mat<eT>(s)) {}
    Col(const subview_elem<eT>& s) : Col(Mat<eT>(s)) {}
    Col& operator=(const Mat<eT>& m) { Mat<eT>::operator=(m); if (this->n_cols != 1) { this->n_rows = this->n_elem; this->n_cols = 1; } return *this; }
};
template <class eT>
class Row : public Mat<eT> {
 public:
    Row() { this->n_rows = 1; }
    explicit Row(uword n) : Mat<eT>(1, n) {}
    Row(const Mat<eT>& m) : Mat<eT>(m) { if (this->n_rows != 1) { this->n_cols = this->n_elem; this->n_rows = 1; } }
};
typedef Mat<double> mat;
typedef Col<double> vec;
typedef Row<double> rowvec;
typedef Col<uword> uvec;

template <class eT>
subview<eT>::operator Mat<eT>() const {
    Mat<eT> o(ridx.size(), cidx.size());
    for (uword c = 0; c < cidx.size(); ++c) for (uword r = 0; r < ridx.size(); ++r) o.at(r, c) = m(ridx[r], cidx[c]);
    return o;
}
template <class eT>
subview<eT>& subview<eT>::operator=(const Mat<eT>& x) {
    if (x.n_rows != ridx.size() || x.n_cols != cidx.size()) throw std::runtime_error("subview: size mismatch in assignment");
    for (uword c = 0; c < cidx.size(); ++c) for (uword r = 0; r < ridx.size(); ++r) m(ridx[r], cidx[c]) = x.at(r, c);
    return *this;
}
template <class eT>
subview_elem<eT>::operator Mat<eT>() const {
    Mat<eT> o(idx.size(), 1);
    for (uword i = 0; i < idx.size(); ++i) o.at(i) = m(idx[i]);
    return o;
}
template <class eT>
subview_elem<eT>& subview_elem<eT>::operator=(const Mat<eT>& x) {
    if (x.n_elem != idx.size()) throw std::runtime_error("elem: size mismatch in assignment");
    for (uword i = 0; i < idx.size(); ++i) m(idx[i]) = x.at(i);
    return *this;
}
template <class eT>
diagview<eT>& diagview<eT>::operator=(const Mat<eT>& x) {
    uword k = std::min(m.n_rows, m.n_cols);
    if (x.n_elem != k) throw std::runtime_error("diag: size mismatch");
    for (uword i = 0; i < k; ++i) m.at(i, i) = x.at(i);
    return *this;
}
template <class eT>
diagview<eT>::operator Mat<eT>() const {
    uword k = std::min(m.n_rows, m.n_cols);
    Mat<eT> o(k, 1);
    for (uword i = 0; i < k; ++i) o.at(i) = m.at(i, i);
    return o;
}
template <class eT>
Mat<eT> each_col_proxy<eT>::operator%(const Mat<eT>& v) const {
    if (v.n_elem != m.n_rows) throw std::runtime_error("each_col: size mismatch");
    Mat<eT> o(m.n_rows, m.n_cols);
    for (uword c = 0; c < m.n_cols; ++c) for (uword r = 0; r < m.n_rows; ++r) o.at(r, c) = m.at(r, c) * v.at(r);
    return o;
}

template <class eT>
class Cube {
 public:
    uword n_rows = 0, n_cols = 0, n_slices = 0;
    std::vector<Mat<eT>> s;
    Cube() {}
    Cube(uword r, uword c, uword n) : n_rows(r), n_cols(c), n_slices(n), s(n, Mat<eT>(r, c)) {}
    Cube(uword r, uword c, uword n, fill::zeros_t) : Cube(r, c, n) {}
    Mat<eT>& slice(uword i) { return s.at(i); }
    const Mat<eT>& slice(uword i) const { return s.at(i); }
};
typedef Cube<double> cube;

template <class T>
class field {
 public:
    std::vector<T> v;
    field() {}
    explicit field(uword n) : v(n) {}
    uword size() const { return v.size(); }
    T& at(uword i) { return v.at(i); }
    const T& at(uword i) const { return v.at(i); }
    T& operator()(uword i) { return v.at(i); }
    const T& operator()(uword i) const { return v.at(i); }
};

// ---- element-wise helpers on mat (non-template so that proxies convert implicitly) ----------------
inline void same_size(const mat& a, const mat& b, const char* op) {
    if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::runtime_error(std::string("size mismatch in ") + op);
}
#define JMBREF_EW(NAME, EXPR)                                                       \
    inline mat NAME(const mat& a, const mat& b) {                                   \
        same_size(a, b, #NAME); mat o(a.n_rows, a.n_cols);                          \
        for (uword i = 0; i < a.n_elem; ++i) { double x = a.mem[i], y = b.mem[i]; o.mem[i] = (EXPR); } \
        return o; }                                                                 \
    inline mat NAME(const mat& a, double y) { mat o(a.n_rows, a.n_cols);            \
        for (uword i = 0; i < a.n_elem; ++i) { double x = a.mem[i]; o.mem[i] = (EXPR); } return o; } \
    inline mat NAME(double x, const mat& b) { mat o(b.n_rows, b.n_cols);            \
        for (uword i = 0; i < b.n_elem; ++i) { double y = b.mem[i]; o.mem[i] = (EXPR); } return o; }
JMBREF_EW(operator+, x + y)
JMBREF_EW(operator-, x - y)
JMBREF_EW(operator%, x * y)
JMBREF_EW(operator/, x / y)
#undef JMBREF_EW
inline mat operator-(const mat& a) { mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = -a.mem[i]; return o; }
inline mat operator*(double s, const mat& a) { mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = s * a.mem[i]; return o; }
inline mat operator*(const mat& a, double s) { return s * a; }
inline mat operator*(const mat& a, const mat& b) {
    if (a.n_cols != b.n_rows) throw std::runtime_error("matrix product: size mismatch");
    mat o(a.n_rows, b.n_cols);
    for (uword c = 0; c < b.n_cols; ++c)
        for (uword k = 0; k < a.n_cols; ++k) {
            const double bk = b.at(k, c);
            for (uword r = 0; r < a.n_rows; ++r) o.at(r, c) += a.at(r, k) * bk;
        }
    return o;
}
inline Mat<uword> operator-(const Mat<uword>& a, int s) { Mat<uword> o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = a.mem[i] - (uword)s; return o; }

#define JMBREF_FN(NAME, EXPR) \
    inline mat NAME(const mat& a) { mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) { double x = a.mem[i]; o.mem[i] = (EXPR); } return o; }
JMBREF_FN(exp, std::exp(x))
JMBREF_FN(log, std::log(x))
JMBREF_FN(log2, std::log2(x))
JMBREF_FN(log10, std::log10(x))
JMBREF_FN(sqrt, std::sqrt(x))
#undef JMBREF_FN
inline mat pow(const mat& a, double p) { mat o(a.n_rows, a.n_cols); for (uword i = 0; i < a.n_elem; ++i) o.mem[i] = std::pow(a.mem[i], p); return o; }
inline mat cumsum(const mat& a) {          // along each column (vectors: along the vector)
    mat o(a.n_rows, a.n_cols);
    if (a.n_rows == 1) { double s = 0; for (uword i = 0; i < a.n_elem; ++i) { s += a.mem[i]; o.mem[i] = s; } return o; }
    for (uword c = 0; c < a.n_cols; ++c) { double s = 0; for (uword r = 0; r < a.n_rows; ++r) { s += a.at(r, c); o.at(r, c) = s; } }
    return o;
}
inline double sum(const mat& a) { double s = 0; for (uword i = 0; i < a.n_elem; ++i) s += a.mem[i]; return s; }   // vectors only
inline mat sum(const mat& a, int dim) {
    if (dim == 1) { mat o(a.n_rows, 1); for (uword c = 0; c < a.n_cols; ++c) for (uword r = 0; r < a.n_rows; ++r) o.mem[r] += a.at(r, c); return o; }
    mat o(1, a.n_cols); for (uword c = 0; c < a.n_cols; ++c) for (uword r = 0; r < a.n_rows; ++r) o.mem[c] += a.at(r, c); return o;
}
inline double mean(const mat& a) { return sum(a) / (double)a.n_elem; }                                            // vectors only
inline mat mean(const mat& a, int dim) { mat s = sum(a, dim); return s * (1.0 / (double)(dim == 1 ? a.n_cols : a.n_rows)); }
inline mat trans(const mat& a) { return a.t(); }
inline double as_scalar(const mat& a) { if (a.n_elem != 1) throw std::runtime_error("as_scalar: not 1x1"); return a.mem[0]; }
inline mat chol(const mat& S) {            // upper triangular R with R'R = S
    if (S.n_rows != S.n_cols) throw std::runtime_error("chol: not square");
    const uword m = S.n_rows; mat H(m, m);
    for (uword j = 0; j < m; ++j)
        for (uword i = 0; i <= j; ++i) {
            double s = S.at(i, j);
            for (uword k = 0; k < i; ++k) s -= H.at(k, i) * H.at(k, j);
            if (i == j) { if (!(s > 0.0)) throw std::runtime_error("chol(): decomposition failed"); H.at(i, j) = std::sqrt(s); }
            else H.at(i, j) = s / H.at(i, i);
        }
    return H;
}
inline double det(const mat& A) {
    if (A.n_rows != A.n_cols) throw std::runtime_error("det: not square");
    const uword m = A.n_rows; mat L = A; double d = 1.0;
    for (uword k = 0; k < m; ++k) {
        uword piv = k; for (uword r = k + 1; r < m; ++r) if (std::fabs(L.at(r, k)) > std::fabs(L.at(piv, k))) piv = r;
        if (L.at(piv, k) == 0.0) return 0.0;
        if (piv != k) { for (uword c = 0; c < m; ++c) std::swap(L.at(k, c), L.at(piv, c)); d = -d; }
        d *= L.at(k, k);
        for (uword r = k + 1; r < m; ++r) { double f = L.at(r, k) / L.at(k, k); for (uword c = k; c < m; ++c) L.at(r, c) -= f * L.at(k, c); }
    }
    return d;
}
inline mat cov(const mat& X) {             // rows = observations, N-1 normalisation
    const uword N = X.n_rows, m = X.n_cols; mat C(m, m); std::vector<double> mu(m, 0.0);
    for (uword c = 0; c < m; ++c) { for (uword r = 0; r < N; ++r) mu[c] += X.at(r, c); mu[c] /= (double)N; }
    for (uword a = 0; a < m; ++a) for (uword b = 0; b < m; ++b) {
        double s = 0; for (uword r = 0; r < N; ++r) s += (X.at(r, a) - mu[a]) * (X.at(r, b) - mu[b]);
        C.at(a, b) = s / (double)(N > 1 ? N - 1 : 1);
    }
    return C;
}
inline mat diagmat(const mat& v) { mat o(v.n_elem, v.n_elem); for (uword i = 0; i < v.n_elem; ++i) o.at(i, i) = v.mem[i]; return o; }
inline mat trimatl(const mat& A) { mat o = A; for (uword c = 0; c < A.n_cols; ++c) for (uword r = 0; r < c && r < A.n_rows; ++r) o.at(r, c) = 0.0; return o; }
inline mat join_cols(const mat& a, const mat& b) {
    if (a.n_elem == 0) return b;
    if (b.n_elem == 0) return a;
    if (a.n_cols != b.n_cols) throw std::runtime_error("join_cols: size mismatch");
    mat o(a.n_rows + b.n_rows, a.n_cols);
    for (uword c = 0; c < a.n_cols; ++c) { for (uword r = 0; r < a.n_rows; ++r) o.at(r, c) = a.at(r, c); for (uword r = 0; r < b.n_rows; ++r) o.at(a.n_rows + r, c) = b.at(r, c); }
    return o;
}
inline mat randn(uword r, uword c) { mat o(r, c); if (!jmbref_fill) throw std::runtime_error("randn: no provider"); jmbref_fill(1, r, c, o.memptr()); return o; }
template <class T> inline T randu(uword n) { T o(n); if (!jmbref_fill) throw std::runtime_error("randu: no provider"); jmbref_fill(0, n, 1, o.memptr()); return o; }
template <class T> inline T randu(uword r, uword c) { T o(r, c); if (!jmbref_fill) throw std::runtime_error("randu: no provider"); jmbref_fill(0, r, c, o.memptr()); return o; }
template <class T> struct conv_to { template <class S> static T from(const S& s) { return T(mat(s)); } };
}  // namespace arma

// ------------------------------------------------------------------------------------------------
// Rcpp
// ------------------------------------------------------------------------------------------------
namespace Rcpp {
class List;
struct Value {
    enum Kind { NUM, INT, LGL, STR, LIST, NIL } kind = NIL;
    std::vector<double> num; std::vector<long long> ints; std::vector<int> lgl; std::vector<std::string> str;
    std::vector<arma::uword> dim;
    std::shared_ptr<List> list;
};
typedef std::shared_ptr<Value> ValuePtr;

class CharacterVector {
 public:
    std::vector<std::string> v;
    std::string& operator[](std::size_t i) { return v.at(i); }
    const std::string& operator[](std::size_t i) const { return v.at(i); }
    std::size_t size() const { return v.size(); }
};
class LogicalVector {
 public:
    std::vector<int> v;
    bool operator()(std::size_t i) const { return v.at(i) != 0; }
    bool operator[](std::size_t i) const { return v.at(i) != 0; }
    std::size_t size() const { return v.size(); }
};
struct NamedItem { std::string name; ValuePtr value; };

inline ValuePtr wrap(double x) { auto v = std::make_shared<Value>(); v->kind = Value::NUM; v->num = {x}; return v; }
inline ValuePtr wrap(const arma::mat& m) {
    auto v = std::make_shared<Value>(); v->kind = Value::NUM; v->num = m.mem; v->dim = {m.n_rows, m.n_cols}; return v;
}
inline ValuePtr wrap(const arma::cube& c) {
    auto v = std::make_shared<Value>(); v->kind = Value::NUM; v->dim = {c.n_rows, c.n_cols, c.n_slices};
    for (const auto& s : c.s) v->num.insert(v->num.end(), s.mem.begin(), s.mem.end());
    return v;
}
ValuePtr wrap(const List& l);

struct Named {
    std::string name;
    explicit Named(const char* n) : name(n) {}
    template <class T> NamedItem operator=(const T& x) const { return NamedItem{name, wrap(x)}; }
};

class List {
 public:
    std::vector<NamedItem> items;
    struct Proxy { ValuePtr v; };
    Proxy operator[](const char* name) const {
        for (const auto& it : items) if (it.name == name) return Proxy{it.value};
        throw std::runtime_error(std::string("Index out of bounds: [index='") + name + "']");
    }
    Proxy operator[](const std::string& name) const { return (*this)[name.c_str()]; }
    Proxy operator[](int i) const { return Proxy{items.at((std::size_t)i).value}; }
    int size() const { return (int)items.size(); }
    bool containsElementNamed(const char* name) const { for (const auto& it : items) if (it.name == name) return true; return false; }
    void push(const std::string& name, ValuePtr v) { items.push_back(NamedItem{name, std::move(v)}); }
    template <class... A> static List create(const A&... a) { List l; (l.items.push_back(a), ...); return l; }
};
inline ValuePtr wrap(const List& l) { auto v = std::make_shared<Value>(); v->kind = Value::LIST; v->list = std::make_shared<List>(l); return v; }

inline double num_at(const Value& v, std::size_t i) {
    switch (v.kind) {
        case Value::NUM: return v.num.at(i);
        case Value::INT: return (double)v.ints.at(i);
        case Value::LGL: return (double)v.lgl.at(i);
        default: throw std::runtime_error("not a numeric value");
    }
}
inline std::size_t len_of(const Value& v) {
    switch (v.kind) { case Value::NUM: return v.num.size(); case Value::INT: return v.ints.size(); case Value::LGL: return v.lgl.size(); case Value::STR: return v.str.size(); default: return 0; }
}
template <class T> struct as_impl;
template <> struct as_impl<double> { static double get(const Value& v) { if (len_of(v) != 1) throw std::runtime_error("Expecting a single value"); return num_at(v, 0); } };
template <> struct as_impl<int> { static int get(const Value& v) { return (int)as_impl<double>::get(v); } };
template <> struct as_impl<bool> { static bool get(const Value& v) { return as_impl<double>::get(v) != 0.0; } };
template <> struct as_impl<List> { static List get(const Value& v) { if (v.kind != Value::LIST) throw std::runtime_error("not a list"); return *v.list; } };
template <> struct as_impl<CharacterVector> { static CharacterVector get(const Value& v) { CharacterVector c; c.v = v.str; return c; } };
template <> struct as_impl<LogicalVector> { static LogicalVector get(const Value& v) { LogicalVector l; for (std::size_t i = 0; i < len_of(v); ++i) l.v.push_back(num_at(v, i) != 0.0); return l; } };
template <> struct as_impl<arma::vec> { static arma::vec get(const Value& v) { arma::vec o(len_of(v)); for (std::size_t i = 0; i < len_of(v); ++i) o.mem[i] = num_at(v, i); return o; } };
template <> struct as_impl<arma::uvec> { static arma::uvec get(const Value& v) { arma::uvec o(len_of(v)); for (std::size_t i = 0; i < len_of(v); ++i) o.mem[i] = (arma::uword)num_at(v, i); return o; } };
template <> struct as_impl<arma::mat> {
    static arma::mat get(const Value& v) {
        arma::uword r = v.dim.size() >= 2 ? v.dim[0] : len_of(v), c = v.dim.size() >= 2 ? v.dim[1] : (len_of(v) ? 1 : 0);
        if (v.dim.size() >= 2 && r * c != len_of(v)) throw std::runtime_error("bad matrix dims");
        arma::mat o(r, c); for (std::size_t i = 0; i < len_of(v); ++i) o.mem[i] = num_at(v, i); return o;
    }
};
template <> struct as_impl<arma::cube> {
    static arma::cube get(const Value& v) {
        if (v.dim.size() != 3) throw std::runtime_error("not a 3-d array");
        arma::cube o(v.dim[0], v.dim[1], v.dim[2]); std::size_t k = 0;
        for (auto& s : o.s) for (auto& x : s.mem) x = num_at(v, k++);
        return o;
    }
};
template <class T> inline T as(const List::Proxy& p) { if (!p.v) throw std::runtime_error("NULL value"); return as_impl<T>::get(*p.v); }

class RNGScope {};
}  // namespace Rcpp
