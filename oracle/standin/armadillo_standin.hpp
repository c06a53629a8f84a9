// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for the subset of Armadillo that
// /root/reference/src/phylomap.cpp uses, so that the UNMODIFIED reference source compiles in a container
// that has neither R nor Armadillo.  Nothing here comes from Armadillo's sources; it is a
// small eager (no expression templates) dense library with the same spelling:
//   arma::mat / vec / colvec / rowvec / imat / irowvec / cube / sp_mat, .row() .col() .slice() .diag() .t()
//   .zeros() .ones() .eye() .insert_slices(), zeros<T>() ones<T>(), trans, sum, accu, min, abs, expmat,
//   operators * (matrix product, scalar), % (element-wise), / + -.
//
// Arithmetic ORDER is part of the contract (the oracle restates the same order):
//   * matrix products: every output element is a left-to-right dot product starting from the first term
//     (what Armadillo's tiny-square gemv emulation evaluates for n <= 4; for n > 4 Armadillo calls BLAS,
//     whose order is implementation-defined, so this is as exact as any restatement can be);
//   * sp_mat products: the same dot products with the structural zeros skipped (Armadillo walks the CSC
//     non-zeros column by column and accumulates into the output, which adds the same terms in the same order);
//   * sum() of a vector = accu(): two running sums over even / odd positions, odd leftover added to the first.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <stdexcept>
#include <vector>

namespace arma {

typedef unsigned long long uword;
typedef long long sword;

template <class T> class Mat;
template <class T> class subview;
template <class T> class diagview;

// Anything that can be evaluated to a dense matrix.
template <class T, class D>
struct Base {
  const D& self() const { return static_cast<const D&>(*this); }
};

template <class T>
class Mat : public Base<T, Mat<T>> {
 public:
  typedef T elem_type;
  uword n_rows = 0, n_cols = 0, n_elem = 0;

 protected:
  std::vector<T> own_;
  T* mem_ = nullptr;  // == own_.data() unless the matrix borrows caller memory
  bool borrowed_ = false;

  void alloc(uword r, uword c) {
    n_rows = r; n_cols = c; n_elem = r * c;
    if (borrowed_) throw std::logic_error("arma stand-in: cannot resize a matrix that borrows memory");
    own_.assign(n_elem, T());
    mem_ = own_.data();
  }

 public:
  Mat() {}
  Mat(uword r, uword c) { alloc(r, c); }
  // aux-memory constructor: copy_aux_mem == false borrows the caller's buffer (phylomap.cpp:917 "B2 aliases B")
  Mat(T* aux, uword r, uword c, bool copy_aux_mem = true, bool /*strict*/ = false) {
    if (copy_aux_mem) { alloc(r, c); std::copy(aux, aux + n_elem, mem_); }
    else { n_rows = r; n_cols = c; n_elem = r * c; mem_ = aux; borrowed_ = true; }
  }
  Mat(const Mat& o) { alloc(o.n_rows, o.n_cols); std::copy(o.mem_, o.mem_ + n_elem, mem_); }
  Mat(Mat&& o) noexcept { *this = std::move(o); }
  template <class D> Mat(const Base<T, D>& e) { *this = e.self().eval(); }

  Mat& operator=(const Mat& o) {
    if (this == &o) return *this;
    if (borrowed_) {
      if (o.n_elem != n_elem) throw std::logic_error("arma stand-in: size mismatch writing to borrowed memory");
      n_rows = o.n_rows; n_cols = o.n_cols;
    } else if (n_rows != o.n_rows || n_cols != o.n_cols) {
      alloc(o.n_rows, o.n_cols);
    }
    std::copy(o.mem_, o.mem_ + n_elem, mem_);
    return *this;
  }
  Mat& operator=(Mat&& o) noexcept {
    if (this == &o) return *this;
    if (borrowed_ || o.borrowed_) { return *this = static_cast<const Mat&>(o); }
    own_ = std::move(o.own_);
    n_rows = o.n_rows; n_cols = o.n_cols; n_elem = o.n_elem;
    mem_ = own_.data();
    o.n_rows = o.n_cols = o.n_elem = 0; o.mem_ = nullptr;
    return *this;
  }
  template <class D> Mat& operator=(const Base<T, D>& e) { return *this = e.self().eval(); }

  const Mat& eval() const { return *this; }

  T* memptr() { return mem_; }
  const T* memptr() const { return mem_; }
  T* begin() { return mem_; }
  T* end() { return mem_ + n_elem; }
  const T* begin() const { return mem_; }
  const T* end() const { return mem_ + n_elem; }
  uword size() const { return n_elem; }

  void chk(uword i) const { if (i >= n_elem) throw std::out_of_range("Mat::operator(): index out of bounds"); }
  void chk(uword r, uword c) const { if (r >= n_rows || c >= n_cols) throw std::out_of_range("Mat::operator(): index out of bounds"); }
  T& operator()(uword i) { chk(i); return mem_[i]; }
  const T& operator()(uword i) const { chk(i); return mem_[i]; }
  T& operator[](uword i) { return mem_[i]; }
  const T& operator[](uword i) const { return mem_[i]; }
  T& operator()(uword r, uword c) { chk(r, c); return mem_[r + c * n_rows]; }
  const T& operator()(uword r, uword c) const { chk(r, c); return mem_[r + c * n_rows]; }
  T& at(uword r, uword c) { return mem_[r + c * n_rows]; }
  const T& at(uword r, uword c) const { return mem_[r + c * n_rows]; }

  Mat& zeros() { std::fill(mem_, mem_ + n_elem, T(0)); return *this; }
  Mat& zeros(uword r, uword c) { alloc(r, c); return *this; }
  Mat& ones() { std::fill(mem_, mem_ + n_elem, T(1)); return *this; }
  Mat& fill(T v) { std::fill(mem_, mem_ + n_elem, v); return *this; }
  Mat& eye() { zeros(); for (uword i = 0; i < std::min(n_rows, n_cols); i++) at(i, i) = T(1); return *this; }
  void set_size(uword r, uword c) { alloc(r, c); }

  subview<T> row(uword r) { if (r >= n_rows) throw std::out_of_range("Mat::row(): index out of bounds"); return subview<T>(this, r, 0, 1, n_cols); }
  subview<T> col(uword c) { if (c >= n_cols) throw std::out_of_range("Mat::col(): index out of bounds"); return subview<T>(this, 0, c, n_rows, 1); }
  const subview<T> row(uword r) const { return const_cast<Mat*>(this)->row(r); }
  const subview<T> col(uword c) const { return const_cast<Mat*>(this)->col(c); }
  diagview<T> diag() { return diagview<T>(this); }

  Mat t() const {
    Mat out(n_cols, n_rows);
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) out.at(c, r) = at(r, c);
    return out;
  }
};

template <class T>
class Col : public Mat<T> {
 public:
  Col() { this->alloc(0, 1); }
  explicit Col(uword n) { this->alloc(n, 1); }
  Col(const Col& o) : Mat<T>(static_cast<const Mat<T>&>(o)) {}
  Col(const Mat<T>& m) : Mat<T>(m) { check(); }
  template <class D> Col(const Base<T, D>& e) : Mat<T>(e) { check(); }
  Col& operator=(const Col& o) { Mat<T>::operator=(static_cast<const Mat<T>&>(o)); return *this; }
  Col& operator=(const Mat<T>& m) { Mat<T>::operator=(m); check(); return *this; }
  template <class D> Col& operator=(const Base<T, D>& e) { Mat<T>::operator=(e.self().eval()); check(); return *this; }
 private:
  void check() const { if (this->n_cols != 1 && this->n_elem != 0) throw std::logic_error("arma stand-in: not a column vector"); }
};

template <class T>
class Row : public Mat<T> {
 public:
  Row() { this->alloc(1, 0); }
  explicit Row(uword n) { this->alloc(1, n); }
  Row(const Row& o) : Mat<T>(static_cast<const Mat<T>&>(o)) {}
  Row(const Mat<T>& m) : Mat<T>(m) { check(); }
  template <class D> Row(const Base<T, D>& e) : Mat<T>(e) { check(); }
  Row& operator=(const Row& o) { Mat<T>::operator=(static_cast<const Mat<T>&>(o)); return *this; }
  Row& operator=(const Mat<T>& m) { Mat<T>::operator=(m); check(); return *this; }
  template <class D> Row& operator=(const Base<T, D>& e) { Mat<T>::operator=(e.self().eval()); check(); return *this; }
 private:
  void check() const { if (this->n_rows != 1 && this->n_elem != 0) throw std::logic_error("arma stand-in: not a row vector"); }
};

// A rectangular window of a matrix (only full rows / columns are ever taken by the reference).
template <class T>
class subview : public Base<T, subview<T>> {
  Mat<T>* m_;
  uword r0_, c0_;
 public:
  uword n_rows, n_cols, n_elem;
  subview(Mat<T>* m, uword r0, uword c0, uword nr, uword nc) : m_(m), r0_(r0), c0_(c0), n_rows(nr), n_cols(nc), n_elem(nr * nc) {}
  subview(const subview&) = default;
  Mat<T> eval() const {
    Mat<T> out(n_rows, n_cols);
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) out.at(r, c) = m_->at(r0_ + r, c0_ + c);
    return out;
  }
  void assign(const Mat<T>& v) {
    if (v.n_rows != n_rows || v.n_cols != n_cols) throw std::logic_error("arma stand-in: copy into submatrix: incompatible dimensions");
    for (uword c = 0; c < n_cols; c++) for (uword r = 0; r < n_rows; r++) m_->at(r0_ + r, c0_ + c) = v.at(r, c);
  }
  subview& operator=(const subview& o) { assign(o.eval()); return *this; }
  subview& operator=(const Mat<T>& v) { assign(v); return *this; }
  template <class D> subview& operator=(const Base<T, D>& e) { assign(e.self().eval()); return *this; }
  uword size() const { return n_elem; }
  T& operator()(uword i) {
    if (i >= n_elem) throw std::out_of_range("subview::operator(): index out of bounds");
    return n_rows == 1 ? m_->at(r0_, c0_ + i) : m_->at(r0_ + i, c0_);
  }
  const T& operator()(uword i) const { return const_cast<subview*>(this)->operator()(i); }
  Mat<T> t() const { return eval().t(); }
};

template <class T>
class diagview : public Base<T, diagview<T>> {
  Mat<T>* m_;
 public:
  explicit diagview(Mat<T>* m) : m_(m) {}
  Mat<T> eval() const {
    uword n = std::min(m_->n_rows, m_->n_cols);
    Mat<T> out(n, 1);
    for (uword i = 0; i < n; i++) out.at(i, 0) = m_->at(i, i);
    return out;
  }
};

typedef Mat<double> mat;
typedef Col<double> vec;
typedef Col<double> colvec;
typedef Row<double> rowvec;
typedef Mat<sword> imat_ll;  // not used
typedef Mat<int> imat;
typedef Row<int> irowvec;
typedef Col<int> ivec;
typedef Col<uword> uvec;

// ---- generators --------------------------------------------------------------------------------------
template <class V> V zeros(uword n) { V v(n); v.zeros(); return v; }
template <class V> V zeros(uword r, uword c) { V v(r, c); v.zeros(); return v; }
template <class V> V ones(uword n) { V v(n); v.ones(); return v; }
template <class V> V ones(uword r, uword c) { V v(r, c); v.ones(); return v; }

// ---- element-wise and scalar operators (eager) -----------------------------------------------------------
template <class T, class F>
Mat<T> zip(const Mat<T>& a, const Mat<T>& b, F f, const char* what) {
  if (a.n_rows != b.n_rows || a.n_cols != b.n_cols) throw std::logic_error(std::string(what) + ": incompatible matrix dimensions");
  Mat<T> out(a.n_rows, a.n_cols);
  for (uword i = 0; i < a.n_elem; i++) out[i] = f(a[i], b[i]);
  return out;
}
template <class T, class D1, class D2> Mat<T> operator%(const Base<T, D1>& a, const Base<T, D2>& b) {
  return zip<T>(a.self().eval(), b.self().eval(), [](T x, T y) { return x * y; }, "element-wise multiplication");
}
template <class T, class D1, class D2> Mat<T> operator+(const Base<T, D1>& a, const Base<T, D2>& b) {
  return zip<T>(a.self().eval(), b.self().eval(), [](T x, T y) { return x + y; }, "addition");
}
template <class T, class D1, class D2> Mat<T> operator-(const Base<T, D1>& a, const Base<T, D2>& b) {
  return zip<T>(a.self().eval(), b.self().eval(), [](T x, T y) { return x - y; }, "subtraction");
}
template <class T, class D> Mat<T> operator/(const Base<T, D>& a, T s) {
  Mat<T> out = a.self().eval();
  for (uword i = 0; i < out.n_elem; i++) out[i] = out[i] / s;
  return out;
}
template <class T, class D> Mat<T> operator*(const Base<T, D>& a, T s) {
  Mat<T> out = a.self().eval();
  for (uword i = 0; i < out.n_elem; i++) out[i] = out[i] * s;
  return out;
}
template <class T, class D> Mat<T> operator*(T s, const Base<T, D>& a) { return a * s; }

// ---- matrix product: left-to-right dot products ----------------------------------------------------------
template <class T, class D1, class D2> Mat<T> operator*(const Base<T, D1>& a_, const Base<T, D2>& b_) {
  const Mat<T> a = a_.self().eval();
  const Mat<T> b = b_.self().eval();
  if (a.n_cols != b.n_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
  Mat<T> out(a.n_rows, b.n_cols);
  for (uword j = 0; j < b.n_cols; j++)
    for (uword i = 0; i < a.n_rows; i++) {
      T acc = a.at(i, 0) * b.at(0, j);
      for (uword k = 1; k < a.n_cols; k++) acc = acc + a.at(i, k) * b.at(k, j);
      out.at(i, j) = acc;
    }
  return out;
}

template <class T, class D> Mat<T> trans(const Base<T, D>& a) { return a.self().eval().t(); }

// accu(): Armadillo's two-accumulator loop.
template <class T> T accu_mem(const T* p, uword n) {
  T acc1 = T(0), acc2 = T(0);
  uword i, j;
  for (i = 0, j = 1; j < n; i += 2, j += 2) { acc1 += p[i]; acc2 += p[j]; }
  if (i < n) acc1 += p[i];
  return acc1 + acc2;
}
template <class T, class D> T accu(const Base<T, D>& a) { const Mat<T> m = a.self().eval(); return accu_mem(m.memptr(), m.n_elem); }
// sum(): the reference only ever sums vectors (a row of PL, a weights column), where Armadillo resolves to accu().
template <class T, class D> T sum(const Base<T, D>& a) {
  const Mat<T> m = a.self().eval();
  if (m.n_rows != 1 && m.n_cols != 1) throw std::logic_error("arma stand-in: sum() of a non-vector is not provided");
  return accu_mem(m.memptr(), m.n_elem);
}
template <class T, class D> T min(const Base<T, D>& a) {
  const Mat<T> m = a.self().eval();
  if (m.n_elem == 0) throw std::logic_error("min(): object has no elements");
  T best = m[0];
  for (uword i = 1; i < m.n_elem; i++) if (m[i] < best) best = m[i];
  return best;
}
template <class T, class D> Mat<T> abs(const Base<T, D>& a) {
  Mat<T> out = a.self().eval();
  for (uword i = 0; i < out.n_elem; i++) out[i] = std::abs(out[i]);
  return out;
}

// expmat(): scaling and squaring with a degree-13 Pade approximant (Higham 2005); Armadillo's own
// scheme is another Pade order, so results agree to rounding (~1e-15 relative), not bitwise.
inline mat solve_dense(mat A, mat Bm) {
  uword n = A.n_rows;
  for (uword k = 0; k < n; k++) {
    uword p = k; double best = std::fabs(A.at(k, k));
    for (uword i = k + 1; i < n; i++) if (std::fabs(A.at(i, k)) > best) { best = std::fabs(A.at(i, k)); p = i; }
    if (best == 0.0) throw std::runtime_error("expmat(): singular denominator");
    if (p != k) {
      for (uword j = 0; j < n; j++) std::swap(A.at(k, j), A.at(p, j));
      for (uword j = 0; j < Bm.n_cols; j++) std::swap(Bm.at(k, j), Bm.at(p, j));
    }
    for (uword i = k + 1; i < n; i++) {
      double f = A.at(i, k) / A.at(k, k);
      if (f == 0.0) continue;
      for (uword j = k; j < n; j++) A.at(i, j) -= f * A.at(k, j);
      for (uword j = 0; j < Bm.n_cols; j++) Bm.at(i, j) -= f * Bm.at(k, j);
    }
  }
  for (uword j = 0; j < Bm.n_cols; j++)
    for (uword ii = n; ii-- > 0;) {
      double s = Bm.at(ii, j);
      for (uword k = ii + 1; k < n; k++) s -= A.at(ii, k) * Bm.at(k, j);
      Bm.at(ii, j) = s / A.at(ii, ii);
    }
  return Bm;
}
template <class D> mat expmat(const Base<double, D>& a_) {
  mat A = a_.self().eval();
  if (A.n_rows != A.n_cols) throw std::logic_error("expmat(): given matrix must be square sized");
  uword n = A.n_rows;
  double norm1 = 0.0;
  for (uword j = 0; j < n; j++) { double s = 0; for (uword i = 0; i < n; i++) s += std::fabs(A.at(i, j)); norm1 = std::max(norm1, s); }
  int sq = 0;
  if (norm1 > 5.371920351148152) { sq = (int)std::ceil(std::log2(norm1 / 5.371920351148152)); A = A * std::ldexp(1.0, -sq); }
  static const double b[] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
                             10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.};
  mat I(n, n); I.eye();
  mat A2 = A * A, A4 = A2 * A2, A6 = A4 * A2;
  mat U = A * (A6 * (A6 * b[13] + A4 * b[11] + A2 * b[9]) + A6 * b[7] + A4 * b[5] + A2 * b[3] + I * b[1]);
  mat V = A6 * (A6 * b[12] + A4 * b[10] + A2 * b[8]) + A6 * b[6] + A4 * b[4] + A2 * b[2] + I * b[0];
  mat R = solve_dense(V - U, V + U);
  for (int s = 0; s < sq; s++) R = R * R;
  return R;
}

// ---- cube: a stack of matrices ------------------------------------------------------------------------
template <class T>
class Cube {
  std::vector<Mat<T>> s_;
 public:
  uword n_rows = 0, n_cols = 0, n_slices = 0;
  Cube() {}
  Cube(uword r, uword c, uword s) : s_(s, Mat<T>(r, c)), n_rows(r), n_cols(c), n_slices(s) {}
  Cube& zeros() { for (auto& m : s_) m.zeros(); return *this; }
  Mat<T>& slice(uword k) { if (k >= n_slices) throw std::out_of_range("Cube::slice(): index out of bounds"); return s_[k]; }
  const Mat<T>& slice(uword k) const { if (k >= n_slices) throw std::out_of_range("Cube::slice(): index out of bounds"); return s_[k]; }
  T& operator()(uword r, uword c, uword k) { return slice(k)(r, c); }
  const T& operator()(uword r, uword c, uword k) const { return slice(k)(r, c); }
  void insert_slices(uword pos, uword n, bool set_to_zero = true) {
    if (pos > n_slices) throw std::out_of_range("Cube::insert_slices(): index out of bounds");
    Mat<T> blank(n_rows, n_cols);
    (void)set_to_zero;  // new slices are zero either way (Armadillo leaves them uninitialised when false)
    s_.insert(s_.begin() + pos, n, blank);
    n_slices += n;
  }
};
typedef Cube<double> cube;

// ---- sp_mat: coordinate writes, dot products that skip structural zeros ----------------------------------
class sp_mat {
  std::vector<double> v_;
  std::vector<unsigned char> set_;
 public:
  uword n_rows = 0, n_cols = 0;
  sp_mat() {}
  sp_mat(uword r, uword c) : v_(r * c, 0.0), set_(r * c, 0), n_rows(r), n_cols(c) {}
  struct ref {
    sp_mat* m; uword idx;
    ref& operator=(double x) { m->v_[idx] = x; m->set_[idx] = (x != 0.0); return *this; }
    operator double() const { return m->v_[idx]; }
  };
  ref operator()(uword r, uword c) { if (r >= n_rows || c >= n_cols) throw std::out_of_range("SpMat::operator(): index out of bounds"); return ref{this, r + c * n_rows}; }
  bool has(uword r, uword c) const { return set_[r + c * n_rows] != 0; }
  double at(uword r, uword c) const { return v_[r + c * n_rows]; }
};
// sparse * dense column(s)
template <class D> mat operator*(const sp_mat& A, const Base<double, D>& x_) {
  const mat x = x_.self().eval();
  if (A.n_cols != x.n_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
  mat out(A.n_rows, x.n_cols);
  for (uword j = 0; j < x.n_cols; j++)
    for (uword i = 0; i < A.n_rows; i++) {
      double acc = 0.0;
      for (uword k = 0; k < A.n_cols; k++) if (A.has(i, k)) acc += A.at(i, k) * x.at(k, j);
      out.at(i, j) = acc;
    }
  return out;
}
// dense row(s) * sparse
template <class D> mat operator*(const Base<double, D>& x_, const sp_mat& A) {
  const mat x = x_.self().eval();
  if (x.n_cols != A.n_rows) throw std::logic_error("matrix multiplication: incompatible matrix dimensions");
  mat out(x.n_rows, A.n_cols);
  for (uword j = 0; j < A.n_cols; j++)
    for (uword i = 0; i < x.n_rows; i++) {
      double acc = 0.0;
      for (uword k = 0; k < A.n_rows; k++) if (A.has(k, j)) acc += x.at(i, k) * A.at(k, j);
      out.at(i, j) = acc;
    }
  return out;
}

}  // namespace arma
