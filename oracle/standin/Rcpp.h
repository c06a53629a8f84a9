// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for the subset of Rcpp (and of R's C API reached
// through it) that /root/reference/src/phylomap.cpp uses, so that the UNMODIFIED reference source compiles
// without R.  Nothing here comes from Rcpp's or R's sources.
//
// Provided: SEXP (a reference-counted tagged vector), NumericVector / IntegerVector / NumericMatrix /
// IntegerMatrix (shallow handles on a SEXP like Rcpp's, so a NumericMatrix& argument is mutated in place
// for the caller: phylomap.cpp:1284 "B2 aliases B"), List with integer / string proxies, Named, as<>, wrap,
// RNGScope, runif / rexp sugar, Rcout, ::Rf_rgamma / ::Rf_dgamma / ::Rf_dpois, unif_rand / exp_rand.
// Random numbers: R's Mersenne-Twister and nmath samplers as restated in oracle/r_rng.hpp (pinned against
// published R outputs by tests/test_oracle_rng.py); the global generator is seeded with standin::set_seed().
#pragma once
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <list>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../r_rng.hpp"

// ---- the object model ----------------------------------------------------------------------------------
struct SEXPREC;
typedef std::shared_ptr<SEXPREC> SEXP;
enum { NILSXP = 0, INTSXP = 13, REALSXP = 14, VECSXP = 19 };
struct SEXPREC {
  int type = NILSXP;
  std::vector<double> real;
  std::vector<int> integer;
  std::vector<SEXP> list;
  std::vector<std::string> names;
  int nrow = -1, ncol = -1;  // dim attribute
  long length() const { return type == REALSXP ? (long)real.size() : type == INTSXP ? (long)integer.size() : type == VECSXP ? (long)list.size() : 0; }
};

namespace standin {
inline SEXP make(int type, long n) {
  SEXP s = std::make_shared<SEXPREC>();
  s->type = type;
  if (type == REALSXP) s->real.assign(n, 0.0);
  if (type == INTSXP) s->integer.assign(n, 0);
  if (type == VECSXP) s->list.assign(n, SEXP());
  return s;
}
inline SEXP coerce(const SEXP& s, int type) {  // Rf_coerceVector for the numeric types
  if (!s) throw std::runtime_error("not compatible with requested type: NULL");
  if (s->type == type) return s;
  SEXP o = make(type, s->length());
  o->nrow = s->nrow; o->ncol = s->ncol;
  if (type == REALSXP && s->type == INTSXP) for (size_t i = 0; i < s->integer.size(); i++) o->real[i] = s->integer[i];
  else if (type == INTSXP && s->type == REALSXP) for (size_t i = 0; i < s->real.size(); i++) o->integer[i] = (int)s->real[i];
  else throw std::runtime_error("not compatible with requested type");
  return o;
}
// R's global generator (RNG.c: one Mersenne-Twister per process).
inline orc::RMersenne& rng() { static orc::RMersenne g(1u); return g; }
inline void set_seed(unsigned seed) { rng().set_seed(seed); }
}  // namespace standin

inline double unif_rand() { return standin::rng().unif(); }
inline double exp_rand() { return orc::exp_rand(standin::rng()); }
inline double norm_rand() { return orc::norm_rand(standin::rng()); }
inline double Rf_rgamma(double shape, double scale) { return orc::rgamma(standin::rng(), shape, scale); }
// R's dpois / dgamma evaluate the same densities through saddle-point expansions (dpois_raw); the
// direct forms below agree to rounding.  dgamma's value never reaches a decision in the reference
// (phylomap.cpp:2179-2188 feed `metropolis` / `hastings`, whose results are unused).
inline double Rf_dpois(double x, double lambda, int give_log) {
  double l = (lambda == 0.0) ? (x == 0.0 ? 0.0 : -INFINITY) : (-lambda + x * std::log(lambda) - std::lgamma(x + 1.0));
  return give_log ? l : std::exp(l);
}
inline double Rf_dgamma(double x, double shape, double scale, int give_log) {
  double l = (x <= 0) ? -INFINITY : ((shape - 1.0) * std::log(x) - x / scale - std::lgamma(shape) - shape * std::log(scale));
  return give_log ? l : std::exp(l);
}

inline bool Rf_isMatrix(const SEXP& s) { return s && s->nrow >= 0; }
inline int Rf_nrows(const SEXP& s) { return s->nrow; }
inline int Rf_ncols(const SEXP& s) { return s->ncol; }
inline int Rf_length(const SEXP& s) { return s ? (int)s->length() : 0; }

namespace Rcpp {

template <int RTYPE> struct storage;
template <> struct storage<REALSXP> { typedef double type; static std::vector<double>& vec(SEXPREC& s) { return s.real; } };
template <> struct storage<INTSXP> { typedef int type; static std::vector<int>& vec(SEXPREC& s) { return s.integer; } };

class index_out_of_bounds : public std::out_of_range {
 public:
  index_out_of_bounds() : std::out_of_range("index out of bounds") {}
};

template <class U> struct standin_is_foreign : std::false_type {};  // RcppArmadillo.h marks the arma types
template <class T> struct as_impl;  // specialised per target; RcppArmadillo.h adds the arma targets

template <int RTYPE>
class Vector {
 public:
  typedef typename storage<RTYPE>::type T;
  typedef T* iterator;
  SEXP s;
  Vector() : s(standin::make(RTYPE, 0)) {}
  template <class I, class = typename std::enable_if<std::is_integral<I>::value>::type>
  Vector(I n) : s(standin::make(RTYPE, (long)n)) {}
  Vector(SEXP x) : s(standin::coerce(x, RTYPE)) {}
  template <class It> Vector(It first, It last) : s(standin::make(RTYPE, 0)) { for (; first != last; ++first) data().push_back((T)*first); }
  std::vector<T>& data() const { return storage<RTYPE>::vec(*s); }
  operator SEXP() const { return s; }
  // implicit conversion to a foreign (Armadillo) type, like Rcpp's templated conversion operator
  template <class U, class = typename std::enable_if<standin_is_foreign<U>::value>::type>
  operator U() const { return as_impl<U>::get(s); }
  long size() const { return (long)data().size(); }
  long length() const { return size(); }
  T* begin() const { return data().data(); }
  T* end() const { return data().data() + data().size(); }
  T& operator()(long i) const { if (i < 0 || i >= size()) throw index_out_of_bounds(); return data()[i]; }
  T& operator[](long i) const { return data()[i]; }
  template <class A> static Vector create(A a) { Vector v(1); v[0] = (T)a; return v; }
  template <class A, class B> static Vector create(A a, B b) { Vector v(2); v[0] = (T)a; v[1] = (T)b; return v; }
  template <class A, class B, class C> static Vector create(A a, B b, C c) { Vector v(3); v[0] = (T)a; v[1] = (T)b; v[2] = (T)c; return v; }
};
typedef Vector<REALSXP> NumericVector;
typedef Vector<INTSXP> IntegerVector;

inline NumericVector operator*(double a, const NumericVector& v) {  // sugar: scalar * vector
  NumericVector o(v.size());
  for (long i = 0; i < v.size(); i++) o[i] = a * v[i];
  return o;
}

template <int RTYPE>
class Matrix : public Vector<RTYPE> {
 public:
  typedef typename storage<RTYPE>::type T;
  Matrix() {}
  Matrix(int r, int c) : Vector<RTYPE>(r * c) { this->s->nrow = r; this->s->ncol = c; }
  Matrix(SEXP x) : Vector<RTYPE>(x) { if (this->s->nrow < 0) throw std::runtime_error("not a matrix"); }
  int nrow() const { return this->s->nrow; }
  int ncol() const { return this->s->ncol; }
  int rows() const { return nrow(); }
  int cols() const { return ncol(); }
  T& operator()(int i, int j) const {
    if (i < 0 || j < 0 || i >= nrow() || j >= ncol()) throw index_out_of_bounds();
    return this->data()[(long)i + (long)j * nrow()];
  }
  T& operator()(long i) const { return Vector<RTYPE>::operator()(i); }
  struct Row {
    const Matrix* m; int r;
    long size() const { return m->ncol(); }
    T& operator()(int j) const { return (*m)(r, j); }
    T& operator[](int j) const { return (*m)(r, j); }
  };
  Row row(int r) const { if (r < 0 || r >= nrow()) throw index_out_of_bounds(); return Row{this, r}; }
};
typedef Matrix<REALSXP> NumericMatrix;
typedef Matrix<INTSXP> IntegerMatrix;

// ---- wrap ------------------------------------------------------------------------------------------------
inline SEXP wrap(const SEXP& s) { return s; }
template <int R> SEXP wrap(const Vector<R>& v) { return v.s; }
inline SEXP wrap(double x) { SEXP s = standin::make(REALSXP, 1); s->real[0] = x; return s; }
inline SEXP wrap(int x) { SEXP s = standin::make(INTSXP, 1); s->integer[0] = x; return s; }
inline SEXP wrap(long x) { return wrap((double)x); }  // `Named("x") = NULL` (g++'s NULL is an integer constant)
inline SEXP wrap(std::nullptr_t) { return SEXP(); }

// ---- as ------------------------------------------------------------------------------------------------
template <> struct as_impl<double> {
  static double get(const SEXP& s) {
    if (!s || s->length() != 1) throw std::runtime_error("expecting a single value");
    return s->type == REALSXP ? s->real[0] : (double)standin::coerce(s, REALSXP)->real[0];
  }
};
template <> struct as_impl<int> {
  static int get(const SEXP& s) {
    if (!s || s->length() != 1) throw std::runtime_error("expecting a single value");
    return s->type == INTSXP ? s->integer[0] : (int)standin::coerce(s, REALSXP)->real[0];
  }
};
template <> struct as_impl<SEXP> { static SEXP get(const SEXP& s) { return s; } };
template <int R> struct as_impl<Vector<R>> { static Vector<R> get(const SEXP& s) { return Vector<R>(s); } };
template <int R> struct as_impl<Matrix<R>> { static Matrix<R> get(const SEXP& s) { return Matrix<R>(s); } };
template <class T> T as(const SEXP& s) { return as_impl<T>::get(s); }

// ---- List ------------------------------------------------------------------------------------------------
class List;
struct Named {
  std::string name; SEXP value;
  explicit Named(const std::string& n) : name(n) {}
  template <class T> Named& operator=(const T& v) { value = wrap(v); return *this; }
};

class ListProxy {
  SEXP owner_; long idx_; std::string name_;  // idx_ < 0: a name not present yet
 public:
  ListProxy(SEXP o, long i, std::string n = std::string()) : owner_(o), idx_(i), name_(n) {}
  SEXP get() const {
    if (idx_ < 0) throw index_out_of_bounds();
    return owner_->list[idx_];
  }
  void set(const SEXP& v) {
    if (idx_ < 0) { owner_->list.push_back(v); owner_->names.resize(owner_->list.size()); owner_->names.back() = name_; idx_ = (long)owner_->list.size() - 1; }
    else owner_->list[idx_] = v;
  }
  template <class T> ListProxy& operator=(const T& v) { set(wrap(v)); return *this; }
  ListProxy& operator=(const ListProxy& o) { set(o.get()); return *this; }
  operator SEXP() const { return get(); }
  operator NumericVector() const { return NumericVector(get()); }
  operator IntegerVector() const { return IntegerVector(get()); }
  operator NumericMatrix() const { return NumericMatrix(get()); }
  operator IntegerMatrix() const { return IntegerMatrix(get()); }
};

class List {
 public:
  SEXP s;
  List() : s(standin::make(VECSXP, 0)) {}
  List(SEXP x) : s(x) { if (!x || x->type != VECSXP) throw std::runtime_error("not compatible with requested type: List"); }
  List(const ListProxy& p) : List(p.get()) {}
  List& operator=(const ListProxy& p) { *this = List(p.get()); return *this; }
  operator SEXP() const { return s; }
  long size() const { return (long)s->list.size(); }
  ListProxy operator[](int i) const { if (i < 0 || i >= size()) throw index_out_of_bounds(); return ListProxy(s, i); }
  ListProxy operator()(int i) const { return (*this)[i]; }
  ListProxy operator[](const std::string& name) const {
    for (size_t i = 0; i < s->names.size(); i++) if (s->names[i] == name) return ListProxy(s, (long)i);
    return ListProxy(s, -1, name);
  }
  ListProxy operator[](const char* name) const { return (*this)[std::string(name)]; }
  bool containsElementNamed(const char* name) const {
    for (size_t i = 0; i < s->names.size(); i++) if (s->names[i] == name) return true;
    return false;
  }
  void push_named(const std::string& n, const SEXP& v) { s->list.push_back(v); s->names.resize(s->list.size()); s->names.back() = n; }
  static List create(const Named& a) { List l; l.push_named(a.name, a.value); return l; }
  static List create(const Named& a, const Named& b) { List l; l.push_named(a.name, a.value); l.push_named(b.name, b.value); return l; }
  static List create(const Named& a, const Named& b, const Named& c) { List l = create(a, b); l.push_named(c.name, c.value); return l; }
};
inline SEXP wrap(const List& l) { return l.s; }
template <> struct as_impl<List> { static List get(const SEXP& s) { return List(s); } };

// Rcpp::stop(): a C++ exception that END_RCPP turns into an R error
class exception : public std::runtime_error {
 public:
  explicit exception(const std::string& m) : std::runtime_error(m) {}
};
inline void stop(const std::string& m) { throw exception(m); }

// ---- random numbers ------------------------------------------------------------------------------------------
struct RNGScope { RNGScope() {} ~RNGScope() {} };  // GetRNGstate / PutRNGstate: the stand-in generator is always live
inline NumericVector runif(int n) { NumericVector v(n); for (int i = 0; i < n; i++) v[i] = unif_rand(); return v; }
// Rcpp sugar rexp(n, rate): scale = 1/rate; a non-finite or non-positive scale gives 0 (scale == 0) or NaN
// without drawing, otherwise scale * exp_rand().
inline NumericVector rexp(int n, double rate) {
  NumericVector v(n);
  for (int i = 0; i < n; i++) v[i] = orc::rexp_rate(standin::rng(), rate);
  return v;
}

static std::ostream& Rcout = std::cerr;  // keep stdout for the caller

// ---- what the generated src/RcppExports.cpp needs ---------------------------------------------------------
namespace traits {
template <class T> struct input_holder {  // by value: as<T>(x)
  T v;
  input_holder(SEXP x) : v(as<T>(x)) {}
  operator T() { return v; }
};
template <class T> struct input_holder<T&> {  // by reference: a handle that shares the caller's object
  T v;
  input_holder(SEXP x) : v(x) {}
  operator T&() { return v; }
};
template <class T> struct input_parameter { typedef input_holder<T> type; };
}  // namespace traits

}  // namespace Rcpp

namespace standin {
inline std::string& last_error() { static thread_local std::string e; return e; }
}
// A C++ exception becomes an R error in the real package (Rf_error); here: a NULL result + standin::last_error().
#define RcppExport extern "C"
#define BEGIN_RCPP try {
#define END_RCPP } catch (std::exception& e__) { standin::last_error() = e__.what(); return SEXP(); } \
                   catch (...) { standin::last_error() = "c++ exception (unknown reason)"; return SEXP(); }
#define PROTECT(x) (x)
#define UNPROTECT(n) ((void)0)

// The reference reports progress with printf("%i \r", i) once per iteration (phylomap.cpp:864 ...); muted here.
#define printf(...) ((void)0)
