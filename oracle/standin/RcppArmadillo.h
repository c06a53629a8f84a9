// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for <RcppArmadillo.h>: the Armadillo subset plus
// the as<> / wrap converters between R vectors and Armadillo objects that src/phylomap.cpp relies on.
#pragma once
#include "Rcpp.h"
#include "armadillo_standin.hpp"

namespace Rcpp {

template <class T> struct standin_is_foreign<arma::Mat<T>> : std::true_type {};
template <class T> struct standin_is_foreign<arma::Row<T>> : std::true_type {};
template <class T> struct standin_is_foreign<arma::Col<T>> : std::true_type {};

template <class T> struct standin_sexp_type;
template <> struct standin_sexp_type<double> { enum { value = REALSXP }; static const std::vector<double>& vec(const SEXPREC& s) { return s.real; } };
template <> struct standin_sexp_type<int> { enum { value = INTSXP }; static const std::vector<int>& vec(const SEXPREC& s) { return s.integer; } };

// as<arma::Mat<T>>: needs a dim attribute (an R matrix); as<Row/Col>: any vector; values are coerced
// (node.states and states may arrive as doubles).
template <class T> struct as_impl<arma::Mat<T>> {
  static arma::Mat<T> get(const SEXP& s0) {
    SEXP s = standin::coerce(s0, standin_sexp_type<T>::value);
    if (s->nrow < 0) throw std::runtime_error("as<arma::Mat>: not a matrix");
    arma::Mat<T> m(s->nrow, s->ncol);
    const std::vector<T>& v = standin_sexp_type<T>::vec(*s);
    std::copy(v.begin(), v.end(), m.memptr());
    return m;
  }
};
template <class T> struct as_impl<arma::Row<T>> {
  static arma::Row<T> get(const SEXP& s0) {
    SEXP s = standin::coerce(s0, standin_sexp_type<T>::value);
    const std::vector<T>& v = standin_sexp_type<T>::vec(*s);
    arma::Row<T> r(v.size());
    std::copy(v.begin(), v.end(), r.memptr());
    return r;
  }
};
template <class T> struct as_impl<arma::Col<T>> {
  static arma::Col<T> get(const SEXP& s0) {
    SEXP s = standin::coerce(s0, standin_sexp_type<T>::value);
    const std::vector<T>& v = standin_sexp_type<T>::vec(*s);
    arma::Col<T> r(v.size());
    std::copy(v.begin(), v.end(), r.memptr());
    return r;
  }
};

// wrap(arma::mat) -> R numeric matrix (column-major doubles with a dim attribute)
inline SEXP wrap(const arma::Mat<double>& m) {
  SEXP s = standin::make(REALSXP, (long)m.n_elem);
  std::copy(m.memptr(), m.memptr() + m.n_elem, s->real.begin());
  s->nrow = (int)m.n_rows; s->ncol = (int)m.n_cols;
  return s;
}

}  // namespace Rcpp
