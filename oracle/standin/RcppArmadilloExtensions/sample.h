// TEST INFRASTRUCTURE ONLY (oracle/_ref build).  Stand-in for <RcppArmadilloExtensions/sample.h>:
// RcppArmadillo::sample(x, size, replace, prob), restated from its published behaviour:
// FixProb (reject non-finite / negative weights, require a positive one, divide by the sum of the positive
// weights), then ProbSampleReplace (sort descending keeping the permutation, cumulative sums, one
// unif_rand() per draw, first position whose cumulative weight is >= u, the last position otherwise).
// The descending sorts go through std::sort like Armadillo's sort / sort_index, so ties fall wherever the
// C++ library puts them (insertion sort, i.e. index order, for n <= 16).
#pragma once
#include <algorithm>
#include <vector>

#include "../RcppArmadillo.h"

namespace Rcpp {
namespace RcppArmadillo {

inline void FixProb(std::vector<double>& p, int require_k, bool replace) {
  double sum = 0.0;
  int npos = 0;
  for (size_t i = 0; i < p.size(); i++) {
    if (!std::isfinite(p[i])) throw std::range_error("NAs not allowed in probability");
    if (p[i] < 0.0) throw std::range_error("Negative probabilities not allowed");
    if (p[i] > 0.0) { npos++; sum += p[i]; }
  }
  if (npos == 0 || (!replace && require_k > npos)) throw std::range_error("Not enough positive probabilities");
  for (size_t i = 0; i < p.size(); i++) p[i] = p[i] / sum;
}

template <int RTYPE>
Vector<RTYPE> sample(const Vector<RTYPE>& x, int size, bool replace, const NumericVector& prob_) {
  int n = (int)x.size();
  if (!replace) throw std::logic_error("sample stand-in: only sampling with replacement is used by the reference");
  if (prob_.size() != n) throw std::range_error("Number of probabilities must equal input vector length");
  if (n >= 200) throw std::logic_error("sample stand-in: the Walker alias branch (n >= 200) is not provided");
  std::vector<double> p(prob_.begin(), prob_.end());
  FixProb(p, size, replace);
  // sort_index(prob, "descend") / sort(prob, "descend")
  struct packet { double val; int index; };
  std::vector<packet> pk(n);
  for (int i = 0; i < n; i++) { pk[i].val = p[i]; pk[i].index = i; }
  std::sort(pk.begin(), pk.end(), [](const packet& a, const packet& b) { return a.val > b.val; });
  std::vector<double> sorted(p);
  std::sort(sorted.begin(), sorted.end(), [](double a, double b) { return a > b; });
  // cumsum
  for (int i = 1; i < n; i++) sorted[i] = sorted[i - 1] + sorted[i];
  Vector<RTYPE> out(size);
  for (int k = 0; k < size; k++) {
    double rU = unif_rand();
    int j;
    for (j = 0; j < n - 1; j++) if (rU <= sorted[j]) break;
    out[k] = x[pk[j].index];
  }
  return out;
}

}  // namespace RcppArmadillo
}  // namespace Rcpp
