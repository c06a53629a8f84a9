// Host-side rate updates of the bf / ks / mt / ksmt samplers: one conjugate-Gamma proposal per rate parameter,
// accepted through the virtual-jump likelihood ratio, rewriting Q and B in the caller's arrays like the reference
// does (B2 aliases B: src/phylomap.cpp:1284).  Runs replicated on every rank from the all-reduced sufficient
// statistics of the sweep, so that Q stays identical everywhere without a broadcast.
//
//   two_state()      updatel01 / updatel10         src/phylomap.cpp:1190-1253   (bf: the accept ratio is computed but
//                    updatel01mtNS / updatel10mtNS  src/phylomap.cpp:2192-2262    never used; mt: Metropolis step)
//   hidden_rates()   updateksl01, updateksl10, updaterkappas, updatelkappas, updategammas   :1435-1786
//                    and their *mt twins :2371-2704 (priors indexed 0:1 / 2:3 / 4:5 / 6:7, no 1e-300 guard, and the
//                    prior(1) slip in the l10 exponent at :2472 — kept on purpose)
//   record_*()       recordQ :1181, recordQks :1789, recordQmtNS :2169, recordQksmt :2709
//
// The statistics row `st` is laid out [R_0..R_{n-1} | N(a,b) at n + a*n + b].
#pragma once
#include <cmath>
#include "pm_rrng.hpp"

namespace pm {
namespace host {

struct RateModel {
  int n;
  double* Q;  // column-major n x n, caller's
  double* B;  // column-major n x n, caller's
  double Omega;
  const double* prior;
  // tallies per rate parameter, in the order of the trace columns (l01, l10, kappa-> x k, kappa<- x k, gamma x k):
  // proposals drawn and proposals installed (null: not kept)
  long long* proposed = nullptr;
  long long* accepted = nullptr;
  void tally(int param, bool ok) { if (proposed) { proposed[param]++; if (ok) accepted[param]++; } }

  double& q(int r, int c) { return Q[r + (size_t)c * n]; }
  double& b(int r, int c) { return B[r + (size_t)c * n]; }
  void sync_b(int r, int c) { b(r, c) = (r == c) ? 1 + q(r, c) / Omega : q(r, c) / Omega; }

  // ---- 2-state ----
  void record_two_state(double* row) { row[6] = q(0, 1); row[7] = q(1, 0); }

  void two_state(const double* st, UniformSource& g, bool metropolis_step) {
    for (int dir = 0; dir < 2; dir++) {
      // dir 0: 0 -> 1 uses (n01, t0, n00); dir 1: 1 -> 0 uses (n10, t1, n11)
      const double jumps = (double)(long long)st[dir == 0 ? 3 : 4];
      const double stays = (double)(long long)st[dir == 0 ? 2 : 5];
      const double dwell = st[dir];
      const double cur = dir == 0 ? q(0, 1) : q(1, 0);
      const double prop = r_rgamma(g, prior[2 * dir] + jumps, 1 / (prior[2 * dir + 1] + dwell));
      if (prop > Omega) { tally(dir, false); continue; }  // no uniform is consumed on this exit (:1205)
      double ratio = std::pow((Omega - prop) / (Omega - cur), stays) * std::exp(dwell * (prop - cur));
      if (ratio > 1) ratio = 1;
      const double u = g.next();
      if (metropolis_step && ratio < u) { tally(dir, false); continue; }
      tally(dir, true);
      if (dir == 0) {
        q(0, 0) = -prop; q(0, 1) = prop;
        b(0, 0) = 1 - prop / Omega; b(0, 1) = prop / Omega;
      } else {
        q(1, 0) = prop; q(1, 1) = -prop;
        b(1, 0) = prop / Omega; b(1, 1) = 1 - prop / Omega;
      }
    }
  }

  // ---- hidden-rate (2(k+1)-state) model, R/sourceme.R:229-246 ----
  struct Hidden {
    int k;
    double lam[2];
    double rk[16], lk[16], ga[17];
  };
  Hidden read_hidden() {
    Hidden h;
    h.k = n / 2 - 1;
    h.lam[0] = q(0, 1); h.lam[1] = q(1, 0);
    for (int i = 0; i < h.k; i++) { h.rk[i] = q(2 * i, 2 * i + 2); h.lk[i] = q(2 * i + 2, 2 * i); }
    h.ga[0] = 1;
    for (int i = 1; i <= h.k; i++) h.ga[i] = q(2 * i, 2 * i + 1) / h.lam[0];
    return h;
  }
  void record_hidden(double* row) {
    const int k = n / 2 - 1, o = n + n * n;
    row[o] = q(0, 1); row[o + 1] = q(1, 0);
    for (int i = 0; i < k; i++) {
      row[o + 2 + i] = q(2 * i, 2 * i + 2);
      row[o + 2 + k + i] = q(2 * i + 2, 2 * i);
      row[o + 2 + 2 * k + i] = q(2 * (i + 1), 2 * (i + 1) + 1) / q(0, 1);
    }
  }

  // trait-change rate of direction d (0: lambda01 on the even states, 1: lambda10 on the odd states)
  void hidden_lambda(const double* st, UniformSource& g, int d, bool multi_tree) {
    const Hidden h = read_hidden();
    const int k = h.k;
    const double Om = Omega;
    const double* R = st;
    const double* Nc = st + n;
    const int pshape = (multi_tree && d == 1) ? 2 : 0, prate = (multi_tree && d == 1) ? 3 : 1;
    auto state = [&](int i) { return 2 * i + d; };
    auto other = [&](int i) { return 2 * i + 1 - d; };
    double shape = prior[pshape];
    for (int i = 0; i <= k; i++) shape = shape + Nc[state(i) * n + other(i)];
    double rate = prior[prate];
    for (int i = 0; i <= k; i++) rate = rate + h.ga[i] * R[state(i)];
    const double prop = r_rgamma(g, shape, 1 / rate);
    const double cur = h.lam[d];
    const double weighted_dwell = rate - prior[1];
    double la = (prop - cur) * weighted_dwell;
    la = la + Nc[state(0) * n + state(0)] * std::log((Om - h.rk[0] - h.ga[0] * prop) / (Om - h.rk[0] - h.ga[0] * cur));
    for (int i = 1; i < k; i++)
      la = la + Nc[state(i) * n + state(i)] *
                    std::log((Om - h.rk[i] - h.lk[i - 1] - h.ga[i] * prop) / (Om - h.rk[i] - h.lk[i - 1] - h.ga[i] * cur));
    la = la + Nc[state(k) * n + state(k)] *
                  std::log((Om - h.lk[k - 1] - h.ga[k] * prop) / (Om - h.lk[k - 1] - h.ga[k] * cur));
    const double u = g.next();
    bool ok = !(prop + h.rk[0] > Om);
    for (int i = 1; i < k && ok; i++) if (h.ga[i] * prop + h.rk[i] + h.lk[i - 1] > Om) ok = false;
    if (ok && h.ga[k] * prop + h.lk[k - 1] > Om) ok = false;
    if (ok && !multi_tree && prop < 1e-300) ok = false;
    if (ok && la < std::log(u)) ok = false;
    tally(d, ok);
    if (!ok) return;
    for (int i = 0; i <= k; i++) {
      double diag;
      if (i == 0) diag = -h.rk[0] - h.ga[0] * prop;
      else if (i < k) diag = -h.lk[i - 1] - h.rk[i] - h.ga[i] * prop;
      else diag = -h.lk[k - 1] - h.ga[k] * prop;
      // k == 0 never reaches here (n >= 4 is enforced by the caller)
      q(state(i), state(i)) = diag;
      q(state(i), other(i)) = h.ga[i] * prop;
      sync_b(state(i), state(i));
      sync_b(state(i), other(i));
    }
  }

  // regime-change rates: towards the next regime (up = true, j = 0..k-1) or the previous one (up = false, j = 1..k)
  void hidden_kappa(const double* st, UniformSource& g, int j, bool up, bool multi_tree) {
    const Hidden h = read_hidden();
    const int k = h.k;
    const double Om = Omega;
    const double* R = st;
    const double* Nc = st + n;
    const int e = 2 * j, o = 2 * j + 1, step = up ? 2 : -2;
    const double shape = prior[multi_tree ? 4 : 2] + Nc[e * n + e + step] + Nc[o * n + o + step];
    const double rate = prior[multi_tree ? 5 : 3] + R[e] + R[o];
    const double prop = r_rgamma(g, shape, 1 / rate);
    const double cur = up ? h.rk[j] : h.lk[j - 1];
    // the opposite-direction regime rate leaving the same pair of states, if there is one
    const bool has_opp = up ? (j > 0) : (j < k);
    const double opp = has_opp ? (up ? h.lk[j - 1] : h.rk[j]) : 0.0;
    double la = (prop - cur) * (R[e] + R[o]);
    for (int side = 0; side < 2; side++) {
      const int s = e + side;
      const double tr = h.ga[j] * h.lam[side];
      const double num = has_opp ? (Om - opp - prop - tr) : (Om - prop - tr);
      double den;
      if (!has_opp) den = Om - cur - tr;
      else den = up ? (Om - h.lk[j - 1] - h.rk[j] - tr) : (Om - h.rk[j] - h.lk[j - 1] - tr);
      la = la + Nc[s * n + s] * std::log(num / den);
    }
    const double u = g.next();
    bool ok = true;
    for (int side = 0; side < 2; side++) {
      const double tr = h.ga[j] * h.lam[side];
      if (has_opp ? (prop + tr + opp > Om) : (prop + tr > Om)) ok = false;
    }
    if (ok && !multi_tree && prop < 1e-300) ok = false;
    if (ok && la < std::log(u)) ok = false;
    tally(up ? 2 + j : 2 + k + (j - 1), ok);
    if (!ok) return;
    q(e, e + step) = prop;
    q(o, o + step) = prop;
    for (int side = 0; side < 2; side++) {
      const double tr = h.ga[j] * h.lam[side];
      q(e + side, e + side) = has_opp ? (-prop - opp - tr) : (-prop - tr);
    }
    sync_b(e, e); sync_b(o, o); sync_b(e, e + step); sync_b(o, o + step);
  }

  // regime multiplier gamma_j (j = 1..k) of the trait-change rates
  void hidden_gamma(const double* st, UniformSource& g, int j, bool multi_tree) {
    const Hidden h = read_hidden();
    const int k = h.k;
    const double Om = Omega;
    const double* R = st;
    const double* Nc = st + n;
    const int e = 2 * j, o = 2 * j + 1;
    const double shape = prior[multi_tree ? 6 : 4] + Nc[e * n + o] + Nc[o * n + e];
    const double rate = prior[multi_tree ? 7 : 5] + R[e] * h.lam[0] + R[o] * h.lam[1];
    const double prop = r_rgamma(g, shape, 1 / rate);
    const bool last = (j == k);
    double la = (prop - h.ga[j]) * (R[e] * h.lam[0] + R[o] * h.lam[1]);
    for (int side = 0; side < 2; side++) {
      const int s = e + side;
      const double l = h.lam[side];
      const double num = last ? (Om - h.lk[j - 1] - prop * l) : (Om - h.lk[j - 1] - h.rk[j] - prop * l);
      const double den = last ? (Om - h.lk[j - 1] - h.ga[j] * l) : (Om - h.rk[j] - h.lk[j - 1] - h.ga[j] * l);
      la = la + Nc[s * n + s] * std::log(num / den);
    }
    const double u = g.next();
    bool ok = true;
    for (int side = 0; side < 2; side++) {
      const double l = h.lam[side];
      if (last ? (h.lk[j - 1] + prop * l > Om) : (h.lk[j - 1] + prop * l + h.rk[j] > Om)) ok = false;
    }
    if (ok && !multi_tree && prop < 1e-300) ok = false;
    if (ok && la < std::log(u)) ok = false;
    tally(2 + 2 * k + (j - 1), ok);
    if (!ok) return;
    q(e, o) = prop * h.lam[0];
    q(o, e) = prop * h.lam[1];
    for (int side = 0; side < 2; side++) {
      const double l = h.lam[side];
      q(e + side, e + side) = last ? (-h.lk[j - 1] - prop * l) : (-h.lk[j - 1] - h.rk[j] - prop * l);
    }
    sync_b(e, e); sync_b(o, o); sync_b(e, o); sync_b(o, e);
  }

  void hidden_rates(const double* st, UniformSource& g, bool multi_tree) {
    const int k = n / 2 - 1;
    hidden_lambda(st, g, 0, multi_tree);
    hidden_lambda(st, g, 1, multi_tree);
    for (int j = 0; j < k; j++) hidden_kappa(st, g, j, true, multi_tree);
    for (int j = 1; j <= k; j++) hidden_kappa(st, g, j, false, multi_tree);
    for (int j = 1; j <= k; j++) hidden_gamma(st, g, j, multi_tree);
  }
};

}  // namespace host
}  // namespace pm
