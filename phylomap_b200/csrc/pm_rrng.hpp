// Host-side random numbers of the rate-updating samplers (bf / ks / mt / ksmt).
//
// The reference draws its conjugate-Gamma proposals and accept/reject uniforms from R's global generator
// (Rf_rgamma / runif under RNGScope: src/phylomap.cpp:1202,1210,1235,1242,1463,1471,...,2332,2347).  These few
// draws per iteration stay on the host, replicated on every rank, so every rank rewrites Q identically without a
// broadcast.  This header is the product's own statement of the published R algorithms it needs:
//   Mersenne-Twister with R's set.seed scrambling -> unif_rand
//   Ahrens-Dieter (1972) exp_rand, inversion norm_rand (Wichura AS241), Ahrens-Dieter GD (1982) / GS (1974) rgamma
// A replay source (uniforms exported from a sequential run) is provided for the deterministic mode.
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace pm {
namespace host {

class UniformSource {
 public:
  virtual ~UniformSource() {}
  virtual double next() = 0;
  virtual bool exhausted() const { return false; }
};

class MersenneR : public UniformSource {
 public:
  explicit MersenneR(uint32_t seed) { reseed(seed); }
  void reseed(uint32_t s) {
    for (int i = 0; i < 50; i++) s = s * 69069u + 1u;
    s = s * 69069u + 1u;  // the word R stores in front of the state vector ("mti"), overwritten with 624
    for (int i = 0; i < 624; i++) { s = s * 69069u + 1u; state_[i] = s; }
    pos_ = 624;
  }
  double next() override {
    if (pos_ >= 624) refill();
    uint32_t y = state_[pos_++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    const double v = y * 2.3283064365386963e-10;
    const double half_ulp = 0.5 * 2.328306437080797e-10;
    if (v <= 0.0) return half_ulp;
    if (1.0 - v <= 0.0) return 1.0 - half_ulp;
    return v;
  }
  // checkpointing (pm_chain_export_state): 624 state words + the read position
  void save(uint32_t out[625]) const { for (int i = 0; i < 624; i++) out[i] = state_[i]; out[624] = (uint32_t)pos_; }
  void load(const uint32_t in[625]) { for (int i = 0; i < 624; i++) state_[i] = in[i]; pos_ = (int)in[624]; }

 private:
  static uint32_t twist(uint32_t hi, uint32_t lo, uint32_t far) {
    const uint32_t y = (hi & 0x80000000u) | (lo & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  void refill() {
    for (int i = 0; i < 624; i++) state_[i] = twist(state_[i], state_[(i + 1) % 624], state_[(i + 397) % 624]);
    pos_ = 0;
  }
  uint32_t state_[624];
  int pos_;
};

class ReplaySource : public UniformSource {
 public:
  ReplaySource(const double* u, int64_t n) : u_(u), n_(n) {}
  double next() override {
    if (pos_ >= n_) { under_ = true; return 0.5; }
    return u_[pos_++];
  }
  bool exhausted() const override { return under_; }

 private:
  const double* u_;
  int64_t n_, pos_ = 0;
  bool under_ = false;
};

// R sexp.c
inline double r_exp_rand(UniformSource& g) {
  static const double qk[16] = {0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
                                0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
                                0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
                                0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
  double u = g.next();
  while (u <= 0.0 || u >= 1.0) u = g.next();
  double base = 0.0;
  for (u += u; u <= 1.0; u += u) base += qk[0];
  u -= 1.0;
  if (u <= qk[0]) return base + u;
  double lowest = g.next();
  int i = 0;
  do {
    const double v = g.next();
    if (v < lowest) lowest = v;
    ++i;
  } while (u > qk[i]);
  return base + lowest * qk[0];
}

// Wichura (1988) AS241 PPND16, as in R's qnorm for the standard normal lower tail
inline double r_qnorm(double p) {
  const double q = p - 0.5;
  if (std::fabs(q) <= 0.425) {
    const double r = 0.180625 - q * q;
    const double num = (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                            45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                         133.14166789178437745) * r + 3.387132872796366608);
    const double den = (((((((r * 5226.495278852854561 + 28729.085735721942674) * r + 39307.89580009271061) * r +
                            21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
                         42.313330701600911252) * r + 1.0);
    return q * num / den;
  }
  double r = std::sqrt(-std::log(q < 0 ? p : 1.0 - p));
  double val;
  if (r <= 5.0) {
    r += -1.6;
    val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
              1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
            4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
               .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
            2.05319162663775882187) * r + 1.0);
  } else {
    r += -5.0;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
              .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
            5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
            .59983220655588793769) * r + 1.0);
  }
  return q < 0 ? -val : val;
}

// R snorm.c, INVERSION
inline double r_norm_rand(UniformSource& g) {
  const double two27 = 134217728.0;
  double u = g.next();
  u = (double)(int)(two27 * u) + g.next();
  return r_qnorm(u / two27);
}

// R rgamma.c (shape a, scale)
inline double r_rgamma(UniformSource& g, double a, double scale) {
  if (std::isnan(a) || std::isnan(scale)) return std::numeric_limits<double>::quiet_NaN();
  if (a <= 0.0 || scale <= 0.0) return (scale == 0.0 || a == 0.0) ? 0.0 : std::numeric_limits<double>::quiet_NaN();
  if (!std::isfinite(a) || !std::isfinite(scale)) return std::numeric_limits<double>::infinity();

  if (a < 1.0) {  // GS
    const double e = 1.0 + 0.36787944117144233 * a;
    double x;
    for (;;) {
      const double p = e * g.next();
      if (p >= 1.0) {
        x = -std::log((e - p) / a);
        if (r_exp_rand(g) >= (1.0 - a) * std::log(x)) break;
      } else {
        x = std::exp(std::log(p) / a);
        if (r_exp_rand(g) >= x) break;
      }
    }
    return scale * x;
  }

  // GD
  const double s2 = a - 0.5, s = std::sqrt(s2), d = 5.656854 - s * 12.0;
  double t = r_norm_rand(g);
  double x = s + 0.5 * t;
  const double first = x * x;
  if (t >= 0.0) return scale * first;
  double u = g.next();
  if (d * u <= t * t * t) return scale * first;

  const double r = 1.0 / a;
  const double q0 = ((((((2.424e-4 * r + 2.4511e-4) * r + -7.388e-5) * r + 0.00144121) * r + 0.00801191) * r +
                      0.02083148) * r + 0.04166669) * r;
  double b, si, c;
  if (a <= 3.686) { b = 0.463 + s + 0.178 * s2; si = 1.235; c = 0.195 / s - 0.079 + 0.16 * s; }
  else if (a <= 13.022) { b = 1.654 + 0.0076 * s2; si = 1.68 / s + 0.275; c = 0.062 / s + 0.024; }
  else { b = 1.77; si = 0.75; c = 0.1515 / s; }

  auto qfun = [&](double tt) {
    const double v = tt / (s + s);
    if (std::fabs(v) <= 0.25)
      return q0 + 0.5 * tt * tt *
                      ((((((0.1233795 * v + -0.1367177) * v + 0.1423657) * v + -0.1662921) * v + 0.2000062) * v +
                        -0.250003) * v + 0.3333333) * v;
    return q0 - s * tt + 0.25 * tt * tt + (s2 + s2) * std::log(1.0 + v);
  };

  if (x > 0.0 && std::log(1.0 - u) <= qfun(t)) return scale * first;
  for (;;) {
    const double e = r_exp_rand(g);
    u = g.next();
    u = u + u - 1.0;
    t = (u < 0.0) ? b - si * e : b + si * e;
    if (t >= -0.71874483771719) {
      const double q = qfun(t);
      if (q > 0.0) {
        const double w = std::expm1(q);
        if (c * std::fabs(u) <= w * std::exp(e - 0.5 * t * t)) break;
      }
    }
  }
  x = s + 0.5 * t;
  return scale * x * x;
}

}  // namespace host
}  // namespace pm
