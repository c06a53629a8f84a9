// TEST INFRASTRUCTURE ONLY (oracle). Not linked into, imported by, or executed from the product path.
//
// Restatement of the third-party random-number arithmetic the reference consumes through Rcpp
// (SURVEY.md §8(c), Appendix A.6).  None of this code lives under /root/reference: it is R's
// nmath / RNG.c, restated from the published algorithms.  Not checked against a real R build (no R in this
// container); it IS pinned against widely published R outputs (tests/test_oracle_rng.py) -- rgamma by
// Kolmogorov-Smirnov tests in every regime of rgamma.c, no R known answers for it were recoverable offline:
//   set.seed(1);   runif(3) = 0.2655087 0.3721239 0.5728534
//   set.seed(42);  runif(2) = 0.9148060 0.9370754
//   set.seed(123); runif(3) = 0.2875775 0.7883051 0.4089769
//   set.seed(1);   rexp(3)  = 0.7551818 1.1816428 0.1457067
//   set.seed(1);   rnorm(3) = -0.6264538 0.1836433 -0.8356286
//   set.seed(123); rnorm(3) = -0.56047565 -0.23017749 1.55870831
//
// Call sites in the reference that reach this arithmetic:
//   runif   src/phylomap.cpp:103,147,151,159,1210,1242,1471,...,2332,2347,2810,2825
//   rexp    src/phylomap.cpp:348,398,1059,2125          (Rcpp sugar: scale*exp_rand(), scale=1/rate)
//   rgamma  src/phylomap.cpp:1202,1235,1463,1538,1612,1681,1748,2204,2241,2399,...
//   sample  19 call sites (RcppArmadillo::sample -> unif_rand), see oracle header.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace orc {

// Source of unif_rand() values.  The sweep code is written once against this interface; the three
// implementations are: R's Mersenne-Twister consumed sequentially (what the reference does), a keyed
// Philox stream (what a GPU can index), and a replay table (uniforms exported from a sequential run).
struct UnifSource {
  virtual ~UnifSource() {}
  virtual double unif() = 0;
};

// R's default generator: MT19937 with set.seed() scrambling (RNG.c: Randomize / MT_sgenrand / MT_genrand / fixup).
struct RMersenne : UnifSource {
  uint32_t mt[624];
  int mti;
  explicit RMersenne(uint32_t seed) { set_seed(seed); }
  void set_seed(uint32_t seed) {
    // initial scrambling: 50 LCG steps, then 625 words of which the first ("mti") is forced to 624.
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    uint32_t dummy0 = 0;
    for (int j = 0; j < 625; j++) {
      seed = 69069u * seed + 1u;
      if (j == 0) dummy0 = seed; else mt[j - 1] = seed;
    }
    (void)dummy0;
    mti = 624;
  }
  uint32_t next_u32() {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t y;
    if (mti >= 624) {
      int kk;
      for (kk = 0; kk < 624 - 397; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
      }
      for (; kk < 623; kk++) {
        y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
        mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
      }
      y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
      mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
      mti = 0;
    }
    y = mt[mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
  }
  double unif() override {
    const double i2_32m1 = 2.328306437080797e-10;  // 1/(2^32 - 1)
    double v = next_u32() * 2.3283064365386963e-10;  // in [0,1)
    if (v <= 0.0) return 0.5 * i2_32m1;
    if ((1.0 - v) <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
  }
};

// R sexp.c: Ahrens & Dieter (1972).  Consumes a data-dependent number of uniforms.
inline double exp_rand(UnifSource& g) {
  static const double q[] = {
      0.6931471805599453, 0.9333736875190459, 0.9888777961838675, 0.9984589039328340,
      0.9998292811061389, 0.9999833164100727, 0.9999985691438767, 0.9999998906925558,
      0.9999999924734159, 0.9999999995283275, 0.9999999999728814, 0.9999999999985598,
      0.9999999999999289, 0.9999999999999968, 0.9999999999999999, 1.0000000000000000};
  double a = 0.;
  double u = g.unif();
  while (u <= 0. || u >= 1.) u = g.unif();
  for (;;) {
    u += u;
    if (u > 1.) break;
    a += q[0];
  }
  u -= 1.;
  if (u <= q[0]) return a + u;
  int i = 0;
  double ustar = g.unif(), umin = ustar;
  do {
    ustar = g.unif();
    if (umin > ustar) umin = ustar;
    i++;
  } while (u > q[i]);
  return a + umin * q[0];
}

// Rcpp sugar rexp(1, rate): scale = 1/rate; non-finite or non-positive scale -> 0 or NaN WITHOUT consuming uniforms.
inline double rexp_rate(UnifSource& g, double rate) {
  double scale = 1.0 / rate;
  if (!std::isfinite(scale) || scale <= 0.0) {
    if (scale == 0.) return 0.0;
    return std::nan("");
  }
  return scale * exp_rand(g);
}

// Wichura AS241 (PPND16), lower tail, as used by R's qnorm5.
inline double qnorm_std(double p) {
  double q = p - 0.5, r, val;
  if (std::fabs(q) <= 0.425) {
    r = 0.180625 - q * q;
    val = q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r +
                    45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r +
                 133.14166789178437745) * r + 3.387132872796366608) /
          (((((((r * 5226.495278852854561 + 28729.085735721942674) * r + 39307.89580009271061) * r +
               21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r +
            42.313330701600911252) * r + 1.);
    return val;
  }
  r = (q < 0) ? p : 1.0 - p;
  r = std::sqrt(-std::log(r));
  if (r <= 5.) {
    r += -1.6;
    val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r +
              1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r +
            4.6303378461565452959) * r + 1.42343711074968357734) /
          (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r +
               .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r +
            2.05319162663775882187) * r + 1.);
  } else {
    r += -5.;
    val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r +
              .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r +
            5.4637849111641143699) * r + 6.6579046435011037772) /
          (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r +
               7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r +
            .59983220655588793769) * r + 1.);
  }
  if (q < 0.0) val = -val;
  return val;
}

// R snorm.c, INVERSION kind (the default): two uniforms per deviate.
inline double norm_rand(UnifSource& g) {
  const double BIG = 134217728;  // 2^27
  double u = g.unif();
  u = (int)(BIG * u) + g.unif();
  return qnorm_std(u / BIG);
}

// R rgamma.c: Ahrens & Dieter GD (1982) for a >= 1, GS (1974) for a < 1.  R caches s,s2,d,q0,b,si,c in statics
// keyed by `a`; recomputing them every call gives identical values.
inline double rgamma(UnifSource& g, double a, double scale) {
  const double sqrt32 = 5.656854;
  const double exp_m1 = 0.36787944117144233;
  const double q1 = 0.04166669, q2 = 0.02083148, q3 = 0.00801191, q4 = 0.00144121, q5 = -7.388e-5, q6 = 2.4511e-4,
               q7 = 2.424e-4;
  const double a1 = 0.3333333, a2 = -0.250003, a3 = 0.2000062, a4 = -0.1662921, a5 = 0.1423657, a6 = -0.1367177,
               a7 = 0.1233795;
  double e, p, q, r, t, u, v, w, x, ret_val;
  if (std::isnan(a) || std::isnan(scale)) return std::nan("");
  if (a <= 0.0 || scale <= 0.0) {
    if (scale == 0. || a == 0.) return 0.;
    return std::nan("");
  }
  if (!std::isfinite(a) || !std::isfinite(scale)) return INFINITY;

  if (a < 1) {
    e = 1.0 + exp_m1 * a;
    for (;;) {
      p = e * g.unif();
      if (p >= 1.0) {
        x = -std::log((e - p) / a);
        if (exp_rand(g) >= (1.0 - a) * std::log(x)) break;
      } else {
        x = std::exp(std::log(p) / a);
        if (exp_rand(g) >= x) break;
      }
    }
    return scale * x;
  }
  double s2 = a - 0.5;
  double s = std::sqrt(s2);
  double d = sqrt32 - s * 12;
  t = norm_rand(g);
  x = s + 0.5 * t;
  ret_val = x * x;
  if (t >= 0) return scale * ret_val;
  u = g.unif();
  if (d * u <= t * t * t) return scale * ret_val;
  r = 1 / a;
  double q0 = ((((((q7 * r + q6) * r + q5) * r + q4) * r + q3) * r + q2) * r + q1) * r;
  double b, si, c;
  if (a <= 3.686) {
    b = 0.463 + s + 0.178 * s2;
    si = 1.235;
    c = 0.195 / s - 0.079 + 0.16 * s;
  } else if (a <= 13.022) {
    b = 1.654 + 0.0076 * s2;
    si = 1.68 / s + 0.275;
    c = 0.062 / s + 0.024;
  } else {
    b = 1.77;
    si = 0.75;
    c = 0.1515 / s;
  }
  if (x > 0.0) {
    v = t / (s + s);
    if (std::fabs(v) <= 0.25)
      q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
    else
      q = q0 - s * t + 0.25 * t * t + (s2 + s2) * std::log(1.0 + v);
    if (std::log(1.0 - u) <= q) return scale * ret_val;
  }
  for (;;) {
    e = exp_rand(g);
    u = g.unif();
    u = u + u - 1.0;
    if (u < 0.0) t = b - si * e; else t = b + si * e;
    if (t >= -0.71874483771719) {
      v = t / (s + s);
      if (std::fabs(v) <= 0.25)
        q = q0 + 0.5 * t * t * ((((((a7 * v + a6) * v + a5) * v + a4) * v + a3) * v + a2) * v + a1) * v;
      else
        q = q0 - s * t + 0.25 * t * t + (s2 + s2) * std::log(1.0 + v);
      if (q > 0.0) {
        w = std::expm1(q);
        if (c * std::fabs(u) <= w * std::exp(e - 0.5 * t * t)) break;
      }
    }
  }
  x = s + 0.5 * t;
  return scale * x * x;
}

// Philox4x32-10 (Salmon et al., SC'11) — the counter-based generator both the oracle's keyed mode and the
// GPU kernels evaluate.  Integer arithmetic only, so CPU and GPU agree bit for bit.
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Keyed uniform U(seed; site, iter, slot, k): two doubles per Philox block, 53 bits each, strictly inside (0,1).
inline double keyed_uniform(uint64_t seed, uint32_t site, uint32_t iter, uint32_t slot, uint32_t k) {
  uint32_t ctr[4] = {k >> 1, slot, iter, site};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  philox4x32_10(ctr, key, o);
  uint32_t hi = (k & 1u) ? o[2] : o[0];
  uint32_t lo = (k & 1u) ? o[3] : o[1];
  uint64_t bits = ((uint64_t)hi << 21) | (uint64_t)(lo >> 11);
  return ((double)bits + 0.5) * (1.0 / 9007199254740992.0);
}

}  // namespace orc
