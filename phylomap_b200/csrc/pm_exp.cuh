// The direct ("matrix exponentiation") sampler maketreelistEXP (reference src/phylomap.cpp:2877-3051 and the branch
// sampler newunifSample :93-208): every iteration draws an INDEPENDENT history — P(t_e) = |L exp(D t_e) R| from the
// caller's eigendecomposition of Q (:2980, the abs() is the reference's), Felsenstein pruning with those matrices, node
// states top-down, then per branch an endpoint-conditioned uniformized path: the number of (real + virtual) jumps by
// inversion of Pois(k; Omega t) (B^k)[a,b] / P_ab(t) (cap 300, :120), jump times uniform on the branch, interior states
// by forward sampling against the backward vectors B^j e_b, virtual jumps dropped.  No state is carried over.
//   k_transprob_eig  thread per branch
//   k_loglik         (pm_loglik.cuh) does the pruning; its log-likelihood by-product is simply not used
//   k_exp_nodes      K2's tiling with rows of P(t_e) instead of rows of B^(m-1)
//   k_exp_branches   thread = (site, chunk of branches); block partial sums like the MCMC path kernels
// Production arithmetic only (the jump times come out of the order-statistics recurrence instead of a sort); the
// statistics are checked against the oracle's restatement of the reference (tests/test_gpu_production.py).
#pragma once
#include "pm_kernels.cuh"

namespace pm {

template <typename Real>
__global__ void k_transprob_eig(const double* __restrict__ L /* n*n row-major */, const double* __restrict__ R,
                                const double* __restrict__ d /* n eigenvalues */, const double* __restrict__ elen, int E, int n,
                                Real* __restrict__ TP) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double t = elen[e];
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int l = 0; l < n; l++) acc += L[i * n + l] * exp(d[l] * t) * R[l * n + j];
      TP[(size_t)e * n * n + i * n + j] = (Real)fabs(acc);
    }
}

template <typename Real, int NS>
__global__ void __launch_bounds__(256) k_exp_nodes(ChainParams<Real> P, const Real* __restrict__ TP, uint32_t iter) {
  constexpr int NC = NS > 0 ? NS : PM_NMAX;
  typedef typename StreamSel<Real, false>::type Stream;
  const int n = NS > 0 ? NS : P.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long S = P.S;
  const long long site_raw = (long long)blockIdx.x * 32 + lane;
  const bool active = site_raw < S;
  const long long site = active ? site_raw : S - 1;
  const bool parity = P.parity_tips != 0;
  const int T = P.T;
  if (warp == 0) {  // root :2925-2933
    Real w[NC], pl[NC];
    VecIO<Real, NS>::load(P.PL + ((long long)(P.root - T) * S + site) * n, n, pl);
#pragma unroll
    for (int j = 0; j < n; j++) w[j] = P.model[2 * n * n + j] * pl[j];
    Stream g; g.open(P.rng, (uint32_t)site, iter, K_NODE, (uint32_t)P.root, P.err_flag);
    const int s = categorical<Real, NC, false>(w, n, g.next(), P.err_flag);
    if (active) {
      P.node_state[(long long)P.root * S + site] = (uint8_t)s;
      if (P.rng.site0 + (uint32_t)site == 0u) *P.root_out = s;
    }
  }
  __syncthreads();
  for (int l = 0; l < P.n_down_levels; l++) {
    const int beg = __ldg(P.down_off + l), end = __ldg(P.down_off + l + 1);
    for (int idx = beg + warp; idx < end; idx += nw) {
      const int* en = P.down_entries + 3 * idx;
      const int v = __ldg(en), pn = __ldg(en + 1), e = __ldg(en + 2);
      const int ps = P.node_state[(long long)pn * S + site];
      Real w[NC], pl[NC];
      if (v < T) tip_partial<Real, NC>(P.tipcode[(long long)v * P.TS + site], n, parity, pl);
      else VecIO<Real, NS>::load(P.PL + ((long long)(v - T) * S + site) * n, n, pl);
#pragma unroll
      for (int j = 0; j < n; j++) w[j] = __ldg(TP + (size_t)e * n * n + ps * n + j) * pl[j];  // :2950
      Stream g; g.open(P.rng, (uint32_t)site, iter, K_NODE, (uint32_t)v, P.err_flag);
      const int s = categorical<Real, NC, false>(w, n, g.next(), P.err_flag);
      if (active) P.node_state[(long long)v * S + site] = (uint8_t)s;
    }
    __syncthreads();
  }
}

template <typename Real, int NS>
__global__ void __launch_bounds__(128) k_exp_branches(ChainParams<Real> P, const Real* __restrict__ TP,
                                                     const double* __restrict__ elen, Real omega, uint32_t iter, int chunk) {
  constexpr int NC = NS > 0 ? NS : PM_NMAX;
  constexpr int NR = NS > 0 ? NS : 1;
  typedef Pin<Real> PN;
  const int n = NS > 0 ? NS : P.n;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_dw = reinterpret_cast<double*>(smem_raw);            // [4 warps][n] (NS>0) or [n] atomics (NS==0)
  unsigned* s_cnt = reinterpret_cast<unsigned*>(s_dw + 4 * n);  // [n*n]
  Real* sB = reinterpret_cast<Real*>(s_cnt + n * n + ((n * n) & 1));
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) { s_cnt[i] = 0; sB[i] = P.model[i]; }
  for (int i = threadIdx.x; i < 4 * n; i += blockDim.x) s_dw[i] = 0.0;
  __syncthreads();
  const long long S = P.S;
  const long long site_raw = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = site_raw < S;
  const long long site = active ? site_raw : S - 1;
  const int e0 = blockIdx.y * chunk, e1 = min(P.E, e0 + chunk);
  double Rsum[NR];
#pragma unroll
  for (int j = 0; j < NR; j++) Rsum[j] = 0;
  unsigned errbits = 0;
  auto add_dwell = [&](int s, Real L) {
    if (NS > 0) {
#pragma unroll
      for (int j = 0; j < NR; j++) Rsum[j] += (s == j) ? (double)L : 0.0;
    } else atomicAdd(&s_dw[s], (double)L);
  };
  // off-diagonal counts only, in the full n x n layout (the host converts to the reference's n(n-1) columns)
  if (active) for (int e = e0; e < e1; e++) {
    const int a = P.node_state[(long long)__ldg(P.e_parent + e) * S + site];
    const int b = P.node_state[(long long)__ldg(P.e_child + e) * S + site];
    const Real t = (Real)__ldg(elen + e);
    const Real Pab = __ldg(TP + (size_t)e * n * n + a * n + b);
    const Real lam = omega * t;
    WordStream g; g.open(P.rng, (uint32_t)site, iter, K_BRSTATE, (uint32_t)e, 0u);
    // number of jumps of the dominating chain, :103-133
    const double rU = (double)PN::u01(g.next());
    double pk = exp(-(double)lam);  // Pois(0; lam); the Poisson recurrence runs in FP64 (lam can reach the hundreds)
    double cum = (a == b) ? pk / (double)Pab : 0.0;
    int nj = 0;
    bool broken = false;
    while (!(cum > rU)) {
      nj++;
      if (nj > 300 || nj >= P.jcap) { broken = true; break; }
      pk = pk * (double)lam / (double)nj;
      cum += pk * (double)__ldg(P.ppow + (size_t)nj * n * n + a * n + b) / (double)Pab;
    }
    if (broken) continue;  // like the reference (:120-125): the branch contributes nothing to this row
    if (nj == 0 || (nj == 1 && a == b)) { add_dwell(a, t); continue; }
    if (nj == 1) {  // one real jump at a uniform time, :147
      const Real tu = t * PN::u01(g.next());
      add_dwell(a, tu); add_dwell(b, t - tu);
      atomicAdd(&s_cnt[a * n + b], 1u);
      continue;
    }
    // nj >= 2: jump times = order statistics of nj uniforms (generated in increasing order), states by forward sampling
    // with weights B[prev, .] * (B^(nj-i) e_b), sampleOnce rule (u < cum) :81-90, :158-160
    Real x = 0, last_change = 0;
    int prev = a;
    for (int i = 1; i <= nj; i++) {
      x = next_order_stat<Real>(x, t, nj - i + 1, g.next());
      int st;
      if (i == nj) st = b;
      else {
        Real w[NC];
        Real tot = 0;
#pragma unroll
        for (int c = 0; c < n; c++) { w[c] = sB[prev * n + c] * __ldg(P.ppow + (size_t)(nj - i) * n * n + c * n + b); tot += w[c]; }
        const Real u = PN::u01(g.next());
        Real cc = 0;
        st = n - 1;
#pragma unroll
        for (int c = 0; c < n; c++) { cc += w[c] / tot; if (st == n - 1 && u < cc) st = c; }
      }
      if (st != prev) {
        add_dwell(prev, x - last_change);
        atomicAdd(&s_cnt[prev * n + st], 1u);
        last_change = x; prev = st;
      }
    }
    add_dwell(prev, t - last_change);
  }
  if (errbits) atomicOr(P.err_flag, errbits);

  if (NS > 0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NR; j++) {
      double v = Rsum[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) s_dw[warp * n + j] = v;
    }
  }
  __syncthreads();
  const long long blk = (long long)blockIdx.y * gridDim.x + blockIdx.x;
  if ((int)threadIdx.x < n) {
    double v;
    if (NS > 0) { v = 0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) v += s_dw[w * n + threadIdx.x]; }
    else v = s_dw[threadIdx.x];
    P.dw_partial[blk * n + threadIdx.x] = v;
  }
  for (int i = threadIdx.x; i < n * n; i += blockDim.x) if (s_cnt[i]) atomicAdd(&P.cnt[i], (unsigned long long)s_cnt[i]);
}

}  // namespace pm
