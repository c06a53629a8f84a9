// log p(y | Q) by matrix exponentiation — the extra column of the DIC chains maketreelistMCMC2sDICt / ksDICt
// (reference src/phylomap.cpp:3135-3178 PPmakePLD, :3268-3297 PPmakePLksD, :3242-3250, :3383-3390).
//   k_transprob  one thread per branch: P(t_e) = exp(Q t_e) in FP64 (scaling and squaring of a degree-18 Taylor series;
//                the reference calls arma::expmat, a Pade scheme — same matrix to ~1e-15)
//   k_loglik     Felsenstein pruning with those matrices, same tiling as k_prune (block = 32 sites x 8 warps, level by
//                level), every node rescaled to sum 1 and the logs of the scale factors accumulated in FP64 per site;
//                block partial = sum over its sites of log(sum_j pid_j PL[root][j]) + S
//   k_reduce_ll  fixed-order sum of the block partials into the statistics row
#pragma once
#include "pm_kernels.cuh"

namespace pm {

template <typename Real>
__global__ void k_transprob(const double* __restrict__ Qrow /* n*n row-major */, const double* __restrict__ elen, int E, int n,
                            Real* __restrict__ TP /* [E][n*n] row-major */) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double X[PM_DIC_NMAX * PM_DIC_NMAX], Ex[PM_DIC_NMAX * PM_DIC_NMAX], T[PM_DIC_NMAX * PM_DIC_NMAX], N2[PM_DIC_NMAX * PM_DIC_NMAX];
  const double t = elen[e];
  double nrm = 0;
  for (int i = 0; i < n; i++) { double r = 0; for (int j = 0; j < n; j++) r += fabs(Qrow[i * n + j] * t); nrm = fmax(nrm, r); }
  int sq = 0;
  while (nrm > 0.5) { nrm *= 0.5; sq++; }
  const double sc = ldexp(1.0, -sq) * t;
  for (int i = 0; i < n * n; i++) { X[i] = Qrow[i] * sc; Ex[i] = 0; T[i] = 0; }
  for (int i = 0; i < n; i++) { Ex[i * n + i] = 1.0; T[i * n + i] = 1.0; }
  for (int k = 1; k <= 18; k++) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int l = 0; l < n; l++) acc += T[i * n + l] * X[l * n + j];
      N2[i * n + j] = acc / k;
    }
    for (int i = 0; i < n * n; i++) { T[i] = N2[i]; Ex[i] += N2[i]; }
  }
  for (int r = 0; r < sq; r++) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      double acc = 0;
      for (int l = 0; l < n; l++) acc += Ex[i * n + l] * Ex[l * n + j];
      N2[i * n + j] = acc;
    }
    for (int i = 0; i < n * n; i++) Ex[i] = N2[i];
  }
  for (int i = 0; i < n * n; i++) TP[(size_t)e * n * n + i] = (Real)Ex[i];
}

template <typename Real, int NS>
__global__ void __launch_bounds__(256) k_loglik(ChainParams<Real> P, const Real* __restrict__ TP, double* __restrict__ ll_partial) {
  constexpr int NC = NS > 0 ? NS : PM_NMAX;  // the direct sampler prunes with this kernel for any n <= PM_NMAX (the DIC chains stop at PM_DIC_NMAX)
  const int n = NS > 0 ? NS : P.n;
  __shared__ double s_S[8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const long long S = P.S;
  const long long site_raw = (long long)blockIdx.x * 32 + lane;
  const bool active = site_raw < S;
  const long long site = active ? site_raw : S - 1;
  const bool parity = P.parity_tips != 0;
  const int T = P.T;
  double Sacc = 0.0;
  for (int l = 0; l < P.n_up_levels; l++) {
    const int beg = __ldg(P.up_off + l), end = __ldg(P.up_off + l + 1);
    for (int idx = beg + warp; idx < end; idx += nw) {
      const int* en = P.up_entries + 5 * idx;
      const int pn = __ldg(en), a = __ldg(en + 1), ea = __ldg(en + 2), b = __ldg(en + 3), eb = __ldg(en + 4);
      Real ca[NC], cb[NC];
      if (a < T) tip_partial<Real, NC>(P.tipcode[(long long)a * P.TS + site], n, parity, ca);
      else VecIO<Real, NS>::load(P.PL + ((long long)(a - T) * S + site) * n, n, ca);
      if (b < T) tip_partial<Real, NC>(P.tipcode[(long long)b * P.TS + site], n, parity, cb);
      else VecIO<Real, NS>::load(P.PL + ((long long)(b - T) * S + site) * n, n, cb);
      const Real* __restrict__ Pa = TP + (size_t)ea * n * n;
      const Real* __restrict__ Pb = TP + (size_t)eb * n * n;
      Real out[NC];
      Real sum = 0;
#pragma unroll
      for (int r = 0; r < n; r++) {
        Real xa = 0, xb = 0;
#pragma unroll
        for (int c = 0; c < n; c++) { xa += __ldg(Pa + r * n + c) * ca[c]; xb += __ldg(Pb + r * n + c) * cb[c]; }
        out[r] = xa * xb;
        sum += out[r];
      }
      Sacc += log((double)sum);
      const Real inv = (Real)1 / sum;
#pragma unroll
      for (int r = 0; r < n; r++) out[r] *= inv;
      if (active) VecIO<Real, NS>::store(P.PL + ((long long)(pn - T) * S + site) * n, n, out);
    }
    __syncthreads();
  }
  s_S[warp][lane] = Sacc;
  __syncthreads();
  if (warp == 0) {
    double tot = 0;
    for (int w = 0; w < nw; w++) tot += s_S[w][lane];
    Real pl[NC];
    VecIO<Real, NS>::load(P.PL + ((long long)(P.root - T) * S + site) * n, n, pl);
    double X = 0;
    for (int j = 0; j < n; j++) X += (double)pl[j] * (double)P.model[2 * n * n + j];  // pid
    double v = active ? log(X) + tot : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (lane == 0) ll_partial[blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(256) k_reduce_ll(const double* __restrict__ ll_partial, long long nblocks, double* out) {
  __shared__ double sh[256];
  double acc = 0;
  for (long long b = threadIdx.x; b < nblocks; b += 256) acc += ll_partial[b];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) *out = sh[0];
}

}  // namespace pm
