// Set-up and reduction kernels (not templated on the arithmetic): included by pm_host.cu only.
#pragma once
#include "pm_device.cuh"

namespace pm {

// ------------------------------------------------------------------------------------------------
// K4: one row of sufficient statistics [R(n) | N(n*n) | root].  One block, fixed summation order.
// ------------------------------------------------------------------------------------------------
// row[err_slot] receives this rank's device error flag (summed over ranks by the all-reduce that follows).
__global__ void __launch_bounds__(256) k_reduce(const double* __restrict__ dw_partial, long long nblocks, int n,
                                                unsigned long long* cnt, const int* root, double* row, int accumulate,
                                                const unsigned* err_flag, int err_slot, uint32_t* ctl = nullptr,
                                                int row_stride = 0) {
  __shared__ double sh[256];
  if (ctl) row += (size_t)ctl[1] * row_stride;  // graph replay: the row index lives in device memory (ChainParams::ctl)
  for (int j = 0; j < n; j++) {
    double acc = 0;
    for (long long b = threadIdx.x; b < nblocks; b += 256) acc += dw_partial[b * n + j];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) row[j] = (accumulate ? row[j] : 0.0) + sh[0];
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n * n; i += 256) { row[n + i] = (accumulate ? row[n + i] : 0.0) + (double)cnt[i]; cnt[i] = 0ull; }
  if (threadIdx.x == 0 && !accumulate) row[n + n * n] = (double)(*root);
  if (threadIdx.x == 0 && err_flag) row[err_slot] = (double)(*err_flag);
  if (ctl) {  // this sweep is complete: the next replay of the graph runs the next sweep into the next row
    __syncthreads();
    if (threadIdx.x == 0) { ctl[0] += 1u; ctl[1] += 1u; }
  }
}

// The rows of a k_small_chain launch (pm_small.cuh): block i sums the per-site dwell times of sweep i in a fixed order and
// assembles row i in k_reduce's layout [R(n) | N(n*n) | root | .. | error flag].
// (the counters and the root slot are handed back zeroed, like k_reduce does: the next launch adds to them atomically)
__global__ void __launch_bounds__(128) k_small_reduce(const double* __restrict__ part, unsigned long long* cnt, int* root, long long S,
                                                      int n, double* rows, int row_stride, const unsigned* err_flag, int err_slot) {
  __shared__ double sh[128];
  const long long i = blockIdx.x;
  double* row = rows + (size_t)i * row_stride;
  for (int j = 0; j < n; j++) {
    double acc = 0;
    for (long long s = threadIdx.x; s < S; s += 128) acc += part[(i * S + s) * n + j];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) { if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) row[j] = sh[0];
    __syncthreads();
  }
  for (int k = threadIdx.x; k < n * n; k += 128) { row[n + k] = (double)cnt[i * n * n + k]; cnt[i * n * n + k] = 0ull; }
  if (threadIdx.x == 0) { row[n + n * n] = (double)root[i]; root[i] = 0; row[err_slot] = (double)(*err_flag); }
}

// ------------------------------------------------------------------------------------------------
// Set-up kernels.
// ------------------------------------------------------------------------------------------------
// Work items of the persistent path kernels, flattened: item i = branch | first ballot word | words - 1 (ChainParams::wk_item).
// Branch e owns the items [wk_off[e], wk_off[e + 1]), g_e = wk_g[e] words each (the last one what is left of the W words).
__global__ void k_build_items(const long long* __restrict__ wk_off, const int* __restrict__ wk_g, int E, long long W, long long total,
                              unsigned long long* items) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int lo = 0, hi = E - 1;  // the last branch with wk_off[e] <= i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (wk_off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  const long long g = wk_g[lo];
  const long long w0 = (i - wk_off[lo]) * g;
  const long long nw = min(g, W - w0);
  items[i] = ((unsigned long long)(unsigned)lo << 32) | ((unsigned long long)w0 << 5) | (unsigned long long)(nw - 1);
}

// %nsmid: the range of the SM identifiers a block can read from %smid (PTX: it may exceed the number of SMs the runtime
// reports, and the numbering need not be contiguous).  The fused prune + node-draw kernel indexes its slots by %smid.
__global__ void k_nsmid(int* out) {
  unsigned v;
  asm volatile("mov.u32 %0, %%nsmid;" : "=r"(v));
  *out = (int)v;
}

// states [ns][T] (a staged block of site rows; int32 1-based, or u8) -> tipcode / node_state [T][S] at site s_base.
// Tiled transpose through shared memory: reads coalesced along T, writes coalesced along S.  Validates the range.
template <typename In>
__global__ void __launch_bounds__(256) k_init_tips(const In* __restrict__ states, int ns, long long s_base, long long S,
                                                   long long TS, int T, int n, int parity, uint8_t* tipcode, uint8_t* node_state,
                                                   unsigned* err_flag) {
  __shared__ uint8_t tile[32][33];
  const int t0 = blockIdx.x * 32, s0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int s = s0 + r, t = t0 + threadIdx.x;
    if (s < ns && t < T) {
      const int v = (int)states[(long long)s * T + t];
      uint8_t code;
      if (parity) code = (uint8_t)(v & 1);
      else if (v < 1 || v > n) { atomicOr(err_flag, PM_DE_BAD_STATE); code = 0; }
      else code = (uint8_t)(v - 1);
      tile[r][threadIdx.x] = code;
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, s = s0 + threadIdx.x;
    if (s < ns && t < T) {
      const uint8_t code = tile[threadIdx.x][r];
      tipcode[(long long)t * TS + s_base + s] = code;
      node_state[(long long)t * S + s_base + s] = parity ? (uint8_t)(code ? 0 : 1) : code;
    }
  }
}

// meta <- number of pieces of the caller's map of every branch; production arithmetic also seeds the position field
// (upper 16 bits) with the only jump point of a two-piece map, so that the first sweep can treat such a branch like any
// later one (k_paths_easy) instead of walking the map in the general kernel.
template <typename Real>
__global__ void __launch_bounds__(256) k_init_meta(const long long* __restrict__ maps_off, const double* __restrict__ maps_len, long long S, int E,
                                                   uint32_t* meta, const Real* __restrict__ e_len, uint16_t* shape) {
  // grid: x over sites, y over branches (strided): the word of a branch is the same for all its sites
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (int e = blockIdx.y; e < E; e += gridDim.y) {
    const long long a = maps_off[e], m = maps_off[e + 1] - a;
    uint32_t w = (uint32_t)m;
    if (shape && m == 2) w |= pos_enc<Real>((Real)maps_len[a], e_len[e]) << 16;
    if (s < S) {
      meta[(long long)e * S + s] = w;
      if (shape) shape[(long long)e * S + s] = 0;
    }
  }
}

}  // namespace pm
