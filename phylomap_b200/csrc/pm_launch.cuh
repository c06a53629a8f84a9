// Launch interface between the host (pm_host.cu) and the sweep kernels, which are instantiated in separate
// translation units (pm_sweep_*.cu) so that they compile in parallel.
#pragma once
#include "pm_kernels.cuh"

namespace pm {

template <typename Real, int NS, bool EXACT>
struct Sweep {
  static void prune(const ChainParams<Real>& P, int grid, size_t smem, cudaStream_t st, int variant);
  static void nodes(const ChainParams<Real>& P, int grid, size_t smem, cudaStream_t st, uint32_t iter);
  static void paths(const ChainParams<Real>& P, dim3 grid, size_t smem, cudaStream_t st, uint32_t iter, int first, int chunk,
                    int hard_blocks);
};

}  // namespace pm
