#pragma once
#include "pm_launch.cuh"

namespace pm {

template <typename Real, int NS, bool EXACT>
void Sweep<Real, NS, EXACT>::prune(const ChainParams<Real>& P, int grid, size_t smem, cudaStream_t st) {
  k_prune<Real, NS, EXACT><<<grid, 256, smem, st>>>(P);
}
template <typename Real, int NS, bool EXACT>
void Sweep<Real, NS, EXACT>::nodes(const ChainParams<Real>& P, int grid, size_t smem, cudaStream_t st, uint32_t iter) {
  k_nodes<Real, NS, EXACT><<<grid, 256, smem, st>>>(P, iter);
}
template <typename Real, int NS, bool EXACT>
void Sweep<Real, NS, EXACT>::paths(const ChainParams<Real>& P, dim3 grid, size_t smem, cudaStream_t st, uint32_t iter,
                                   int first, int chunk) {
  k_paths<Real, NS, EXACT><<<grid, 128, smem, st>>>(P, iter, first, chunk);
}

}  // namespace pm
