// Shared by the inst_*.cu translation units: launch one instantiated kernel.
#pragma once
#include "launch.hpp"

namespace enterp {

template <class K, class... Args>
cudaError_t launch_grid_stride(K kernel, unsigned long long work_items, int items_per_block, cudaStream_t st,
                               Args... args) {
  if (work_items == 0) return cudaSuccess;
  const int grid = grid_for(kernel, work_items, items_per_block);
  kernel<<<grid, BLOCK, 0, st>>>(args...);
  count_launch();
  return cudaGetLastError();
}

// Launch for kernels that may stage `smem` bytes of pivot levels: picks the block size (one fat CTA per SM
// when the table is big, as far as the kernel's launch bounds and registers allow) and the matching grid.
template <class K, class... Args>
cudaError_t launch_searching(K kernel, unsigned long long work_items, size_t smem, cudaStream_t st, Args... args) {
  if (work_items == 0) return cudaSuccess;
  int block = BLOCK;
  if (smem > 0) {
    cudaFuncAttributes fa{};
    cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return e;
    if (smem > 32u * 1024u) {  // few CTAs per SM fit: make them fat
      int cap = fa.maxThreadsPerBlock < MAX_BLOCK ? fa.maxThreadsPerBlock : MAX_BLOCK;
      const int by_regs = fa.numRegs > 0 ? (65536 / fa.numRegs) / 256 * 256 : cap;
      if (by_regs < cap) cap = by_regs;
      block = cap >= 256 ? cap : 256;
    }
  }
  cudaError_t eo = cudaSuccess;
  const int per_sm =
      cached_occupancy(reinterpret_cast<const void*>(kernel), block, smem, smem > 48u * 1024u ? SEARCH_SMEM_MAX : 0, &eo);
  if (eo != cudaSuccess) return eo;
  unsigned long long want = (work_items + block - 1) / block;
  const unsigned long long cap_grid = (unsigned long long)per_sm * (unsigned long long)sm_count();
  const int grid = (int)(want < cap_grid ? want : cap_grid);
  kernel<<<grid, block, smem, st>>>(args...);
  count_launch();
  return cudaGetLastError();
}

}  // namespace enterp
