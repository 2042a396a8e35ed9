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

}  // namespace enterp
