// Host-callable launchers for the kernel families in kernels.cuh. Each family is instantiated in
// its own translation unit (inst_*.cu) so that nvcc can compile them in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.cuh"

namespace enterp {

// Total kernels launched by this library in this process (enterp_launch_count()).
void count_launch();
// SM count of the current device (cached per device).
int sm_count();

// Resident CTAs per SM of `kernel` at (block, smem) on the current device, cached: the occupancy query (and the
// opt-in to > 48 KB of dynamic shared memory it needs first) costs microseconds, which is most of a small launch.
// `opt_in_smem` > 0 raises cudaFuncAttributeMaxDynamicSharedMemorySize to that value once per (kernel, device).
int cached_occupancy(const void* kernel, int block, size_t smem, size_t opt_in_smem, cudaError_t* err);

// Grid for a grid-stride kernel: enough CTAs to fill every SM at the kernel's own occupancy, never
// more than the work needs. `kernel` is the __global__ function pointer.
template <class K>
int grid_for(K kernel, unsigned long long work_items, int items_per_block) {
  cudaError_t e = cudaSuccess;
  const int per_sm = cached_occupancy(reinterpret_cast<const void*>(kernel), BLOCK, 0, 0, &e);
  unsigned long long want = (work_items + items_per_block - 1) / items_per_block;
  unsigned long long cap = (unsigned long long)per_sm * (unsigned long long)sm_count();
  if (want < 1) want = 1;
  return (int)(want < cap ? want : cap);
}

// dh: components in the recursion (dim + weighted); proj: weighted -> divide by the last component.
// deg: 1..7 compile-time degree, 0 = runtime degree (<= MAX_DYN_DEGREE).
template <class R>
cudaError_t launch_bspline(int dh, bool proj, int deg, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                           bool vec_ok, cudaStream_t st);
// Span-row fast path (rows.cuh): deg 1..7, table <= ROWS_MAX_BYTES. POW2: exact inverse multiply.
// One translation unit per (R, POW2, PROJ, EQUI) so that nvcc compiles them in parallel.
template <class R, bool POW2, bool PROJ, bool EQUI>
cudaError_t launch_rows_part(int dh, int deg, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                             const RowsDev<R>& rw, cudaStream_t st);
template <class R>
inline cudaError_t launch_bspline_rows(int dh, bool proj, int deg, bool equi, bool pow2, const CurveDev<R>& c,
                                       const SampleSrc<R>& s, R* out, const RowsDev<R>& rw, cudaStream_t st) {
  const int key = (pow2 ? 4 : 0) | (proj ? 2 : 0) | (equi ? 1 : 0);
  switch (key) {
    case 0: return launch_rows_part<R, false, false, false>(dh, deg, c, s, out, rw, st);
    case 1: return launch_rows_part<R, false, false, true>(dh, deg, c, s, out, rw, st);
    case 2: return launch_rows_part<R, false, true, false>(dh, deg, c, s, out, rw, st);
    case 3: return launch_rows_part<R, false, true, true>(dh, deg, c, s, out, rw, st);
    case 4: return launch_rows_part<R, true, false, false>(dh, deg, c, s, out, rw, st);
    case 5: return launch_rows_part<R, true, false, true>(dh, deg, c, s, out, rw, st);
    case 6: return launch_rows_part<R, true, true, false>(dh, deg, c, s, out, rw, st);
    default: return launch_rows_part<R, true, true, true>(dh, deg, c, s, out, rw, st);
  }
}
template <class R>
cudaError_t launch_linear(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                          cudaStream_t st);
// n: 1..16 compile-time element count, 0 = runtime (<= MAX_DYN_BEZIER)
template <class R>
cudaError_t launch_bezier(int dh, bool proj, int n, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                          cudaStream_t st);
// Lean kernels (lean.cuh): launch and return true when the call qualifies (status in *err), else false.
// host_elems: the curve's lifted control values on the host ([n][dh]). One translation unit per (R, BATCH).
template <class R, bool BATCH>
bool lean_linear_part(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                      cudaStream_t st, cudaError_t* err);
template <class R, bool BATCH>
bool lean_bezier_part(int dh, bool proj, int n, const R* host_elems, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                      bool vec_ok, cudaStream_t st, cudaError_t* err);
template <class R>
inline bool lean_linear(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                        cudaStream_t st, cudaError_t* err) {
  return s.t ? lean_linear_part<R, true>(dh, proj, equi, c, s, out, vec_ok, st, err)
             : lean_linear_part<R, false>(dh, proj, equi, c, s, out, vec_ok, st, err);
}
template <class R>
inline bool lean_bezier(int dh, bool proj, int n, const R* host_elems, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                        bool vec_ok, cudaStream_t st, cudaError_t* err) {
  return s.t ? lean_bezier_part<R, true>(dh, proj, n, host_elems, c, s, out, vec_ok, st, err)
             : lean_bezier_part<R, false>(dh, proj, n, host_elems, c, s, out, vec_ok, st, err);
}
template <class R>
cudaError_t launch_bezier_many(int dim, int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl,
                               uint32_t samples, R step, R* out, cudaStream_t st);
template <class R>
cudaError_t launch_span(int kind, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, uint32_t* idx, R* factor,
                        cudaStream_t st);
// Bezier::gen_with_tangent (tangent = 1, K = 2) / gen_with_deriatives::<K>
template <class R>
cudaError_t launch_bezier_deriv(int dim, const CurveDev<R>& c, const SampleSrc<R>& s, int K, int tangent, R* out,
                                cudaStream_t st);
template <class R>
cudaError_t launch_stepper(const SampleSrc<R>& s, R* out, cudaStream_t st);
// out[j] = step * fl((s.first + j) << ls) + start for j in [0, s.count)
template <class R>
cudaError_t launch_stepper_strided(const SampleSrc<R>& s, unsigned ls, R* out, cudaStream_t st);

// out[i] = vals[m(a + i)] (kernels.cuh runs_expand_kernel); words = 32-bit words per element (1..4)
cudaError_t launch_runs_expand(int words, const void* vals, void* out, unsigned long long a, unsigned long long count,
                               unsigned long long vb, unsigned ls, cudaStream_t st);

}  // namespace enterp
