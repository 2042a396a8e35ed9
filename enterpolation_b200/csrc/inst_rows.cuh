// Span-row B-spline kernel instantiations for one (scalar type, division mode, projection, knot base).
#pragma once
#include "inst_common.cuh"
#include "rows.cuh"

namespace enterp {

template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2>
cudaError_t rows_launch_one(const CurveDev<R>& c, const SampleSrc<R>& s, R* out, const RowsDev<R>& rw, cudaStream_t st) {
  auto kernel = bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2>;
  static bool configured = false;  // per instantiation; benign race (idempotent attribute)
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROWS_MAX_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (s.count == 0) return cudaSuccess;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BLOCK, rw.bytes) != cudaSuccess || per_sm < 1)
    per_sm = 1;
  const unsigned long long want = (s.count + ROWS_S * BLOCK - 1) / (ROWS_S * BLOCK);
  const unsigned long long cap = (unsigned long long)per_sm * (unsigned long long)sm_count();
  const int grid = (int)(want < cap ? want : cap);
  kernel<<<grid, BLOCK, rw.bytes, st>>>(c, s, out, rw);
  count_launch();
  return cudaGetLastError();
}

template <class R, int DH, bool PROJ, bool EQUI, bool POW2>
cudaError_t rows_pick_degree(int deg, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, const RowsDev<R>& rw,
                             cudaStream_t st) {
  switch (deg) {
    case 1: return rows_launch_one<R, DH, PROJ, 1, EQUI, POW2>(c, s, out, rw, st);
    case 2: return rows_launch_one<R, DH, PROJ, 2, EQUI, POW2>(c, s, out, rw, st);
    case 3: return rows_launch_one<R, DH, PROJ, 3, EQUI, POW2>(c, s, out, rw, st);
    case 4: return rows_launch_one<R, DH, PROJ, 4, EQUI, POW2>(c, s, out, rw, st);
    case 5: return rows_launch_one<R, DH, PROJ, 5, EQUI, POW2>(c, s, out, rw, st);
    case 6: return rows_launch_one<R, DH, PROJ, 6, EQUI, POW2>(c, s, out, rw, st);
    case 7: return rows_launch_one<R, DH, PROJ, 7, EQUI, POW2>(c, s, out, rw, st);
  }
  return cudaErrorInvalidValue;
}

template <class R, bool POW2, bool PROJ, bool EQUI>
cudaError_t launch_rows_part_impl(int dh, int deg, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                                  const RowsDev<R>& rw, cudaStream_t st) {
  constexpr int B = PROJ ? 1 : 0;  // weighted curves carry dim + 1 components
  switch (dh - B) {
    case 1: return rows_pick_degree<R, 1 + B, PROJ, EQUI, POW2>(deg, c, s, out, rw, st);
    case 2: return rows_pick_degree<R, 2 + B, PROJ, EQUI, POW2>(deg, c, s, out, rw, st);
    case 3: return rows_pick_degree<R, 3 + B, PROJ, EQUI, POW2>(deg, c, s, out, rw, st);
    case 4: return rows_pick_degree<R, 4 + B, PROJ, EQUI, POW2>(deg, c, s, out, rw, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace enterp

#define ENTERP_DEFINE_ROWS_PART(R, POW2, PROJ, EQUI)                                                                  \
  namespace enterp {                                                                                                  \
  template <>                                                                                                         \
  cudaError_t launch_rows_part<R, POW2, PROJ, EQUI>(int dh, int deg, const CurveDev<R>& c, const SampleSrc<R>& s,     \
                                                    R* out, const RowsDev<R>& rw, cudaStream_t st) {                  \
    return launch_rows_part_impl<R, POW2, PROJ, EQUI>(dh, deg, c, s, out, rw, st);                                    \
  }                                                                                                                   \
  }
