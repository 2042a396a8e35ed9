// Span-row B-spline kernel instantiations for one (scalar type, division mode) pair.
#pragma once
#include "inst_common.cuh"
#include "rows.cuh"

namespace enterp {


template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2>
cudaError_t rows_launch_one(const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok, const RowsDev<R>& rw,
                            cudaStream_t st) {
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
  unsigned long long want = (s.count + ROWS_S * BLOCK - 1) / (ROWS_S * BLOCK);
  const unsigned long long cap = (unsigned long long)per_sm * (unsigned long long)sm_count();
  const int grid = (int)(want < cap ? want : cap);
  kernel<<<grid, BLOCK, rw.bytes, st>>>(c, s, out, rw);
  count_launch();
  return cudaGetLastError();
}

template <class R, int DH, bool PROJ, bool POW2>
cudaError_t rows_pick_degree(int deg, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                             const RowsDev<R>& rw, cudaStream_t st) {
#define ENTERP_ROWS(DD)                                                                          \
  case DD:                                                                                       \
    return equi ? rows_launch_one<R, DH, PROJ, DD, true, POW2>(c, s, out, vec_ok, rw, st)        \
                : rows_launch_one<R, DH, PROJ, DD, false, POW2>(c, s, out, vec_ok, rw, st);
  switch (deg) {
    ENTERP_ROWS(1) ENTERP_ROWS(2) ENTERP_ROWS(3) ENTERP_ROWS(4) ENTERP_ROWS(5) ENTERP_ROWS(6) ENTERP_ROWS(7)
  }
#undef ENTERP_ROWS
  return cudaErrorInvalidValue;
}

template <class R, bool POW2>
cudaError_t launch_rows_impl(int dh, bool proj, int deg, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                             bool vec_ok, const RowsDev<R>& rw, cudaStream_t st) {
  if (!proj) {
    switch (dh) {
      case 1: return rows_pick_degree<R, 1, false, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 2: return rows_pick_degree<R, 2, false, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 3: return rows_pick_degree<R, 3, false, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 4: return rows_pick_degree<R, 4, false, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
    }
  } else {
    switch (dh) {
      case 2: return rows_pick_degree<R, 2, true, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 3: return rows_pick_degree<R, 3, true, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 4: return rows_pick_degree<R, 4, true, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
      case 5: return rows_pick_degree<R, 5, true, POW2>(deg, equi, c, s, out, vec_ok, rw, st);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace enterp
