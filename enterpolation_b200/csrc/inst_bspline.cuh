// BSpline kernel instantiations for one scalar type R (included by inst_bspline_f32.cu / _f64.cu).
#pragma once
#include "inst_common.cuh"

namespace enterp {

template <class R, int DH, bool PROJ, int DEG>
cudaError_t bspline_pick_knots(bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                               cudaStream_t st) {
  if (equi) return launch_grid_stride(bspline_kernel<R, DH, PROJ, DEG, true>, s.count, BLOCK, st, c, s, out, vec_ok);
  if constexpr (knot_slots_ok<R>(DEG)) {
    if (c.kslots != nullptr)  // >= 4096 sorted knots: knot-row slots (resolve.hpp)
      return launch_searching(bspline_slots_kernel<R, DH, PROJ, DEG>, s.count, c.top_bytes, st, c, s, out, vec_ok);
  }
  return launch_searching(bspline_kernel<R, DH, PROJ, DEG, false>, s.count, c.top_bytes, st, c, s, out, vec_ok);
}

template <class R, int DH, bool PROJ>
cudaError_t bspline_pick_degree(int deg, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                                cudaStream_t st) {
  switch (deg) {
    case 1: return bspline_pick_knots<R, DH, PROJ, 1>(equi, c, s, out, vec_ok, st);
    case 2: return bspline_pick_knots<R, DH, PROJ, 2>(equi, c, s, out, vec_ok, st);
    case 3: return bspline_pick_knots<R, DH, PROJ, 3>(equi, c, s, out, vec_ok, st);
    case 4: return bspline_pick_knots<R, DH, PROJ, 4>(equi, c, s, out, vec_ok, st);
    case 5: return bspline_pick_knots<R, DH, PROJ, 5>(equi, c, s, out, vec_ok, st);
    case 6: return bspline_pick_knots<R, DH, PROJ, 6>(equi, c, s, out, vec_ok, st);
    case 7: return bspline_pick_knots<R, DH, PROJ, 7>(equi, c, s, out, vec_ok, st);
    default: return bspline_pick_knots<R, DH, PROJ, 0>(equi, c, s, out, vec_ok, st);
  }
}

template <class R>
cudaError_t launch_bspline_impl(int dh, bool proj, int deg, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s,
                                R* out, bool vec_ok, cudaStream_t st) {
  if (!proj) {
    switch (dh) {
      case 1: return bspline_pick_degree<R, 1, false>(deg, equi, c, s, out, vec_ok, st);
      case 2: return bspline_pick_degree<R, 2, false>(deg, equi, c, s, out, vec_ok, st);
      case 3: return bspline_pick_degree<R, 3, false>(deg, equi, c, s, out, vec_ok, st);
      case 4: return bspline_pick_degree<R, 4, false>(deg, equi, c, s, out, vec_ok, st);
    }
  } else {
    switch (dh) {
      case 2: return bspline_pick_degree<R, 2, true>(deg, equi, c, s, out, vec_ok, st);
      case 3: return bspline_pick_degree<R, 3, true>(deg, equi, c, s, out, vec_ok, st);
      case 4: return bspline_pick_degree<R, 4, true>(deg, equi, c, s, out, vec_ok, st);
      case 5: return bspline_pick_degree<R, 5, true>(deg, equi, c, s, out, vec_ok, st);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace enterp
