// Bezier (single curve) kernel instantiations for one scalar type R.
#pragma once
#include "inst_common.cuh"

namespace enterp {

template <class R, int DH, bool PROJ>
cudaError_t bezier_pick_n(int n, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok, cudaStream_t st) {
#define ENTERP_BZ(NN) \
  case NN: return launch_grid_stride(bezier_kernel<R, DH, PROJ, NN>, s.count, BLOCK, st, c, s, out, vec_ok);
  switch (n) {
    ENTERP_BZ(1) ENTERP_BZ(2) ENTERP_BZ(3) ENTERP_BZ(4) ENTERP_BZ(5) ENTERP_BZ(6) ENTERP_BZ(7) ENTERP_BZ(8)
    ENTERP_BZ(9) ENTERP_BZ(10) ENTERP_BZ(11) ENTERP_BZ(12) ENTERP_BZ(13) ENTERP_BZ(14) ENTERP_BZ(15) ENTERP_BZ(16)
    default: return launch_grid_stride(bezier_kernel<R, DH, PROJ, 0>, s.count, BLOCK, st, c, s, out, vec_ok);
  }
#undef ENTERP_BZ
}

template <class R>
cudaError_t launch_bezier_impl(int dh, bool proj, int n, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                               bool vec_ok, cudaStream_t st) {
  if (!proj) {
    switch (dh) {
      case 1: return bezier_pick_n<R, 1, false>(n, c, s, out, vec_ok, st);
      case 2: return bezier_pick_n<R, 2, false>(n, c, s, out, vec_ok, st);
      case 3: return bezier_pick_n<R, 3, false>(n, c, s, out, vec_ok, st);
      case 4: return bezier_pick_n<R, 4, false>(n, c, s, out, vec_ok, st);
    }
  } else {
    switch (dh) {
      case 2: return bezier_pick_n<R, 2, true>(n, c, s, out, vec_ok, st);
      case 3: return bezier_pick_n<R, 3, true>(n, c, s, out, vec_ok, st);
      case 4: return bezier_pick_n<R, 4, true>(n, c, s, out, vec_ok, st);
      case 5: return bezier_pick_n<R, 5, true>(n, c, s, out, vec_ok, st);
    }
  }
  return cudaErrorInvalidValue;
}

}  // namespace enterp
