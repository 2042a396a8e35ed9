// Linear, span, stepper and many-Bezier instantiations (both scalar types).
#include "inst_common.cuh"

namespace enterp {

template <class R, int DH, bool PROJ>
static cudaError_t linear_pick(bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                               cudaStream_t st) {
  if (equi) return launch_grid_stride(linear_kernel<R, DH, PROJ, true>, s.count, BLOCK, st, c, s, out, vec_ok);
  return launch_searching(linear_kernel<R, DH, PROJ, false>, s.count, c.top_bytes, st, c, s, out, vec_ok);
}

template <class R>
static cudaError_t launch_linear_impl(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                                      bool vec_ok, cudaStream_t st) {
  if (!proj) {
    switch (dh) {
      case 1: return linear_pick<R, 1, false>(equi, c, s, out, vec_ok, st);
      case 2: return linear_pick<R, 2, false>(equi, c, s, out, vec_ok, st);
      case 3: return linear_pick<R, 3, false>(equi, c, s, out, vec_ok, st);
      case 4: return linear_pick<R, 4, false>(equi, c, s, out, vec_ok, st);
    }
  } else {
    switch (dh) {
      case 2: return linear_pick<R, 2, true>(equi, c, s, out, vec_ok, st);
      case 3: return linear_pick<R, 3, true>(equi, c, s, out, vec_ok, st);
      case 4: return linear_pick<R, 4, true>(equi, c, s, out, vec_ok, st);
      case 5: return linear_pick<R, 5, true>(equi, c, s, out, vec_ok, st);
    }
  }
  return cudaErrorInvalidValue;
}

template <class R>
static cudaError_t launch_span_impl(int kind, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, uint32_t* idx,
                                    R* factor, cudaStream_t st) {
  if (kind == 0) {
    if (equi) return launch_grid_stride(span_kernel<R, 0, true>, s.count, BLOCK, st, c, s, idx, factor);
    return launch_grid_stride(span_kernel<R, 0, false>, s.count, BLOCK, st, c, s, idx, factor);
  }
  if (kind == 2) {
    if (equi) return launch_grid_stride(span_kernel<R, 2, true>, s.count, BLOCK, st, c, s, idx, factor);
    return launch_grid_stride(span_kernel<R, 2, false>, s.count, BLOCK, st, c, s, idx, factor);
  }
  return launch_grid_stride(span_kernel<R, 1, true>, s.count, BLOCK, st, c, s, idx, factor);
}

template <class R, int D>
static cudaError_t many_pick_n(int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl, uint32_t samples,
                               R step, R* out, cudaStream_t st) {
  // one warp per curve -> BLOCK/32 curves per block
#define ENTERP_BM(NN) \
  case NN: return launch_grid_stride(bezier_many_kernel<R, D, NN>, n_curves, BLOCK / 32, st, ctrl, n_curves, n_ctrl, samples, step, out);
  switch (n) {
    ENTERP_BM(1) ENTERP_BM(2) ENTERP_BM(3) ENTERP_BM(4) ENTERP_BM(5) ENTERP_BM(6) ENTERP_BM(7) ENTERP_BM(8)
    ENTERP_BM(9) ENTERP_BM(10) ENTERP_BM(11) ENTERP_BM(12) ENTERP_BM(13) ENTERP_BM(14) ENTERP_BM(15) ENTERP_BM(16)
    default: return launch_grid_stride(bezier_many_kernel<R, D, 0>, n_curves, BLOCK / 32, st, ctrl, n_curves, n_ctrl, samples, step, out);
  }
#undef ENTERP_BM
}

template <class R>
static cudaError_t launch_many_impl(int dim, int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl,
                                    uint32_t samples, R step, R* out, cudaStream_t st) {
  switch (dim) {
    case 1: return many_pick_n<R, 1>(n, ctrl, n_curves, n_ctrl, samples, step, out, st);
    case 2: return many_pick_n<R, 2>(n, ctrl, n_curves, n_ctrl, samples, step, out, st);
    case 3: return many_pick_n<R, 3>(n, ctrl, n_curves, n_ctrl, samples, step, out, st);
    case 4: return many_pick_n<R, 4>(n, ctrl, n_curves, n_ctrl, samples, step, out, st);
  }
  return cudaErrorInvalidValue;
}

template <class R>
static cudaError_t launch_deriv_impl(int dim, const CurveDev<R>& c, const SampleSrc<R>& s, int K, int tangent, R* out,
                                     cudaStream_t st) {
  switch (dim) {
    case 1: return launch_grid_stride(bezier_deriv_kernel<R, 1>, s.count, BLOCK, st, c, s, K, tangent, out);
    case 2: return launch_grid_stride(bezier_deriv_kernel<R, 2>, s.count, BLOCK, st, c, s, K, tangent, out);
    case 3: return launch_grid_stride(bezier_deriv_kernel<R, 3>, s.count, BLOCK, st, c, s, K, tangent, out);
    case 4: return launch_grid_stride(bezier_deriv_kernel<R, 4>, s.count, BLOCK, st, c, s, K, tangent, out);
  }
  return cudaErrorInvalidValue;
}

#define ENTERP_SPECIALIZE(R)                                                                                           \
  template <>                                                                                                          \
  cudaError_t launch_linear<R>(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,      \
                               bool vec_ok, cudaStream_t st) {                                                         \
    return launch_linear_impl<R>(dh, proj, equi, c, s, out, vec_ok, st);                                               \
  }                                                                                                                    \
  template <>                                                                                                          \
  cudaError_t launch_span<R>(int kind, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, uint32_t* idx,          \
                             R* factor, cudaStream_t st) {                                                             \
    return launch_span_impl<R>(kind, equi, c, s, idx, factor, st);                                                     \
  }                                                                                                                    \
  template <>                                                                                                          \
  cudaError_t launch_bezier_deriv<R>(int dim, const CurveDev<R>& c, const SampleSrc<R>& s, int K, int tangent, R* out, \
                                     cudaStream_t st) {                                                                \
    return launch_deriv_impl<R>(dim, c, s, K, tangent, out, st);                                                       \
  }                                                                                                                    \
  template <>                                                                                                          \
  cudaError_t launch_stepper<R>(const SampleSrc<R>& s, R* out, cudaStream_t st) {                                      \
    return launch_grid_stride(stepper_kernel<R>, s.count, BLOCK, st, s, out);                                          \
  }                                                                                                                    \
  template <>                                                                                                          \
  cudaError_t launch_bezier_many<R>(int dim, int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl,       \
                                    uint32_t samples, R step, R* out, cudaStream_t st) {                               \
    return launch_many_impl<R>(dim, n, ctrl, n_curves, n_ctrl, samples, step, out, st);                                \
  }

ENTERP_SPECIALIZE(float)
ENTERP_SPECIALIZE(double)

}  // namespace enterp
