// Linear, span, stepper and many-Bezier instantiations (both scalar types).
#include <cstdint>

#include "inst_common.cuh"

namespace enterp {

template <class R, int DH, bool PROJ>
static cudaError_t linear_pick(bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                               cudaStream_t st) {
  if (equi) return launch_grid_stride(linear_kernel<R, DH, PROJ, true>, s.count, BLOCK, st, c, s, out, vec_ok);
  if (c.slots != nullptr && c.records != nullptr && c.bkt != nullptr)  // big sorted knot vectors: slot table (resolve.hpp)
    return launch_grid_stride(linear_slots_kernel<R, DH, PROJ>, s.count, BLOCK * slot_samples<R>(), st, c, s, out, vec_ok);
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

// Flattened (curve, sample) mapping for few samples per curve (kernels.cuh). Returns false when it does not apply.
template <class R, int D>
static bool many_flat(int n, const R* ctrl, unsigned long long n_curves, uint32_t samples, R step, R* out, cudaStream_t st,
                      cudaError_t* err) {
  const unsigned long long total = n_curves * (unsigned long long)samples;
  if (n < 1 || n > 16 || total >= (1ull << 32) - (1ull << 24)) return false;
  // One warp per curve pays a dependent control-point load per curve and idles lanes when samples % 32 != 0; it wins
  // only where that is amortised: long sample runs, or big curves whose arithmetic dwarfs the flat kernel's per-sample
  // control loads (measured: C3 16x3x256 31.7 vs 28.9 G samples/s, 8x1x1024 348 vs 291; but 4x3x256 235 vs 342,
  // 16x3x16 13 vs 29, 4x3x4 29 vs 234).
  if (samples >= 128 && (n * D > 16 || samples >= 1024)) return false;
  if ((reinterpret_cast<uintptr_t>(ctrl) % 16u) != 0) return false;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(out) % 16u) == 0;
  // multiply-shift division by `samples` (Granlund & Montgomery, unsigned 32-bit)
  uint32_t l = 0;
  while ((1ull << l) < samples) ++l;
  const uint32_t magic = (uint32_t)(((1ull << 32) * ((1ull << l) - samples)) / samples + 1);
  const uint32_t sh1 = l < 1 ? l : 1, sh2 = l > 0 ? l - 1 : 0;
#define ENTERP_BF(NN)                                                                                                 \
  case NN:                                                                                                            \
    *err = launch_grid_stride(bezier_many_flat_kernel<R, D, NN>, total, BLOCK, st, ctrl, (uint32_t)total, samples,    \
                              magic, sh1, sh2, step, out, vec_ok);                                                    \
    return true;
  switch (n) {
    ENTERP_BF(1) ENTERP_BF(2) ENTERP_BF(3) ENTERP_BF(4) ENTERP_BF(5) ENTERP_BF(6) ENTERP_BF(7) ENTERP_BF(8)
    ENTERP_BF(9) ENTERP_BF(10) ENTERP_BF(11) ENTERP_BF(12) ENTERP_BF(13) ENTERP_BF(14) ENTERP_BF(15) ENTERP_BF(16)
  }
#undef ENTERP_BF
  return false;
}

template <class R, int D>
static cudaError_t many_pick_n(int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl, uint32_t samples,
                               R step, R* out, cudaStream_t st) {
  cudaError_t fe = cudaSuccess;
  if (many_flat<R, D>(n, ctrl, n_curves, samples, step, out, st, &fe)) return fe;
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
  cudaError_t launch_stepper_strided<R>(const SampleSrc<R>& s, unsigned ls, R* out, cudaStream_t st) {                 \
    return launch_grid_stride(stepper_strided_kernel<R>, s.count, BLOCK, st, s, ls, out);                              \
  }                                                                                                                    \
  template <>                                                                                                          \
  cudaError_t launch_bezier_many<R>(int dim, int n, const R* ctrl, unsigned long long n_curves, uint32_t n_ctrl,       \
                                    uint32_t samples, R step, R* out, cudaStream_t st) {                               \
    return launch_many_impl<R>(dim, n, ctrl, n_curves, n_ctrl, samples, step, out, st);                                \
  }

ENTERP_SPECIALIZE(float)
ENTERP_SPECIALIZE(double)

cudaError_t launch_runs_expand(int words, const void* vals, void* out, unsigned long long a, unsigned long long count,
                               unsigned long long vb, unsigned ls, cudaStream_t st) {
  const uint32_t* v = static_cast<const uint32_t*>(vals);
  uint32_t* o = static_cast<uint32_t*>(out);
  switch (words) {
    case 1: return launch_grid_stride(runs_expand_kernel<1>, count, BLOCK * 4, st, v, o, a, count, vb, ls);
    case 2: return launch_grid_stride(runs_expand_kernel<2>, count, BLOCK * 4, st, v, o, a, count, vb, ls);
    case 3: return launch_grid_stride(runs_expand_kernel<3>, count, BLOCK * 4, st, v, o, a, count, vb, ls);
    case 4: return launch_grid_stride(runs_expand_kernel<4>, count, BLOCK * 4, st, v, o, a, count, vb, ls);
  }
  return cudaErrorInvalidValue;
}

}  // namespace enterp
