// Launchers of the lean kernels (lean.cuh) for one (scalar type R, batch-or-take): verify the preconditions on
// the host, pick the instantiation, launch. Each returns false (nothing launched) when the call does not qualify.
#pragma once
#include <cstring>

#include "inst_common.cuh"
#include "lean.cuh"

namespace enterp {

constexpr unsigned long long LEAN_INDEX_LIMIT = (1ull << 32) - (1ull << 24);  // headroom for the 32-bit grid stride

template <class R, bool BATCH>
bool lean_src(const SampleSrc<R>& s, bool vec_ok, LeanTake<R>& t) {
  if ((s.t != nullptr) != BATCH || !vec_ok || s.count == 0) return false;
  if (s.first > LEAN_INDEX_LIMIT || s.count > LEAN_INDEX_LIMIT - s.first) return false;
  t.step = s.step;
  t.start = s.start;
  t.first = (uint32_t)s.first;
  t.count = (uint32_t)s.count;
  return true;
}

// Host replica of the device's stepper value (this file is compiled with -ffp-contract=off)
template <class R>
R lean_param(const LeanTake<R>& t, uint32_t i) {
  const R prod = t.step * (R)(t.first + i);
  return prod + t.start;
}

template <class R, int DH, bool PROJ, bool EASE, bool BATCH>
cudaError_t lean_linear_pick_mode(int mode, const LeanLinear<R>& c, const LeanTake<R>& t, const R* ts, R* out,
                                  cudaStream_t st) {
  switch (mode) {
    case 0: return launch_grid_stride(linear_equi_lean_kernel<R, DH, PROJ, 0, EASE, BATCH>, t.count, BLOCK, st, c, t, ts, out);
    case 1: return launch_grid_stride(linear_equi_lean_kernel<R, DH, PROJ, 1, EASE, BATCH>, t.count, BLOCK, st, c, t, ts, out);
    case 2: return launch_grid_stride(linear_equi_lean_kernel<R, DH, PROJ, 2, EASE, BATCH>, t.count, BLOCK, st, c, t, ts, out);
    default: return launch_grid_stride(linear_equi_lean_kernel<R, DH, PROJ, 3, EASE, BATCH>, t.count, BLOCK, st, c, t, ts, out);
  }
}

template <class R, int DH, bool PROJ, bool BATCH>
cudaError_t lean_linear_pick(int mode, const LeanLinear<R>& c, const LeanTake<R>& t, const R* ts, R* out, cudaStream_t st) {
  if (c.easing) return lean_linear_pick_mode<R, DH, PROJ, true, BATCH>(mode, c, t, ts, out, st);
  return lean_linear_pick_mode<R, DH, PROJ, false, BATCH>(mode, c, t, ts, out, st);
}

template <class R, int DH, bool PROJ>
cudaError_t lean_linear_sorted_pick(const CurveDev<R>& c, const LeanLinear<R>& l, const LeanTake<R>& t, R* out,
                                    cudaStream_t st) {
  if (l.easing) return launch_grid_stride(linear_sorted_take_kernel<R, DH, PROJ, true>, t.count, BLOCK, st, c, l, t, out);
  return launch_grid_stride(linear_sorted_take_kernel<R, DH, PROJ, false>, t.count, BLOCK, st, c, l, t, out);
}

// take() over sorted knots: needs the interval-record table (resolve.hpp) and nothing else
template <class R>
bool lean_linear_sorted_take(int dh, bool proj, const CurveDev<R>& c, const LeanTake<R>& t, R* out, cudaStream_t st,
                             cudaError_t* err) {
  if (c.records == nullptr || c.rec_stride != (uint32_t)linear_rec_stride<R>(dh)) return false;
  LeanLinear<R> l{};
  l.easing = c.easing;
  l.easing_n = c.easing_n;
  l.ease_min = c.ease_min;
  l.ease_max = c.ease_max;
  *err = cudaErrorInvalidValue;
  if (!proj) {
    switch (dh) {
      case 1: *err = lean_linear_sorted_pick<R, 1, false>(c, l, t, out, st); break;
      case 2: *err = lean_linear_sorted_pick<R, 2, false>(c, l, t, out, st); break;
      case 3: *err = lean_linear_sorted_pick<R, 3, false>(c, l, t, out, st); break;
      case 4: *err = lean_linear_sorted_pick<R, 4, false>(c, l, t, out, st); break;
    }
  } else {
    switch (dh) {
      case 2: *err = lean_linear_sorted_pick<R, 2, true>(c, l, t, out, st); break;
      case 3: *err = lean_linear_sorted_pick<R, 3, true>(c, l, t, out, st); break;
      case 4: *err = lean_linear_sorted_pick<R, 4, true>(c, l, t, out, st); break;
      case 5: *err = lean_linear_sorted_pick<R, 5, true>(c, l, t, out, st); break;
    }
  }
  return true;
}

template <class R, bool BATCH>
bool lean_linear_impl(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, bool vec_ok,
                      cudaStream_t st, cudaError_t* err) {
  LeanTake<R> t{};
  if (!lean_src<R, BATCH>(s, vec_ok, t)) return false;
  if (c.n_pre != 0 || c.inner_len < 2 || c.inner_len > 0x7fffffffu) return false;
  if (!equi) {
    if constexpr (BATCH) return false;  // random parameters over sorted knots: the generic kernel (staged pivot levels)
    else return lean_linear_sorted_take<R>(dh, proj, c, t, out, st, err);
  }
  const int mode = c.eq_const ? 0 : (c.eq_mode == 1 ? 1 : (c.eq_mode == 2 ? 2 : 3));
  if (!BATCH) {
    if (!(c.eq_step > R(0)) || !(t.step >= R(0))) return false;  // also rejects NaN
    const R x0 = lean_param(t, 0), x1 = lean_param(t, t.count - 1);
    const R offset = c.eq_const ? R(0) : c.eq_offset;
    if (!(x0 >= offset)) return false;
    const R scaled1 = c.eq_const ? x1 * c.eq_step : (x1 - c.eq_offset) / c.eq_step;
    if (!(scaled1 >= R(0) && scaled1 < R(4294967040.0))) return false;
  }
  LeanLinear<R> l{};
  l.elems = c.elems;
  l.len = c.inner_len;
  l.eq_step = c.eq_step;
  l.eq_offset = c.eq_offset;
  l.eq_rstep = c.eq_rstep;
  l.eq_lenm2 = c.eq_lenm2;
  l.easing = c.easing;
  l.easing_n = c.easing_n;
  l.ease_min = c.ease_min;
  l.ease_max = c.ease_max;
  l.invalid = c.invalid;
  *err = cudaErrorInvalidValue;
  if (!proj) {
    switch (dh) {
      case 1: *err = lean_linear_pick<R, 1, false, BATCH>(mode, l, t, s.t, out, st); break;
      case 2: *err = lean_linear_pick<R, 2, false, BATCH>(mode, l, t, s.t, out, st); break;
      case 3: *err = lean_linear_pick<R, 3, false, BATCH>(mode, l, t, s.t, out, st); break;
      case 4: *err = lean_linear_pick<R, 4, false, BATCH>(mode, l, t, s.t, out, st); break;
    }
  } else {
    switch (dh) {
      case 2: *err = lean_linear_pick<R, 2, true, BATCH>(mode, l, t, s.t, out, st); break;
      case 3: *err = lean_linear_pick<R, 3, true, BATCH>(mode, l, t, s.t, out, st); break;
      case 4: *err = lean_linear_pick<R, 4, true, BATCH>(mode, l, t, s.t, out, st); break;
      case 5: *err = lean_linear_pick<R, 5, true, BATCH>(mode, l, t, s.t, out, st); break;
    }
  }
  return true;
}

constexpr int LEAN_BEZIER_MAX_N = 8;

template <class R, int DH, bool PROJ, int N, bool BATCH>
cudaError_t lean_bezier_launch(const R* host_elems, const CurveDev<R>& c, const LeanTake<R>& t, const R* ts, R* out,
                               cudaStream_t st) {
  LeanBezier<R, N * DH> b{};
  b.affine = c.n_pre;
  b.aff_a = c.n_pre ? c.pre_a[0] : R(1);
  b.aff_b = c.n_pre ? c.pre_b[0] : R(0);
  std::memcpy(b.el, host_elems, sizeof(R) * N * DH);
  return launch_grid_stride(bezier_lean_kernel<R, DH, PROJ, N, BATCH>, t.count, BLOCK, st, b, t, ts, out);
}

template <class R, int DH, bool PROJ, bool BATCH>
cudaError_t lean_bezier_pick_n(int n, const R* host_elems, const CurveDev<R>& c, const LeanTake<R>& t, const R* ts,
                               R* out, cudaStream_t st) {
  switch (n) {
    case 2: return lean_bezier_launch<R, DH, PROJ, 2, BATCH>(host_elems, c, t, ts, out, st);
    case 3: return lean_bezier_launch<R, DH, PROJ, 3, BATCH>(host_elems, c, t, ts, out, st);
    case 4: return lean_bezier_launch<R, DH, PROJ, 4, BATCH>(host_elems, c, t, ts, out, st);
    case 5: return lean_bezier_launch<R, DH, PROJ, 5, BATCH>(host_elems, c, t, ts, out, st);
    case 6: return lean_bezier_launch<R, DH, PROJ, 6, BATCH>(host_elems, c, t, ts, out, st);
    case 7: return lean_bezier_launch<R, DH, PROJ, 7, BATCH>(host_elems, c, t, ts, out, st);
    case 8: return lean_bezier_launch<R, DH, PROJ, 8, BATCH>(host_elems, c, t, ts, out, st);
  }
  return cudaErrorInvalidValue;
}

template <class R, bool BATCH>
bool lean_bezier_impl(int dh, bool proj, int n, const R* host_elems, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,
                      bool vec_ok, cudaStream_t st, cudaError_t* err) {
  LeanTake<R> t{};
  if (!lean_src<R, BATCH>(s, vec_ok, t)) return false;
  if (n < 2 || n > LEAN_BEZIER_MAX_N || host_elems == nullptr) return false;
  if (c.n_pre > 1 || (c.n_pre == 1 && c.pre_kind[0] != 1)) return false;
  *err = cudaErrorInvalidValue;
  if (!proj) {
    switch (dh) {
      case 1: *err = lean_bezier_pick_n<R, 1, false, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 2: *err = lean_bezier_pick_n<R, 2, false, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 3: *err = lean_bezier_pick_n<R, 3, false, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 4: *err = lean_bezier_pick_n<R, 4, false, BATCH>(n, host_elems, c, t, s.t, out, st); break;
    }
  } else {
    switch (dh) {
      case 2: *err = lean_bezier_pick_n<R, 2, true, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 3: *err = lean_bezier_pick_n<R, 3, true, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 4: *err = lean_bezier_pick_n<R, 4, true, BATCH>(n, host_elems, c, t, s.t, out, st); break;
      case 5: *err = lean_bezier_pick_n<R, 5, true, BATCH>(n, host_elems, c, t, s.t, out, st); break;
    }
  }
  return true;
}

}  // namespace enterp

#define ENTERP_DEFINE_LEAN(R, BATCH)                                                                                   \
  namespace enterp {                                                                                                   \
  template <>                                                                                                          \
  bool lean_linear_part<R, BATCH>(int dh, bool proj, bool equi, const CurveDev<R>& c, const SampleSrc<R>& s, R* out,   \
                                  bool vec_ok, cudaStream_t st, cudaError_t* err) {                                    \
    return lean_linear_impl<R, BATCH>(dh, proj, equi, c, s, out, vec_ok, st, err);                                     \
  }                                                                                                                    \
  template <>                                                                                                          \
  bool lean_bezier_part<R, BATCH>(int dh, bool proj, int n, const R* host_elems, const CurveDev<R>& c,                 \
                                  const SampleSrc<R>& s, R* out, bool vec_ok, cudaStream_t st, cudaError_t* err) {     \
    return lean_bezier_impl<R, BATCH>(dh, proj, n, host_elems, c, s, out, vec_ok, st, err);                            \
  }                                                                                                                    \
  }
