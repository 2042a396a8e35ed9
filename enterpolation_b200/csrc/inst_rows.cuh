// Span-row B-spline kernel instantiations for one (scalar type, division mode, projection, knot base).
#pragma once
#include <cstdlib>

#include "inst_common.cuh"
#include "rows.cuh"

namespace enterp {

// smem: dynamic shared memory of the launch (the staged row table; 0 for the global-rows variant)
template <class K, class R>
cudaError_t rows_launch_kernel(K kernel, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, const RowsDev<R>& rw,
                               size_t smem, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  const int per_sm = cached_occupancy(reinterpret_cast<const void*>(kernel), BLOCK, smem, smem > 0 ? ROWS_MAX_BYTES : 0, &e);
  if (e != cudaSuccess) return e;
  const unsigned long long want = (s.count + ROWS_S * BLOCK - 1) / (ROWS_S * BLOCK);
  const unsigned long long cap = (unsigned long long)per_sm * (unsigned long long)sm_count();
  const int grid = (int)(want < cap ? want : cap);
  kernel<<<grid, BLOCK, smem, st>>>(c, s, out, rw);
  count_launch();
  return cudaGetLastError();
}

// Ablation switch (env ENTERP_ROWS_CHUNK=1): every check-free take() through the chunked global-rows variant.
inline bool rows_force_chunk() {
  static const bool v = [] {
    const char* e = std::getenv("ENTERP_ROWS_CHUNK");
    return e && e[0] == '1';
  }();
  return v;
}

// Ablation switch (env ENTERP_NO_RUNS=1): f32 takes beyond index 2^24 through the per-sample kernels.
inline bool rows_no_runs() {
  static const bool v = [] {
    const char* e = std::getenv("ENTERP_NO_RUNS");
    return e && e[0] == '1';
  }();
  return v;
}

// VARIANT 4 (rows.cuh RUNS): CTAs walk index tiles of 2^lt samples round-robin. The tile shrinks until every CTA
// gets enough of them for an even split (the work per tile varies with the binade: 1 / run length).
template <class K, class R>
cudaError_t rows_launch_runs(K kernel, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, RowsDev<R> rw, cudaStream_t st) {
  cudaError_t e = cudaSuccess;
  const int per_sm = cached_occupancy(reinterpret_cast<const void*>(kernel), BLOCK, 0, 0, &e);
  if (e != cudaSuccess) return e;
  const unsigned long long cap = (unsigned long long)per_sm * (unsigned long long)sm_count();
  const unsigned long long last = s.first + s.count - 1;
  unsigned lt = 16;
  while (lt > 12 && ((last >> lt) - (s.first >> lt) + 1) < 64 * cap) --lt;
  const unsigned long long tiles = (last >> lt) - (s.first >> lt) + 1;
  rw.runs_lt = lt;
  kernel<<<(int)(tiles < cap ? tiles : cap), BLOCK, 0, st>>>(c, s, out, rw);
  count_launch();
  return cudaGetLastError();
}

// Host replica of the panic test of rows_span for one stepper value (same IEEE operations; this file is
// compiled with -ffp-contract=off): Equidistant::strict_upper_bound_clamped panics unless the parameter is
// left of the first searched knot or 0 <= (x - offset) / step < 2^64 (list.rs:476-488).
template <class R>
bool take_param_valid(const CurveDev<R>& c, const SampleSrc<R>& s, unsigned long long i) {
  const R prod = s.step * (R)(s.first + i);
  const R x = prod + s.start;
  if (x < c.eq_first) return true;
  const R num = x - c.eq_offset;
  const R scaled = num / c.eq_step;
  return scaled >= R(0) && scaled < R(18446744073709551616.0);
}

template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2>
cudaError_t rows_launch_one(const CurveDev<R>& c, const SampleSrc<R>& s, R* out, const RowsDev<R>& rw, cudaStream_t st) {
  if (s.count == 0) return cudaSuccess;
  const bool fits = rw.bytes <= ROWS_MAX_BYTES;  // bigger tables serve the global-rows take variant only
  bool need_check = c.n_pre > 0;  // input adaptors ride on the checked variant
  if (s.t == nullptr && EQUI && !need_check)
    need_check = !(take_param_valid(c, s, 0) && take_param_valid(c, s, s.count - 1));
  // Check-free take(): the chunked variant (rows from global memory, span computation skipped while the stepper
  // stays inside the cached row's hold interval) wins wherever the span computation is not nearly free — sorted
  // knots (a pivot search), exact-division rows (C1: 205 vs 182 G samples/s) — and is the only one for tables
  // beyond the shared-memory budget. Power-of-two equidistant rows that fit (C5) measure the same either way
  // (309 vs 307 G samples/s at 2^32, 261 vs 268 at 2^24) and keep the shared-memory variant.
  // f32 take() reaching past index 2^24: from there on the stepper's from_usize(i) repeats in runs, and VARIANT 4
  // evaluates every distinct parameter once (rows.cuh RUNS). The head below 2^24 keeps the per-sample kernels.
  if constexpr (sizeof(R) == 4) {
    constexpr unsigned long long RUN0 = 1ull << 24;
    const unsigned long long end = s.first + s.count;
    if (s.t == nullptr && !rows_no_runs() && end >= RUN0 + 4096 && end < (1ull << 62)) {
      const unsigned long long split = s.first > RUN0 ? s.first : RUN0;
      constexpr int D = PROJ ? DH - 1 : DH;
      if (split > s.first) {
        SampleSrc<R> head = s;
        head.count = split - s.first;
        const cudaError_t e = rows_launch_one<R, DH, PROJ, DEG, EQUI, POW2>(c, head, out, rw, st);
        if (e != cudaSuccess) return e;
      }
      SampleSrc<R> tail = s;
      tail.first = split;
      tail.count = end - split;
      if (need_check)  // input adaptors / possible panics
        return rows_launch_runs(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 6>, c, tail, out + (split - s.first) * D, rw, st);
      return rows_launch_runs(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 4>, c, tail, out + (split - s.first) * D, rw, st);
    }
  }
  constexpr bool HAS_V2 = EQUI && POW2;  // the only rows for which variant 2 is ever chosen: not instantiated elsewhere
  if (s.t == nullptr && !need_check && (!HAS_V2 || !fits || rows_force_chunk()))
    return rows_launch_kernel(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 3>, c, s, out, rw, 0, st);
  if (s.t == nullptr && need_check)  // adaptors / possible panics: same chunked walk with the per-sample checks
    return rows_launch_kernel(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 1>, c, s, out, rw, 0, st);
  if (!fits) return cudaErrorNotSupported;  // big table and a batch: the caller falls back to the generic kernel
  if (s.t != nullptr)
    return rows_launch_kernel(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 0>, c, s, out, rw, rw.bytes, st);
  if constexpr (HAS_V2)
    return rows_launch_kernel(bspline_rows_kernel<R, DH, PROJ, DEG, EQUI, POW2, 2>, c, s, out, rw, rw.bytes, st);
  else
    return cudaErrorNotSupported;  // unreachable: handled by variant 3 above
}

template <class R, int DH, bool PROJ, bool EQUI, bool POW2>
cudaError_t rows_pick_degree(int deg, const CurveDev<R>& c, const SampleSrc<R>& s, R* out, const RowsDev<R>& rw,
                             cudaStream_t st) {
  switch (deg) {
    case 1: return rows_launch_one<R, DH, PROJ, 1, EQUI, POW2>(c, s, out, rw, st);
    case 2: return rows_launch_one<R, DH, PROJ, 2, EQUI, POW2>(c, s, out, rw, st);
    case 3: return rows_launch_one<R, DH, PROJ, 3, EQUI, POW2>(c, s, out, rw, st);
    case 4: return rows_launch_one<R, DH, PROJ, 4, EQUI, POW2>(c, s, out, rw, st);
#if ENTERP_ROWS_MAX_DEGREE >= 5
    case 5: return rows_launch_one<R, DH, PROJ, 5, EQUI, POW2>(c, s, out, rw, st);
#endif
#if ENTERP_ROWS_MAX_DEGREE >= 6
    case 6: return rows_launch_one<R, DH, PROJ, 6, EQUI, POW2>(c, s, out, rw, st);
#endif
#if ENTERP_ROWS_MAX_DEGREE >= 7
    case 7: return rows_launch_one<R, DH, PROJ, 7, EQUI, POW2>(c, s, out, rw, st);
#endif
  }
  return cudaErrorNotSupported;  // no span-row kernel for this degree in this build: the caller takes the generic one
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
