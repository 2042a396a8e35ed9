// Lean `Curve::take(n)` kernels (signal.rs:221-228) for the two curve families whose generic kernels are
// bound by instruction issue rather than by HBM: Linear over Equidistant / ConstEquidistant knots
// (linear/mod.rs:122-128, list.rs:532-551, 702-720) and single Bezier curves (bezier/mod.rs:44-56).
//
// ncu on the generic kernels (profiles/README.md, "lean take"): 127 and 118 warp instructions per sample at
// 83 % / 77 % issue utilisation, most of them uniform bookkeeping — adaptor slots, easing dispatch, batch-or-take
// selection, 64-bit index arithmetic with I2F.U64 on the XU pipe, the panic test of `to_usize().unwrap()`,
// and (Bezier) one scalar LDG per control value per sample. The launchers in inst_lean.cuh check on the host,
// with the same IEEE operations the device uses, that a given take() needs none of that:
//   * no input adaptor (Bezier: at most one affine map), 16-byte aligned output, sample indices below
//     2^32 - 2^24 (I2FP.F32.U32 instead of I2F.U64); easing and batch-vs-take are template parameters;
//   * Linear take(): eq_step > 0, stepper step >= 0 (so x_i, x_i - offset and the quotient are monotone in i), the
//     first parameter is not left of the first knot and the last quotient is below 2^32 - 256 — hence no sample of
//     the launch can take the `x < offset` exit, fail `to_usize`, or overflow the 32-bit index. Batches
//     (`Signal::sample`, arbitrary parameters) keep those tests, as selects.
// Everything else falls back to the generic kernels, which stay the reference-complete path.
// Arithmetic is the same sequence of IEEE operations as kernels.cuh (same helpers), so results are bit-identical.
#pragma once
#include "kernels.cuh"

namespace enterp {

__device__ __forceinline__ float from_u32(uint32_t i, float) { return __uint2float_rn(i); }
__device__ __forceinline__ double from_u32(uint32_t i, double) { return __uint2double_rn(i); }
// saturating conversions (negative -> 0, huge -> 2^32 - 1); NaN results are overridden by the callers
__device__ __forceinline__ uint32_t to_u32(float v) { return __float2uint_rz(v); }
__device__ __forceinline__ uint32_t to_u32(double v) { return __double2uint_rz(v); }

// Three-component elements (12 / 24 bytes) cannot be stored with one vector instruction per thread: three narrow
// stores per warp write every 32-byte sector of the warp's 384 / 768 contiguous bytes in two or three partial pieces.
// When the whole warp is converged on 32 consecutive samples, the 96 values are transposed through shared memory
// and leave as full 128-bit stores. `stage`: this warp's 96 values; `i0`: sample index of lane 0.
template <class R>
__device__ __forceinline__ void store_warp3(R* __restrict__ out, uint32_t i0, const R (&v)[3], R* stage) {
  const unsigned lane = threadIdx.x & 31u;
  stage[3 * lane] = v[0];
  stage[3 * lane + 1] = v[1];
  stage[3 * lane + 2] = v[2];
  __syncwarp();
  constexpr unsigned N16 = 96u * (unsigned)sizeof(R) / 16u;  // 24 (f32) / 48 (f64) 128-bit pieces
  float4* dst = reinterpret_cast<float4*>(out + (size_t)i0 * 3);
  const float4* src = reinterpret_cast<const float4*>(stage);
#pragma unroll
  for (unsigned k = 0; k < (N16 + 31u) / 32u; ++k) {
    const unsigned q = lane + 32u * k;
    if (q < N16) __stcs(dst + q, src[q]);
  }
  __syncwarp();
}
// Measured (take(2^26..28), tools/sweep_perf.py): plain [f64;3] gains 20-55 % (Linear 210 -> 251, sorted-knot Linear
// 157 -> 243, cubic Bezier 196 -> 236 G samples/s); [f32;3] loses 8-16 % (its three 32-bit stores already reach 86-90 % of
// the HBM roof) and weighted curves lose 5-10 % (bound by the projection divisions) — so only plain [f64;3] uses it.
template <class R, int D, bool PROJ>
__host__ __device__ constexpr bool lean_transposed_store() { return D == 3 && sizeof(R) == 8 && !PROJ; }
// store of one sample in the lean kernels (output 16-byte aligned, host-checked); `stage3`: BLOCK/32 * 96 values
template <class R, int D, bool PROJ>
__device__ __forceinline__ void lean_store(R* __restrict__ out, uint32_t i, const R (&v)[D], R* stage3) {
  if constexpr (lean_transposed_store<R, D, PROJ>()) {
    if (__activemask() == 0xffffffffu) {
      store_warp3<R>(out, i - (threadIdx.x & 31u), v, stage3 + (threadIdx.x >> 5) * 96);
      return;
    }
  }
  store_elem<R, D>(out, i, v, true);
}
template <class R, int D, bool PROJ>
struct LeanStage {
  static constexpr int N = lean_transposed_store<R, D, PROJ>() ? (BLOCK / 32) * 96 : 4;
};

template <class R>
struct LeanTake {
  R step, start;    // t_i = step * fl(first + i) + start (list.rs:416-418)
  uint32_t first;   // global index of sample 0 of this launch
  uint32_t count;   // samples of this launch; first + count <= 2^32 - 2^24
};

// ---------------------------------------------------------------------------------------------------------------
// Linear over (Const)Equidistant knots. MODE selects how (x - offset) / step is formed, all bit-identical to the
// IEEE quotient: 0 ConstEquidistant (x * fl(N-1), list.rs:702-720), 1 multiply by the exact inverse (power-of-two
// step), 2 reciprocal + FMA correction inside the guard window with an IEEE fallback, 3 IEEE divide.
// ---------------------------------------------------------------------------------------------------------------
template <class R>
struct LeanLinear {
  const R* elems;  // [len][DH]
  uint32_t len;
  R eq_step, eq_offset, eq_rstep, eq_lenm2;
  // easing of the lerp factor (linear/mod.rs:127), EASE kernels only
  int easing, easing_n;
  R ease_min, ease_max;
  unsigned long long* invalid;  // BATCH kernels only: samples where the reference would panic
};

template <class R, int MODE>
__device__ __forceinline__ R lean_scaled(const LeanLinear<R>& c, R x) {
  if constexpr (MODE == 0) {
    return xmul(x, c.eq_step);
  } else {
    const R num = xsub(x, c.eq_offset);
    if constexpr (MODE == 1) return xmul(num, c.eq_rstep);
    if constexpr (MODE == 2) {
      if (in_guard(num)) return exact_div_scalar<R>(num, c.eq_rstep, -c.eq_step);
    }
    return xdiv(num, c.eq_step);
  }
}

template <class R>
__device__ __forceinline__ R lean_easing(const LeanLinear<R>& c, R f) {
  switch (c.easing) {  // uniform; same expressions as apply_easing (kernels.cuh)
    case 1: return xsub(R(1), f);
    case 2: return smoothstart_exact(f, c.easing_n);
    case 3: return xsub(R(1), smoothstart_exact(xsub(R(1), f), c.easing_n));
    case 4: return smoothstep_exact(f);
    case 5: return xmul(xmul(xmul(f, f), f), xadd(xmul(f, xsub(xmul(f, R(6)), R(15))), R(10)));
    default: {
      R t;
      if (f < c.ease_min) t = R(0);
      else if (f > c.ease_max) t = R(1);
      else t = xdiv(xsub(f, c.ease_min), xsub(c.ease_max, c.ease_min));
      return smoothstep_exact(t);
    }
  }
}

// BATCH: parameters come from an array and may be anything, so the kernel keeps the `x < offset` exit and the
// `to_usize` panic test of upper_border as selects (no divergent branches); otherwise the launcher has proven them
// unreachable for this take().
template <class R, int DH, bool PROJ, int MODE, bool EASE, bool BATCH>
__global__ void __launch_bounds__(BLOCK) linear_equi_lean_kernel(const LeanLinear<R> c, const LeanTake<R> s,
                                                                 const R* __restrict__ t, R* __restrict__ out) {
  __shared__ __align__(16) R stage3[LeanStage<R, PROJ ? DH - 1 : DH, PROJ>::N];
  const uint32_t stride = gridDim.x * BLOCK;
  for (uint32_t i = blockIdx.x * BLOCK + threadIdx.x; i < s.count; i += stride) {
    const R x = BATCH ? __ldcs(t + i) : xadd(xmul(s.step, from_u32(s.first + i, R(0))), s.start);
    const R scaled = lean_scaled<R, MODE>(c, x);
    // Equidistant::upper_border (list.rs:532-551)
    const R fl = floor(scaled);
    const bool big = BATCH && fl >= R(4294967040.0);
    const uint32_t lo = big ? 0xFFFFFFFEu : to_u32(fl);
    const uint32_t hi = fl == scaled ? lo : lo + 1u;  // ceil == floor exactly on a knot
    const bool last = big || hi >= c.len;
    uint32_t imin = last ? c.len - 2u : lo;
    uint32_t imax = last ? c.len - 1u : hi;
    R f = xsub(scaled, last ? c.eq_lenm2 : fl);
    bool valid = true;
    if constexpr (BATCH) {
      const bool below = x < (MODE == 0 ? R(0) : c.eq_offset);
      valid = below || to_usize_ok(scaled);
      if (below || !valid) {
        imin = 0;
        imax = 1;
        f = scaled;
      }
    }
    if constexpr (EASE) f = lean_easing(c, f);
    const R omf = xsub(R(1), f);
    const R* __restrict__ a = c.elems + (size_t)imin * DH;
    const R* __restrict__ b = c.elems + (size_t)imax * DH;
    R w[DH];
    if constexpr ((DH * (int)sizeof(R)) % 16 == 0) {
      R av[DH], bv[DH];
      load_vec<DH, (DH * (int)sizeof(R)) % 32 == 0>(a, av);
      load_vec<DH, (DH * (int)sizeof(R)) % 32 == 0>(b, bv);
#pragma unroll
      for (int k = 0; k < DH; ++k) w[k] = merge1(av[k], bv[k], f, omf);
    } else {
#pragma unroll
      for (int k = 0; k < DH; ++k) w[k] = merge1(__ldg(a + k), __ldg(b + k), f, omf);
    }
    constexpr int D = PROJ ? DH - 1 : DH;
    R v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = PROJ ? xdiv(w[k], w[DH - 1]) : w[k];  // Homogeneous::project
    if constexpr (BATCH) {
      if (!valid) {
#pragma unroll
        for (int k = 0; k < D; ++k) v[k] = qnan(R(0));
        atomicAdd(c.invalid, 1ull);
      }
    }
    lean_store<R, D, PROJ>(out, i, v, stage3);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Linear over sorted knots, take(): every CTA walks one contiguous range of samples, so a thread's parameter moves
// by BLOCK stepper steps per trip and usually stays inside the knot interval whose record (both knots, both
// elements; resolve.hpp) it already holds in registers: the pivot search (list.rs:69-86, 169-203) and the record
// load run only when k_j <= x < k_{j+1} fails. That test is exactly the interior case of upper_border (m - 1 == j),
// so the edge rules (left of / at or right of all knots, NaN) are always decided by the search itself.
// ---------------------------------------------------------------------------------------------------------------
// (BLOCK, 4): without a min-blocks hint ptxas settles on 32 registers and spills the tile counter into the loop
template <class R, int DH, bool PROJ, bool EASE>
__global__ void __launch_bounds__(BLOCK, 4) linear_sorted_take_kernel(const CurveDev<R> c, const LeanLinear<R> l,
                                                                   const LeanTake<R> s, R* __restrict__ out) {
  constexpr int STRIDE = linear_rec_stride<R>(DH);
  __shared__ __align__(16) R stage3[LeanStage<R, PROJ ? DH - 1 : DH, PROJ>::N];
  const uint32_t tiles = (s.count + BLOCK - 1) / BLOCK;
  const uint32_t per = (tiles + gridDim.x - 1) / gridDim.x;
  const uint32_t t0 = blockIdx.x * per;
  const uint32_t t1 = min(tiles, t0 + per);
  R rec[STRIDE];
  R div = R(0);
  bool have = false;
  for (uint32_t tile = t0; tile < t1; ++tile) {
    const uint32_t i = tile * BLOCK + threadIdx.x;
    if (i >= s.count) break;  // ragged last tile of the launch
    const R x = xadd(xmul(s.step, from_u32(s.first + i, R(0))), s.start);
    bool edge = false;
    if (!(have && x >= rec[0] && x < rec[1])) {
      const uint32_t len = c.inner_len;
      const uint32_t m = count_le(c, x);
      edge = (m == len) || (m == 0);
      const uint32_t imax = m == len ? len - 1 : (m == 0 ? 1u : m);
      load_vec<STRIDE, (STRIDE * (int)sizeof(R)) % 32 == 0>(c.records + (size_t)(imax - 1u) * STRIDE, rec);
      div = xsub(rec[1], rec[0]);
      have = true;
    }
    // linear_factor (checked, list.rs:232-248) on the edges, linear_factor_unchecked (212-224) inside
    R f = (edge && div == R(0)) ? div : xdiv(xsub(x, rec[0]), div);
    if constexpr (EASE) f = lean_easing(l, f);
    const R omf = xsub(R(1), f);
    R w[DH];
#pragma unroll
    for (int k = 0; k < DH; ++k) w[k] = merge1(rec[2 + k], rec[2 + DH + k], f, omf);
    constexpr int D = PROJ ? DH - 1 : DH;
    R v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = PROJ ? xdiv(w[k], w[DH - 1]) : w[k];
    lean_store<R, D, PROJ>(out, i, v, stage3);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Bezier: the control values travel in the kernel parameter block (constant bank), so the first de Casteljau
// level reads them as instruction operands and no load is issued per sample.
// ---------------------------------------------------------------------------------------------------------------
template <class R, int NV>
struct LeanBezier {
  int affine;   // TransformInput of Bezier .domain(a, b) (adaptors.rs:141-152): x * aff_a + aff_b, two roundings
  R aff_a, aff_b;
  R el[NV];     // [N][DH], homogeneous-lifted when weighted
};

template <class R, int DH, bool PROJ, int N, bool BATCH>
__global__ void __launch_bounds__(BLOCK) bezier_lean_kernel(const LeanBezier<R, N * DH> c, const LeanTake<R> s,
                                                            const R* __restrict__ t, R* __restrict__ out) {
  __shared__ __align__(16) R stage3[LeanStage<R, PROJ ? DH - 1 : DH, PROJ>::N];
  const uint32_t stride = gridDim.x * BLOCK;
  for (uint32_t i = blockIdx.x * BLOCK + threadIdx.x; i < s.count; i += stride) {
    R x = BATCH ? __ldcs(t + i) : xadd(xmul(s.step, from_u32(s.first + i, R(0))), s.start);
    if (c.affine) x = xadd(xmul(x, c.aff_a), c.aff_b);
    const R omx = xsub(R(1), x);
    R ws[N * DH];
#pragma unroll
    for (int m = 0; m < N * DH; ++m) ws[m] = c.el[m];
#pragma unroll
    for (int k = 1; k <= N - 1; ++k)
#pragma unroll
      for (int j = 0; j < N - k; ++j)
#pragma unroll
        for (int q = 0; q < DH; ++q) ws[j * DH + q] = merge1(ws[j * DH + q], ws[(j + 1) * DH + q], x, omx);
    constexpr int D = PROJ ? DH - 1 : DH;
    R v[D];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = PROJ ? xdiv(ws[k], ws[DH - 1]) : ws[k];
    lean_store<R, D, PROJ>(out, i, v, stage3);
  }
}

}  // namespace enterp
