// "Span-row" B-spline kernel: the fast path for curves whose per-span tables fit in shared memory
// (configs C1 and C5 of BASELINE.json, the README/gradient/NURBS-circle shapes, ...).
//
// For every reachable span index the host (resolve.cpp, build_rows) lays out ONE row holding
// everything de Boor needs for that span: the left knots k[i-1], the denominators
// k[i+p-r]-k[i-1] with their correctly rounded reciprocals, and the degree+1 (lifted) elements.
// The rows are TMA-bulk-copied (cp.async.bulk + mbarrier) into shared memory once per CTA; a
// thread keeps the row of its current span in registers and only reloads it when the span changes.
//
// Arithmetic stays bit-identical to the reference (bspline/mod.rs:160-166):
//   * x - k[i-1], 1 - f and the three roundings of merge() are executed as such. For f32 they are
//     issued as packed FADD2/FMUL2/FFMA2 (two IEEE lanes per instruction; halves the issue slots).
//     ptxas would contract mul.f32x2 + add.f32x2 into one FFMA2 even under --fmad=false, so the
//     merge's final addition is written fma(t1, ONE, t2) with ONE = 1.0 passed as an OPAQUE kernel
//     argument: RN(t1*1 + t2) == RN(t1 + t2) and nothing is left to contract.
//   * the division (x - kl)/den is replaced by an exact equivalent:
//       POW2 rows : den is a power of two  ->  num * (1/den), bit-identical to num / den;
//       otherwise : q0 = RN(num*y), r0 = num - den*q0 (exact, FMA), q1 = RN(q0 + r0*y),
//                   r1 = num - den*q1, q = RN(q1 + r1*y) with y = RN(1/den) tabulated by the host.
//                   q1 is faithful, so by Markstein's theorem the last step is the correctly rounded
//                   quotient. The sequence needs num and den away from the under/overflow range:
//                   the host only selects these rows when every reachable denominator is within
//                   [2^-20, 2^20] (f64: [2^-200, 2^200]) and the kernel falls back to the IEEE
//                   division when |num| leaves [2^-40, 2^40] (f64: [2^-300, 2^300]; this includes
//                   num == +-0, inf and NaN).
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "rows_layout.hpp"

namespace enterp {

// compile-time loop: f(std::integral_constant<int, I>{}) for I in [I0, N)
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, N>(f);
  }
}

// 128-bit shared-memory word -> VEC values of R
__device__ __forceinline__ void unpack128(const float4& v, float* dst) {
  dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}
__device__ __forceinline__ void unpack128(const float4& v, double* dst) {
  dst[0] = __hiloint2double(__float_as_int(v.y), __float_as_int(v.x));
  dst[1] = __hiloint2double(__float_as_int(v.w), __float_as_int(v.z));
}

// ---- packed arithmetic ---------------------------------------------------------------------------
template <class R>
struct PK;

template <>
struct PK<float> {
  static constexpr int W = 2;
  using T = unsigned long long;
  static __device__ __forceinline__ T make(float a, float b) {
    T r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
  }
  static __device__ __forceinline__ T bcast(float a) { return make(a, a); }
  static __device__ __forceinline__ float lane(T v, int k) {
    float a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
    return k ? b : a;
  }
  static __device__ __forceinline__ T mul(T a, T b) {
    T r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
  }
  static __device__ __forceinline__ T sub(T a, T b) {
    T r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
  }
  static __device__ __forceinline__ T fma(T a, T b, T c) {
    T r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
  }
  static __device__ __forceinline__ float guard_lo() { return 9.094947017729282e-13f; }  // 2^-40
  static __device__ __forceinline__ float guard_hi() { return 1099511627776.0f; }        // 2^40
};

template <>
struct PK<double> {
  static constexpr int W = 1;
  using T = double;
  static __device__ __forceinline__ T make(double a, double) { return a; }
  static __device__ __forceinline__ T bcast(double a) { return a; }
  static __device__ __forceinline__ double lane(T v, int) { return v; }
  static __device__ __forceinline__ T mul(T a, T b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ T sub(T a, T b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ T fma(T a, T b, T c) { return __fma_rn(a, b, c); }
  static __device__ __forceinline__ double guard_lo() { return 4.909093465297727e-91; }   // 2^-300
  static __device__ __forceinline__ double guard_hi() { return 2.037035976334486e+90; }   // 2^300
};

__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double xfma(double a, double b, double c) { return __fma_rn(a, b, c); }

// Correctly rounded num / den from y = RN(1/den) and nd = -den (see header). Caller guards ranges.
template <class R>
__device__ __forceinline__ typename PK<R>::T exact_div(typename PK<R>::T num, typename PK<R>::T y,
                                                       typename PK<R>::T nd) {
  using P = PK<R>;
  const typename P::T q0 = P::mul(num, y);
  const typename P::T r0 = P::fma(nd, q0, num);
  const typename P::T q1 = P::fma(r0, y, q0);
  const typename P::T r1 = P::fma(nd, q1, num);
  return P::fma(r1, y, q1);
}
template <class R>
__device__ __forceinline__ R exact_div_scalar(R num, R y, R nd) {
  const R q0 = xmul(num, y);
  const R r0 = xfma(nd, q0, num);
  const R q1 = xfma(r0, y, q0);
  const R r1 = xfma(nd, q1, num);
  return xfma(r1, y, q1);
}
template <class R>
__device__ __forceinline__ bool in_guard(R a) {
  const R m = fabs(a);
  return (m >= PK<R>::guard_lo()) && (m <= PK<R>::guard_hi());
}

// Equidistant::strict_upper_bound_clamped with the division replaced by its exact equivalent.
template <class R>
__device__ __forceinline__ bool equi_inner_index_fast(const CurveDev<R>& c, const RowsDev<R>& rw, R x, uint32_t& idx) {
  if (x < c.eq_first) {
    idx = c.imin;
    return true;
  }
  const R num = xsub(x, c.eq_offset);
  R scaled;
  if (rw.eq_mode == 1) {
    scaled = xmul(num, rw.eq_rstep);
  } else if (rw.eq_mode == 2 && in_guard(num)) {
    scaled = exact_div_scalar<R>(num, rw.eq_rstep, -c.eq_step);
  } else {
    scaled = xdiv(num, c.eq_step);
  }
  if (!to_usize_ok(scaled)) return false;
  if (scaled >= R(4294967040.0)) {
    idx = c.imax;
  } else {
    const uint32_t u = (uint32_t)scaled;  // truncation == floor for scaled >= 0
    idx = u >= c.imax ? c.imax : u + 1u;
  }
  return true;
}

// ---- TMA bulk staging ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void stage_rows(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* mbar) {
  if (threadIdx.x == 0) {
    const uint32_t bar = smem_u32(mbar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(bar)
                 : "memory");
  }
  __syncthreads();  // barrier initialised and copy issued before anyone waits
  const uint32_t bar = smem_u32(mbar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar)
      : "memory");
}

// =================================================================================================
template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2>
__global__ void __launch_bounds__(BLOCK) bspline_rows_kernel(const CurveDev<R> c, const SampleSrc<R> s,
                                                             R* __restrict__ out, const bool vec_ok,
                                                             const RowsDev<R> rw) {
  using P = PK<R>;
  using T = typename P::T;
  constexpr int W = P::W;
  constexpr int NOPS = sched_nops(DEG, W);
  constexpr int DHP = round_up(DH, W);
  constexpr int ROW = row_len<R>(DEG, DH);
  constexpr int VEC = 16 / (int)sizeof(R);  // R values per 128-bit shared load
  constexpr int EL0 = 3 * NOPS * W;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  stage_rows(smem_raw, rw.table, rw.bytes, &mbar);
  const R* __restrict__ tab = reinterpret_cast<const R*>(smem_raw);

  R row[ROW];
  uint32_t cached = 0xFFFFFFFFu;
  const T one2 = P::bcast(rw.one);

  for (unsigned long long base = (unsigned long long)blockIdx.x * BLOCK; base < s.count;
       base += (unsigned long long)gridDim.x * BLOCK) {
    const unsigned long long i = base + threadIdx.x;
    const bool active = i < s.count;
    const R x = active ? sample_param(s, i) : c.eq_first;
    uint32_t index;
    bool valid = true;
    if (EQUI) {
      uint32_t inner;
      valid = equi_inner_index_fast(c, rw, x, inner);
      if (!valid) inner = c.imin;
      index = map_from_inner(c, inner);
    } else {
      index = map_from_inner(c, c.imin + coop_count_le(c, x));
    }
    if (!active) continue;
    const uint32_t ri = index - rw.lo;
    if (ri != cached) {
      const float4* __restrict__ src = reinterpret_cast<const float4*>(tab + (size_t)ri * ROW);
#pragma unroll
      for (int k = 0; k < ROW / VEC; ++k) {
        unpack128(src[k], &row[k * VEC]);
      }
      cached = ri;
    }
    // ---- factors -------------------------------------------------------------------------------
    T fac[NOPS], omf[NOPS];
    bool fast_ok = true;
#pragma unroll
    for (int t = 0; t < NOPS; ++t) {
      const T kl = P::make(row[t * W], row[t * W + (W - 1)]);
      const T y = P::make(row[NOPS * W + t * W], row[NOPS * W + t * W + (W - 1)]);
      const T num = P::sub(P::bcast(x), kl);
      if (POW2) {
        fac[t] = P::mul(num, y);
      } else {
        const T nd = P::make(row[2 * NOPS * W + t * W], row[2 * NOPS * W + t * W + (W - 1)]);
        fac[t] = exact_div<R>(num, y, nd);
#pragma unroll
        for (int l = 0; l < W; ++l) fast_ok = fast_ok && in_guard(P::lane(num, l));
      }
    }
    if (!POW2 && !fast_ok) {
      // rare: x (almost) on a knot, far outside the domain, inf or NaN -> the IEEE division itself
#pragma unroll
      for (int t = 0; t < NOPS; ++t) {
        R q[2];
#pragma unroll
        for (int l = 0; l < W; ++l) {
          const R num = xsub(x, row[t * W + l]);
          q[l] = xdiv(num, -row[2 * NOPS * W + t * W + l]);
        }
        fac[t] = P::make(q[0], q[W - 1]);
      }
    }
#pragma unroll
    for (int t = 0; t < NOPS; ++t) omf[t] = P::sub(one2, fac[t]);

    // ---- de Boor triangle ------------------------------------------------------------------------
    R ws[DEG * DHP];  // levels r >= 1 (level 0 is read straight from the cached row)
    static_for<1, DEG + 1>([&](auto rc) {
      constexpr int r = decltype(rc)::value;
      static_for<0, DEG - r + 1>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        constexpr int where = sched_find(DEG, W, r, j);
        static_assert(where >= 0, "factor not scheduled");
        const R f = P::lane(fac[where >> 1], where & 1);
        const R om = P::lane(omf[where >> 1], where & 1);
        const R* a = (r == 1) ? &row[EL0 + j * DHP] : &ws[j * DHP];
        const R* b = (r == 1) ? &row[EL0 + (j + 1) * DHP] : &ws[(j + 1) * DHP];
        R* o = &ws[j * DHP];
        if constexpr (W == 2) {
#pragma unroll
          for (int k = 0; k + 1 < DH; k += 2) {
            const T t1 = P::mul(P::make(a[k], a[k + 1]), P::bcast(om));
            const T t2 = P::mul(P::make(b[k], b[k + 1]), P::bcast(f));
            const T v = P::fma(t1, one2, t2);
            o[k] = P::lane(v, 0);
            o[k + 1] = P::lane(v, 1);
          }
          if (DH % 2) o[DH - 1] = merge1(a[DH - 1], b[DH - 1], f, om);
        } else {
#pragma unroll
          for (int k = 0; k < DH; ++k) o[k] = merge1(a[k], b[k], f, om);
        }
      });
    });
    finish_and_store<R, DH, PROJ>(out, i, ws, valid, vec_ok, c.invalid);
  }
}

}  // namespace enterp
