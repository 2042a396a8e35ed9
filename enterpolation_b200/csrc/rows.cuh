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

// =================================================================================================
// Branch-light equidistant span (Equidistant::strict_upper_bound_clamped, list.rs:476-488) returning
// the ROW index directly: ri = map_from(inner) - lo = inner + ri_off when neither map_from special
// case (inner == 0, inner == inner_len) can occur (host flag rw.map_simple).
template <class R>
__device__ __forceinline__ uint32_t sat_u32(R v);
template <>
__device__ __forceinline__ uint32_t sat_u32<float>(float v) { return __float2uint_rz(v); }   // saturating; NaN is never relied on
template <>
__device__ __forceinline__ uint32_t sat_u32<double>(double v) { return __double2uint_rz(v); }

// The host only selects this kernel when (resolve.cpp / capi.cu):
//   * map_from(inner) has no reachable special case, so the row index is inner + ri_off;
//   * EQUI: eq_step is a power of two for POW2 rows (scaled = num * (1/step) exactly), and within the
//     safe window otherwise (exact division from eq_rstep = RN(1/step), IEEE fallback off-window).
//   * CHECK = false: the caller has proven that no sample of the launch can make the reference panic
//     (take mode: both ends of the monotone stepper range are valid), so the test is skipped.
template <class R, bool EQUI, bool POW2, bool CHECK>
__device__ __forceinline__ uint32_t rows_span(const CurveDev<R>& c, const RowsDev<R>& rw, R x, bool& valid) {
  uint32_t inner;
  if (EQUI) {
    const R num = xsub(x, c.eq_offset);
    R scaled;
    if (POW2) {
      scaled = xmul(num, rw.eq_rstep);
    } else if (in_guard(num)) {
      scaled = exact_div_scalar<R>(num, rw.eq_rstep, -c.eq_step);
    } else {
      scaled = xdiv(num, c.eq_step);
    }
    const bool left = x < c.eq_first;
    // floor(scaled).to_usize().unwrap() panics unless 0 <= scaled < 2^64 (NaN fails both tests)
    valid = !CHECK || left || ((scaled >= R(0)) && (scaled < R(18446744073709551616.0)));
    const uint32_t u = sat_u32<R>(scaled);              // trunc == floor for scaled >= 0
    const uint32_t up = min(u, c.imax - 1u) + 1u;        // min(imax, floor + 1); imax >= 1
    inner = (left || !valid) ? c.imin : up;
  } else {
    valid = true;
    inner = c.imin + count_le(c, x);
  }
  return min(inner + rw.ri_off, rw.n_rows - 1u);  // the clamp is a no-op for every valid sample
}

// Chains of input adaptors run out of line on purpose: the hot kernel must stay small (its adaptor-free loops are
// I-cache sensitive). The list travels BY VALUE: a reference to the kernel's CurveDev parameter would make the
// compiler copy the whole ~470-byte struct into every thread's local memory to have an address for it.
template <class R>
struct PreList {
  int n;
  int kind[MAX_PRE];
  R a[MAX_PRE], b[MAX_PRE];
};
template <class R>
__device__ __noinline__ R apply_pre_list(const PreList<R> p, R x) {
#pragma unroll
  for (int i = 0; i < MAX_PRE; ++i) {
    if (i < p.n) {
      if (p.kind[i] == 0)
        x = x < p.a[i] ? p.a[i] : (x > p.b[i] ? p.b[i] : x);  // Clamp::eval, adaptors.rs:26-32
      else
        x = xadd(xmul(x, p.a[i]), p.b[i]);                     // TransformInput::eval, adaptors.rs:141-152
    }
  }
  return x;
}
template <class R>
__device__ __forceinline__ R apply_pre_outlined(const CurveDev<R>& c, R x) {
  PreList<R> p;
  p.n = c.n_pre;
#pragma unroll
  for (int i = 0; i < MAX_PRE; ++i) {
    p.kind[i] = c.pre_kind[i];
    p.a[i] = c.pre_a[i];
    p.b[i] = c.pre_b[i];
  }
  return apply_pre_list<R>(p, x);
}

// factor operations all of whose lanes belong to de Boor levels r >= 2
__host__ __device__ constexpr uint32_t upper_ops(int P, int W) {
  const FactorSched sch = make_sched(P, W);
  uint32_t m = 0;
  for (int t = 0; t < sch.nops; ++t) {
    bool up = true;
    for (int l = 0; l < W; ++l) up = up && sch.r[t][l] >= 2;
    if (up) m |= 1u << t;
  }
  return m;
}

constexpr int ROWS_S = 2;  // samples per thread per loop trip (block-strided, so stores stay coalesced)

// CHUNK: every CTA walks ONE contiguous range of tiles instead of striding the whole grid, so a thread's
// parameter advances by only TILE stepper steps per trip: its span (and the row cached in registers) changes
// rarely, row reloads can come from global memory / L2 (tables of any size), and the span computation (sorted
// knots: the pivot search; equidistant: the exact division and floor) is skipped while the parameter stays inside
// the cached row's [hold_lo, hold_hi) interval, which the host tabulated with the kernel's own arithmetic.
template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2, bool TAKE, bool CHECK, bool CHUNK, bool RUNS = false>
__device__ __forceinline__ void rows_body(const CurveDev<R>& c, const SampleSrc<R>& s, R* __restrict__ out,
                                          const RowsDev<R>& rw, const R* __restrict__ tab) {
  using P = PK<R>;
  using T = typename P::T;
  constexpr int W = P::W;
  constexpr int NOPS = sched_nops(DEG, W);
  constexpr int DHP = round_up(DH, W);
  constexpr int ROW = row_len<R>(DEG, DH);
  constexpr bool HOLD = CHUNK;
  constexpr int HOLD_OFF = row_hold_off<R>(DEG, DH);
  constexpr int VEC = 16 / (int)sizeof(R);  // R values per 128-bit shared load
  constexpr int EL0 = 3 * NOPS * W;
  constexpr int D = PROJ ? DH - 1 : DH;
  constexpr unsigned long long TILE = (unsigned long long)ROWS_S * BLOCK;

  unsigned long long stride, block0, rest;
  if (CHUNK) {
    const unsigned long long tiles = (s.count + TILE - 1) / TILE;
    const unsigned long long per = (tiles + gridDim.x - 1) / gridDim.x;
    const unsigned long long t0 = (unsigned long long)blockIdx.x * per;
    if (t0 >= tiles) return;
    const unsigned long long t1 = t0 + per < tiles ? t0 + per : tiles;
    stride = TILE;
    block0 = t0 * TILE;
    rest = (t1 * TILE < s.count ? t1 * TILE : s.count) - block0;
  } else if (RUNS) {
    stride = TILE;  // unused: the run walk below has its own decomposition
    block0 = 0;
    rest = 0;
  } else {
    stride = (unsigned long long)gridDim.x * TILE;
    block0 = (unsigned long long)blockIdx.x * TILE;
    if (block0 >= s.count) return;
    rest = s.count - block0;
  }
  // trips in which the whole tile is in range need no bounds checks; at most one ragged trip follows
  const uint32_t n_full = rest >= TILE ? (uint32_t)((rest - TILE) / stride) + 1u : 0u;
  const uint32_t n_trips = (uint32_t)((rest + stride - 1) / stride);
  const unsigned long long i0 = block0 + threadIdx.x;
  unsigned long long gi = s.first + i0;  // global stepper index of this thread's first sample (take mode)
  const R* __restrict__ tp = TAKE ? nullptr : s.t + i0;
  R* __restrict__ op = out + i0 * D;

  const bool upper_pow2 = !POW2 && upper_ops(DEG, W) != 0u && (rw.pow2_ops & upper_ops(DEG, W)) == upper_ops(DEG, W);
  R row[ROW];
  uint32_t cached = 0xFFFFFFFFu;
  const T one2 = P::bcast(rw.one);

  // de Boor for one sample whose span row index is ri (bspline/mod.rs:156-168)
  auto eval_to = [&](const R x, const bool valid, const uint32_t ri, auto&& sink) {
    if (ri != cached) {
      const float4* __restrict__ src = reinterpret_cast<const float4*>(tab + (size_t)ri * ROW);
#pragma unroll
      for (int q = 0; q < ROW / VEC; ++q) unpack128(src[q], &row[q * VEC]);
      cached = ri;
    }
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
        // Uniform knots with a power-of-two step: the upper de Boor levels (r >= 2) divide by powers of two in
        // every row (C1: three of its six factors); those multiply by the exact inverse. ONE launch-uniform test
        // per sample (upper_pow2) selects between two unrolled schedules, not one test per operation.
        if (upper_pow2 && ((upper_ops(DEG, W) >> t) & 1u))
          fac[t] = P::mul(num, y);
        else
          fac[t] = exact_div<R>(num, y, nd);
        // Range guard of the exact division with two compares per sample: the left knots of a span are
        // ascending (kl_0 <= .. <= kl_{p-1}), so x - kl_{p-1} >= 2^-40 bounds every numerator from below
        // and x - kl_0 <= 2^40 from above. Anything else (x on or left of a knot, far right, NaN, inf)
        // takes the IEEE division below.
        constexpr int w_first = sched_find(DEG, W, 1, 0);    // factor (r=1, j=0): numerator x - kl_0
        constexpr int w_last = sched_find(DEG, W, DEG, 0);   // factor (r=p, j=0): numerator x - kl_{p-1}
        if (t == (w_first >> 1)) fast_ok = fast_ok && (P::lane(num, w_first & 1) <= P::guard_hi());
        if (t == (w_last >> 1)) fast_ok = fast_ok && (P::lane(num, w_last & 1) >= P::guard_lo());
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
          for (int q = 0; q + 1 < DH; q += 2) {
            const T t1 = P::mul(P::make(a[q], a[q + 1]), P::bcast(om));
            const T t2 = P::mul(P::make(b[q], b[q + 1]), P::bcast(f));
            const T v = P::fma(t1, one2, t2);
            o[q] = P::lane(v, 0);
            o[q + 1] = P::lane(v, 1);
          }
          if (DH % 2) o[DH - 1] = merge1(a[DH - 1], b[DH - 1], f, om);
        } else {
#pragma unroll
          for (int q = 0; q < DH; ++q) o[q] = merge1(a[q], b[q], f, om);
        }
      });
    });
    // project + store (dst is 16-byte aligned on this path; the host checks)
    R v[D];
    if (PROJ) {
      const R rat = ws[DH - 1];  // Homogeneous::project (weights/homogeneous.rs:134-136)
#pragma unroll
      for (int q = 0; q < D; ++q) v[q] = xdiv(ws[q], rat);
    } else {
#pragma unroll
      for (int q = 0; q < D; ++q) v[q] = ws[q];
    }
    if (CHECK && !valid) {
#pragma unroll
      for (int q = 0; q < D; ++q) v[q] = qnan(R(0));
      atomicAdd(c.invalid, 1ull);
    }
    sink(v);
  };
  auto eval_one = [&](const R x, const bool valid, const uint32_t ri, R* __restrict__ dst) {
    eval_to(x, valid, ri, [&](const R(&v)[D]) { store_elem<R, D>(dst, 0, v, true); });
  };

  // Input adaptors (Clamp / Slice / Bezier-style affine maps, base/adaptors.rs) ride on the CHECK variant
  // only, so the adaptor-free take loop stays as lean as it is.
  auto pre_only = [&](R x) -> R {
    if (CHECK && c.n_pre) {
      if (c.n_pre == 1) {
        const R a = c.pre_a[0], b = c.pre_b[0];
        x = c.pre_kind[0] == 0 ? (x < a ? a : (x > b ? b : x)) : xadd(xmul(x, a), b);
      } else {
        x = apply_pre_outlined(c, x);
      }
    }
    return x;
  };
  auto param = [&](const int slot, const bool active) -> R {
    R x;
    if (TAKE)
      x = xadd(xmul(s.step, from_index(gi + (unsigned long long)slot * BLOCK, R(0))), s.start);
    else
      x = active ? __ldcs(tp + slot * BLOCK) : c.eq_first;
    if (CHECK && c.n_pre) {
      if (c.n_pre == 1) {  // the common single Clamp / affine map inline (uniform branches); chains out of line
        const R a = c.pre_a[0], b = c.pre_b[0];
        x = c.pre_kind[0] == 0 ? (x < a ? a : (x > b ? b : x)) : xadd(xmul(x, a), b);
      } else {
        x = apply_pre_outlined(c, x);
      }
    }
    return x;
  };

  // RUNS (f32 take() beyond index 2^24): the stepper's fl(i) (Equidistant::eval, list.rs:416-418: `step *
  // R::from_usize(i) + offset`) has a 24-bit mantissa, so from 2^24 on consecutive indices share their parameter in
  // runs of 2^(floor(log2 i) - 23) samples (256 in the upper half of a take(2^32)) and the reference evaluates the
  // same t over and over. Here a CTA walks tiles of 2^lt consecutive indices; per sub-tile it evaluates each DISTINCT
  // parameter once (thread m takes the m-th representable index v = vb + m * 2^ls, exactly a float), parks the
  // results in shared memory, and then all threads stream the samples out: sample i reads the value of
  // RN-even(i / 2^ls), which is what from_usize(i) rounds to, computed in integers. Bit-exact by construction
  // (same t, same arithmetic), ~1/2^ls of the FP work, and the kernel runs at the store roof.
  if constexpr (RUNS) {
    static_assert(sizeof(R) == 4 && TAKE && !CHUNK, "f32 takes only");
    constexpr int NV = BLOCK;
    __shared__ __align__(16) R vals[2][(NV + 1) * D];
    // CHECK (input adaptors, possible panics): a value the reference would panic on is NaN for every sample of its
    // run, and the invalid counter advances by the run's samples inside this launch, as the per-sample kernels count
    __shared__ unsigned char bad[CHECK ? 2 : 1][CHECK ? NV + 1 : 1];
    unsigned long long n_bad = 0;
    const unsigned lt = rw.runs_lt;
    const unsigned long long A = s.first, B = s.first + s.count;
    const unsigned long long tile1 = (B - 1) >> lt;
    int buf = 0;
    for (unsigned long long tile = (A >> lt) + blockIdx.x; tile <= tile1; tile += gridDim.x) {
      const unsigned long long base = tile << lt;            // >= 2^24 (host), a multiple of 2^lt
      const int ls = 40 - __clzll((long long)base);          // floor(log2 base) - 23 >= 1: fl() spacing of the binade
      const unsigned long long T = 1ull << lt;
      const unsigned long long SUB = (ls >= (int)lt || ((unsigned long long)NV << ls) > T) ? T : ((unsigned long long)NV << ls);
      for (unsigned long long sb = base; sb < base + T; sb += SUB) {
        if (sb >= B || sb + SUB <= A) continue;              // uniform: nothing of the launch in this sub-tile
        const unsigned long long vb = (sb >> ls) << ls;      // representable index at or below the sub-tile
        const uint32_t nvals = (uint32_t)((sb + SUB - 1 - vb) >> ls) + 2u;  // <= NV + 1
        R* __restrict__ vbuf = vals[buf];
        for (uint32_t m = threadIdx.x; m < nvals; m += BLOCK) {
          const unsigned long long v = vb + ((unsigned long long)m << ls);
          const R x = pre_only(xadd(xmul(s.step, from_index(v, R(0))), s.start));  // from_index(v) is exact
          bool valid;
          const uint32_t ri = rows_span<R, EQUI, POW2, CHECK>(c, rw, x, valid);
          eval_to(x, true, ri, [&](const R(&v4)[D]) {
#pragma unroll
            for (int q = 0; q < D; ++q) vbuf[m * D + q] = (CHECK && !valid) ? qnan(R(0)) : v4[q];
          });
          if (CHECK) bad[CHECK ? buf : 0][CHECK ? m : 0] = valid ? 0 : 1;
        }
        __syncthreads();  // one barrier per sub-tile: the other buffer is free once everybody has passed this one
        const unsigned long long lo_i = sb > A ? sb : A;
        const unsigned long long hi_i = sb + SUB < B ? sb + SUB : B;
        // m(i) = RN-even((i - vb) / 2^ls) = (dd + 2^(ls-1) - 1 + parity) >> ls with parity = bit 0 of the truncated
        // quotient i >> ls (the mantissa from_usize would keep): branch-free, 32-bit unless the run length is huge.
        auto stream_out = [&](auto zero) {
          using U = decltype(zero);
          const U hm1 = (U)(((1ull << ls) >> 1) - 1ull);
          const U vpar = (U)((vb >> ls) & 1ull);
          U dd = (U)(lo_i - vb) + (U)threadIdx.x;
          const U end = (U)(hi_i - vb);
          R* __restrict__ dst = out + (lo_i - A + threadIdx.x) * D;
#pragma unroll 4
          for (; dd < end; dd += (U)BLOCK, dst += (size_t)BLOCK * D) {
            const U par = ((dd >> ls) + vpar) & (U)1;
            const uint32_t m = (uint32_t)((dd + hm1 + par) >> ls);
            if (CHECK) n_bad += bad[CHECK ? buf : 0][CHECK ? m : 0];
            R v4[D];
            if constexpr (D == 4) {
              const float4 t4 = reinterpret_cast<const float4*>(vbuf)[m];
              v4[0] = t4.x; v4[1] = t4.y; v4[2] = t4.z; v4[3] = t4.w;
            } else if constexpr (D == 2) {
              const float2 t2 = reinterpret_cast<const float2*>(vbuf)[m];
              v4[0] = t2.x; v4[1] = t2.y;
            } else {
#pragma unroll
              for (int q = 0; q < D; ++q) v4[q] = vbuf[m * D + q];
            }
            store_elem<R, D>(dst, 0, v4, true);
          }
        };
        if (ls <= 20) stream_out(uint32_t(0)); else stream_out((unsigned long long)0);
        buf ^= 1;
      }
    }
    if (CHECK && n_bad) atomicAdd(c.invalid, n_bad);
    return;
  }

  // Batches: the parameters of the NEXT trip are requested before this trip's arithmetic, so their DRAM latency
  // hides behind ~130 instructions of de Boor instead of stalling the head of every trip.
  R ahead[ROWS_S];
  if (!TAKE && n_full > 0) {
#pragma unroll
    for (int u = 0; u < ROWS_S; ++u) ahead[u] = __ldcs(tp + u * BLOCK);
  }

  uint32_t k = 0;
  for (; k < n_full; ++k, gi += stride, op += stride * D) {
    R x[ROWS_S];
    bool valid[ROWS_S];
    uint32_t ri[ROWS_S];
    if (!TAKE) {
#pragma unroll
      for (int u = 0; u < ROWS_S; ++u) x[u] = ahead[u];
      if (k + 1 < n_full) {
#pragma unroll
        for (int u = 0; u < ROWS_S; ++u) ahead[u] = __ldcs(tp + stride + u * BLOCK);
      }
    }
#pragma unroll
    for (int u = 0; u < ROWS_S; ++u) {
      x[u] = TAKE ? param(u, true) : pre_only(x[u]);
      if (HOLD && cached != 0xFFFFFFFFu && x[u] >= row[HOLD_OFF] && x[u] < row[HOLD_OFF + 1]) {
        ri[u] = cached;  // still inside the cached span: the search would return it again
        valid[u] = true;
      } else {
        ri[u] = rows_span<R, EQUI, POW2, CHECK>(c, rw, x[u], valid[u]);
      }
    }
    if (!TAKE) tp += stride;
#pragma unroll
    for (int u = 0; u < ROWS_S; ++u) eval_one(x[u], valid[u], ri[u], op + (size_t)u * BLOCK * D);
  }
  if (k < n_trips) {  // ragged last tile of the launch
    const unsigned long long here = i0 + (unsigned long long)k * stride;
#pragma unroll 1
    for (int u = 0; u < ROWS_S; ++u) {
      const bool active = here + (unsigned long long)u * BLOCK < s.count;
      const R x = param(u, active);
      bool valid;
      const uint32_t ri = rows_span<R, EQUI, POW2, CHECK>(c, rw, x, valid);
      if (active) eval_one(x, valid, ri, op + (size_t)u * BLOCK * D);
    }
  }
}

// Occupancy hint: small rows (cubic [f32;4] needs 36 row registers) fit 4 CTAs/SM under a 64-register
// cap; bigger rows keep the default so that they do not spill.
template <class R>
__host__ __device__ constexpr int rows_min_blocks(int P, int DH, bool POW2) {
  // exact-division rows keep y and -den of every factor live as well: 3 CTAs/SM (85 registers) instead of spilling
  return row_len<R>(P, DH) * (int)(sizeof(R) / 4) <= 40 ? (POW2 ? 4 : 3) : 1;
}

// VARIANT: 0 = batch (parameters read from memory, per-sample panic checks), 1 = take with per-sample
// checks and input adaptors (CHUNKed, rows from global memory, like 3), 2 = take, check-free (the host has proven that no sample of the launch can
// make the reference panic: t_i = step*fl(i)+start is monotone in i and the set of parameters the
// reference accepts is an interval unbounded below, so it suffices that both ends of the launch's range
// are valid), 3 = take, check-free, CHUNKed, rows read from global memory (any table size; sorted knots skip
// the search while a thread stays in its span), 4 = f32 take beyond index 2^24, check-free: one evaluation per
// distinct stepper value, broadcast through shared memory (RUNS above), 6 = the same with input adaptors and the
// per-sample panic accounting. Separate kernels rather than paths in one:
// the hot loop is I-cache sensitive.
template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool POW2, int VARIANT>
__global__ void __launch_bounds__(BLOCK, rows_min_blocks<R>(DEG, DH, POW2)) bspline_rows_kernel(const CurveDev<R> c,
                                                                                               const SampleSrc<R> s,
                                                                                               R* __restrict__ out,
                                                                                               const RowsDev<R> rw) {
  if constexpr (VARIANT == 6) {
    rows_body<R, DH, PROJ, DEG, EQUI, POW2, true, true, false, true>(c, s, out, rw, rw.table);
  } else if constexpr (VARIANT == 4) {
    rows_body<R, DH, PROJ, DEG, EQUI, POW2, true, false, false, true>(c, s, out, rw, rw.table);
  } else if constexpr (VARIANT == 3) {
    rows_body<R, DH, PROJ, DEG, EQUI, POW2, true, false, true>(c, s, out, rw, rw.table);
  } else if constexpr (VARIANT == 1) {
    rows_body<R, DH, PROJ, DEG, EQUI, POW2, true, true, true>(c, s, out, rw, rw.table);
  } else {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    tma_stage(smem_raw, rw.table, rw.bytes, &mbar);
    const R* __restrict__ tab = reinterpret_cast<const R*>(smem_raw);
    rows_body<R, DH, PROJ, DEG, EQUI, POW2, VARIANT != 0, VARIANT != 2, false>(c, s, out, rw, tab);
  }
}

}  // namespace enterp
