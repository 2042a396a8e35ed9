// Hand-written sm_100a kernels of the curve-evaluation hot path.
//
// Arithmetic contract: every floating-point operation the reference performs is performed here
// with the same operands, in the same order, with IEEE round-to-nearest and NO contraction
// (intrinsics __f*_rn / __d*_rn are never fused by nvcc/ptxas). Table lookups replace the
// reference's recomputation of knots (resolve.cpp builds the tables with the same operations),
// and `1 - factor` is hoisted once per merge because all components share it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace enterp {

constexpr int MAX_LEVELS = 16;    // pivot-tree depth: fan-out 8 (f32) / 4 (f64), 4^16 > 2^31 searched knots
constexpr int MAX_DYN_DEGREE = 63;  // runtime-degree de Boor keeps its workspace in local memory
constexpr int MAX_DYN_BEZIER = 256;
constexpr int MAX_PRE = 4;   // nested input adaptors carried to the device
constexpr int BLOCK = 256;
constexpr int MAX_BLOCK = 1024;        // kernels that stage big pivot tables run one fat CTA per SM
constexpr unsigned SEARCH_SMEM_MAX = 160u * 1024u;  // shared-memory budget for staged pivot levels

template <class R>
struct CurveDev {
  const R* vk;     // virtual knot vector (BorderBuffer / BorderDeletion already applied)
  const R* ik;     // innermost knot chain (what the search runs on)
  const R* elems;  // [n_elem][DH], homogeneous-lifted when weighted
  const R* knot_rows;   // B-spline, degree 1..7: padded per-span knot rows (resolve.hpp), or nullptr
  uint32_t krow_lo;     // span index of knot row 0
  const R* records;     // Linear over sorted knots: interval records (resolve.hpp), or nullptr
  uint32_t rec_stride;  // in units of R
  const R* slots;       // Linear over sorted knots: the records again, indexed by an equal-width grid (resolve.hpp), or nullptr
  uint32_t n_slots;     // grid cells (n_slots + 1 entries)
  R slot_inv;           // cell(x) = clamp(trunc((x - bkt_first) * slot_inv), 0, n_slots - 1)
  R k_last;             // last inner knot
  const R* kslots;      // B-splines over sorted knots: the knot rows again, indexed by an equal-width grid (resolve.hpp), or nullptr
  uint32_t n_kslots;    // grid cells (n_kslots + 1 entries); a slot's last value holds its span index as raw bits
  R kslot_inv;          // cell(x) = clamp(trunc((x - bkt_first) * kslot_inv), 0, n_kslots - 1)
  const R* lvl[MAX_LEVELS];  // 32-ary pivot pyramid, top level first, NaN padded (one contiguous buffer)
  uint32_t lvl_off[MAX_LEVELS];  // offset of each level inside that buffer, in elements
  uint32_t lvl_last[MAX_LEVELS]; // index of the last node of each level
  int n_lvl;
  int n_top;           // levels [0, n_top) are staged into shared memory by the kernel (0: none)
  uint32_t top_bytes;  // their total size (multiple of 32)
  // Bucket index of the searched knot range (resolve.hpp), or nullptr: n_bkt buckets, one 16-byte word per 32 of them
  const uint32_t* bkt;
  uint32_t n_bkt;
  R bkt_first, bkt_inv;
  int ballot;  // ablation (ENTERP_BALLOT_SEARCH=1): the warp-cooperative ballot/shuffle 32-ary search instead of index/tree
  uint32_t n_elem, inner_len, virt_len, degree;
  uint32_t imin, imax, map0, mapL;
  int32_t map_add;
  R eq_step, eq_offset, eq_first, eq_lenm2;
  int eq_const;  // ConstEquidistant (list.rs:566-721): scaled = x * eq_step with eq_step = fl(N-1), offset 0
  int eq_mode;   // exact replacement of the divide by eq_step: 0 none, 1 multiply by eq_rstep, 2 exact division
  R eq_rstep;    // RN(1 / eq_step)
  // input adaptors applied to the raw parameter, outermost first (base/adaptors.rs): Clamp(lo, hi) or the
  // affine map t * a + b of TransformInput / Slice (two roundings, never fused)
  int n_pre;
  int pre_kind[MAX_PRE];  // 0 clamp, 1 affine
  R pre_a[MAX_PRE], pre_b[MAX_PRE];
  // Linear easing of the lerp factor (easing/mod.rs, plateau.rs)
  int easing, easing_n;
  R ease_min, ease_max;
  unsigned long long* invalid;  // device counter: samples where the reference would panic
};

template <class R>
struct SampleSrc {
  const R* t;                // batch mode: parameters; take mode: nullptr
  R step, start;             // take mode: t_i = step * fl(i) + start (list.rs:416-418)
  unsigned long long first;  // take mode: global index of this launch's sample 0
  unsigned long long count;  // samples in this launch
};

// Span-row table of the fast B-spline kernel (rows.cuh).
template <class R>
struct RowsDev {
  const R* table;     // global copy of the row table (staged into shared memory by TMA)
  uint32_t n_rows;    // hi - lo + 1
  uint32_t lo;        // span index of row 0
  uint32_t bytes;     // n_rows * row_len * sizeof(R), multiple of 16
  R one;              // 1.0, opaque to the compiler (see header comment)
  R eq_rstep;         // RN(1 / eq_step) (exact when eq_step is a power of two)
  uint32_t ri_off;    // row index = inner index + ri_off (mod 2^32): map_add - lo
  uint32_t pow2_ops;  // exact-division rows: bit t set = factor operation t divides by powers of two in every row
  uint32_t runs_lt;   // VARIANT 4 (stepper runs): log2 of the index tile a CTA walks at a time (launch-specific)
};

// ---- exact arithmetic -------------------------------------------------------------------------
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ float from_index(unsigned long long i, float) { return __ull2float_rn(i); }
__device__ __forceinline__ double from_index(unsigned long long i, double) { return __ull2double_rn(i); }
__device__ __forceinline__ float qnan(float) { return __int_as_float(0x7fc00000); }
__device__ __forceinline__ double qnan(double) { return __longlong_as_double(0x7ff8000000000000ll); }

// Merge::merge (topology-traits 0.1.2 blanket impl == utils.rs:6-12): a*(1-f) + b*f.
template <class R>
__device__ __forceinline__ R merge1(R a, R b, R f, R omf) {
  return xadd(xmul(a, omf), xmul(b, f));
}

// ---- input adaptors and easing ---------------------------------------------------------------------
// Clamp::eval: num_traits::clamp (x < min -> min, x > max -> max, NaN passes) (adaptors.rs:26-32);
// TransformInput::eval: input * multiplication + addition (adaptors.rs:141-152).
template <class R>
__device__ __forceinline__ R apply_pre(const CurveDev<R>& c, R x) {
  if (c.n_pre == 0) return x;  // uniform: one branch instead of MAX_PRE predicated slots
#pragma unroll
  for (int i = 0; i < MAX_PRE; ++i) {
    if (i < c.n_pre) {
      if (c.pre_kind[i] == 0)
        x = x < c.pre_a[i] ? c.pre_a[i] : (x > c.pre_b[i] ? c.pre_b[i] : x);
      else
        x = xadd(xmul(x, c.pre_a[i]), c.pre_b[i]);
    }
  }
  return x;
}

// easing/mod.rs:86-132: flip = 1 - x; smoothstart<N> = x^N by repeated multiplication; smoothend<N> =
// flip(smoothstart<N>(flip(x))); smoothstep = x*x*(3 - 2*x); smootherstep = x*x*x*(x*(x*6 - 15) + 10);
// Plateau (plateau.rs:27-48) = smoothstep(over_clamp(x, min, max)). Rust evaluates left to right.
template <class R>
__device__ __forceinline__ R smoothstep_exact(R x) {
  return xmul(xmul(x, x), xsub(R(3), xmul(R(2), x)));
}
template <class R>
__device__ __forceinline__ R smoothstart_exact(R x, int n) {
  R m = x;
  for (int i = 1; i < n; ++i) m = xmul(m, x);
  return m;
}
template <class R>
__device__ __forceinline__ R apply_easing(const CurveDev<R>& c, R f) {
  switch (c.easing) {
    case 1: return xsub(R(1), f);
    case 2: return smoothstart_exact(f, c.easing_n);
    case 3: return xsub(R(1), smoothstart_exact(xsub(R(1), f), c.easing_n));
    case 4: return smoothstep_exact(f);
    case 5: return xmul(xmul(xmul(f, f), f), xadd(xmul(f, xsub(xmul(f, R(6)), R(15))), R(10)));
    case 6: {
      R t;
      if (f < c.ease_min) t = R(0);
      else if (f > c.ease_max) t = R(1);
      else t = xdiv(xsub(f, c.ease_min), xsub(c.ease_max, c.ease_min));
      return smoothstep_exact(t);
    }
    default: return f;
  }
}

// ---- sample parameter --------------------------------------------------------------------------
template <class R>
__device__ __forceinline__ R sample_param(const SampleSrc<R>& s, unsigned long long i) {
  if (s.t) return __ldcs(s.t + i);
  return xadd(xmul(s.step, from_index(s.first + i, R(0))), s.start);
}

// ---- exact division by a build-time constant (see rows.cuh header for the argument) ---------------
__device__ __forceinline__ float xfma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double xfma(double a, double b, double c) { return __fma_rn(a, b, c); }
template <class R>
__device__ __forceinline__ R exact_div_scalar(R num, R y, R nd) {
  const R q0 = xmul(num, y);
  const R r0 = xfma(nd, q0, num);
  const R q1 = xfma(r0, y, q0);
  const R r1 = xfma(nd, q1, num);
  return xfma(r1, y, q1);
}
__device__ __forceinline__ bool in_guard(float a) {
  const float m = fabsf(a);
  return (m >= 9.094947017729282e-13f) && (m <= 1099511627776.0f);  // [2^-40, 2^40]
}
__device__ __forceinline__ bool in_guard(double a) {
  const double m = fabs(a);
  return (m >= 4.909093465297727e-91) && (m <= 2.037035976334486e+90);  // [2^-300, 2^300]
}
// (x - offset) / step of an Equidistant chain, bit-identical to the IEEE division
template <class R>
__device__ __forceinline__ R equi_scaled(const CurveDev<R>& c, R x) {
  if (c.eq_const) return xmul(x, c.eq_step);
  const R num = xsub(x, c.eq_offset);
  if (c.eq_mode == 1) return xmul(num, c.eq_rstep);
  if (c.eq_mode == 2 && in_guard(num)) return exact_div_scalar<R>(num, c.eq_rstep, -c.eq_step);
  return xdiv(num, c.eq_step);
}

// ---- span search -------------------------------------------------------------------------------
// `floor(scaled).to_usize().unwrap()` (list.rs:456, 486): valid iff 0 <= scaled < 2^64.
template <class R>
__device__ __forceinline__ bool to_usize_ok(R v) {
  return (v >= R(0)) && (v < R(18446744073709551616.0));
}

// Equidistant::strict_upper_bound_clamped (list.rs:476-488) on the inner chain.
template <class R>
__device__ __forceinline__ bool equi_inner_index(const CurveDev<R>& c, R x, uint32_t& idx) {
  if (x < c.eq_first) {
    idx = c.imin;
    return true;
  }
  const R scaled = equi_scaled(c, x);
  if (!to_usize_ok(scaled)) return false;
  const R fl = floor(scaled);
  if (fl >= R(4294967040.0)) {
    idx = c.imax;
  } else {
    const uint32_t u = (uint32_t)fl;
    idx = u >= c.imax ? c.imax : u + 1u;
  }
  return true;
}

// ---- TMA bulk staging: one elected thread issues cp.async.bulk global->shared, everyone waits on the
// mbarrier (SASS: UBLKCP + SYNCS). `bytes` must be a multiple of 16 and < 2^20.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_stage(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* mbar) {
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

// Number of knots of ik[imin, imax) that are <= x: the reference's branchy binary search
// (list.rs:69-86) returns imin + this count on sorted knots, including ties (x == knot goes up),
// repeated knots and NaN (-> 0).
//
// Layout: a pivot tree whose nodes are single 32-byte sectors (FAN = 8 f32 / 4 f64 keys, NaN padded;
// NaN <= x is false, so padding never counts); level l+1 holds the last key of every node of level l.
// A thread descends on its own, fetching ONE sector per level with a 256-bit load (LDG.E.256) and
// counting the keys <= x in registers. Measured history on B200, C2 (2^20 knots, 2^28 samples):
//   warp-cooperative 32-ary ballot search      9.5 G samples/s (a warp instruction per level PER SAMPLE)
//   per-thread 6-probe bisection, 32-ary       21.5 G samples/s (25.7 L1 sector lookups, 8.0 L2 sectors per sample)
//   + upper levels TMA-staged in shared memory 20.3 G samples/s (upper levels were L1 hits anyway)
//   sector tree (this code)                    see profiles/README.md
template <class R>
struct Node;
template <>
struct Node<float> {
  static constexpr int FAN = 8;
  template <bool SMEM>
  static __device__ __forceinline__ uint32_t count_le(const float* __restrict__ p, float x) {
    float v[8];
    if (SMEM) {
      const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
      asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
          : "l"(p));
    }
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) c += (v[i] <= x) ? 1u : 0u;
    return c;
  }
};
template <>
struct Node<double> {
  static constexpr int FAN = 4;
  template <bool SMEM>
  static __device__ __forceinline__ uint32_t count_le(const double* __restrict__ p, double x) {
    double v[4];
    if (SMEM) {
      const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
    }
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) c += (v[i] <= x) ? 1u : 0u;
    return c;
  }
};

// `top`: shared-memory copy of levels [0, n_top) (nullptr when nothing is staged).
__device__ __forceinline__ uint32_t bucket_sat_u32(float v) { return __float2uint_rz(v); }   // saturating above 2^32
__device__ __forceinline__ uint32_t bucket_sat_u32(double v) { return __double2uint_rz(v); }
constexpr uint32_t BUCKET_SCAN = 3;  // knots compared inside one bucket (2-bit counts are exact up to 3; resolve.cpp)

// Pivot-tree descent (below). `top`: shared-memory copy of levels [0, n_top) (nullptr when nothing is staged).
template <class R>
__device__ __forceinline__ uint32_t count_le_tree(const CurveDev<R>& c, R x, const R* __restrict__ top);

// Big sorted knot vectors: the bucket index turns the descent into two dependent loads. bucket(x) is monotone, so
// every knot of a lower bucket is < x and every knot of a higher bucket is > x; only the knots sharing x's bucket
// (usually none or one) are compared. NaN lands in bucket 0 and compares false everywhere (count 0, as the
// reference's search returns its lower bound for NaN).
// One 16-byte word per block of 32 buckets: { knots in lower blocks, 2-bit counts of buckets 0..15, of 16..31, crowded }.
// Returns false when there is no index or x falls into a crowded block (the caller descends the tree); otherwise
// lo = knots in lower buckets, cnt = knots in x's bucket (0..3).
template <class R>
__device__ __forceinline__ bool bucket_lookup(const CurveDev<R>& c, R x, uint32_t& lo, uint32_t& cnt) {
  if (c.bkt == nullptr) return false;
  const R tq = xmul(xsub(x, c.bkt_first), c.bkt_inv);
  // NaN must land in bucket 0 explicitly: the f64 -> u32 conversion does not map NaN to 0 (measured on sm_100a)
  const uint32_t b = min(tq >= R(0) ? bucket_sat_u32(tq) : 0u, c.n_bkt - 1u);
  const uint4 w = __ldg(reinterpret_cast<const uint4*>(c.bkt) + (b >> 5));
  const uint32_t p = b & 31u;
  // fields below p: all of w.y when p >= 16, and the low 2*(p & 15) bits of the word that holds p
  const uint32_t own = p < 16u ? w.y : w.z;
  const uint32_t part = own & ((1u << (2u * (p & 15u))) - 1u);
  const uint32_t full = p < 16u ? 0u : w.y;
  lo = w.x + __popc(part & 0x55555555u) + 2u * __popc(part & 0xAAAAAAAAu) + __popc(full & 0x55555555u) +
       2u * __popc(full & 0xAAAAAAAAu);
  cnt = (own >> (2u * (p & 15u))) & 3u;
  return w.w == 0u;
}

// Ablation (ENTERP_BALLOT_SEARCH=1): the warp-cooperative 32-ary search BASELINE.json's north star names. The warp
// serves its 32 samples one after the other: sample j's parameter is broadcast by shuffle, every lane probes one of 32
// equidistant pivots of the current range, a ballot narrows the range 32-fold (4 rounds for 2^20 knots). It costs a
// gather instruction per round PER SAMPLE where the per-thread searches cost one per round per WARP, which is why it
// is not the product path (C2: profiles/r02_ncu_c2_ballot_search.txt). Must be called by all 32 lanes.
template <class R>
__device__ __noinline__ uint32_t count_le_ballot(const R* __restrict__ k, uint32_t n, R x) {
  const unsigned lane = threadIdx.x & 31u;
  uint32_t mine = 0;
  for (int j = 0; j < 32; ++j) {
    const R xj = __shfl_sync(0xFFFFFFFFu, x, j);
    uint32_t lo = 0, len = n, result = 0;
    for (;;) {
      if (len <= 32u) {
        const bool le = lane < len && __ldg(k + lo + lane) <= xj;
        result = lo + __popc(__ballot_sync(0xFFFFFFFFu, le));
        break;
      }
      const uint32_t step = (len + 31u) / 32u;
      const uint32_t end = min(len, (lane + 1u) * step);          // this lane's block is [lane * step, end)
      const bool le = lane * step < len && __ldg(k + lo + end - 1u) <= xj;
      const uint32_t t = __popc(__ballot_sync(0xFFFFFFFFu, le));  // blocks whose last knot is <= x: a prefix (sorted knots)
      const uint32_t skip = min(len, t * step);
      lo += skip;
      len = min(step, len - skip);
    }
    if ((int)lane == j) mine = result;
  }
  return mine;
}

template <class R>
__device__ __forceinline__ uint32_t count_le(const CurveDev<R>& c, R x, const R* __restrict__ top = nullptr) {
  if (c.ballot) return count_le_ballot<R>(c.ik + c.imin, c.imax - c.imin, x);
  uint32_t lo, hi;
  if (!bucket_lookup(c, x, lo, hi)) return count_le_tree(c, x, top);  // one call site: the unrolled descent is big
  hi += lo;
  const R* __restrict__ k = c.ik + c.imin + lo;
  uint32_t cnt = lo;
#pragma unroll
  for (uint32_t q = 0; q < BUCKET_SCAN; ++q)
    if (lo + q < hi) cnt += (__ldg(k + q) <= x) ? 1u : 0u;
  return cnt;
}

template <class R>
__device__ __forceinline__ uint32_t count_le_tree(const CurveDev<R>& c, R x, const R* __restrict__ top) {
  constexpr uint32_t FAN = Node<R>::FAN;
  uint32_t cnt = 0;
#pragma unroll
  for (int l = 0; l < MAX_LEVELS; ++l) {
    if (l < c.n_lvl) {
      // every key above was <= x: all nodes below are <= x too and the last one holds the remainder
      cnt = min(cnt, c.lvl_last[l]);
      if (top != nullptr && l < c.n_top)
        cnt = cnt * FAN + Node<R>::template count_le<true>(top + c.lvl_off[l] + (size_t)cnt * FAN, x);
      else
        cnt = cnt * FAN + Node<R>::template count_le<false>(c.lvl[l] + (size_t)cnt * FAN, x);
    }
  }
  return cnt;
}

// Stages the top pyramid levels into dynamic shared memory when the host asked for it; returns the
// shared-memory base or nullptr. Must be called by every thread of the block.
template <class R>
__device__ __forceinline__ const R* stage_top_levels(const CurveDev<R>& c, unsigned char* smem_raw,
                                                     unsigned long long* mbar) {
  if (c.top_bytes == 0) return nullptr;
  tma_stage(smem_raw, c.lvl[0], c.top_bytes, mbar);
  return reinterpret_cast<const R*>(smem_raw);
}

// inner index -> virtual index: BorderBuffer::map_from (adaptors.rs:31-39) / BorderDeletion (-1).
template <class R>
__device__ __forceinline__ uint32_t map_from_inner(const CurveDev<R>& c, uint32_t inner_idx) {
  if (inner_idx == c.inner_len) return c.mapL;
  if (inner_idx == 0) return c.map0;
  return (uint32_t)((int32_t)inner_idx + c.map_add);
}

// BSpline span: knots.strict_upper_bound_clamped(x, degree, len - degree) (bspline/mod.rs:148-154).
template <class R, bool EQUI>
__device__ __forceinline__ bool bspline_span(const CurveDev<R>& c, R x, uint32_t& index, const R* __restrict__ top = nullptr) {
  uint32_t inner;
  bool ok = true;
  if (EQUI) {
    ok = equi_inner_index(c, x, inner);
    if (!ok) inner = c.imin;
  } else {
    inner = c.imin + count_le(c, x, top);
  }
  index = map_from_inner(c, inner);
  return ok;
}

// Linear span: knots.upper_border(x) -> (min, max, factor) (list.rs:169-248 / 532-551).
template <class R, bool EQUI>
__device__ __forceinline__ bool linear_border(const CurveDev<R>& c, R x, uint32_t& imin, uint32_t& imax, R& factor,
                                              const R* __restrict__ top = nullptr) {
  const uint32_t len = c.inner_len;
  if (EQUI) {
    // Equidistant::upper_border (list.rs:532-551) / ConstEquidistant::upper_border (702-720)
    const R scaled = equi_scaled(c, x);
    if (x < c.eq_offset) {
      imin = 0;
      imax = 1;
      factor = scaled;
      return true;
    }
    // floor/ceil(scaled).to_usize().unwrap(): both succeed iff 0 <= scaled < 2^64 (floats that large are
    // integers, so ceil adds nothing)
    if (!to_usize_ok(scaled)) {
      imin = 0;
      imax = 1;
      factor = qnan(R(0));
      return false;
    }
    const R fl = floor(scaled);            // == trunc(scaled) for scaled >= 0
    const bool whole = fl == scaled;       // ceil == floor exactly on a knot (then min == max, factor 0)
    const bool big = fl >= R(4294967040.0);
    const uint32_t lo = big ? 0xFFFFFFFEu : (uint32_t)fl;
    const uint32_t hi = whole ? lo : lo + 1u;
    if (big || hi >= len) {
      imin = len - 2;
      imax = len - 1;
      factor = xsub(scaled, c.eq_lenm2);
    } else {
      imin = lo;
      imax = hi;
      factor = xsub(scaled, fl);           // Real::fract = self - self.trunc()
    }
    return true;
  } else {
    const uint32_t m = count_le(c, x, top);  // strict_upper_bound over the whole chain
    const bool edge = (m == len) || (m == 0);
    imax = m == len ? len - 1 : (m == 0 ? 1u : m);
    imin = imax - 1;
    const R kmin = __ldg(c.ik + imin), kmax = __ldg(c.ik + imax);
    const R div = xsub(kmax, kmin);
    // linear_factor (checked, 232-248) on the edges, linear_factor_unchecked (212-224) inside
    factor = (edge && div == R(0)) ? div : xdiv(xsub(x, kmin), div);
    return true;
  }
}

// ---- output stores -----------------------------------------------------------------------------
template <class R, int D>
__device__ __forceinline__ void store_elem(R* __restrict__ out, unsigned long long i, const R (&v)[D], bool vec_ok) {
  R* p = out + i * D;
  constexpr int BYTES = D * (int)sizeof(R);
  if constexpr (BYTES == 32) {
    // [f64;4]: one 256-bit store per sample. Two 128-bit stores per thread leave every 32-byte sector half written
    // by each instruction (measured: every [f64;4] kernel capped at 3.5 TB/s)
    if (vec_ok && (reinterpret_cast<uintptr_t>(p) & 31u) == 0) {
      const double* d = reinterpret_cast<const double*>(v);
      asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(d[0]), "d"(d[1]), "d"(d[2]), "d"(d[3]) : "memory");
      return;
    }
  }
  if (BYTES % 16 == 0 && vec_ok) {
    constexpr int N4 = BYTES / 16;
    const float4* src = reinterpret_cast<const float4*>(v);
#pragma unroll
    for (int k = 0; k < N4; ++k) __stcs(reinterpret_cast<float4*>(p) + k, src[k]);
  } else if (BYTES % 8 == 0 && vec_ok) {
    constexpr int N2 = BYTES / 8;
    const float2* src = reinterpret_cast<const float2*>(v);
#pragma unroll
    for (int k = 0; k < N2; ++k) __stcs(reinterpret_cast<float2*>(p) + k, src[k]);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) __stcs(p + k, v[k]);
  }
}

// 12- and 24-byte elements ([R;3]): a per-thread store instruction of a warp strides its lanes over 3 (6) lines and
// half-fills every sector it touches, three times over (C2, ncu: 1.9 GB through the L1->XBAR port for 0.8 GB of
// output, on the port that also carries the gathers and sat at 88 % busy). Full warps therefore park their 32
// elements in shared memory and write them back as whole 16-byte words: 384 (768) contiguous bytes, every line once.
template <class R, int D>
__host__ __device__ constexpr bool stage_stores() { return D == 3; }  // 12 B ([f32;3]) and 24 B ([f64;3]); everything else stores whole sectors per lane pair
template <class R, int D>
__device__ __forceinline__ void store_warp_staged(R* __restrict__ out_warp, const R (&v)[D], R* __restrict__ stage_warp) {
  const unsigned lane = threadIdx.x & 31u;
#pragma unroll
  for (int k = 0; k < D; ++k) stage_warp[lane * D + k] = v[k];
  __syncwarp();
  constexpr int N16 = 32 * D * (int)sizeof(R) / 16;
  const float4* __restrict__ src = reinterpret_cast<const float4*>(stage_warp);
  float4* __restrict__ dst = reinterpret_cast<float4*>(out_warp);
#pragma unroll
  for (int q = 0; q < (N16 + 31) / 32; ++q)
    if (q * 32 + (int)lane < N16) __stcs(dst + q * 32 + lane, src[q * 32 + lane]);
  __syncwarp();
}

// `stage`: shared-memory scratch of blockDim.x * D values (only touched when stage_stores<R, D>()), `full_warp`:
// warp-uniform, all 32 lanes hold consecutive samples i .. i + 31 of the launch (lane 0's i a multiple of 32).
template <class R, int DH, bool PROJ>
__device__ __forceinline__ void finish_and_store(R* __restrict__ out, unsigned long long i, const R* w0, bool valid,
                                                 bool vec_ok, unsigned long long* invalid, R* __restrict__ stage = nullptr,
                                                 bool full_warp = false) {
  constexpr int D = PROJ ? DH - 1 : DH;
  R v[D];
  if (PROJ) {
    // Homogeneous::project (weights/homogeneous.rs:134-136): element / rational, true divisions
    const R rat = w0[DH - 1];
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = xdiv(w0[k], rat);
  } else {
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = w0[k];
  }
  if (!valid) {
#pragma unroll
    for (int k = 0; k < D; ++k) v[k] = qnan(R(0));
    atomicAdd(invalid, 1ull);
  }
  if constexpr (stage_stores<R, D>()) {
    if (full_warp && vec_ok) {  // vec_ok: out is 16-byte aligned, and so is every 32-sample group of it
      const unsigned long long i0 = i - (threadIdx.x & 31u);
      store_warp_staged<R, D>(out + i0 * D, v, stage + (threadIdx.x & ~31u) * D);
      return;
    }
  }
  store_elem<R, D>(out, i, v, vec_ok);
}

// Vector loads of `N` values from a 16-byte aligned address (N * sizeof(R) a multiple of 16); 256-bit
// loads (LDG.E.256) when the caller guarantees 32-byte alignment (ALIGN32).
template <int N, bool ALIGN32>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[N]) {
  static_assert(N % 4 == 0, "16-byte multiples");
  if constexpr (ALIGN32 && N % 8 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 8)
      asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
          : "=f"(v[i]), "=f"(v[i + 1]), "=f"(v[i + 2]), "=f"(v[i + 3]), "=f"(v[i + 4]), "=f"(v[i + 5]), "=f"(v[i + 6]), "=f"(v[i + 7])
          : "l"(p + i));
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
      v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
    }
  }
}
template <int N, bool ALIGN32>
__device__ __forceinline__ void load_vec(const double* __restrict__ p, double (&v)[N]) {
  static_assert(N % 2 == 0, "16-byte multiples");
  if constexpr (ALIGN32 && N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < N; i += 4)
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[i]), "=d"(v[i + 1]), "=d"(v[i + 2]), "=d"(v[i + 3]) : "l"(p + i));
  } else {
#pragma unroll
    for (int i = 0; i < N; i += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2*>(p + i));
      v[i] = t.x; v[i + 1] = t.y;
    }
  }
}
// span index carried as raw bits in the padding of a knot-row slot
__device__ __forceinline__ uint32_t raw_index(float v) { return __float_as_uint(v); }
__device__ __forceinline__ uint32_t raw_index(double v) { return (uint32_t)__double_as_longlong(v); }

// Largest block the generic B-spline kernel may be launched with: 1024 threads (one fat CTA per SM, used
// when big pivot tables are staged in shared memory) only where the register working set fits the
// 64-register cap that implies; otherwise 256.
template <class R>
__host__ __device__ constexpr int bspline_block_cap(int DEG, int DH) {
  return (DEG == 0 || (int)(sizeof(R) / 4) * (2 * DEG + (DEG + 1) * DH + 8) <= 48) ? MAX_BLOCK : BLOCK;
}

// =================================================================================================
// BSpline: span -> gather degree+1 elements and 2*degree knots -> de Boor (bspline/mod.rs:145-169)
// DEG == 0: runtime degree (DynSpace semantics; workspace in local memory).
// =================================================================================================
// SLOTS: the knot-row slot variant (sorted knots, degrees 1..3 whose rows have padding to spare). A separate
// instantiation: the extra live row costs registers that the plain variant must not pay.
template <class R>
__host__ __device__ constexpr int knot_row_stride(int deg) {
  return (2 * deg + (32 / (int)sizeof(R)) - 1) / (32 / (int)sizeof(R)) * (32 / (int)sizeof(R));
}
template <class R>
__host__ __device__ constexpr bool knot_slots_ok(int deg) {
  // f32 only: a whole slot is ONE 32-byte sector there. For f64 (64-byte rows) the variant measured slower than the
  // bucket index on C4 (27.0 vs 32.7 G samples/s: 40 bytes of spills at the 3-CTA register cap, or 2 CTAs without it).
  return sizeof(R) == 4 && deg >= 1 && deg <= 3 && 2 * deg < knot_row_stride<R>(deg);
}

template <class R, int DH, bool PROJ, int DEG, bool EQUI, bool SLOTS>
__device__ __forceinline__ void bspline_body(const CurveDev<R>& c, const SampleSrc<R>& s, R* __restrict__ out, const bool vec_ok) {
  static_assert(!SLOTS || (!EQUI && knot_slots_ok<R>(DEG)), "slot variant: sorted knots, rows with padding");
  constexpr int P_CAP = DEG > 0 ? DEG : MAX_DYN_DEGREE;
  const int p = DEG > 0 ? DEG : (int)c.degree;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const R* __restrict__ top = EQUI ? nullptr : stage_top_levels(c, smem_raw, &mbar);
  constexpr int D_OUT = PROJ ? DH - 1 : DH;
  __shared__ __align__(16) R stage[stage_stores<R, D_OUT>() ? bspline_block_cap<R>(DEG, DH) * D_OUT : 1];
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < s.count;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = base + threadIdx.x;
    const bool active = i < s.count;
    const bool full_warp = base + (threadIdx.x | 31u) < s.count;
    R x = active ? apply_pre(c, sample_param(s, i)) : c.eq_first;
    uint32_t index = 0;
    bool valid = true;
    // knot-row stride of the compile-time degrees: 2 * degree knots padded to whole 32-byte sectors
    constexpr int KSTR = (2 * P_CAP + (32 / (int)sizeof(R)) - 1) / (32 / (int)sizeof(R)) * (32 / (int)sizeof(R));
    R kr[DEG > 0 ? KSTR : 1];
    bool have_row = false;
    if constexpr (SLOTS) {
      // Knot-row slots (resolve.hpp; the Linear slot table's idea, linear_slots_kernel): cell(x) is monotone, so slot
      // b (the span whose upper knot is the first knot of cell b or beyond) serves x unless x has passed that knot;
      // slot b + 1 (the span whose lower knot is the last knot of cell b) serves it if x has passed that one too.
      // The row brings its span index along, so the bucket-index gather disappears for ~97 % of the samples (C4:
      // 7.16 -> 6.5 missing sectors per sample on the L1 -> XBAR port that bounds this kernel).
      {
        const R tq = xmul(xsub(x, c.bkt_first), c.kslot_inv);
        const uint32_t b = min(tq >= R(0) ? bucket_sat_u32(tq) : 0u, c.n_kslots - 1u);
        if (b != 0u) {
          load_vec<KSTR, true>(c.kslots + (size_t)b * KSTR, kr);
          have_row = !(x >= kr[DEG]);  // kr[p] = knots[index]
          if (!have_row) {
            load_vec<KSTR, true>(c.kslots + (size_t)(b + 1u) * KSTR, kr);
            have_row = x >= kr[DEG - 1];  // kr[p - 1] = knots[index - 1]
          }
          if (have_row) index = raw_index(kr[KSTR - 1]);
        }
      }
    }
    if (!have_row) valid = bspline_span<R, EQUI>(c, x, index, top);
    if constexpr (EQUI) {
      // Equidistant knots with a NEGATIVE step (the builders take domain(a, b) with a > b as it comes): list.rs:476-488
      // then returns max.min(min_index + 1) BELOW its `min`, and the reference panics on the element access at
      // `index - degree + i` (bspline/mod.rs:131, usize underflow). Same outcome here, and no gather in front of the tables.
      if (index < (uint32_t)p) {
        valid = false;
        index = (uint32_t)p;
      }
    }
    if (!active) continue;
    const uint32_t first = index - (uint32_t)p;
    R kn[2 * P_CAP];
    R ws[(P_CAP + 1) * DH];
    const R* __restrict__ vk = c.vk + first;
    const R* __restrict__ el = c.elems + (size_t)first * DH;
    if constexpr (DEG > 0) {
      // gathers: vector loads (LDG.E.256 / .128) where the layout guarantees alignment
      if (have_row) {
#pragma unroll
        for (int m = 0; m < 2 * P_CAP; ++m) kn[m] = kr[m];
      } else if (c.knot_rows != nullptr) {
        load_vec<KSTR, true>(c.knot_rows + (size_t)(index - c.krow_lo) * KSTR, kr);
#pragma unroll
        for (int m = 0; m < 2 * P_CAP; ++m) kn[m] = kr[m];
      } else {
#pragma unroll
        for (int m = 0; m < 2 * P_CAP; ++m) kn[m] = __ldg(vk + m);
      }
      if constexpr ((DH * (int)sizeof(R)) % 16 == 0) {
        R ev[(P_CAP + 1) * DH];
        load_vec<(P_CAP + 1) * DH, (DH * (int)sizeof(R)) % 32 == 0>(el, ev);  // el is DH*sizeof(R)-aligned
#pragma unroll
        for (int m = 0; m < (P_CAP + 1) * DH; ++m) ws[m] = ev[m];
      } else {
#pragma unroll
        for (int m = 0; m < (P_CAP + 1) * DH; ++m) ws[m] = __ldg(el + m);
      }
#pragma unroll
      for (int r = 1; r <= P_CAP; ++r) {
#pragma unroll
        for (int j = 0; j <= P_CAP - r; ++j) {
          // i = j + r + index - p ; knots[i-1] = kn[j+r-1] ; knots[i+p-r] = kn[p+j]
          const R left = kn[j + r - 1];
          const R f = xdiv(xsub(x, left), xsub(kn[P_CAP + j], left));
          const R omf = xsub(R(1), f);
#pragma unroll
          for (int k = 0; k < DH; ++k) ws[j * DH + k] = merge1(ws[j * DH + k], ws[(j + 1) * DH + k], f, omf);
        }
      }
    } else {
      for (int m = 0; m < 2 * p; ++m) kn[m] = __ldg(vk + m);
      for (int m = 0; m < (p + 1) * DH; ++m) ws[m] = __ldg(el + m);
      for (int r = 1; r <= p; ++r) {
        for (int j = 0; j <= p - r; ++j) {
          const R left = kn[j + r - 1];
          const R f = xdiv(xsub(x, left), xsub(kn[p + j], left));
          const R omf = xsub(R(1), f);
#pragma unroll
          for (int k = 0; k < DH; ++k) ws[j * DH + k] = merge1(ws[j * DH + k], ws[(j + 1) * DH + k], f, omf);
        }
      }
    }
    finish_and_store<R, DH, PROJ>(out, i, ws, valid, vec_ok, c.invalid, stage, full_warp);
  }
}

template <class R, int DH, bool PROJ, int DEG, bool EQUI>
__global__ void __launch_bounds__(bspline_block_cap<R>(DEG, DH)) bspline_kernel(const CurveDev<R> c, const SampleSrc<R> s, R* __restrict__ out,
                                                        const bool vec_ok) {
  bspline_body<R, DH, PROJ, DEG, EQUI, false>(c, s, out, vec_ok);
}
template <class R, int DH, bool PROJ, int DEG>
__global__ void __launch_bounds__(bspline_block_cap<R>(DEG, DH), 3) bspline_slots_kernel(const CurveDev<R> c, const SampleSrc<R> s,
                                                                                         R* __restrict__ out, const bool vec_ok) {
  bspline_body<R, DH, PROJ, DEG, false, true>(c, s, out, vec_ok);
}

// record stride (in R) of the Linear interval table for DH components — must match resolve.cpp
template <class R>
__host__ __device__ constexpr int linear_rec_stride(int DH) {
  return (2 + 2 * DH + (16 / (int)sizeof(R)) - 1) / (16 / (int)sizeof(R)) * (16 / (int)sizeof(R));
}

// =================================================================================================
// Linear: upper_border -> two element fetches -> merge (linear/mod.rs:122-128), Identity easing.
// =================================================================================================
template <class R, int DH, bool PROJ, bool EQUI>
__global__ void __launch_bounds__(MAX_BLOCK) linear_kernel(const CurveDev<R> c, const SampleSrc<R> s, R* __restrict__ out,
                                                       const bool vec_ok) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  const R* __restrict__ top = EQUI ? nullptr : stage_top_levels(c, smem_raw, &mbar);
  constexpr int D_OUT = PROJ ? DH - 1 : DH;
  __shared__ __align__(16) R stage[stage_stores<R, D_OUT>() ? MAX_BLOCK * D_OUT : 1];
  for (unsigned long long base = (unsigned long long)blockIdx.x * blockDim.x; base < s.count;
       base += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = base + threadIdx.x;
    const bool active = i < s.count;
    const bool full_warp = base + (threadIdx.x | 31u) < s.count;
    R x = active ? apply_pre(c, sample_param(s, i)) : R(0);
    uint32_t imin, imax;
    R f;
    R w[DH];
    bool valid = true;
    if (!EQUI && c.records != nullptr) {
      // sorted knots with an interval-record table: search, then ONE vector load brings both knots and
      // both elements of the interval (list.rs:169-203 upper_border; linear/mod.rs:122-128)
      constexpr int STRIDE = linear_rec_stride<R>(DH);
      constexpr bool A32 = (STRIDE * (int)sizeof(R)) % 32 == 0;
      const uint32_t len = c.inner_len;
      // interval that serves a search result m (list.rs:169-203): m - 1, clamped to [0, len - 2]
      auto interval_of = [len](uint32_t m) { return m == 0u ? 0u : (m >= len ? len - 2u : m - 1u); };
      R rec[STRIDE];
      uint32_t m, cn;
      if (bucket_lookup(c, x, m, cn)) {
        // The interval records carry the knots themselves: record interval_of(m) holds knot m (slot 1, or slot 0 for
        // m == 0), so the compare against the <= 3 knots of x's bucket walks the records it would fetch anyway and the
        // knot array is never touched: one index word + one record for the ~95 % of samples whose bucket is empty.
        const uint32_t stop = m + cn;
        uint32_t j = interval_of(m);
        load_vec<STRIDE, A32>(c.records + (size_t)j * STRIDE, rec);
#pragma unroll
        for (uint32_t q = 0; q < BUCKET_SCAN; ++q) {
          if (m < stop && ((m == 0u ? rec[0] : rec[1]) <= x)) {
            ++m;
            const uint32_t jn = interval_of(m);
            if (jn != j) {
              j = jn;
              load_vec<STRIDE, A32>(c.records + (size_t)j * STRIDE, rec);
            }
          } else {
            break;
          }
        }
        imin = j;
      } else {
        m = count_le_tree(c, x, top);
        imin = interval_of(m);
        load_vec<STRIDE, A32>(c.records + (size_t)imin * STRIDE, rec);
      }
      const bool edge = (m == len) || (m == 0);
      imax = imin + 1;
      if (!active) continue;
      const R div = xsub(rec[1], rec[0]);
      f = (edge && div == R(0)) ? div : xdiv(xsub(x, rec[0]), div);
      if (c.easing) f = apply_easing(c, f);
      const R omf = xsub(R(1), f);
#pragma unroll
      for (int k = 0; k < DH; ++k) w[k] = merge1(rec[2 + k], rec[2 + DH + k], f, omf);
    } else {
      valid = linear_border<R, EQUI>(c, x, imin, imax, f, top);
      if (!active) continue;
      if (c.easing) f = apply_easing(c, f);  // easing.eval(factor), linear/mod.rs:127
      const R omf = xsub(R(1), f);
      const R* __restrict__ a = c.elems + (size_t)imin * DH;
      const R* __restrict__ b = c.elems + (size_t)imax * DH;
      if constexpr ((DH * (int)sizeof(R)) % 16 == 0) {
        R av[DH], bv[DH];  // elements are DH*sizeof(R)-aligned: one or two vector loads each
        load_vec<DH, (DH * (int)sizeof(R)) % 32 == 0>(a, av);
        load_vec<DH, (DH * (int)sizeof(R)) % 32 == 0>(b, bv);
#pragma unroll
        for (int k = 0; k < DH; ++k) w[k] = merge1(av[k], bv[k], f, omf);
      } else {
#pragma unroll
        for (int k = 0; k < DH; ++k) w[k] = merge1(__ldg(a + k), __ldg(b + k), f, omf);
      }
    }
    finish_and_store<R, DH, PROJ>(out, i, w, valid, vec_ok, c.invalid, stage, full_warp);
  }
}

// =================================================================================================
// Linear over big sorted knot vectors, random batches (C2): the slot table (resolve.hpp). cell(x) is monotone, so
// every knot of a lower cell is < x and every knot of a higher cell is > x. Slot b is the interval ENDING at the
// first knot of cell b (or beyond): it serves x unless x has passed that knot; slot b + 1 is the interval STARTING
// at the last knot of cell b: it serves x if x has passed that one too. What is left (x between two knots of its
// own cell), cell 0 (left of the first knot, NaN) takes the bucket index + interval records like linear_kernel.
// A thread carries SLOT_S independent samples through the phases (all first gathers in flight together, then the
// second ones, then the rare index walks): the kernel is bound by the latency of dependent L2/DRAM gathers and by
// the L1 -> XBAR request port (one missing sector per clock), not by arithmetic.
// =================================================================================================
// measured on C2 (2^20 knots, 2 cells per knot; G samples/s): 1 sample per thread 127.2, 2: 129.1, 4: 111.3 (72 registers)
template <class R>
__host__ __device__ constexpr int slot_samples() { return sizeof(R) == 4 ? 2 : 1; }

template <class R, int DH, bool PROJ>
__global__ void __launch_bounds__(BLOCK, slot_samples<R>() == 1 ? 8 : (slot_samples<R>() == 2 ? 5 : 3)) linear_slots_kernel(const CurveDev<R> c, const SampleSrc<R> s, R* __restrict__ out,
                                                                const bool vec_ok) {
  constexpr int S = slot_samples<R>();
  constexpr int STRIDE = linear_rec_stride<R>(DH);
  constexpr bool A32 = (STRIDE * (int)sizeof(R)) % 32 == 0;
  constexpr int D_OUT = PROJ ? DH - 1 : DH;
  __shared__ __align__(16) R stage[stage_stores<R, D_OUT>() ? BLOCK * D_OUT : 1];
  const uint32_t len = c.inner_len;
  auto interval_of = [len](uint32_t m) { return m == 0u ? 0u : (m >= len ? len - 2u : m - 1u); };
  constexpr unsigned long long TILE = (unsigned long long)S * BLOCK;
  for (unsigned long long base = (unsigned long long)blockIdx.x * TILE; base < s.count;
       base += (unsigned long long)gridDim.x * TILE) {
    R x[S];
    R rec[S][STRIDE];
    uint32_t cell[S];
    bool served[S], edge[S];
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const unsigned long long i = base + (unsigned long long)u * BLOCK + threadIdx.x;
      x[u] = i < s.count ? apply_pre(c, sample_param(s, i)) : c.bkt_first;
    }
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const R tq = xmul(xsub(x[u], c.bkt_first), c.slot_inv);
      cell[u] = min(tq >= R(0) ? bucket_sat_u32(tq) : 0u, c.n_slots - 1u);
      load_vec<STRIDE, A32>(c.slots + (size_t)cell[u] * STRIDE, rec[u]);
    }
    bool second[S];
#pragma unroll
    for (int u = 0; u < S; ++u) second[u] = cell[u] != 0u && x[u] >= rec[u][1];
#pragma unroll
    for (int u = 0; u < S; ++u)
      if (second[u]) load_vec<STRIDE, A32>(c.slots + (size_t)(cell[u] + 1u) * STRIDE, rec[u]);
#pragma unroll
    for (int u = 0; u < S; ++u) {
      served[u] = cell[u] != 0u && (!second[u] || x[u] >= rec[u][0]);
      edge[u] = x[u] >= c.k_last;  // served samples have m >= 1: "edge" (m == len) <=> the last knot is <= x
    }
#pragma unroll
    for (int u = 0; u < S; ++u) {
      if (!served[u]) {
        uint32_t m, cn;
        if (bucket_lookup(c, x[u], m, cn)) {
          const uint32_t stop = m + cn;
          uint32_t j = interval_of(m);
          load_vec<STRIDE, A32>(c.records + (size_t)j * STRIDE, rec[u]);
#pragma unroll
          for (uint32_t q = 0; q < BUCKET_SCAN; ++q) {
            if (m < stop && ((m == 0u ? rec[u][0] : rec[u][1]) <= x[u])) {
              ++m;
              const uint32_t jn = interval_of(m);
              if (jn != j) {
                j = jn;
                load_vec<STRIDE, A32>(c.records + (size_t)j * STRIDE, rec[u]);
              }
            } else {
              break;
            }
          }
        } else {
          m = count_le_tree<R>(c, x[u], static_cast<const R*>(nullptr));
          load_vec<STRIDE, A32>(c.records + (size_t)interval_of(m) * STRIDE, rec[u]);
        }
        edge[u] = (m == len) || (m == 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < S; ++u) {
      const unsigned long long i = base + (unsigned long long)u * BLOCK + threadIdx.x;
      const bool full_warp = base + (unsigned long long)u * BLOCK + (threadIdx.x | 31u) < s.count;
      if (i >= s.count) continue;
      const R div = xsub(rec[u][1], rec[u][0]);
      R f = (edge[u] && div == R(0)) ? div : xdiv(xsub(x[u], rec[u][0]), div);
      if (c.easing) f = apply_easing(c, f);
      const R omf = xsub(R(1), f);
      R w[DH];
#pragma unroll
      for (int k = 0; k < DH; ++k) w[k] = merge1(rec[u][2 + k], rec[u][2 + DH + k], f, omf);
      finish_and_store<R, DH, PROJ>(out, i, w, true, vec_ok, c.invalid, stage, full_warp);
    }
  }
}

// =================================================================================================
// Bezier (one curve): copy all elements -> de Casteljau triangle (bezier/mod.rs:44-56, 78-91, 240-246)
// N == 0: runtime element count (workspace in local memory).
// =================================================================================================
template <class R, int DH, bool PROJ, int N>
__global__ void __launch_bounds__(BLOCK) bezier_kernel(const CurveDev<R> c, const SampleSrc<R> s, R* __restrict__ out,
                                                       const bool vec_ok) {
  constexpr int N_CAP = N > 0 ? N : MAX_DYN_BEZIER;
  const int n = N > 0 ? N : (int)c.n_elem;
  for (unsigned long long i = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; i < s.count;
       i += (unsigned long long)gridDim.x * BLOCK) {
    const R x = apply_pre(c, sample_param(s, i));  // Bezier .domain(a,b) arrives as an affine adaptor
    const R omx = xsub(R(1), x);
    R ws[N_CAP * DH];
    if (N > 0) {
#pragma unroll
      for (int m = 0; m < N_CAP * DH; ++m) ws[m] = __ldg(c.elems + m);
#pragma unroll
      for (int k = 1; k <= N_CAP - 1; ++k)
#pragma unroll
        for (int j = 0; j < N_CAP - k; ++j)
#pragma unroll
          for (int q = 0; q < DH; ++q) ws[j * DH + q] = merge1(ws[j * DH + q], ws[(j + 1) * DH + q], x, omx);
    } else {
      for (int m = 0; m < n * DH; ++m) ws[m] = __ldg(c.elems + m);
      for (int k = 1; k <= n - 1; ++k)
        for (int j = 0; j < n - k; ++j)
#pragma unroll
          for (int q = 0; q < DH; ++q) ws[j * DH + q] = merge1(ws[j * DH + q], ws[(j + 1) * DH + q], x, omx);
    }
    finish_and_store<R, DH, PROJ>(out, i, ws, true, vec_ok, c.invalid);
  }
}

// =================================================================================================
// Many Bezier curves x take(samples) on [0,1]: one warp per curve, lanes stride the samples.
// ctrl [n_curves][N][D] -> out [n_curves][samples][D].
// =================================================================================================
template <class R, int D, int N>
__global__ void __launch_bounds__(BLOCK) bezier_many_kernel(const R* __restrict__ ctrl, unsigned long long n_curves,
                                                            uint32_t n_ctrl, uint32_t samples, R step,
                                                            R* __restrict__ out) {
  constexpr int N_CAP = N > 0 ? N : MAX_DYN_BEZIER;
  const int n = N > 0 ? N : (int)n_ctrl;
  const unsigned lane = threadIdx.x & 31u;
  const unsigned long long warps = ((unsigned long long)gridDim.x * BLOCK) >> 5;
  for (unsigned long long curve = (((unsigned long long)blockIdx.x * BLOCK + threadIdx.x) >> 5); curve < n_curves;
       curve += warps) {
    const R* __restrict__ cp = ctrl + curve * (unsigned long long)n * D;
    // small curves: the control values are read once per curve, not once per sample (a cubic has 54 flops per
    // sample; twelve loads beside them made the kernel LSU-bound)
    constexpr bool HOIST = N > 0 && N * D <= 16;
    R held[HOIST ? N_CAP * D : 1];
    if constexpr (HOIST) {
#pragma unroll
      for (int m = 0; m < N_CAP * D; ++m) held[m] = __ldg(cp + m);
    }
    for (uint32_t sidx = lane; sidx < samples; sidx += 32u) {
      const R x = xadd(xmul(step, from_index(sidx, R(0))), R(0));  // Stepper over [0,1]
      const R omx = xsub(R(1), x);
      R ws[N_CAP * D];
      if (N > 0) {
#pragma unroll
        for (int m = 0; m < N_CAP * D; ++m) ws[m] = HOIST ? held[HOIST ? m : 0] : __ldg(cp + m);
#pragma unroll
        for (int k = 1; k <= N_CAP - 1; ++k)
#pragma unroll
          for (int j = 0; j < N_CAP - k; ++j)
#pragma unroll
            for (int q = 0; q < D; ++q) ws[j * D + q] = merge1(ws[j * D + q], ws[(j + 1) * D + q], x, omx);
      } else {
        for (int m = 0; m < n * D; ++m) ws[m] = __ldg(cp + m);
        for (int k = 1; k <= n - 1; ++k)
          for (int j = 0; j < n - k; ++j)
#pragma unroll
            for (int q = 0; q < D; ++q) ws[j * D + q] = merge1(ws[j * D + q], ws[(j + 1) * D + q], x, omx);
      }
      R* p = out + (curve * samples + sidx) * D;
#pragma unroll
      for (int q = 0; q < D; ++q) __stcs(p + q, ws[q]);
    }
  }
}

// Few samples per curve: one warp per curve would idle most lanes (4 samples: 28 of 32), so the (curve, sample)
// pairs are flattened and every thread takes one pair; neighbouring threads share a curve's control values through
// L1. `total` = n_curves * samples < 2^32 - 2^24; the division by `samples` is the multiply-shift form
// (Granlund & Montgomery): q = (t + ((i - t) >> sh1)) >> sh2 with t = umulhi(magic, i).
template <class R, int D, int N>
__global__ void __launch_bounds__(BLOCK) bezier_many_flat_kernel(const R* __restrict__ ctrl, uint32_t total,
                                                                 uint32_t samples, uint32_t magic, uint32_t sh1,
                                                                 uint32_t sh2, R step, R* __restrict__ out, bool vec_ok) {
  static_assert(N > 0, "compile-time control point count");
  const uint32_t stride = gridDim.x * BLOCK;
  for (uint32_t i = blockIdx.x * BLOCK + threadIdx.x; i < total; i += stride) {
    const uint32_t t = __umulhi(magic, i);
    const uint32_t curve = (t + ((i - t) >> sh1)) >> sh2;
    const uint32_t sidx = i - curve * samples;
    const R x = xadd(xmul(step, from_index(sidx, R(0))), R(0));  // Stepper over [0,1]
    const R omx = xsub(R(1), x);
    const R* __restrict__ cp = ctrl + (size_t)curve * (N * D);
    R ws[N * D];
    if constexpr ((N * D * (int)sizeof(R)) % 16 == 0) {
      load_vec<N * D, false>(cp, ws);  // curves are N*D*sizeof(R) apart: 16-byte aligned when the base is
    } else {
#pragma unroll
      for (int m = 0; m < N * D; ++m) ws[m] = __ldg(cp + m);
    }
#pragma unroll
    for (int k = 1; k <= N - 1; ++k)
#pragma unroll
      for (int j = 0; j < N - k; ++j)
#pragma unroll
        for (int q = 0; q < D; ++q) ws[j * D + q] = merge1(ws[j * D + q], ws[(j + 1) * D + q], x, omx);
    R v[D];
#pragma unroll
    for (int q = 0; q < D; ++q) v[q] = ws[q];
    store_elem<R, D>(out, i, v, vec_ok);
  }
}

// =================================================================================================
// Bezier::gen_with_tangent / gen_with_deriatives::<K> (bezier/mod.rs:96-153, 270-285), restated as
// written upstream: the k-th derivative is prod(len-k..len) * (k-th forward difference) of the control
// polygon after len-1-k folding steps, EXCEPT that the highest one (k == min(K, len-1)) is taken from the
// still-zeroed output array and is therefore always zero; K <= len-1 panics upstream (host rejects it).
// Not a throughput path: runtime sizes, workspace in local memory. out [n][K][D].
// =================================================================================================
constexpr int MAX_DERIV_N = 64;

template <class R, int D>
__global__ void __launch_bounds__(BLOCK) bezier_deriv_kernel(const CurveDev<R> c, const SampleSrc<R> s, int K, int tangent,
                                                             R* __restrict__ out) {
  const int len = (int)c.n_elem;
  for (unsigned long long i = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; i < s.count;
       i += (unsigned long long)gridDim.x * BLOCK) {
    const R x = sample_param(s, i);
    R ws[MAX_DERIV_N * D];
    R grad[MAX_DERIV_N * D];
    for (int m = 0; m < len * D; ++m) ws[m] = __ldg(c.elems + m);
    auto fold = [&](int count) {  // one triangle_folding_inline step over the first `count` + 1 points
      const R omx = xsub(R(1), x);
      for (int j = 0; j < count; ++j)
#pragma unroll
        for (int q = 0; q < D; ++q) ws[j * D + q] = merge1(ws[j * D + q], ws[(j + 1) * D + q], x, omx);
    };
    R* o = out + i * (unsigned long long)K * D;
    if (tangent) {  // bezier_with_tangent, K == 2
      if (len < 2) {
#pragma unroll
        for (int q = 0; q < D; ++q) {
          o[q] = ws[q];
          o[D + q] = xmul(ws[q], R(0));
        }
        continue;
      }
      for (int k = 1; k <= len - 2; ++k) fold(len - k);
      const R scale = from_index((unsigned long long)(len - 1), R(0));
      const R omx = xsub(R(1), x);
#pragma unroll
      for (int q = 0; q < D; ++q) {
        o[D + q] = xmul(xsub(ws[D + q], ws[q]), scale);
        o[q] = merge1(ws[q], ws[D + q], x, omx);
      }
      continue;
    }
    const int deg = K < len - 1 ? K : len - 1;
    for (int k = 1; k <= len - deg - 1; ++k) fold(len - k);
    for (int m = 0; m < K; ++m)
#pragma unroll
      for (int q = 0; q < D; ++q) grad[m * D + q] = xmul(ws[q], R(0));
    for (int k = deg; k >= 1; --k) {
      for (int st = 1; st <= k; ++st)  // lower_triangle_folding_inline(grad[..=k], second - first, k)
        for (int j = st; j <= k; ++j)
#pragma unroll
          for (int q = 0; q < D; ++q) grad[j * D + q] = xsub(grad[j * D + q], grad[(j - 1) * D + q]);
      unsigned long long prod = 1;
      for (int f = len - k; f < len; ++f) prod *= (unsigned long long)f;
      const R pr = from_index(prod, R(0));
#pragma unroll
      for (int q = 0; q < D; ++q) grad[k * D + q] = xmul(grad[k * D + q], pr);
      fold(len - 1);  // "one step of the normal folding" over the whole slice
      for (int j = 0; j < k; ++j)
#pragma unroll
        for (int q = 0; q < D; ++q) grad[j * D + q] = ws[j * D + q];
    }
    for (int m = 0; m < K * D; ++m) o[m] = grad[m];
  }
}

// =================================================================================================
// Parity / debug kernels: span decision only, and the stepper values.
// kind: 0 linear, 1 bezier, 2 bspline
// =================================================================================================
template <class R, int KIND, bool EQUI>
__global__ void __launch_bounds__(BLOCK) span_kernel(const CurveDev<R> c, const SampleSrc<R> s, uint32_t* __restrict__ idx,
                                                     R* __restrict__ factor) {
  for (unsigned long long base = (unsigned long long)blockIdx.x * BLOCK; base < s.count;
       base += (unsigned long long)gridDim.x * BLOCK) {
    const unsigned long long i = base + threadIdx.x;
    const bool active = i < s.count;
    R x = active ? apply_pre(c, sample_param(s, i)) : R(0);
    uint32_t index = 0;
    R f = R(0);
    bool valid = true;
    if (KIND == 2) {
      valid = bspline_span<R, EQUI>(c, x, index);
    } else if (KIND == 0) {
      uint32_t lo;
      valid = linear_border<R, EQUI>(c, x, lo, index, f);
    }
    if (!active) continue;
    if (!valid) {
      index = 0xFFFFFFFFu;
      f = qnan(R(0));
    }
    idx[i] = index;
    if (factor) factor[i] = f;
  }
}

// Stepper values at the representable indices (s.first + j) << ls only: the distinct parameters of an f32 take beyond
// index 2^24 (rows.cuh RUNS), for the run-length host transfer (capi.cu).
template <class R>
__global__ void __launch_bounds__(BLOCK) stepper_strided_kernel(const SampleSrc<R> s, unsigned ls, R* __restrict__ out) {
  for (unsigned long long j = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; j < s.count;
       j += (unsigned long long)gridDim.x * BLOCK)
    out[j] = xadd(xmul(s.step, from_index((s.first + j) << ls, R(0))), s.start);
}

// Generic stepper runs (capi.cu generic_runs_take): out[i] = vals[m(a + i)] for i in [0, count), m = the index of the
// distinct value sample a + i reads: RN-even((a + i - vb) / 2^ls), the integer form of from_usize's rounding (rows.cuh
// RUNS). W = 32-bit words per element. The 32 lanes of a warp read one or two values (L1 broadcasts) and store
// coalesced: the kernel runs at the store roof whatever produced `vals`.
template <int W>
__global__ void __launch_bounds__(BLOCK) runs_expand_kernel(const uint32_t* __restrict__ vals, uint32_t* __restrict__ out,
                                                            unsigned long long a, unsigned long long count,
                                                            unsigned long long vb, unsigned ls) {
  const unsigned long long hm1 = ((1ull << ls) >> 1) - 1ull, vq = (vb >> ls) & 1ull;
  for (unsigned long long i = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; i < count;
       i += (unsigned long long)gridDim.x * BLOCK) {
    const unsigned long long dd = a + i - vb;
    const unsigned long long m = (dd + hm1 + (((dd >> ls) + vq) & 1ull)) >> ls;
    if constexpr (W == 4) {
      __stcs(reinterpret_cast<float4*>(out) + i, __ldg(reinterpret_cast<const float4*>(vals) + m));
    } else if constexpr (W == 2) {
      __stcs(reinterpret_cast<float2*>(out) + i, __ldg(reinterpret_cast<const float2*>(vals) + m));
    } else {
#pragma unroll
      for (int q = 0; q < W; ++q) __stcs(out + i * W + q, __ldg(vals + m * W + q));
    }
  }
}

template <class R>
__global__ void __launch_bounds__(BLOCK) stepper_kernel(const SampleSrc<R> s, R* __restrict__ out) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; i < s.count;
       i += (unsigned long long)gridDim.x * BLOCK)
    out[i] = sample_param(s, i);
}

}  // namespace enterp
