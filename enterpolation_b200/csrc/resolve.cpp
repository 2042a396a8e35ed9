// See resolve.hpp. Compiled with -ffp-contract=off: table entries must carry exactly the bits the
// reference's per-access arithmetic would produce.
#include "resolve.hpp"

#include "rows_layout.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>

namespace enterp {
namespace {

struct Fail {
  enterp_status status;
  uint64_t found, necessary, index;
  int32_t mode;
};
inline Fail ok() { return {ENTERP_OK, 0, 0, 0, 0}; }
inline Fail fail(enterp_status s, uint64_t found = 0, uint64_t necessary = 0, int32_t mode = 0, uint64_t index = 0) {
  return {s, found, necessary, index, mode};
}

// Sorted::new — src/base/list.rs:263-278: error index = first i whose predecessor is greater or
// unordered (NaN).
template <class R>
Fail sorted_check(const void* knots, uint64_t n) {
  const R* k = static_cast<const R*>(knots);
  for (uint64_t i = 1; i < n; ++i) {
    const R a = k[i - 1], b = k[i];
    const bool ordered = (a < b) || (a == b);
    if (!ordered) return fail(ENTERP_NOT_SORTED, 0, 0, 0, i);
  }
  return ok();
}
Fail sorted_check(const enterp_curve_desc& d) {
  if (d.n_knots && !d.knots) return fail(ENTERP_INVALID_ARGUMENT);
  return d.scalar == ENTERP_F32 ? sorted_check<float>(d.knots, d.n_knots) : sorted_check<double>(d.knots, d.n_knots);
}

// The shape a chain resolves to (filled progressively, stage by stage).
struct Shape {
  int32_t adaptor = AD_NONE;
  uint64_t border_n = 0, inner_len = 0, virt_len = 0, degree = 0, space_len = 0;
};

// ---- stage 0: .elements(..) / .elements_with_weights(..) ---------------------------------------
Fail check_elements(const enterp_curve_desc& d) {
  if (d.kind != ENTERP_LINEAR && d.kind != ENTERP_BEZIER && d.kind != ENTERP_BSPLINE) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.scalar != ENTERP_F32 && d.scalar != ENTERP_F64) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.dim < 1 || d.dim > 4) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.n_elements && !d.elements) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.weighted && d.n_elements && !d.weights) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.n_elements >= (1ull << 31)) return fail(ENTERP_UNSUPPORTED);
  if (d.kind == ENTERP_BEZIER) {
    // src/bezier/builder.rs:141, 178
    return d.n_elements < 1 ? fail(ENTERP_EMPTY) : ok();
  }
  // src/linear/builder.rs:172, 230; src/bspline/builder.rs:232, 269
  return d.n_elements < 2 ? fail(ENTERP_TOO_FEW_ELEMENTS, d.n_elements) : ok();
}

// ---- stage 1: knots --------------------------------------------------------------------------
Fail check_knots(const enterp_curve_desc& d, Shape& s) {
  const uint64_t e = d.n_elements, k = d.n_knots, c = d.eq_count;
  if (d.kind == ENTERP_BEZIER) {
    s.degree = e - 1;
    return ok();
  }
  if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
    // Linear::equidistant_unchecked (src/linear/mod.rs:193-207), and BSpline::new(elements, ConstEquidistant::<R, N>,
    // space) (src/bspline/mod.rs:212-235) — the generic constructor takes any SortedChain; no builder reaches it
    if (d.kind != ENTERP_LINEAR && d.kind != ENTERP_BSPLINE) return fail(ENTERP_UNSUPPORTED);
  } else if (d.knot_base != ENTERP_KNOTS_SORTED && d.knot_base != ENTERP_KNOTS_EQUIDISTANT) {
    return fail(ENTERP_INVALID_ARGUMENT);
  }
  if (d.knot_base == ENTERP_KNOTS_SORTED && k >= (1ull << 31)) return fail(ENTERP_UNSUPPORTED);
  if (d.kind == ENTERP_LINEAR) {
    s.degree = 1;
    if (d.easing < ENTERP_EASE_IDENTITY || d.easing > ENTERP_EASE_PLATEAU) return fail(ENTERP_INVALID_ARGUMENT);
    if ((d.easing == ENTERP_EASE_SMOOTHSTART || d.easing == ENTERP_EASE_SMOOTHEND) && d.easing_n < 1)
      return fail(ENTERP_INVALID_ARGUMENT);
    if (d.knot_base == ENTERP_KNOTS_SORTED) {
      // src/linear/builder.rs:331 (count) then 335 (Sorted::new)
      if (k != e) return fail(ENTERP_KNOT_ELEMENT_INEQUALITY, e, k);
      Fail f = sorted_check(d);
      if (f.status) return f;
      s.inner_len = s.virt_len = k;
    } else if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
      s.inner_len = s.virt_len = e;  // ConstEquidistant<R, N> with N = elements.len()
    } else {
      // src/linear/builder.rs:427-453: as many equidistant knots as elements
      if (d.eq_form < ENTERP_EQ_DOMAIN || d.eq_form > ENTERP_EQ_DISTANCE) return fail(ENTERP_INVALID_ARGUMENT);
      s.inner_len = s.virt_len = e;
    }
    return ok();
  }
  // BSpline
  if (d.mode < ENTERP_OPEN || d.mode > ENTERP_LEGACY) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
    // BSpline::new (src/bspline/mod.rs:212-235): N = n_knots knots i/(N-1); the constructor's own checks, in its order
    if (d.mode != ENTERP_OPEN) return fail(ENTERP_UNSUPPORTED);
    if (k >= (1ull << 31)) return fail(ENTERP_UNSUPPORTED);
    if (k < e || e <= k - e + 1) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
    s.inner_len = s.virt_len = k;
    s.adaptor = AD_NONE;
  } else if (d.knot_base == ENTERP_KNOTS_SORTED) {
    switch (d.mode) {
      case ENTERP_OPEN:  // src/bspline/builder.rs:377-401
        if (k < 2) return fail(ENTERP_TOO_FEW_KNOTS, k);
        if (k < e || e <= k - e + 1) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
        break;
      case ENTERP_CLAMPED:  // 447-466
        if (k < 2) return fail(ENTERP_TOO_FEW_KNOTS, k);
        if (e < k) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_CLAMPED);
        break;
      default:  // ENTERP_LEGACY 512-534
        if (k < 4) return fail(ENTERP_TOO_FEW_KNOTS, k);
        if (k <= e + 1 || e < k - e) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_LEGACY);
        break;
    }
    // Open/Clamped: `Sorted::new(knots)?`; Legacy unwraps (builder.rs:529, a panic upstream) —
    // surfaced as NotSorted through the ABI instead of aborting the process.
    Fail f = sorted_check(d);
    if (f.status) return f;
    s.inner_len = k;
    if (d.mode == ENTERP_OPEN) {
      s.adaptor = AD_NONE;
      s.virt_len = k;
    } else if (d.mode == ENTERP_CLAMPED) {
      s.adaptor = AD_BORDER_BUFFER;
      s.border_n = e - k;  // builder.rs:459
      s.virt_len = k + 2 * s.border_n;
    } else {
      s.adaptor = AD_BORDER_DELETION;
      s.virt_len = k - 2;
    }
  } else {
    if (d.mode == ENTERP_LEGACY) return fail(ENTERP_UNSUPPORTED);  // no impl block upstream
    if (d.eq_by != ENTERP_EQ_BY_DEGREE && d.eq_by != ENTERP_EQ_BY_QUANTITY) return fail(ENTERP_INVALID_ARGUMENT);
    if (d.eq_form < ENTERP_EQ_DOMAIN || d.eq_form > ENTERP_EQ_DISTANCE) return fail(ENTERP_INVALID_ARGUMENT);
    if (c >= (1ull << 31)) return fail(ENTERP_UNSUPPORTED);
    const bool open = d.mode == ENTERP_OPEN;
    uint64_t deg = 0;
    if (d.eq_by == ENTERP_EQ_BY_DEGREE) {
      // Open 633-655 / Clamped 780-798
      if (c < 1) return fail(ENTERP_INVALID_DEGREE, c);
      if (e <= c) return fail(ENTERP_INCONGRUOUS_ELEMENTS_DEGREE, e, c, d.mode);
      deg = c;
      s.inner_len = open ? e - 1 + c : e - c + 1;
    } else {
      // Open 673-692 (line 680 reports a *legacy*-flavoured error) / Clamped 820-836
      if (c < 2) return fail(ENTERP_TOO_FEW_KNOTS, c);
      if (open) {
        if (c < e) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_LEGACY);
        if (e <= c - e + 1) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_OPEN);
        deg = c - e + 1;
      } else {
        if (e < c) return fail(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_CLAMPED);
        deg = e - c + 1;
      }
      s.inner_len = c;
    }
    if (open) {  // 903-933
      s.adaptor = AD_NONE;
      s.virt_len = s.inner_len;
    } else {  // 972-1017: BorderBuffer::new(Equidistant, deg - 1)
      s.adaptor = AD_BORDER_BUFFER;
      s.border_n = deg - 1;
      s.virt_len = s.inner_len + 2 * s.border_n;
    }
  }
  s.degree = s.virt_len - e + 1;  // src/bspline/mod.rs:254
  return ok();
}

// ---- stage 2: workspace ------------------------------------------------------------------------
Fail check_space(const enterp_curve_desc& d, Shape& s) {
  const uint64_t e = d.n_elements;
  if (d.kind == ENTERP_LINEAR) return ok();
  if (d.space < ENTERP_SPACE_CONSTANT || d.space > ENTERP_SPACE_WORKSPACE) return fail(ENTERP_INVALID_ARGUMENT);
  if (d.kind == ENTERP_BEZIER) {
    if (d.input_form < ENTERP_INPUT_NORMALIZED || d.input_form > ENTERP_INPUT_TRANSFORM) return fail(ENTERP_INVALID_ARGUMENT);
    if (d.space == ENTERP_SPACE_WORKSPACE) {
      // src/bezier/builder.rs:354-367 — the error is built with swapped arguments (line 359)
      if (d.workspace_n < e) return fail(ENTERP_TOO_SMALL_WORKSPACE, e, d.workspace_n);
      s.space_len = d.workspace_n;
    } else {
      s.space_len = e;  // dynamic 317-324; constant::<N>() pins N == len by type (330-340)
    }
    return ok();
  }
  const uint64_t deg = s.virt_len - e + 1;
  if (d.space == ENTERP_SPACE_DYNAMIC) {
    s.space_len = deg + 1;  // src/bspline/builder.rs:1071-1078
    return ok();
  }
  // constant::<N>() 1087-1103, workspace(S) 1114-1130
  if (d.workspace_n <= deg) return fail(ENTERP_TOO_SMALL_WORKSPACE, d.workspace_n, deg);
  s.space_len = d.workspace_n;
  return ok();
}

Fail run_stages(const enterp_curve_desc& d, int stage, Shape& s) {
  Fail f = check_elements(d);
  if (f.status || stage <= STAGE_ELEMENTS) return f;
  f = check_knots(d, s);
  if (f.status || stage <= STAGE_KNOTS) return f;
  return check_space(d, s);
}

void report(const Fail& f, enterp_error_info* err) {
  if (!err) return;
  err->status = f.status;
  err->mode = f.mode;
  err->found = f.found;
  err->necessary = f.necessary;
  err->index = f.index;
}

// ---- table construction ------------------------------------------------------------------------
template <class R>
bool is_pow2(R v) {
  if (!(v > R(0)) || !std::isfinite(v)) return false;
  int ex;
  return std::frexp(v, &ex) == R(0.5);
}

// Safe magnitude window of denominators for the FMA-based exact division in rows.cuh.
template <class R>
bool den_in_window(R v) {
  const R lo = sizeof(R) == 4 ? R(9.5367431640625e-07) : R(6.223015277861142e-61);   // 2^-20 / 2^-200
  const R hi = sizeof(R) == 4 ? R(1048576.0) : R(1.6069380442589903e+60);            // 2^20  / 2^200
  return std::isfinite(v) && v >= lo && v <= hi;
}

// Span rows (see rows_layout.hpp): one row per span index in [lo, hi].
// Monotone integer key of a float (ordered like the floats themselves; -0 and +0 share key 0).
template <class R>
struct FloatKey;
template <>
struct FloatKey<float> {
  using I = int64_t;
  static I key(float v) {
    int32_t i;
    std::memcpy(&i, &v, 4);
    return i >= 0 ? (I)i : -(I)(i & 0x7fffffff);
  }
  static float value(I k) {
    const uint32_t u = k >= 0 ? (uint32_t)k : (0x80000000u | (uint32_t)(-k));
    float v;
    std::memcpy(&v, &u, 4);
    return v;
  }
};
template <>
struct FloatKey<double> {
  using I = int64_t;
  static I key(double v) {
    int64_t i;
    std::memcpy(&i, &v, 8);
    return i >= 0 ? i : -(I)(i & 0x7fffffffffffffffll);
  }
  static double value(I k) {
    const uint64_t u = k >= 0 ? (uint64_t)k : (0x8000000000000000ull | (uint64_t)(-k));
    double v;
    std::memcpy(&v, &u, 8);
    return v;
  }
};

template <class R>
void build_rows(const enterp_curve_desc& d, const Shape& s, Resolved& r, const R* vk) {
  r.rows_mode = 0;
  const int p = (int)s.degree;
  if (s.degree < 1 || s.degree > (uint64_t)ROWS_MAX_DEGREE) return;
  // ConstEquidistant spans come from floor(x * fl(N - 1)) (list.rs:652-664), not from comparisons with the tabulated
  // knots: the generic kernel carries that arithmetic; no span rows
  if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) return;
  const uint64_t lo = s.degree, hi = s.virt_len - s.degree;
  const uint64_t n_rows = hi - lo + 1;
  const int W = lanes_of<R>();
  const bool equi = d.knot_base == ENTERP_KNOTS_EQUIDISTANT;
  const int len = row_len<R>(p, r.dh);
  if (n_rows * (uint64_t)len * sizeof(R) > ROWS_GLOBAL_MAX_BYTES) return;
  if (equi) {
    const R step = (R)r.eq_step;
    if (!(step > R(0)) || !std::isfinite(step)) return;  // degenerate chains keep the generic kernel
  }
  const FactorSched sch = make_sched(p, W);
  const int nops = sch.nops;
  const int dhp = round_up(r.dh, W);
  const R nan = std::numeric_limits<R>::quiet_NaN();
  std::vector<R> tab((size_t)n_rows * len, R(0));
  const R* el = reinterpret_cast<const R*>(r.elems.data());
  bool all_pow2 = true, all_window = true;
  uint32_t op_pow2 = (1u << nops) - 1u;  // bit t: every needed row has power-of-two denominators in all lanes of op t
  for (uint64_t idx = lo; idx <= hi; ++idx) {
    R* row = tab.data() + (idx - lo) * len;
    // a span strictly inside (lo, hi) is only ever selected when it has positive width
    const bool needed = equi || idx == lo || idx == hi || vk[idx - 1] < vk[idx];
    const R* kn = vk + (idx - p);  // kn[m] = knots[index - p + m], m = 0 .. 2p-1
    for (int t = 0; t < nops; ++t)
      for (int l = 0; l < W; ++l) {
        const int rr = sch.r[t][l], j = sch.j[t][l];
        const R left = kn[j + rr - 1];                 // knots[i - 1]
        const R den = kn[p + j] - left;                // knots[i + p - r] - knots[i - 1]
        if (needed) {
          if (!(is_pow2(den) && den >= R(1e-30) && den <= R(1e30))) {
            all_pow2 = false;
            op_pow2 &= ~(1u << t);
          }
          if (!den_in_window(den)) all_window = false;
        }
        row[t * W + l] = left;
        row[nops * W + t * W + l] = needed ? R(1) / den : nan;
        row[2 * nops * W + t * W + l] = needed ? -den : nan;
      }
    for (int i = 0; i <= p; ++i)
      for (int c = 0; c < r.dh; ++c) row[3 * nops * W + i * dhp + c] = el[(idx - p + i) * r.dh + c];
    if (!equi) {
      // the search (list.rs:69-86 restricted to [lo, hi]) returns idx exactly for knots[idx-1] <= x < knots[idx];
      // everything below knots[lo] maps to lo, everything from knots[hi-1] up to hi
      const R inf = std::numeric_limits<R>::infinity();
      row[row_hold_off<R>(p, r.dh)] = idx == lo ? -inf : vk[idx - 1];
      row[row_hold_off<R>(p, r.dh) + 1] = idx == hi ? inf : vk[idx];
    }
  }
  if (equi) {
    // Equidistant spans come from floor((x - offset) / step), whose switch-over points are NOT the tabulated
    // knots once rounding is involved. Find them exactly: row_of() replays the kernel's span arithmetic
    // (rows.cuh rows_span, EQUI; its multiply-by-inverse and reciprocal+FMA forms are bit-identical to this IEEE
    // division) and is monotone in x for step > 0, so the first float of every row is found by bisection over the
    // float ordering.
    using I = typename FloatKey<R>::I;
    const R step = (R)r.eq_step, offset = (R)r.eq_offset, first = (R)r.eq_first;
    const uint32_t imin = (uint32_t)r.imin, imax = (uint32_t)r.imax;
    const uint32_t ri_off = (uint32_t)((int64_t)r.map_add - (int64_t)lo);
    auto row_of = [&](R x) -> uint32_t {
      const R num = x - offset;
      const R scaled = num / step;
      uint32_t inner;
      if (x < first) {
        inner = imin;
      } else {
        const uint32_t u = !(scaled >= R(0)) ? 0u : (scaled >= R(4294967295.0) ? 0xFFFFFFFFu : (uint32_t)scaled);
        inner = (u < imax - 1u ? u : imax - 1u) + 1u;
      }
      const uint32_t ri = inner + ri_off;
      return ri < (uint32_t)n_rows - 1u ? ri : (uint32_t)n_rows - 1u;
    };
    const R inf = std::numeric_limits<R>::infinity();
    const int off = row_hold_off<R>(p, r.dh);
    I from = FloatKey<R>::key(-std::numeric_limits<R>::max());
    const I top = FloatKey<R>::key(std::numeric_limits<R>::max());
    tab[off] = -inf;
    for (uint64_t rr = 1; rr < n_rows; ++rr) {
      R bound = inf;  // no finite parameter reaches this row
      if (row_of(FloatKey<R>::value(top)) >= rr) {
        I a = from, b = top;  // invariant: row_of(a) < rr (or a == from) <= row_of(b)
        if (row_of(FloatKey<R>::value(a)) >= rr) {
          b = a;
        } else {
          while ((uint64_t)b - (uint64_t)a > 1) {  // key differences can exceed int64 for doubles
            const I m = a + (I)(((uint64_t)b - (uint64_t)a) / 2);
            if (row_of(FloatKey<R>::value(m)) >= rr) b = m; else a = m;
          }
        }
        bound = FloatKey<R>::value(b);
        from = b;
      }
      tab[(rr - 1) * len + off + 1] = bound;
      tab[rr * len + off] = bound;
    }
    // The last row ends where the reference starts to panic (floor(scaled).to_usize() fails from 2^64 on,
    // list.rs:486), so that "inside a hold interval" also means "valid" for the checked kernel variant.
    auto panics = [&](R x) { return !((x - offset) / step < R(18446744073709551616.0)); };
    R last_hi = inf;
    if (panics(FloatKey<R>::value(top))) {
      I a = from, b = top;
      if (panics(FloatKey<R>::value(a))) {
        b = a;
      } else {
        while ((uint64_t)b - (uint64_t)a > 1) {
          const I m = a + (I)(((uint64_t)b - (uint64_t)a) / 2);
          if (panics(FloatKey<R>::value(m))) b = m; else a = m;
        }
      }
      last_hi = FloatKey<R>::value(b);
    }
    tab[(n_rows - 1) * len + off + 1] = last_hi;
  }
  // mode 1 multiplies by exact inverses everywhere (the equidistant span division included), mode 2
  // divides exactly from tabulated reciprocals; both need the span division to be covered.
  if (all_pow2 && (!equi || r.eq_mode == 1))
    r.rows_mode = 1;
  else if (all_window && (!equi || r.eq_mode >= 1))
    r.rows_mode = 2;
  else
    return;
  r.rows_pow2_ops = op_pow2;
  r.rows_n = (uint32_t)n_rows;
  r.rows_lo = (uint32_t)lo;
  r.rows_len = (uint32_t)len;
  r.rows.resize(tab.size() * sizeof(R));
  std::memcpy(r.rows.data(), tab.data(), r.rows.size());
}

// Bucket index of the sorted search (resolve.hpp), for 32 knots and up — small tables gain too (256 knots, random batch:
// Linear 114 -> 160, cubic B-spline 67 -> 84 G samples/s: one L1-resident lookup instead of a three-level descent) — and only
// when the knots spread over the buckets: the kernel compares at most BUCKET_SCAN knots of a bucket and falls back to the
// pivot tree for blocks that hold a more crowded bucket.
//
// Succinct layout (round 2): one 16-byte word per block of 32 buckets = { prefix: knots in all lower blocks, 32 two-bit
// knot counts (0..3, exact), flag: some bucket of the block holds more than 3 knots }. The knots below bucket b of its
// block are a masked popcount away, so 16 buckets per knot cost 8 bits per knot instead of the 64 of the 4-byte-per-bucket
// table of round 1: C2's index shrinks from 32 MB (8 buckets per knot, capped) to 8 MB (16 per knot) and shares the L2
// with the 32 MB of interval records behind it (r01: L2 hit rate 51 %, 42 DRAM bytes per sample for 16 algorithmic).
constexpr uint32_t BUCKET_SCAN = 3;
template <class R>
uint32_t bucket_of(R x, R first, R inv, uint32_t n_buckets) {
  const R t = (x - first) * inv;  // same two IEEE operations as the kernel (bucket_lookup, kernels.cuh)
  const uint32_t b = !(t >= R(0)) ? 0u : (t >= R(4294967295.0) ? 0xFFFFFFFFu : (uint32_t)t);
  return b < n_buckets - 1u ? b : n_buckets - 1u;
}
template <class R>
void build_buckets(Resolved& r, const R* knots, uint64_t n) {
  r.buckets.clear();
  r.n_buckets = 0;
  if (n < 32 || n >= (1ull << 29)) return;
  const R first = knots[0], last = knots[n - 1];
  if (!(last > first) || !std::isfinite(first) || !std::isfinite(last)) return;
  uint64_t nb = 32;
  // Many buckets per knot: with 16 per knot 94 % of the buckets are empty for evenly spread knots, so most lookups
  // touch no knot at all (r01, 4-byte entries, C2 / C4 in G samples/s: 2 per knot 74.4 / 30.2, 4: 82.6 / 31.1,
  // 8: 94.7 / 31.4). The cap keeps the index at 16 MB.
  uint64_t per_knot = 16;
  if (const char* e = std::getenv("ENTERP_BUCKETS_PER_KNOT")) {  // ablation: 1, 2, 4, 8, 16, ...
    const long v = std::atol(e);
    if (v >= 1 && v <= 64) per_knot = (uint64_t)v;
  }
  while (nb < per_knot * n) nb <<= 1;
  if (nb > (1ull << 25)) nb = 1ull << 25;
  const R width = last - first;
  const R inv = (R)nb / width;
  if (!std::isfinite(width) || !std::isfinite(inv) || !(inv > R(0))) return;
  const uint64_t n_blocks = nb / 32;
  std::vector<uint32_t> tab(n_blocks * 4, 0);  // per block: prefix, counts[0..16), counts[16..32), crowded flag
  std::vector<uint32_t> cnt(nb, 0);
  uint32_t prev = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const uint32_t b = bucket_of<R>(knots[i], first, inv, (uint32_t)nb);
    if (b < prev) return;  // cannot happen for sorted knots (bucket_of is monotone); be safe
    prev = b;
    ++cnt[b];
  }
  uint64_t crowded = 0;  // knots living in blocks the kernel hands to the tree
  uint32_t run = 0;
  for (uint64_t blk = 0; blk < n_blocks; ++blk) {
    uint32_t* w = tab.data() + blk * 4;
    w[0] = run;
    uint32_t in_block = 0;
    bool over = false;
    for (uint32_t q = 0; q < 32; ++q) {
      const uint32_t c = cnt[blk * 32 + q];
      in_block += c;
      if (c > BUCKET_SCAN) over = true;
      w[1 + (q >> 4)] |= (c < 3u ? c : 3u) << (2u * (q & 15u));
    }
    w[3] = over ? 1u : 0u;
    if (over) crowded += in_block;
    run += in_block;
  }
  if (crowded * 50 > n) return;  // more than 2 % of the knots are clustered: keep the pivot tree alone
  r.buckets = std::move(tab);
  r.n_buckets = (uint32_t)nb;
  r.bkt_first = first;
  r.bkt_inv = inv;
}

template <class R>
void build_tables(const enterp_curve_desc& d, const Shape& s, Resolved& r) {
  const R nan = std::numeric_limits<R>::quiet_NaN();
  const uint64_t e = d.n_elements;
  // elements -> [e][dh]; the (T, w) -> Homogeneous lift is one multiply per component with the
  // same operands at every fetch (src/weights/mod.rs:38-40, homogeneous.rs:83-88, 119-124), so
  // doing it once here is bit-identical. w == 0 keeps the raw element ("direction").
  {
    const R* el = static_cast<const R*>(d.elements);
    const R* w = static_cast<const R*>(d.weights);
    r.elems.resize(e * r.dh * sizeof(R));
    R* out = reinterpret_cast<R*>(r.elems.data());
    for (uint64_t i = 0; i < e; ++i) {
      if (d.weighted) {
        const R wi = w[i];
        for (int c = 0; c < d.dim; ++c) out[i * r.dh + c] = (wi == R(0)) ? el[i * d.dim + c] : el[i * d.dim + c] * wi;
        out[i * r.dh + d.dim] = (wi == R(0)) ? R(0) : wi;  // Homogeneous::infinity stores R::zero(): a -0.0 weight loses its sign
      } else {
        for (int c = 0; c < d.dim; ++c) out[i * r.dh + c] = el[i * d.dim + c];
      }
    }
  }
  if (d.kind == ENTERP_BEZIER) {
    r.domain[0] = 0.0;
    r.domain[1] = 1.0;
    if (d.input_form != ENTERP_INPUT_NORMALIZED) {
      // TransformInput::normalized_to_domain (src/base/adaptors.rs:137-139) and its domain()
      // (161-166): note t -> t * 1/(end-start) + (-start), only "right" for start == 0.
      // INPUT_TRANSFORM: the two stored fields verbatim (adaptors.rs:113-117).
      const bool verbatim = d.input_form == ENTERP_INPUT_TRANSFORM;
      const R start = (R)d.in_a, end = (R)d.in_b;
      const R add = verbatim ? (R)d.in_a : -start;
      const R mul = verbatim ? (R)d.in_b : R(1) / (end - start);
      r.has_transform = 1;
      r.in_add = add;
      r.in_mul = mul;
      r.domain[0] = (R)((R(0) - add) / mul);
      r.domain[1] = (R)((R(1) - add) / mul);
    }
    return;
  }
  // inner knot chain values
  std::vector<R> inner(s.inner_len);
  if (d.knot_base == ENTERP_KNOTS_SORTED) {
    std::memcpy(inner.data(), d.knots, s.inner_len * sizeof(R));
  } else if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
    // ConstEquidistant::eval (list.rs:588-590): fl(i) / fl(N - 1); the span code multiplies by fl(N - 1)
    const R nm1 = (R)(s.inner_len - 1);
    for (uint64_t i = 0; i < s.inner_len; ++i) inner[i] = (R)i / nm1;
    r.eq_step = nm1;
    r.eq_offset = 0;
  } else {
    // Equidistant::{new, normalized, step} (src/base/list.rs:380-408) and eval (416-418):
    // step * fl(i) + offset — two roundings per knot, exactly as recomputed upstream.
    R step, offset;
    const R len1 = (R)(s.inner_len - 1);
    if (d.eq_form == ENTERP_EQ_NORMALIZED) {
      step = R(1) / len1;
      offset = R(0);
    } else if (d.eq_form == ENTERP_EQ_DISTANCE) {
      step = (R)d.eq_b;
      offset = (R)d.eq_a;
    } else {
      const R start = (R)d.eq_a, end = (R)d.eq_b;
      step = (end - start) / len1;
      offset = start;
    }
    for (uint64_t i = 0; i < s.inner_len; ++i) {
      const R prod = step * (R)i;
      inner[i] = prod + offset;
    }
    r.eq_step = step;
    r.eq_offset = offset;
  }
  // knot store + offsets (see resolve.hpp)
  std::vector<R> store;
  if (s.adaptor == AD_BORDER_BUFFER) {
    // BorderBuffer::eval (src/bspline/adaptors.rs:47-50)
    store.resize(s.virt_len);
    for (uint64_t i = 0; i < s.virt_len; ++i) {
      uint64_t c = i < s.border_n ? s.border_n : i;
      if (c > s.inner_len + s.border_n - 1) c = s.inner_len + s.border_n - 1;
      store[i] = inner[c - s.border_n];
    }
    r.vk_off = 0;
    r.ik_off = (uint32_t)s.border_n;
  } else if (s.adaptor == AD_BORDER_DELETION) {
    store = inner;  // BorderDeletion::eval(i) = inner[i + 1] (adaptors.rs:116-118)
    r.vk_off = 1;
    r.ik_off = 0;
  } else {
    store = inner;
    r.vk_off = r.ik_off = 0;
  }
  r.n_store = store.size();
  r.store.resize(store.size() * sizeof(R));
  std::memcpy(r.store.data(), store.data(), r.store.size());
  const R* vk = store.data() + r.vk_off;

  // search bounds on the inner chain + map back (adaptors.rs:21-39, 66-77, 137-145)
  if (d.kind == ENTERP_BSPLINE) {
    const uint64_t lo = s.degree, hi = s.virt_len - s.degree;  // src/bspline/mod.rs:148-149
    if (s.adaptor == AD_BORDER_BUFFER) {
      auto map_into = [&](uint64_t i) -> uint64_t {
        if (i < s.border_n) return 0;
        if (i - s.border_n >= s.inner_len) return s.inner_len;
        return i - s.border_n;
      };
      r.imin = map_into(lo);
      r.imax = map_into(hi);
      r.map_add = (int64_t)s.border_n;
      r.map0 = 0;
      r.mapL = s.virt_len;
    } else if (s.adaptor == AD_BORDER_DELETION) {
      r.imin = lo + 1;
      r.imax = hi + 1;
      r.map_add = -1;
      r.map0 = 0;  // unreachable: imin >= 1
      r.mapL = s.inner_len - 1;
    } else {
      r.imin = lo;
      r.imax = hi;
      r.map_add = 0;
      r.map0 = 0;
      r.mapL = s.inner_len;
    }
    r.domain[0] = vk[s.degree - 1];           // src/bspline/mod.rs:180-185
    r.domain[1] = vk[s.virt_len - s.degree];
  } else {
    r.imin = 0;
    r.imax = s.inner_len;  // strict_upper_bound over the whole chain (list.rs:104-109)
    r.map_add = 0;
    r.map0 = 0;
    r.mapL = s.inner_len;
    r.domain[0] = vk[0];                      // src/linear/mod.rs:139-141
    r.domain[1] = vk[s.virt_len - 1];
  }
  if (d.knot_base == ENTERP_KNOTS_EQUIDISTANT || d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT)
    r.eq_first = inner[r.imin < s.inner_len ? r.imin : s.inner_len - 1];  // eval(min) of strict_upper_bound_clamped

  // Pivot tree for the sorted search: nodes are 32-byte sectors (FAN keys), bottom level = the searched
  // knot range, each upper level = last key of every node below; all NaN padded (see kernels.cuh).
  if (d.knot_base == ENTERP_KNOTS_SORTED) {
    const uint64_t FAN = 32 / sizeof(R);
    const uint64_t W = r.imax - r.imin;
    std::vector<std::vector<R>> lv;
    std::vector<R> cur((W + FAN - 1) / FAN * FAN, nan);
    if (cur.empty()) cur.assign(FAN, nan);
    for (uint64_t i = 0; i < W; ++i) cur[i] = inner[r.imin + i];
    lv.push_back(cur);
    while (lv.back().size() > FAN) {
      const std::vector<R>& below = lv.back();
      const uint64_t nodes = below.size() / FAN;
      std::vector<R> up((nodes + FAN - 1) / FAN * FAN, nan);
      for (uint64_t b = 0; b < nodes; ++b) up[b] = below[FAN * b + FAN - 1];  // NaN if the node is partial
      lv.push_back(up);
    }
    for (size_t i = lv.size(); i-- > 0;) {
      std::vector<unsigned char> raw(lv[i].size() * sizeof(R));
      std::memcpy(raw.data(), lv[i].data(), raw.size());
      r.levels.level.push_back(std::move(raw));
    }
  }
  if (d.knot_base == ENTERP_KNOTS_SORTED) build_buckets<R>(r, inner.data() + r.imin, r.imax - r.imin);

  // How the kernels may compute (x - offset) / step exactly without the IEEE divide (rows.cuh header):
  // 1 = step is a power of two (multiply by the exact inverse), 2 = exact division from RN(1/step).
  if (d.knot_base == ENTERP_KNOTS_EQUIDISTANT) {
    const R step = (R)r.eq_step;
    r.eq_mode = 0;
    r.eq_rstep = 0;
    if (step > R(0) && std::isfinite(step)) {
      if (is_pow2(step) && step >= R(1e-30) && step <= R(1e30)) {
        r.eq_mode = 1;
        r.eq_rstep = R(1) / step;
      } else if (den_in_window(step)) {
        r.eq_mode = 2;
        r.eq_rstep = R(1) / step;  // correctly rounded by the host division
      }
    }
  }
  if (d.kind == ENTERP_BSPLINE) build_rows<R>(d, s, r, vk);
  if (d.kind == ENTERP_BSPLINE && s.degree >= 1 && s.degree <= 7) {
    const uint32_t per = 32 / sizeof(R);
    const uint32_t stride = (uint32_t)((2 * s.degree + per - 1) / per * per);
    const uint64_t lo = s.degree, hi = s.virt_len - s.degree;
    if ((hi - lo + 1) * stride * sizeof(R) <= (1ull << 28)) {
      std::vector<R> kr((size_t)(hi - lo + 1) * stride, R(0));
      for (uint64_t idx = lo; idx <= hi; ++idx)
        for (uint64_t m = 0; m < 2 * s.degree; ++m) kr[(idx - lo) * stride + m] = vk[idx - s.degree + m];
      r.krow_stride = stride;
      r.knot_rows.resize(kr.size() * sizeof(R));
      std::memcpy(r.knot_rows.data(), kr.data(), r.knot_rows.size());
      // knot-row slots (resolve.hpp): cell(x) -> the row (and span index) that serves x, or its neighbour
      const uint64_t n_s = r.imax - r.imin;  // searched knots
      const bool map_simple = (r.imin > 0 && r.imax < s.inner_len) || s.adaptor == AD_NONE;
      uint64_t per_knot = 2;
      if (const char* e = std::getenv("ENTERP_SLOTS_PER_KNOT")) {
        const long v = std::atol(e);
        if (v >= 0 && v <= 16) per_knot = (uint64_t)v;
      }
      if (d.knot_base == ENTERP_KNOTS_SORTED && per_knot > 0 && n_s >= 4096 && sizeof(R) == 4 && s.degree <= 3 && 2 * s.degree < stride && map_simple &&
          !r.buckets.empty()) {
        const R* sk = inner.data() + r.imin;
        const R first = sk[0], last = sk[n_s - 1];
        uint64_t ns = 1;
        while (ns < per_knot * n_s) ns <<= 1;
        const R inv = (R)ns / (last - first);
        if ((ns + 1) * stride * sizeof(R) <= (64ull << 20) + stride * sizeof(R) && last > first && std::isfinite(inv) && inv > R(0)) {
          std::vector<R> sl((size_t)(ns + 1) * stride);
          uint64_t pcount = 0;  // searched knots whose cell is below b
          for (uint64_t b = 0; b <= ns; ++b) {
            while (b < ns && pcount < n_s && bucket_of<R>(sk[pcount], first, inv, (uint32_t)ns) < b) ++pcount;
            if (b == ns) pcount = n_s;
            const uint64_t index = (uint64_t)((int64_t)(r.imin + pcount) + r.map_add);  // span index, in [lo, hi]
            R* dst = sl.data() + b * stride;
            std::memcpy(dst, kr.data() + (index - lo) * stride, stride * sizeof(R));
            const uint32_t bits = (uint32_t)index;
            if (sizeof(R) == 4) {
              std::memcpy(dst + stride - 1, &bits, 4);
            } else {
              const uint64_t wide = bits;
              std::memcpy(dst + stride - 1, &wide, 8);
            }
          }
          r.n_kslots = (uint32_t)ns;
          r.kslot_inv = inv;
          r.kslots.resize(sl.size() * sizeof(R));
          std::memcpy(r.kslots.data(), sl.data(), r.kslots.size());
        }
      }
    }
  }
  if (d.kind == ENTERP_LINEAR && d.knot_base == ENTERP_KNOTS_SORTED && s.inner_len >= 2) {
    const uint32_t per = 16 / sizeof(R);
    const uint32_t stride = (uint32_t)((2 + 2 * r.dh + per - 1) / per * per);
    // up to [f64;4]+weight: 2 knots + 2 x 5 components = 96 bytes; capped at 256 MiB like the knot rows (beyond that
    // the kernel gathers knots and elements separately: linear_kernel handles records == nullptr)
    if (stride * sizeof(R) <= 128 && (s.inner_len - 1) * stride * sizeof(R) <= (1ull << 28)) {
      const R* el = reinterpret_cast<const R*>(r.elems.data());
      std::vector<R> rec((size_t)(s.inner_len - 1) * stride, R(0));
      for (uint64_t j = 0; j + 1 < s.inner_len; ++j) {
        R* p = rec.data() + j * stride;
        p[0] = inner[j];
        p[1] = inner[j + 1];
        for (int c = 0; c < r.dh; ++c) {
          p[2 + c] = el[j * r.dh + c];
          p[2 + r.dh + c] = el[(j + 1) * r.dh + c];
        }
      }
      r.rec_stride = stride;
      r.records.resize(rec.size() * sizeof(R));
      std::memcpy(r.records.data(), rec.data(), r.records.size());
      // slot table (resolve.hpp): only worth it when the search is the cost, i.e. for tables far beyond L1
      const R first = inner[0], last = inner[s.inner_len - 1];
      // Two cells per knot: ~75 % of the samples are served by their first gather, ~22 % by the second, ~3 % walk the
      // index (C2, G samples/s: no slots 99.9, 1 cell per knot 102.5, 2: 127.2, 4: 81.6 — a 128 MB table no longer
      // shares the L2; cells as a percentage of the knots, ENTERP_SLOTS_CELLS_PERCENT: 125 % 111.4, 150 % 122.1, 175 % 129.6,
      // 200 % 128.9, 250 % 117.3, 300 % 100.7). The table is 2 x the records; beyond 64 MiB it is not built (the index +
      // records path stays).
      uint64_t per_knot = 2;
      if (const char* e = std::getenv("ENTERP_SLOTS_PER_KNOT")) {  // ablation: 0 (off), 1, 2, 4
        const long v = std::atol(e);
        if (v >= 0 && v <= 16) per_knot = (uint64_t)v;
      }
      if (per_knot > 0 && s.inner_len >= 4096 && last > first && std::isfinite(first) && std::isfinite(last) &&
          !r.buckets.empty()) {
        uint64_t ns = 1;
        while (ns < per_knot * s.inner_len) ns <<= 1;
        if (const char* e = std::getenv("ENTERP_SLOTS_CELLS_PERCENT")) {  // ablation: cells as a percentage of the knots (any count works)
          const long v = std::atol(e);
          if (v >= 50 && v <= 800) ns = (uint64_t)((double)s.inner_len * (double)v / 100.0);
        }
        const bool ablation = std::getenv("ENTERP_SLOTS_PER_KNOT") != nullptr || std::getenv("ENTERP_SLOTS_CELLS_PERCENT") != nullptr;
        if (!ablation && (ns + 1) * stride * sizeof(R) > (64ull << 20) + stride * sizeof(R)) ns = 0;
        while (ns > 1 && (ns + 1) * stride * sizeof(R) > (1ull << 28)) ns >>= 1;
        const R inv = (R)ns / (last - first);
        if (ns > 0 && std::isfinite(inv) && inv > R(0)) {
          std::vector<R> sl((size_t)(ns + 1) * stride);
          uint64_t p = 0;  // knots whose cell is below b
          const uint64_t len = s.inner_len;
          for (uint64_t b = 0; b <= ns; ++b) {
            while (b < ns && p < len && bucket_of<R>(inner[p], first, inv, (uint32_t)ns) < b) ++p;
            if (b == ns) p = len;
            const uint64_t j = p == 0 ? 0 : (p >= len ? len - 2 : p - 1);
            std::memcpy(sl.data() + b * stride, rec.data() + j * stride, stride * sizeof(R));
          }
          r.n_slots = (uint32_t)ns;
          r.slot_inv = inv;
          r.slots.resize(sl.size() * sizeof(R));
          std::memcpy(r.slots.data(), sl.data(), r.slots.size());
        }
      }
    }
  }
  if (d.kind == ENTERP_LINEAR) {
    r.easing = d.easing;
    r.easing_n = d.easing_n;
    if (d.easing == ENTERP_EASE_PLATEAU) {
      // Plateau::new (easing/plateau.rs:19-25): halfed = strength / 2; min = 0 + halfed; max = 1 - halfed
      const R halfed = (R)d.easing_param / R(2);
      r.ease_min = R(0) + halfed;
      r.ease_max = R(1) - halfed;
    }
  }
}

}  // namespace

enterp_status validate(const enterp_curve_desc& d, int stage, enterp_error_info* err) {
  Shape s;
  Fail f = run_stages(d, stage, s);
  report(f, err);
  return f.status;
}

enterp_status resolve(const enterp_curve_desc& d, Resolved& out, enterp_error_info* err) {
  Shape s;
  Fail f = run_stages(d, STAGE_SPACE, s);
  report(f, err);
  if (f.status) return f.status;
  out = Resolved{};
  out.kind = d.kind;
  out.scalar = d.scalar;
  out.dim = d.dim;
  out.weighted = d.weighted ? 1 : 0;
  out.dh = d.dim + out.weighted;
  out.knot_base = d.kind == ENTERP_BEZIER ? -1 : d.knot_base;
  out.adaptor = s.adaptor;
  out.n_elem = d.n_elements;
  out.border_n = s.border_n;
  out.inner_len = s.inner_len;
  out.virt_len = s.virt_len;
  out.degree = s.degree;
  out.space_len = s.space_len;
  if (d.scalar == ENTERP_F32)
    build_tables<float>(d, s, out);
  else
    build_tables<double>(d, s, out);
  return ENTERP_OK;
}

}  // namespace enterp
