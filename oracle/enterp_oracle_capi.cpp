// TEST INFRASTRUCTURE — NOT PRODUCT CODE. C entry points over enterp_oracle.hpp for ctypes
// (tests/, smoke(), bench.py cpu_baseline / --impl reference only). Parity: PINNED, see header of
// enterp_oracle.hpp. Shares only the descriptor/status *declarations* with include/enterp.h.
//
// The builder restatement (`resolve`) follows src/bspline/builder.rs, src/linear/builder.rs and
// src/bezier/builder.rs line by line; citations inline.
#include <algorithm>
#include <cstring>
#include <memory>
#include <chrono>
#include <thread>
#include <type_traits>

#include "../include/enterp.h"
#include "enterp_oracle.hpp"

using namespace ora;

namespace {

struct Err {
  enterp_status status = ENTERP_OK;
  int mode = 0;
  uint64_t found = 0, necessary = 0, index = 0;
};
inline Err err(enterp_status s, uint64_t found = 0, uint64_t necessary = 0, int mode = 0, uint64_t index = 0) {
  Err e;
  e.status = s;
  e.found = found;
  e.necessary = necessary;
  e.mode = mode;
  e.index = index;
  return e;
}

enum Adaptor { AD_NONE = 0, AD_BUFFER = 1, AD_DELETION = 2 };

// What the typestate builder has resolved to just before `.build()`.
struct Resolved {
  int kind = 0, scalar = 0, dim = 0, weighted = 0;
  size_t e = 0;
  int knot_base = -1;  // SORTED | EQUIDISTANT | -1
  int adaptor = AD_NONE;
  size_t border_n = 0;   // BorderBuffer n
  size_t inner_len = 0;  // length of the innermost knot chain
  size_t virt_len = 0;   // K::len() of the built curve
  size_t degree = 0;
  int eq_form = 0;
  double eq_a = 0, eq_b = 0;
  int space = 0;
  size_t space_len = 0;
  int input_form = 0;
  double in_a = 0, in_b = 0;
  int easing = 0, easing_n = 1;
  double easing_param = 0;
};

template <class R>
Err check_sorted(const enterp_curve_desc& d) {
  size_t bad = first_unsorted(static_cast<const R*>(d.knots), (size_t)d.n_knots);
  if (bad != SIZE_MAX) return err(ENTERP_NOT_SORTED, 0, 0, 0, bad);
  return Err{};
}
Err check_sorted_any(const enterp_curve_desc& d) {
  return d.scalar == ENTERP_F32 ? check_sorted<float>(d) : check_sorted<double>(d);
}

Err resolve(const enterp_curve_desc& d, Resolved& r) {
  if (d.dim < 1 || d.dim > 4) return err(ENTERP_INVALID_ARGUMENT);
  if (d.scalar != ENTERP_F32 && d.scalar != ENTERP_F64) return err(ENTERP_INVALID_ARGUMENT);
  if (d.n_elements > 0 && !d.elements) return err(ENTERP_INVALID_ARGUMENT);
  if (d.weighted && d.n_elements > 0 && !d.weights) return err(ENTERP_INVALID_ARGUMENT);
  r.kind = d.kind;
  r.scalar = d.scalar;
  r.dim = d.dim;
  r.weighted = d.weighted ? 1 : 0;
  r.e = (size_t)d.n_elements;
  const size_t e = r.e;
  const size_t k = (size_t)d.n_knots;
  switch (d.kind) {
    case ENTERP_LINEAR: {
      // linear/builder.rs:172,230 — elements()/elements_with_weights(): fewer than 2 elements
      if (e < 2) return err(ENTERP_TOO_FEW_ELEMENTS, e);
      if (d.knot_base == ENTERP_KNOTS_SORTED) {
        if (k > 0 && !d.knots) return err(ENTERP_INVALID_ARGUMENT);
        // linear/builder.rs:331 then Sorted::new (335 -> list.rs:263-278)
        if (e != k) return err(ENTERP_KNOT_ELEMENT_INEQUALITY, e, k);
        Err s = check_sorted_any(d);
        if (s.status) return s;
        r.knot_base = ENTERP_KNOTS_SORTED;
        r.inner_len = r.virt_len = k;
      } else if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
        // Linear::equidistant_unchecked (linear/mod.rs:193-207): ConstEquidistant<R, N>, N = elements.len()
        r.knot_base = ENTERP_KNOTS_CONST_EQUIDISTANT;
        r.inner_len = r.virt_len = e;
      } else if (d.knot_base == ENTERP_KNOTS_EQUIDISTANT) {
        // linear/builder.rs:427-453: Equidistant::{new,normalized,step}(elements.len(), ..)
        r.knot_base = ENTERP_KNOTS_EQUIDISTANT;
        r.inner_len = r.virt_len = e;
        r.eq_form = d.eq_form;
        r.eq_a = d.eq_a;
        r.eq_b = d.eq_b;
      } else {
        return err(ENTERP_INVALID_ARGUMENT);
      }
      r.degree = 1;
      if (d.easing < ENTERP_EASE_IDENTITY || d.easing > ENTERP_EASE_PLATEAU) return err(ENTERP_INVALID_ARGUMENT);
      if ((d.easing == ENTERP_EASE_SMOOTHSTART || d.easing == ENTERP_EASE_SMOOTHEND) && d.easing_n < 1)
        return err(ENTERP_INVALID_ARGUMENT);
      r.easing = d.easing;
      r.easing_n = d.easing_n;
      r.easing_param = d.easing_param;
      return Err{};
    }
    case ENTERP_BEZIER: {
      // bezier/builder.rs:141,178
      if (e < 1) return err(ENTERP_EMPTY);
      r.knot_base = -1;
      r.degree = e - 1;
      r.input_form = d.input_form;
      r.in_a = d.in_a;
      r.in_b = d.in_b;
      r.space = d.space;
      if (d.space == ENTERP_SPACE_DYNAMIC) {
        r.space_len = e;  // 317-324
      } else if (d.space == ENTERP_SPACE_CONSTANT) {
        r.space_len = e;  // 330-340: `E: ConstChain<N>` pins N == len at compile time
      } else if (d.space == ENTERP_SPACE_WORKSPACE) {
        // 354-367; note the swapped arguments at 359: new(elements.len(), space.len())
        if (d.workspace_n < e) return err(ENTERP_TOO_SMALL_WORKSPACE, e, d.workspace_n);
        r.space_len = (size_t)d.workspace_n;
      } else {
        return err(ENTERP_INVALID_ARGUMENT);
      }
      return Err{};
    }
    case ENTERP_BSPLINE: {
      // bspline/builder.rs:232,269
      if (e < 2) return err(ENTERP_TOO_FEW_ELEMENTS, e);
      if (d.knot_base == ENTERP_KNOTS_SORTED) {
        if (k > 0 && !d.knots) return err(ENTERP_INVALID_ARGUMENT);
        r.knot_base = ENTERP_KNOTS_SORTED;
        r.inner_len = k;
        if (d.mode == ENTERP_OPEN) {  // 377-401
          if (k < 2) return err(ENTERP_TOO_FEW_KNOTS, k);
          if (k < e) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
          if (e <= k - e + 1) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
          Err s = check_sorted_any(d);
          if (s.status) return s;
          r.adaptor = AD_NONE;
          r.virt_len = k;
        } else if (d.mode == ENTERP_CLAMPED) {  // 447-466
          if (k < 2) return err(ENTERP_TOO_FEW_KNOTS, k);
          if (e < k) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_CLAMPED);
          Err s = check_sorted_any(d);
          if (s.status) return s;
          r.adaptor = AD_BUFFER;
          r.border_n = e - k;  // "deg-1"
          r.virt_len = k + 2 * r.border_n;
        } else if (d.mode == ENTERP_LEGACY) {  // 512-534
          if (k < 4) return err(ENTERP_TOO_FEW_KNOTS, k);
          if (k <= e + 1) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_LEGACY);
          if (e < k - e) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_LEGACY);
          // 529: Sorted::new(knots).unwrap() PANICS on unsorted input; reported as NotSorted here.
          Err s = check_sorted_any(d);
          if (s.status) return s;
          r.adaptor = AD_DELETION;
          r.virt_len = k - 2;
        } else {
          return err(ENTERP_INVALID_ARGUMENT);
        }
      } else if (d.knot_base == ENTERP_KNOTS_EQUIDISTANT) {
        r.knot_base = ENTERP_KNOTS_EQUIDISTANT;
        r.eq_form = d.eq_form;
        r.eq_a = d.eq_a;
        r.eq_b = d.eq_b;
        const size_t c = (size_t)d.eq_count;
        if (d.mode == ENTERP_OPEN) {
          if (d.eq_by == ENTERP_EQ_BY_DEGREE) {  // 633-655
            if (c < 1) return err(ENTERP_INVALID_DEGREE, c);
            if (e <= c) return err(ENTERP_INCONGRUOUS_ELEMENTS_DEGREE, e, c, ENTERP_OPEN);
            r.inner_len = e - 1 + c;
          } else if (d.eq_by == ENTERP_EQ_BY_QUANTITY) {  // 673-692 (680: legacy-flavoured error)
            if (c < 2) return err(ENTERP_TOO_FEW_KNOTS, c);
            if (c < e) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_LEGACY);
            if (e <= c - e + 1) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_OPEN);
            r.inner_len = c;
          } else {
            return err(ENTERP_INVALID_ARGUMENT);
          }
          r.adaptor = AD_NONE;  // 903-933
          r.virt_len = r.inner_len;
        } else if (d.mode == ENTERP_CLAMPED) {
          size_t deg;
          if (d.eq_by == ENTERP_EQ_BY_DEGREE) {  // 780-798
            if (c < 1) return err(ENTERP_INVALID_DEGREE, c);
            if (e <= c) return err(ENTERP_INCONGRUOUS_ELEMENTS_DEGREE, e, c, ENTERP_CLAMPED);
            r.inner_len = e - c + 1;
            deg = c;
          } else if (d.eq_by == ENTERP_EQ_BY_QUANTITY) {  // 820-836
            if (c < 2) return err(ENTERP_TOO_FEW_KNOTS, c);
            if (e < c) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, c, ENTERP_CLAMPED);
            r.inner_len = c;
            deg = e - c + 1;
          } else {
            return err(ENTERP_INVALID_ARGUMENT);
          }
          r.adaptor = AD_BUFFER;  // 972-1017: BorderBuffer::new(Equidistant, deg - 1)
          r.border_n = deg - 1;
          r.virt_len = r.inner_len + 2 * r.border_n;
        } else {
          return err(ENTERP_UNSUPPORTED);  // no `impl` block for Legacy + equidistant
        }
      } else if (d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
        // BSpline::new(elements, ConstEquidistant::<R, N>::new(), space) (bspline/mod.rs:212-235): N = n_knots; no
        // builder path exists, the constructor's own checks apply (no TooFewKnots, no sort check, no adaptor)
        if (d.mode != ENTERP_OPEN) return err(ENTERP_UNSUPPORTED);
        if (k < e) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
        if (e <= k - e + 1) return err(ENTERP_INCONGRUOUS_ELEMENTS_KNOTS, e, k, ENTERP_OPEN);
        r.knot_base = ENTERP_KNOTS_CONST_EQUIDISTANT;
        r.inner_len = r.virt_len = k;
        r.adaptor = AD_NONE;
      } else {
        return err(ENTERP_INVALID_ARGUMENT);
      }
      r.degree = r.virt_len - e + 1;  // bspline/mod.rs:254
      r.space = d.space;
      if (d.space == ENTERP_SPACE_DYNAMIC) {
        r.space_len = r.virt_len - e + 2;  // 1071-1078
      } else if (d.space == ENTERP_SPACE_CONSTANT || d.space == ENTERP_SPACE_WORKSPACE) {
        // 1087-1103, 1114-1130
        if (d.workspace_n <= r.virt_len - e + 1)
          return err(ENTERP_TOO_SMALL_WORKSPACE, d.workspace_n, r.virt_len - e + 1);
        r.space_len = (size_t)d.workspace_n;
      } else {
        return err(ENTERP_INVALID_ARGUMENT);
      }
      return Err{};
    }
    default:
      return err(ENTERP_INVALID_ARGUMENT);
  }
}

// ---------------------------------------------------------------------------------------------
// Type-erased holder; the sample loop lives inside the monomorphised implementation.
// ---------------------------------------------------------------------------------------------
struct CurveIface {
  Resolved res;
  std::vector<unsigned char> elem_store, weight_store, knot_store;
  virtual ~CurveIface() {}
  virtual void domain(double out[2]) const = 0;
  virtual uint64_t eval_many(const void* t, size_t n, void* out) const = 0;
  virtual uint64_t take(size_t n_total, size_t first, size_t count, void* out) const = 0;
  virtual uint64_t span_many(const void* t, size_t n, uint32_t* idx, void* factor) const = 0;
  virtual uint64_t span_take(size_t n_total, size_t first, size_t count, uint32_t* idx, void* factor) const = 0;
  // Bezier::gen_with_tangent / gen_with_deriatives::<K> (bezier/mod.rs:270-285); -1 when this is not a plain Bezier
  virtual int64_t derivatives_many(const void*, size_t, size_t, bool, void*) const { return -1; }
};

template <class R, int D>
inline void store(const Vec<R, D>& v, R* out) {
  for (int i = 0; i < D; ++i) out[i] = v.v[i];
}
template <class R, int D>
inline void store_nan(R* out) {
  for (int i = 0; i < D; ++i) out[i] = std::numeric_limits<R>::quiet_NaN();
}

// span hooks -----------------------------------------------------------------------------------
template <class K, class E>
inline void span_of(const Linear<K, E>& c, typename K::Output x, uint32_t& idx, typename K::Output& f) {
  auto b = c.border(x);
  idx = (uint32_t)b.max_index;
  f = b.factor;
}
template <class R, class E, class S>
inline void span_of(const Bezier<R, E, S>&, R, uint32_t& idx, R& f) {
  idx = 0;
  f = R(0);
}
template <class K, class E, class S>
inline void span_of(const BSpline<K, E, S>& c, typename K::Output x, uint32_t& idx, typename K::Output& f) {
  idx = (uint32_t)c.index_of(x);
  f = typename K::Output(0);
}
template <class G>
inline void span_of(const Weighted<G>& c, typename G::R x, uint32_t& idx, typename G::R& f) {
  span_of(c.inner, x, idx, f);
}
template <class G>
inline void span_of(const TransformInput<G>& c, typename G::R x, uint32_t& idx, typename G::R& f) {
  span_of(c.inner, x * c.multiplication + c.addition, idx, f);
}

template <class C, class R, int D>
struct CurveImpl : CurveIface {
  C curve;
  explicit CurveImpl(const C& c) : curve(c) {}
  void domain(double out[2]) const override {
    R d[2];
    curve.domain(d);
    out[0] = (double)d[0];
    out[1] = (double)d[1];
  }
  uint64_t eval_many(const void* t_, size_t n, void* out_) const override {
    const R* t = static_cast<const R*>(t_);
    R* out = static_cast<R*>(out_);
    uint64_t invalid = 0;
    for (size_t i = 0; i < n; ++i) {
      try {
        store<R, D>(curve.eval(t[i]), out + i * D);
      } catch (const Panic&) {
        store_nan<R, D>(out + i * D);
        ++invalid;
      }
    }
    return invalid;
  }
  // Curve::take(n_total) (signal.rs:221-228), global indices [first, first+count)
  uint64_t take(size_t n_total, size_t first, size_t count, void* out_) const override {
    R* out = static_cast<R*>(out_);
    R d[2];
    curve.domain(d);
    Stepper<R> st(n_total, d[0], d[1]);
    uint64_t invalid = 0;
    for (size_t i = 0; i < count; ++i) {
      try {
        store<R, D>(curve.eval(st.at(first + i)), out + i * D);
      } catch (const Panic&) {
        store_nan<R, D>(out + i * D);
        ++invalid;
      }
    }
    return invalid;
  }
  uint64_t span_many(const void* t_, size_t n, uint32_t* idx, void* factor_) const override {
    const R* t = static_cast<const R*>(t_);
    R* factor = static_cast<R*>(factor_);
    uint64_t invalid = 0;
    for (size_t i = 0; i < n; ++i) {
      R f = R(0);
      try {
        span_of(curve, t[i], idx[i], f);
      } catch (const Panic&) {
        idx[i] = 0xFFFFFFFFu;
        f = std::numeric_limits<R>::quiet_NaN();
        ++invalid;
      }
      if (factor) factor[i] = f;
    }
    return invalid;
  }
  uint64_t span_take(size_t n_total, size_t first, size_t count, uint32_t* idx, void* factor_) const override {
    R* factor = static_cast<R*>(factor_);
    R d[2];
    curve.domain(d);
    Stepper<R> st(n_total, d[0], d[1]);
    uint64_t invalid = 0;
    for (size_t i = 0; i < count; ++i) {
      R f = R(0);
      try {
        span_of(curve, st.at(first + i), idx[i], f);
      } catch (const Panic&) {
        idx[i] = 0xFFFFFFFFu;
        f = std::numeric_limits<R>::quiet_NaN();
        ++invalid;
      }
      if (factor) factor[i] = f;
    }
    return invalid;
  }
};

// Clamp<G> (base/adaptors.rs:14-44) and Slice<G, R> (55-106) over an already built curve. The adaptor only
// rewrites the parameter, so it is applied to the parameter array before the wrapped curve evaluates it.
struct Adapted final : CurveIface {
  const CurveIface* inner;
  int op;         // 0 clamp, 1 slice
  double a, b;    // clamp: [min, max] = inner domain; slice: multiplication, addition (exact R values)
  Adapted(const CurveIface* in, int op_, double a_, double b_) : inner(in), op(op_), a(a_), b(b_) { res = in->res; }
  void domain(double out[2]) const override { inner->domain(out); }  // Clamp 40-43; Slice 99-105
  template <class R>
  R map(R x) const {
    if (op == 0) {  // num_traits::clamp
      const R lo = (R)a, hi = (R)b;
      if (x < lo) return lo;
      if (x > hi) return hi;
      return x;
    }
    return x * (R)a + (R)b;  // TransformInput::eval (adaptors.rs:141-152)
  }
  std::vector<unsigned char> mapped(const void* t, size_t n) const {
    const size_t rs = res.scalar == ENTERP_F32 ? 4 : 8;
    std::vector<unsigned char> buf(n * rs);
    if (res.scalar == ENTERP_F32)
      for (size_t i = 0; i < n; ++i) ((float*)buf.data())[i] = map<float>(((const float*)t)[i]);
    else
      for (size_t i = 0; i < n; ++i) ((double*)buf.data())[i] = map<double>(((const double*)t)[i]);
    return buf;
  }
  std::vector<unsigned char> stepped(size_t n_total, size_t first, size_t count) const {
    double d[2];
    domain(d);
    const size_t rs = res.scalar == ENTERP_F32 ? 4 : 8;
    std::vector<unsigned char> buf(count * rs);
    if (res.scalar == ENTERP_F32) {
      Stepper<float> st(n_total, (float)d[0], (float)d[1]);
      for (size_t i = 0; i < count; ++i) ((float*)buf.data())[i] = st.at(first + i);
    } else {
      Stepper<double> st(n_total, d[0], d[1]);
      for (size_t i = 0; i < count; ++i) ((double*)buf.data())[i] = st.at(first + i);
    }
    return buf;
  }
  uint64_t eval_many(const void* t, size_t n, void* out) const override {
    auto m = mapped(t, n);
    return inner->eval_many(m.data(), n, out);
  }
  uint64_t take(size_t n_total, size_t first, size_t count, void* out) const override {
    auto t = stepped(n_total, first, count);  // Curve::take over the adaptor's own domain()
    return eval_many(t.data(), count, out);
  }
  uint64_t span_many(const void* t, size_t n, uint32_t* idx, void* factor) const override {
    auto m = mapped(t, n);
    return inner->span_many(m.data(), n, idx, factor);
  }
  uint64_t span_take(size_t n_total, size_t first, size_t count, uint32_t* idx, void* factor) const override {
    auto t = stepped(n_total, first, count);
    return span_many(t.data(), count, idx, factor);
  }
};

// Plain (unweighted, untransformed) Bezier: adds the derivative entry points.
template <class R, int D, class E, class S>
struct BezierImpl final : CurveImpl<Bezier<R, E, S>, R, D> {
  using Base = CurveImpl<Bezier<R, E, S>, R, D>;
  explicit BezierImpl(const Bezier<R, E, S>& c) : Base(c) {}
  // out [n][K][D]; tangent: K == 2 through bezier_with_tangent
  int64_t derivatives_many(const void* t_, size_t n, size_t K, bool tangent, void* out_) const override {
    using T = Vec<R, D>;
    const R* t = static_cast<const R*>(t_);
    R* out = static_cast<R*>(out_);
    const size_t len = this->curve.elements.len();
    std::vector<T> ws(len), grad(K);
    int64_t invalid = 0;
    for (size_t i = 0; i < n; ++i) {
      for (size_t q = 0; q < len; ++q) ws[q] = this->curve.elements.eval(q);
      try {
        if (tangent)
          bezier_with_tangent<R, T>(ws, t[i], grad.data());
        else
          bezier_with_derivatives<R, T>(ws, t[i], K, grad.data());
        for (size_t q = 0; q < K; ++q) store<R, D>(grad[q], out + (i * K + q) * D);
      } catch (const Panic&) {
        for (size_t q = 0; q < K * D; ++q) out[i * K * D + q] = std::numeric_limits<R>::quiet_NaN();
        ++invalid;
      }
    }
    return invalid;
  }
};

template <class R, int D, class C>
CurveIface* finish(const C& c) {
  return new CurveImpl<C, R, D>(c);
}
template <class R, int D, class S>
CurveIface* finish(const Bezier<R, SliceChain<Vec<R, D>>, S>& c) {
  return new BezierImpl<R, D, SliceChain<Vec<R, D>>, S>(c);
}

// Wrap with Weighted / TransformInput as the builders' build() do
// (linear/builder.rs:565-571; bspline/builder.rs:1262-1268; bezier/builder.rs:452-458, 486-488, 519-525).
template <class R, int D, bool W, class C>
CurveIface* wrap(const Resolved& r, const C& c) {
  if constexpr (W) {
    Weighted<C> w{c};
    if (r.kind == ENTERP_BEZIER && r.input_form == ENTERP_INPUT_DOMAIN)
      return finish<R, D>(TransformInput<Weighted<C>>::normalized_to_domain(w, (R)r.in_a, (R)r.in_b));
    if (r.kind == ENTERP_BEZIER && r.input_form == ENTERP_INPUT_TRANSFORM)  // stored fields, adaptors.rs:113-117
      return finish<R, D>(TransformInput<Weighted<C>>{w, (R)r.in_a, (R)r.in_b});
    return finish<R, D>(w);
  } else {
    if (r.kind == ENTERP_BEZIER && r.input_form == ENTERP_INPUT_DOMAIN)
      return finish<R, D>(TransformInput<C>::normalized_to_domain(c, (R)r.in_a, (R)r.in_b));
    if (r.kind == ENTERP_BEZIER && r.input_form == ENTERP_INPUT_TRANSFORM)
      return finish<R, D>(TransformInput<C>{c, (R)r.in_a, (R)r.in_b});
    return finish<R, D>(c);
  }
}

template <class R>
Equidistant<R> make_equidistant(const Resolved& r) {
  switch (r.eq_form) {
    case ENTERP_EQ_NORMALIZED:
      return Equidistant<R>::normalized(r.inner_len);
    case ENTERP_EQ_DISTANCE:
      return Equidistant<R>::step(r.inner_len, (R)r.eq_a, (R)r.eq_b);
    default:
      return Equidistant<R>::make(r.inner_len, (R)r.eq_a, (R)r.eq_b);
  }
}

// Space choice. Constant spaces are instantiated at a few capacities (the zeroed tail beyond
// degree+1 is never read, so any N > degree is semantically identical to the user's N).
template <class T, class F>
CurveIface* with_space(const Resolved& r, size_t need, F&& f) {
  if (r.space == ENTERP_SPACE_DYNAMIC) return f(DynSpace<T>{r.space_len});
  if (need <= 2) return f(ConstSpace<T, 2>{});
  if (need <= 4) return f(ConstSpace<T, 4>{});
  if (need <= 8) return f(ConstSpace<T, 8>{});
  if (need <= 16) return f(ConstSpace<T, 16>{});
  if (need <= 32) return f(ConstSpace<T, 32>{});
  return f(DynSpace<T>{std::max(need, r.space_len)});  // huge constant spaces: heap, same arithmetic
}

template <class R, int D, bool W>
CurveIface* build_typed(const enterp_curve_desc& d, const Resolved& r) {
  using T = Vec<R, D>;
  using EChain = std::conditional_t<W, WeightsChain<T, R>, SliceChain<T>>;
  using Elem = typename EChain::Output;
  // own the inputs
  auto holder = std::make_unique<std::vector<unsigned char>[]>(3);
  holder[0].assign((const unsigned char*)d.elements, (const unsigned char*)d.elements + r.e * D * sizeof(R));
  if (W) holder[1].assign((const unsigned char*)d.weights, (const unsigned char*)d.weights + r.e * sizeof(R));
  if (r.knot_base == ENTERP_KNOTS_SORTED)
    holder[2].assign((const unsigned char*)d.knots, (const unsigned char*)d.knots + (size_t)d.n_knots * sizeof(R));
  EChain elems;
  if constexpr (W)
    elems = EChain{reinterpret_cast<const T*>(holder[0].data()), reinterpret_cast<const R*>(holder[1].data()), r.e};
  else
    elems = EChain{reinterpret_cast<const T*>(holder[0].data()), r.e};
  const R* kn = reinterpret_cast<const R*>(holder[2].data());

  CurveIface* out = nullptr;
  if (r.kind == ENTERP_LINEAR) {
    Easing<R> ease;
    if (r.easing == ENTERP_EASE_PLATEAU) {
      ease = Easing<R>::plateau((R)r.easing_param);
    } else {
      ease.kind = r.easing;
      ease.n = r.easing_n;
    }
    if (r.knot_base == ENTERP_KNOTS_SORTED)
      out = wrap<R, D, W>(r, Linear<Sorted<R>, EChain>{Sorted<R>(kn, r.inner_len), elems, ease});
    else if (r.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT)
      out = wrap<R, D, W>(r, Linear<ConstEquidistant<R>, EChain>{ConstEquidistant<R>(r.inner_len), elems, ease});
    else
      out = wrap<R, D, W>(r, Linear<Equidistant<R>, EChain>{make_equidistant<R>(r), elems, ease});
  } else if (r.kind == ENTERP_BEZIER) {
    out = with_space<Elem>(r, r.e, [&](auto space) {
      return wrap<R, D, W>(r, Bezier<R, EChain, decltype(space)>{elems, space});
    });
  } else {
    auto with_knots = [&](auto knots) {
      return with_space<Elem>(r, r.degree + 1, [&](auto space) {
        return wrap<R, D, W>(r, BSpline<decltype(knots), EChain, decltype(space)>{knots, elems, space, r.degree});
      });
    };
    if (r.knot_base == ENTERP_KNOTS_SORTED) {
      Sorted<R> s(kn, r.inner_len);
      if (r.adaptor == AD_NONE) out = with_knots(s);
      else if (r.adaptor == AD_BUFFER) out = with_knots(BorderBuffer<Sorted<R>>(s, r.border_n));
      else out = with_knots(BorderDeletion<Sorted<R>>(s));
    } else if (r.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT) {
      out = with_knots(ConstEquidistant<R>(r.inner_len));
    } else {
      Equidistant<R> q = make_equidistant<R>(r);
      if (r.adaptor == AD_NONE) out = with_knots(q);
      else out = with_knots(BorderBuffer<Equidistant<R>>(q, r.border_n));
    }
  }
  out->res = r;
  out->elem_store = std::move(holder[0]);
  out->weight_store = std::move(holder[1]);
  out->knot_store = std::move(holder[2]);
  return out;
}

template <class R>
CurveIface* build_scalar(const enterp_curve_desc& d, const Resolved& r) {
  switch (r.dim * 2 + r.weighted) {
    case 2: return build_typed<R, 1, false>(d, r);
    case 3: return build_typed<R, 1, true>(d, r);
    case 4: return build_typed<R, 2, false>(d, r);
    case 5: return build_typed<R, 2, true>(d, r);
    case 6: return build_typed<R, 3, false>(d, r);
    case 7: return build_typed<R, 3, true>(d, r);
    case 8: return build_typed<R, 4, false>(d, r);
    default: return build_typed<R, 4, true>(d, r);
  }
}

// Many independent Bezier curves, each `.take(samples)` on [0,1] with a constant workspace
// (bezier/mod.rs:240-246; C3 of BASELINE.json). ctrl [n_curves][n_ctrl][dim].
template <class R, int D>
void bezier_many(uint64_t c0, uint64_t c1, uint32_t n_ctrl, const R* ctrl, uint32_t samples, R* out) {
  using T = Vec<R, D>;
  Stepper<R> st(samples, R(0), R(1));
  std::vector<T> ws(n_ctrl);
  for (uint64_t c = c0; c < c1; ++c) {
    const T* cp = reinterpret_cast<const T*>(ctrl) + c * n_ctrl;
    for (uint32_t s = 0; s < samples; ++s) {
      const R x = st.at(s);
      for (uint32_t i = 0; i < n_ctrl; ++i) ws[i] = cp[i];
      for (uint32_t k = 1; k + 1 <= n_ctrl; ++k)
        for (uint32_t i = 0; i < n_ctrl - k; ++i) ws[i] = merge(ws[i], ws[i + 1], x);
      store<R, D>(ws[0], out + (c * samples + s) * D);
    }
  }
}
}  // namespace

// NB: std::vector move keeps the heap buffer address, so the chains' raw pointers stay valid.

extern "C" {

#define ORA_API __attribute__((visibility("default")))

ORA_API int ora_create(const enterp_curve_desc* desc, void** out, enterp_error_info* info) {
  if (!desc || !out) return ENTERP_INVALID_ARGUMENT;
  *out = nullptr;
  Resolved r;
  Err e = resolve(*desc, r);
  if (info) {
    info->status = e.status;
    info->mode = e.mode;
    info->found = e.found;
    info->necessary = e.necessary;
    info->index = e.index;
  }
  if (e.status) return e.status;
  *out = desc->scalar == ENTERP_F32 ? build_scalar<float>(*desc, r) : build_scalar<double>(*desc, r);
  return ENTERP_OK;
}
ORA_API void ora_destroy(void* c) { delete static_cast<CurveIface*>(c); }
ORA_API int64_t ora_bezier_derivatives(void* c, const void* t, uint64_t n, uint64_t k, int tangent, void* out) {
  return static_cast<CurveIface*>(c)->derivatives_many(t, (size_t)n, (size_t)k, tangent != 0, out);
}
// Curve::clamp() / Curve::slice(bounds). The caller keeps `base` alive for the lifetime of the result.
ORA_API void* ora_clamp(void* base) {
  const CurveIface* b = static_cast<const CurveIface*>(base);
  double d[2];
  b->domain(d);
  return new Adapted(b, 0, d[0], d[1]);
}
ORA_API void* ora_slice(void* base, int has_start, double start, int has_end, double end) {
  const CurveIface* b = static_cast<const CurveIface*>(base);
  double d[2];
  b->domain(d);
  double mul, add;
  if (b->res.scalar == ENTERP_F32) {  // Slice::new, adaptors.rs:61-81, in R arithmetic
    const float s0 = (float)d[0], s1 = (float)d[1];
    const float b0 = has_start ? (float)start : s0, b1 = has_end ? (float)end : s1;
    const float scale = (b1 - b0) / (s1 - s0);
    mul = scale;
    add = b0 - s0;
  } else {
    const double s0 = d[0], s1 = d[1];
    const double b0 = has_start ? start : s0, b1 = has_end ? end : s1;
    const double scale = (b1 - b0) / (s1 - s0);
    mul = scale;
    add = b0 - s0;
  }
  return new Adapted(b, 1, mul, add);
}
ORA_API void ora_domain(void* c, double out[2]) { static_cast<CurveIface*>(c)->domain(out); }
ORA_API void ora_info(void* c, enterp_curve_info* info) {
  const Resolved& r = static_cast<CurveIface*>(c)->res;
  info->kind = r.kind;
  info->scalar = r.scalar;
  info->dim = r.dim;
  info->weighted = r.weighted;
  info->knot_base = r.knot_base;
  info->adaptor = r.adaptor;
  info->n_elements = r.e;
  info->n_knots = r.virt_len;
  info->degree = r.degree;
  info->border = r.border_n;
  static_cast<CurveIface*>(c)->domain(info->domain);
}
ORA_API uint64_t ora_eval(void* c, const void* t, uint64_t n, void* out) {
  return static_cast<CurveIface*>(c)->eval_many(t, (size_t)n, out);
}
ORA_API uint64_t ora_take(void* c, uint64_t n_total, uint64_t first, uint64_t count, void* out) {
  return static_cast<CurveIface*>(c)->take((size_t)n_total, (size_t)first, (size_t)count, out);
}
ORA_API uint64_t ora_span(void* c, const void* t, uint64_t n, uint32_t* idx, void* factor) {
  return static_cast<CurveIface*>(c)->span_many(t, (size_t)n, idx, factor);
}
ORA_API uint64_t ora_span_take(void* c, uint64_t n_total, uint64_t first, uint64_t count, uint32_t* idx, void* factor) {
  return static_cast<CurveIface*>(c)->span_take((size_t)n_total, (size_t)first, (size_t)count, idx, factor);
}
// All-cores variants: contiguous index ranges per std::thread, one shared read-only curve
// (the reference has no threads; this is the "all host cores" CPU baseline of BASELINE.md §2).
ORA_API uint64_t ora_take_mt(void* c, uint64_t n_total, uint64_t first, uint64_t count, void* out, int threads) {
  CurveIface* cv = static_cast<CurveIface*>(c);
  if (threads <= 1) return cv->take(n_total, first, count, out);
  const size_t stride = (size_t)cv->res.dim * (cv->res.scalar == ENTERP_F32 ? 4 : 8);
  std::vector<std::thread> pool;
  std::vector<uint64_t> inv(threads, 0);
  for (int t = 0; t < threads; ++t) {
    uint64_t lo = count * t / threads, hi = count * (t + 1) / threads;
    pool.emplace_back([=, &inv] { inv[t] = cv->take(n_total, first + lo, hi - lo, (char*)out + lo * stride); });
  }
  uint64_t total = 0;
  for (int t = 0; t < threads; ++t) {
    pool[t].join();
    total += inv[t];
  }
  return total;
}
ORA_API uint64_t ora_eval_mt(void* c, const void* tin, uint64_t n, void* out, int threads) {
  CurveIface* cv = static_cast<CurveIface*>(c);
  if (threads <= 1) return cv->eval_many(tin, n, out);
  const size_t rs = cv->res.scalar == ENTERP_F32 ? 4 : 8;
  const size_t stride = (size_t)cv->res.dim * rs;
  std::vector<std::thread> pool;
  std::vector<uint64_t> inv(threads, 0);
  for (int t = 0; t < threads; ++t) {
    uint64_t lo = n * t / threads, hi = n * (t + 1) / threads;
    pool.emplace_back(
        [=, &inv] { inv[t] = cv->eval_many((const char*)tin + lo * rs, hi - lo, (char*)out + lo * stride); });
  }
  uint64_t total = 0;
  for (int t = 0; t < threads; ++t) {
    pool[t].join();
    total += inv[t];
  }
  return total;
}

// README bench shape (benches/benches.rs:62-93: `b.iter::<Vec<f64>, _>(|| curve.by_ref().take(n).collect())`):
// `reps` sequential collects of take(n_total) into a freshly allocated vector each, single thread; returns seconds.
ORA_API double ora_bench_take(void* c, uint64_t n_total, uint64_t reps) {
  CurveIface* cv = static_cast<CurveIface*>(c);
  const size_t stride = (size_t)cv->res.dim * (cv->res.scalar == ENTERP_F32 ? 4 : 8);
  volatile unsigned char sink = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (uint64_t r = 0; r < reps; ++r) {
    std::vector<unsigned char> v(stride * n_total);  // the Vec collect() allocates
    cv->take(n_total, 0, n_total, v.data());
    sink = sink ^ v[(size_t)(r % (stride * n_total))];
  }
  const auto t1 = std::chrono::steady_clock::now();
  (void)sink;
  return std::chrono::duration<double>(t1 - t0).count();
}

// Stepper values (signal.rs:607-609)
ORA_API void ora_stepper(int scalar, double start, double end, uint64_t n_total, uint64_t first, uint64_t count, void* out) {
  if (scalar == ENTERP_F32) {
    Stepper<float> st(n_total, (float)start, (float)end);
    for (uint64_t i = 0; i < count; ++i) ((float*)out)[i] = st.at(first + i);
  } else {
    Stepper<double> st(n_total, start, end);
    for (uint64_t i = 0; i < count; ++i) ((double*)out)[i] = st.at(first + i);
  }
}

ORA_API int ora_bezier_many_take(int scalar, int dim, uint64_t n_curves, uint32_t n_ctrl, const void* ctrl,
                                 uint32_t samples, void* out, int threads) {
  if (dim < 1 || dim > 4 || n_ctrl < 1) return ENTERP_INVALID_ARGUMENT;
  auto run = [&](uint64_t c0, uint64_t c1) {
#define ORA_BM(RT, DD) bezier_many<RT, DD>(c0, c1, n_ctrl, (const RT*)ctrl, samples, (RT*)out)
    if (scalar == ENTERP_F32) {
      switch (dim) { case 1: ORA_BM(float, 1); break; case 2: ORA_BM(float, 2); break; case 3: ORA_BM(float, 3); break; default: ORA_BM(float, 4); }
    } else {
      switch (dim) { case 1: ORA_BM(double, 1); break; case 2: ORA_BM(double, 2); break; case 3: ORA_BM(double, 3); break; default: ORA_BM(double, 4); }
    }
#undef ORA_BM
  };
  if (threads <= 1) {
    run(0, n_curves);
    return ENTERP_OK;
  }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; ++t) {
    uint64_t lo = n_curves * t / threads, hi = n_curves * (t + 1) / threads;
    pool.emplace_back([=] { run(lo, hi); });
  }
  for (auto& th : pool) th.join();
  return ENTERP_OK;
}

// ---- knot-chain KAT hooks (list.rs / adaptors.rs doctests and unit tests) ---------------------
// chain: 0 Sorted(values), 1 Equidistant{len,step,offset}, 2 ConstEquidistant<N=len>;
// adaptor: 0 none, 1 BorderBuffer(n), 2 BorderDeletion. op: 0 strict_upper_bound(x),
// 1 strict_upper_bound_clamped(x,min,max), 2 upper_border(x) -> (min,max,factor), 3 eval(min), 4 len.
// Returns 0, or -1 if the reference would panic.
ORA_API int ora_chain_op(int chain, int adaptor, uint64_t border_n, const double* values, uint64_t len, double step,
                         double offset, int op, double x, uint64_t min, uint64_t max, uint64_t* out_idx,
                         double* out_val) {
  auto run = [&](const auto& ch) -> int {
    try {
      switch (op) {
        case 0: out_idx[0] = ch.strict_upper_bound(x); break;
        case 1: out_idx[0] = ch.strict_upper_bound_clamped(x, (size_t)min, (size_t)max); break;
        case 2: {
          auto b = ch.upper_border(x);
          out_idx[0] = b.min_index;
          out_idx[1] = b.max_index;
          out_val[0] = b.factor;
          break;
        }
        case 3: out_val[0] = ch.eval((size_t)min); break;
        default: out_idx[0] = ch.len(); break;
      }
    } catch (const Panic&) {
      return -1;
    }
    return 0;
  };
  auto adapt = [&](const auto& base) -> int {
    using G = std::decay_t<decltype(base)>;
    if (adaptor == 1) return run(BorderBuffer<G>(base, (size_t)border_n));
    if (adaptor == 2) return run(BorderDeletion<G>(base));
    return run(base);
  };
  if (chain == 0) return adapt(Sorted<double>(values, (size_t)len));
  if (chain == 1) return adapt(Equidistant<double>((size_t)len, step, offset));
  return adapt(ConstEquidistant<double>((size_t)len));
}
ORA_API double ora_equidistant_step(int form, uint64_t len, double a, double b) {
  if (form == ENTERP_EQ_NORMALIZED) return Equidistant<double>::normalized((size_t)len).step_;
  if (form == ENTERP_EQ_DISTANCE) return b;
  return Equidistant<double>::make((size_t)len, a, b).step_;
}

}  // extern "C"
