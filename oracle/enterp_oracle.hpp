// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
//
// CPU restatement of enterpolation 0.3.0's eval path (the reference is Rust and cannot be built
// in this image: no cargo/rustc, dependencies not vendored — see DESIGN.md §oracle). Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may load this code.
//
// Parity status: PINNED. tests/test_oracle_kat.py replays every literal known-answer test the
// reference holds for this path through this file.
//
// Build: g++ -O2 -ffp-contract=off -fno-fast-math (GCC contracts a*b+c into FMA by default; the
// reference never fuses, so contraction must stay off).
//
// Structure mirrors the reference's generics (static dispatch inside the sample loop, as rustc
// monomorphises it); every function cites the reference lines it follows (paths relative to
// /root/reference/src).
#pragma once
#include <array>
#include <cmath>
#include <cstddef>
#include <algorithm>
#include <cstdint>
#include <limits>
#include <vector>

namespace ora {

// Raised where the reference panics (`to_usize().unwrap()` on NaN/negative/huge: base/list.rs:456,
// 486, 539-540). The sample loop turns it into "all components NaN + invalid counter".
struct Panic {};

// num-traits 0.2.19 `ToPrimitive::to_usize` for floats: Some(trunc) iff -1 < x < 2^64, else None.
template <class R>
inline size_t to_usize_unwrap(R x) {
  if (!(x > R(-1)) || !(x < R(18446744073709551616.0))) throw Panic{};
  return static_cast<size_t>(x);
}
// `R::from_usize(n).unwrap()`: `n as R`, round-to-nearest-even.
template <class R>
inline R from_usize(size_t n) {
  return static_cast<R>(n);
}

// ---------------------------------------------------------------------------------------------
// Elements: `T: Add<Output=T> + Mul<R,Output=T> + Copy` (README "Requirements for Elements").
// ---------------------------------------------------------------------------------------------
template <class R, int D>
struct Vec {
  R v[D];
  Vec() {
    for (int i = 0; i < D; ++i) v[i] = R(0);  // Default::default()
  }
  friend Vec operator+(Vec a, const Vec& b) {
    for (int i = 0; i < D; ++i) a.v[i] = a.v[i] + b.v[i];
    return a;
  }
  friend Vec operator-(Vec a, const Vec& b) {
    for (int i = 0; i < D; ++i) a.v[i] = a.v[i] - b.v[i];
    return a;
  }
  friend Vec operator*(Vec a, R s) {
    for (int i = 0; i < D; ++i) a.v[i] = a.v[i] * s;
    return a;
  }
  friend Vec operator/(Vec a, R s) {
    for (int i = 0; i < D; ++i) a.v[i] = a.v[i] / s;
    return a;
  }
};

// weights/homogeneous.rs:14-17; Add 139-151; Mul<R> 195-207; project 134-136.
template <class E, class R>
struct Homogeneous {
  E element;
  R rational;
  Homogeneous() : element(), rational(R(0)) {}
  Homogeneous(E e, R r) : element(e), rational(r) {}
  friend Homogeneous operator+(const Homogeneous& a, const Homogeneous& b) {
    return Homogeneous(a.element + b.element, a.rational + b.rational);
  }
  friend Homogeneous operator*(const Homogeneous& a, R s) {
    return Homogeneous(a.element * s, a.rational * s);
  }
  E project() const { return element / rational; }
  // weighted_or_infinite 83-88 (w == 0 keeps the raw element as a direction), weighted_unchecked 119-124
  static Homogeneous weighted_or_infinite(E e, R w) {
    if (w == R(0)) return Homogeneous(e, R(0));
    return Homogeneous(e * w, w);
  }
};

// topology-traits 0.1.2 `Merge` blanket impl == utils.rs:6-12:
//   first * (1 - factor) + second * factor, every operation separately rounded.
template <class T, class R>
inline T merge(const T& self, const T& other, R factor) {
  return self * (R(1) - factor) + other * factor;
}

// ---------------------------------------------------------------------------------------------
// Chains (base/mod.rs:19-70): slice-backed element / knot containers.
// ---------------------------------------------------------------------------------------------
template <class T>
struct SliceChain {
  using Output = T;
  const T* data;
  size_t n;
  T eval(size_t i) const { return data[i]; }
  size_t len() const { return n; }
};

// weights/mod.rs:31-41 + IntoWeight for (T,R) 83-93: the lift happens at EVERY fetch.
template <class T, class R>
struct WeightsChain {
  using Output = Homogeneous<T, R>;
  const T* data;
  const R* w;
  size_t n;
  Output eval(size_t i) const { return Output::weighted_or_infinite(data[i], w[i]); }
  size_t len() const { return n; }
};

// ---------------------------------------------------------------------------------------------
// SortedChain default methods (base/list.rs:50-250), CRTP so overrides dispatch like the trait.
// ---------------------------------------------------------------------------------------------
template <class R>
struct Border {
  size_t min_index, max_index;
  R factor;
};

template <class Self, class R>
struct SortedChainDefaults {
  const Self& self() const { return static_cast<const Self&>(*this); }
  // list.rs:69-86
  size_t strict_upper_bound_clamped(R element, size_t min, size_t max) const {
    size_t pointer = min;
    size_t dist = max - min;
    while (dist > 0) {
      size_t step = dist / 2;
      size_t sample = pointer + step;
      if (element >= self().eval(sample)) {
        pointer = sample + 1;
        dist -= step + 1;
      } else {
        dist = step;
      }
    }
    return pointer;
  }
  // list.rs:104-109
  size_t strict_upper_bound(R element) const {
    return self().strict_upper_bound_clamped(element, 0, self().len());
  }
  // list.rs:212-224
  R linear_factor_unchecked(size_t min_index, size_t max_index, R element) const {
    R max = self().eval(max_index);
    R min = self().eval(min_index);
    return (element - min) / (max - min);
  }
  // list.rs:232-248
  R linear_factor(size_t min_index, size_t max_index, R element) const {
    R max = self().eval(max_index);
    R min = self().eval(min_index);
    R div = max - min;
    if (div == R(0)) return div;
    return (element - min) / div;
  }
  // list.rs:169-203
  Border<R> upper_border(R element) const {
    size_t max_index = self().strict_upper_bound(element);
    if (self().len() == max_index) {
      size_t mx = self().len() - 1, mn = mx - 1;
      return {mn, mx, self().linear_factor(mn, mx, element)};
    }
    if (max_index == 0) {
      return {0, 1, self().linear_factor(0, 1, element)};
    }
    return {max_index - 1, max_index, self().linear_factor_unchecked(max_index - 1, max_index, element)};
  }
  // signal.rs Chain::first/last
  R first() const { return self().eval(0); }
  R last() const { return self().eval(self().len() - 1); }
};

// Sorted<C> (list.rs:255-310): pure wrapper, keeps every default.
template <class R>
struct Sorted : SortedChainDefaults<Sorted<R>, R> {
  using Output = R;
  const R* k;
  size_t n;
  Sorted(const R* k_, size_t n_) : k(k_), n(n_) {}
  R eval(size_t i) const { return k[i]; }
  size_t len() const { return n; }
};
// Sorted::new (list.rs:263-278): returns the index i at which k[i-1] > k[i] or they are unordered
// (NaN), or SIZE_MAX when sorted.
template <class R>
inline size_t first_unsorted(const R* k, size_t n) {
  if (n == 0) return SIZE_MAX;
  R last = k[0];
  for (size_t i = 1; i < n; ++i) {
    R cur = k[i];
    if (!(last <= cur)) return i;  // partial_cmp None | Greater
    last = cur;
  }
  return SIZE_MAX;
}

// Equidistant<R> (list.rs:354-552)
template <class R>
struct Equidistant : SortedChainDefaults<Equidistant<R>, R> {
  using Output = R;
  size_t len_;
  R step_, offset_;
  Equidistant(size_t len, R step, R offset) : len_(len), step_(step), offset_(offset) {}
  // 380-386
  static Equidistant normalized(size_t len) { return Equidistant(len, R(1) / from_usize<R>(len - 1), R(0)); }
  // 393-399
  static Equidistant make(size_t len, R start, R end) {
    return Equidistant(len, (end - start) / from_usize<R>(len - 1), start);
  }
  // 402-408
  static Equidistant step(size_t len, R start, R step) { return Equidistant(len, step, start); }
  // 416-418: two roundings, never fused
  R eval(size_t i) const { return step_ * from_usize<R>(i) + offset_; }
  size_t len() const { return len_; }
  // 446-458
  size_t strict_upper_bound(R element) const {
    if (element < offset_) return 0;
    R scaled = (element - offset_) / step_;
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    return std::min(len_, min_index + 1);
  }
  // 476-488
  size_t strict_upper_bound_clamped(R element, size_t min, size_t max) const {
    if (element < eval(min)) return min;
    R scaled = (element - offset_) / step_;
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    return std::min(max, min_index + 1);
  }
  // 532-551
  Border<R> upper_border(R element) const {
    R scaled = (element - offset_) / step_;
    if (element < offset_) return {0, 1, scaled};
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    size_t max_index = to_usize_unwrap(std::ceil(scaled));
    if (max_index >= len_) {
      return {len_ - 2, len_ - 1, scaled - from_usize<R>(len_ - 2)};
    }
    R factor = scaled - std::trunc(scaled);  // Real::fract
    return {min_index, max_index, factor};
  }
};

// ConstEquidistant<R,N> (list.rs:566-721) — different arithmetic: i/(N-1) and x*(N-1).
template <class R>
struct ConstEquidistant : SortedChainDefaults<ConstEquidistant<R>, R> {
  using Output = R;
  size_t N;
  explicit ConstEquidistant(size_t n) : N(n) {}
  R eval(size_t i) const { return from_usize<R>(i) / from_usize<R>(N - 1); }  // 588-590
  size_t len() const { return N; }
  size_t strict_upper_bound(R element) const {  // 622-634
    if (element < R(0)) return 0;
    R scaled = element * from_usize<R>(N - 1);
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    return std::min(N, min_index + 1);
  }
  size_t strict_upper_bound_clamped(R element, size_t min, size_t max) const {  // 652-664
    if (element < eval(min)) return min;
    R scaled = element * from_usize<R>(N - 1);
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    return std::min(max, min_index + 1);
  }
  Border<R> upper_border(R element) const {  // 702-720
    R scaled = element * from_usize<R>(N - 1);
    if (element < R(0)) return {0, 1, scaled};
    size_t min_index = to_usize_unwrap(std::floor(scaled));
    size_t max_index = to_usize_unwrap(std::ceil(scaled));
    if (max_index >= N) return {N - 2, N - 1, scaled - from_usize<R>(N - 2)};
    return {min_index, max_index, scaled - std::trunc(scaled)};
  }
};

// BorderBuffer<G> (bspline/adaptors.rs:7-85)
template <class G>
struct BorderBuffer : SortedChainDefaults<BorderBuffer<G>, typename G::Output> {
  using R = typename G::Output;
  using Output = R;
  G inner;
  size_t n;
  BorderBuffer(G g, size_t n_) : inner(g), n(n_) {}
  size_t map_into(size_t index) const {  // 21-29
    if (index < n) return 0;
    if (index - n >= inner.len()) return inner.len();
    return index - n;
  }
  size_t map_from(size_t index) const {  // 31-39
    if (index == inner.len()) return len();
    if (index == 0) return 0;
    return index + n;
  }
  R eval(size_t input) const {  // 47-50
    size_t clamped = std::min(std::max(input, n), inner.len() + n - 1);
    return inner.eval(clamped - n);
  }
  size_t len() const { return inner.len() + 2 * n; }  // 57-59
  size_t strict_upper_bound_clamped(R element, size_t min, size_t max) const {  // 66-77
    size_t inner_min = map_into(min);
    size_t inner_max = map_into(max);
    return map_from(inner.strict_upper_bound_clamped(element, inner_min, inner_max));
  }
  size_t strict_upper_bound(R element) const {  // 78-84
    return map_from(inner.strict_upper_bound(element));
  }
};

// BorderDeletion<G> (bspline/adaptors.rs:94-146)
template <class G>
struct BorderDeletion : SortedChainDefaults<BorderDeletion<G>, typename G::Output> {
  using R = typename G::Output;
  using Output = R;
  G inner;
  explicit BorderDeletion(G g) : inner(g) {}
  R eval(size_t input) const { return inner.eval(input + 1); }  // 116-118
  size_t len() const { return inner.len() - 2; }                // 128-130
  size_t strict_upper_bound_clamped(R element, size_t min, size_t max) const {  // 137-145
    return inner.strict_upper_bound_clamped(element, min + 1, max + 1) - 1;
  }
};

// ---------------------------------------------------------------------------------------------
// Spaces (base/space.rs): constant = zeroed stack array, dynamic = fresh heap Vec per eval.
// ---------------------------------------------------------------------------------------------
template <class T, int N>
struct ConstSpace {
  using Work = std::array<T, N>;
  size_t len() const { return N; }
  Work workspace() const { return Work{}; }  // [Default::default(); N]  41-43
};
template <class T>
struct DynSpace {
  using Work = std::vector<T>;
  size_t len_;
  size_t len() const { return len_; }
  Work workspace() const { return Work(len_); }  // vec![Default::default(); len]  84-86
};

// ---------------------------------------------------------------------------------------------
// Curves
// ---------------------------------------------------------------------------------------------
// Built-in easing functions (easing/mod.rs:67-72 Identity, 86-132 flip/smoothstart/smoothend/smoothstep/
// smootherstep; easing/plateau.rs:10-57 Plateau). One runtime-tagged struct stands for the `F` parameter.
template <class R>
struct Easing {
  int kind = 0;  // ENTERP_EASE_*: 0 identity, 1 flip, 2 smoothstart<N>, 3 smoothend<N>, 4 smoothstep, 5 smootherstep, 6 plateau
  int n = 1;
  R min = R(0), max = R(1);  // Plateau
  static Easing plateau(R strength) {  // plateau.rs:19-25
    Easing e;
    e.kind = 6;
    const R halfed = strength / from_usize<R>(2);
    e.min = R(0) + halfed;
    e.max = R(1) - halfed;
    return e;
  }
  static R flip(R x) { return R(1) - x; }                       // mod.rs:86-91
  static R smoothstart(R x, int n) {                            // mod.rs:93-102
    R mul = x;
    for (int i = 1; i < n; ++i) mul = mul * x;
    return mul;
  }
  static R smoothstep(R x) { return x * x * (from_usize<R>(3) - from_usize<R>(2) * x); }  // mod.rs:114-121
  static R smootherstep(R x) {                                  // mod.rs:123-132
    return x * x * x * (x * (x * from_usize<R>(6) - from_usize<R>(15)) + from_usize<R>(10));
  }
  static R over_clamp(R input, R min, R max) {                  // plateau.rs:27-38
    if (input < min) return R(0);
    if (input > max) return R(1);
    return (input - min) / (max - min);
  }
  R eval(R x) const {
    switch (kind) {
      case 1: return flip(x);
      case 2: return smoothstart(x, n);
      case 3: return flip(smoothstart(flip(x), n));             // mod.rs:104-112
      case 4: return smoothstep(x);
      case 5: return smootherstep(x);
      case 6: return smoothstep(over_clamp(x, min, max));       // plateau.rs:45-47
      default: return x;
    }
  }
};

// linear/mod.rs:122-128 (eval), 139-141 (domain).
template <class K, class E>
struct Linear {
  using R = typename K::Output;
  using Output = typename E::Output;
  K knots;
  E elements;
  Easing<R> easing;
  Output eval(R scalar) const {
    Border<R> b = knots.upper_border(scalar);
    Output min_point = elements.eval(b.min_index);
    Output max_point = elements.eval(b.max_index);
    return merge(min_point, max_point, easing.eval(b.factor));
  }
  void domain(R out[2]) const {
    out[0] = knots.first();
    out[1] = knots.last();
  }
  // parity/debug view of the span decision
  Border<R> border(R scalar) const { return knots.upper_border(scalar); }
};

// bezier/mod.rs:44-56 (triangle_folding_inline), 78-91 (bezier), 218-229 (workspace), 240-246
// (eval), 257-259 (domain).
template <class R_, class E, class S>
struct Bezier {
  using R = R_;
  using Output = typename E::Output;
  E elements;
  S space;
  Output eval(R scalar) const {
    auto ws = space.workspace();
    const size_t len = elements.len();
    for (size_t i = 0; i < len; ++i) ws[i] = elements.eval(i);
    for (size_t k = 1; k <= len - 1; ++k)
      for (size_t i = 0; i < len - k; ++i) ws[i] = merge(ws[i], ws[i + 1], scalar);
    return ws[0];
  }
  void domain(R out[2]) const {
    out[0] = R(0);
    out[1] = R(1);
  }
};

// bezier/mod.rs:58-70 (lower_triangle_folding_inline), 96-116 (bezier_with_tangent), 122-153
// (bezier_with_deriatives) — restated as written upstream, including that the highest derivative
// (k == deg) is taken from the still-zeroed `grad` and therefore always comes out as zero, and that
// K <= len - 1 indexes out of bounds (a panic upstream, `Panic` here).
template <class R, class T>
inline void bezier_with_tangent(std::vector<T>& ws, R x, T out[2]) {
  const size_t len = ws.size();
  if (len < 2) {
    out[0] = ws[0];
    out[1] = ws[0] * R(0);
    return;
  }
  for (size_t k = 1; k <= len - 2; ++k)
    for (size_t i = 0; i < len - k; ++i) ws[i] = merge(ws[i], ws[i + 1], x);
  const T tangent = (ws[1] - ws[0]) * from_usize<R>(len - 1);
  out[0] = merge(ws[0], ws[1], x);
  out[1] = tangent;
}
template <class R, class T>
inline void bezier_with_derivatives(std::vector<T>& ws, R x, size_t K, T* grad) {
  const size_t len = ws.size();
  const size_t deg = std::min(K, len - 1);
  for (size_t k = 1; k <= len - deg - 1; ++k)
    for (size_t i = 0; i < len - k; ++i) ws[i] = merge(ws[i], ws[i + 1], x);
  for (size_t i = 0; i < K; ++i) grad[i] = ws[0] * R(0);
  for (size_t k = deg; k >= 1; --k) {
    if (k >= K) throw Panic{};  // `&mut grad[..=k]` out of range
    for (size_t s = 1; s <= k; ++s)          // lower_triangle_folding_inline(grad[..=k], second - first, k)
      for (size_t i = s; i <= k; ++i) grad[i] = grad[i] - grad[i - 1];
    size_t prod = 1;
    for (size_t q = len - k; q < len; ++q) prod *= q;
    grad[k] = grad[k] * from_usize<R>(prod);
    for (size_t i = 0; i < len - 1; ++i) ws[i] = merge(ws[i], ws[i + 1], x);  // one normal folding step
    for (size_t i = 0; i < k; ++i) grad[i] = ws[i];
  }
}

// bspline/mod.rs:126-133 (workspace), 145-169 (eval), 180-185 (domain), 254 (degree).
template <class K, class E, class S>
struct BSpline {
  using R = typename K::Output;
  using Output = typename E::Output;
  K knots;
  E elements;
  S space;
  size_t degree;  // knots.len() - elements.len() + 1
  size_t index_of(R scalar) const {
    return knots.strict_upper_bound_clamped(scalar, degree, knots.len() - degree);
  }
  Output eval(R scalar) const {
    const size_t index = index_of(scalar);
    // `index - degree + i` underflows when the search returns an index below its lower cut (Equidistant with a negative
    // step, list.rs:486-487): the element access at bspline/mod.rs:131 is then out of bounds, a panic upstream
    if (index < degree) throw Panic{};
    auto ws = space.workspace();
    for (size_t i = 0; i <= degree; ++i) ws[i] = elements.eval(index - degree + i);
    for (size_t r = 1; r <= degree; ++r) {
      for (size_t j = 0; j <= degree - r; ++j) {
        const size_t i = j + r + index - degree;
        const R factor =
            (scalar - knots.eval(i - 1)) / (knots.eval(i + degree - r) - knots.eval(i - 1));
        ws[j] = merge(ws[j], ws[j + 1], factor);
      }
    }
    return ws[0];
  }
  void domain(R out[2]) const {
    out[0] = knots.eval(degree - 1);
    out[1] = knots.eval(knots.len() - degree);
  }
};

// weights/weighted.rs:28-37 (+ Curve 40-48)
template <class G>
struct Weighted {
  using R = typename G::R;
  G inner;
  auto eval(R x) const { return inner.eval(x).project(); }
  void domain(R out[2]) const { inner.domain(out); }
};

// base/adaptors.rs:113-167: eval(input * multiplication + addition); normalized_to_domain 137-139
// sets addition = -start, multiplication = 1/(end-start); domain 161-166.
template <class G>
struct TransformInput {
  using R = typename G::R;
  G inner;
  R addition, multiplication;
  static TransformInput normalized_to_domain(G g, R start, R end) {
    return TransformInput{g, -start, R(1) / (end - start)};
  }
  auto eval(R x) const { return inner.eval(x * multiplication + addition); }
  void domain(R out[2]) const {
    R orig[2];
    inner.domain(orig);
    out[0] = (orig[0] - addition) / multiplication;
    out[1] = (orig[1] - addition) / multiplication;
  }
};

// Stepper / Curve::take (base/signal.rs:221-228, 607-609; Equidistant::new list.rs:393-399).
template <class R>
struct Stepper {
  Equidistant<R> eq;
  Stepper(size_t steps, R start, R end) : eq(Equidistant<R>::make(steps, start, end)) {}
  R at(size_t i) const { return eq.eval(i); }
};

}  // namespace ora
