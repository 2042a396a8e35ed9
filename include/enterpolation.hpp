// enterpolation.hpp — header-only C++17 mirror of enterpolation's builder + Signal/Curve surface on
// top of the C ABI (enterp.h). The reference is Rust; no Rust toolchain exists in this image, so the
// host side above the ABI is C++ (and Python, enterpolation_b200/). Method names, call order and error
// taxonomy follow the reference:
//
//   auto spline = enterp::BSpline::builder()          // src/bspline/mod.rs:115
//                     .clamped()                       // src/bspline/builder.rs:177-209
//                     .elements<double>({0,0,0,6,0,0,0})
//                     .equidistant().degree(3).domain(-2.0, 2.0)
//                     .constant(4)
//                     .build();                        // throws enterp::Error (LinearError/BezierError/BSplineError)
//   std::vector<double> y = spline.take(1 << 20).collect();      // Curve::take  src/base/signal.rs:221-228
//   auto z = spline.sample(std::vector<double>{-1.5, 0.0});       // Signal::sample  signal.rs:166-173
//
// `builder()` defers the first error to build() (the *Builder types); pass `enterp::immediate` to get
// the *Director behaviour (throw at the offending call). All rules are checked by the library
// (enterp_curve_validate / enterp_curve_create); this header only records the chain.
#pragma once
#include <array>
#include <cstdint>
#include <initializer_list>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "enterp.h"

namespace enterp {

struct immediate_t {};
inline constexpr immediate_t immediate{};

// Error carrying the reference's error struct payload.
class Error : public std::runtime_error {
 public:
  Error(enterp_status s, const enterp_error_info& i)
      : std::runtime_error(std::string(enterp_status_str(s))), status(s), info(i) {}
  enterp_status status;
  enterp_error_info info;
  bool is_too_few_elements() const { return status == ENTERP_TOO_FEW_ELEMENTS; }
  bool is_not_sorted() const { return status == ENTERP_NOT_SORTED; }
};
class DeviceError : public std::runtime_error {
 public:
  explicit DeviceError(enterp_status s)
      : std::runtime_error(std::string(enterp_status_str(s)) + ": " + enterp_last_cuda_error()), status(s) {}
  enterp_status status;
};

template <class R>
constexpr int scalar_of() {
  static_assert(std::is_same<R, float>::value || std::is_same<R, double>::value, "R must be f32 or f64");
  return std::is_same<R, float>::value ? ENTERP_F32 : ENTERP_F64;
}

template <class R>
class Curve;

// `Take<C, R>` (signal.rs:527-576)
template <class R>
class Take {
 public:
  Take(const Curve<R>* c, uint64_t samples, uint64_t first, uint64_t count)
      : curve_(c), samples_(samples), first_(first), count_(count) {}
  uint64_t size() const { return count_; }
  // contiguous global range of rank `rank` of `world` (the stepper keeps the global index)
  Take shard(int rank, int world) const {
    const uint64_t lo = samples_ * (uint64_t)rank / (uint64_t)world, hi = samples_ * (uint64_t)(rank + 1) / (uint64_t)world;
    return Take(curve_, samples_, lo, hi - lo);
  }
  std::vector<R> collect(int device = 0) const;                       // host result, [count][dim]
  void collect_device(R* out_dev, int device = 0, void* stream = nullptr) const;

 private:
  const Curve<R>* curve_;
  uint64_t samples_, first_, count_;
};

// A built curve: `impl Signal<R> + Curve<R>`.
template <class R>
class Curve {
 public:
  Curve() = default;
  explicit Curve(enterp_curve* h) : h_(h) { enterp_curve_get_info(h_, &info_); }
  Curve(Curve&& o) noexcept : h_(o.h_), info_(o.info_) { o.h_ = nullptr; }
  Curve& operator=(Curve&& o) noexcept {
    if (this != &o) {
      reset();
      h_ = o.h_;
      info_ = o.info_;
      o.h_ = nullptr;
    }
    return *this;
  }
  Curve(const Curve&) = delete;
  Curve& operator=(const Curve&) = delete;
  ~Curve() { reset(); }

  std::array<R, 2> domain() const { return {(R)info_.domain[0], (R)info_.domain[1]}; }  // Curve::domain
  int dim() const { return info_.dim; }
  uint64_t degree() const { return info_.degree; }
  const enterp_curve_info& info() const { return info_; }
  enterp_curve* handle() const { return h_; }

  Take<R> take(uint64_t samples) const { return Take<R>(this, samples, 0, samples); }  // Curve::take
  // Curve::clamp() (signal.rs:271-290) and Curve::slice(start..end) (signal.rs:241-262): derived handles that
  // share this curve's device tables.
  Curve clamp() const {
    enterp_curve* h = nullptr;
    check(enterp_curve_clamp(h_, &h));
    return Curve(h);
  }
  Curve slice(R start, R end) const {
    enterp_curve* h = nullptr;
    check(enterp_curve_slice(h_, 1, (double)start, 1, (double)end, &h));
    return Curve(h);
  }
  Curve slice_from(R start) const {  // start..
    enterp_curve* h = nullptr;
    check(enterp_curve_slice(h_, 1, (double)start, 0, 0.0, &h));
    return Curve(h);
  }
  Curve slice_to(R end) const {  // ..end
    enterp_curve* h = nullptr;
    check(enterp_curve_slice(h_, 0, 0.0, 1, (double)end, &h));
    return Curve(h);
  }
  std::vector<R> sample(const std::vector<R>& t, int device = 0) const {                // Signal::sample
    std::vector<R> out(t.size() * (size_t)info_.dim);
    check(enterp_sample_host(h_, device, t.data(), t.size(), out.data()));
    return out;
  }
  std::vector<R> eval(R t, int device = 0) const { return sample(std::vector<R>{t}, device); }  // Signal::eval
  void sample_device(const R* t_dev, uint64_t n, R* out_dev, int device = 0, void* stream = nullptr) const {
    check(enterp_eval_batch(h_, device, t_dev, n, out_dev, stream));
  }
  unsigned long long invalid_count(int device = 0) const {
    unsigned long long n = 0;
    check(enterp_invalid_count(h_, device, &n));
    return n;
  }
  static void check(enterp_status s) {
    if (s == ENTERP_OK) return;
    if (s == ENTERP_NO_DEVICE || s == ENTERP_CUDA_ERROR) throw DeviceError(s);
    enterp_error_info i{};
    i.status = s;
    throw Error(s, i);
  }

 private:
  void reset() {
    if (h_) enterp_curve_destroy(h_);
    h_ = nullptr;
  }
  enterp_curve* h_ = nullptr;
  enterp_curve_info info_{};
};

template <class R>
std::vector<R> Take<R>::collect(int device) const {
  std::vector<R> out(count_ * (size_t)curve_->dim());
  Curve<R>::check(enterp_take_host(curve_->handle(), device, samples_, first_, count_, out.data()));
  return out;
}
template <class R>
void Take<R>::collect_device(R* out_dev, int device, void* stream) const {
  Curve<R>::check(enterp_eval_take(curve_->handle(), device, samples_, first_, count_, out_dev, stream));
}

namespace detail {

// Records one builder chain into an enterp_curve_desc.
template <class R>
struct Chain {
  enterp_curve_desc d{};
  std::vector<R> elements, weights, knots;
  uint64_t const_knots = 0;  // N of ConstEquidistant<R, N> knots given to BSpline::new
  bool immediate = false;
  bool failed = false;
  enterp_status first_status = ENTERP_OK;
  enterp_error_info first_info{};

  void sync() {
    d.scalar = scalar_of<R>();
    d.n_elements = d.dim > 0 ? elements.size() / (size_t)d.dim : 0;
    d.elements = elements.empty() ? nullptr : elements.data();
    d.weights = d.weighted ? (weights.empty() ? nullptr : weights.data()) : nullptr;
    d.n_knots = d.knot_base == ENTERP_KNOTS_CONST_EQUIDISTANT ? const_knots : knots.size();
    d.knots = knots.empty() ? nullptr : knots.data();
  }
  void validate(int stage) {
    if (failed) return;
    sync();
    enterp_error_info info{};
    const enterp_status s = enterp_curve_validate(&d, stage, &info);
    if (s != ENTERP_OK) {
      if (immediate) throw Error(s, info);
      failed = true;
      first_status = s;
      first_info = info;
    }
  }
  Curve<R> build() {
    if (failed) throw Error(first_status, first_info);
    sync();
    enterp_curve* h = nullptr;
    enterp_error_info info{};
    const enterp_status s = enterp_curve_create(&d, &h, &info);
    if (s != ENTERP_OK) throw Error(s, info);
    return Curve<R>(h);
  }
};

}  // namespace detail

// Common part of the three builders (CRTP so that chaining keeps the concrete type).
template <class Self, class R>
class BuilderBase {
 public:
  Self& self() { return static_cast<Self&>(*this); }
  // `.elements(E)`: n elements of `dim` components each, AoS
  Self& elements(std::vector<R> flat, int dim = 1) {
    c_.d.dim = dim;
    c_.d.weighted = 0;
    c_.elements = std::move(flat);
    c_.validate(0);
    return self();
  }
  Self& elements(std::initializer_list<R> scalars) { return elements(std::vector<R>(scalars), 1); }
  // `.elements_with_weights(G)` given as elements.stack(weights) (signal.rs:71-78)
  Self& elements_with_weights(std::vector<R> flat, std::vector<R> weights, int dim = 1) {
    c_.d.dim = dim;
    c_.d.weighted = 1;
    c_.elements = std::move(flat);
    c_.weights = std::move(weights);
    c_.validate(0);
    return self();
  }
  Curve<R> build() { return c_.build(); }
  const enterp_curve_desc& desc() {
    c_.sync();
    return c_.d;
  }

 protected:
  detail::Chain<R> c_;
};

template <class R = double>
class LinearBuilder : public BuilderBase<LinearBuilder<R>, R> {
 public:
  explicit LinearBuilder(bool imm = false) {
    this->c_.d.kind = ENTERP_LINEAR;
    this->c_.immediate = imm;
  }
  LinearBuilder& knots(std::vector<R> k) {  // src/linear/builder.rs:325-340
    this->c_.d.knot_base = ENTERP_KNOTS_SORTED;
    this->c_.knots = std::move(k);
    this->c_.validate(1);
    return *this;
  }
  LinearBuilder& equidistant() {  // 362-369
    this->c_.d.knot_base = ENTERP_KNOTS_EQUIDISTANT;
    this->c_.d.eq_form = ENTERP_EQ_NORMALIZED;
    return *this;
  }
  // Linear::equidistant_unchecked (src/linear/mod.rs:193-207): ConstEquidistant knots i/(N-1) on [0,1]
  LinearBuilder& const_equidistant() {
    this->c_.d.knot_base = ENTERP_KNOTS_CONST_EQUIDISTANT;
    this->c_.validate(1);
    return *this;
  }
  // .easing(F) (src/linear/builder.rs:496-503) for the built-in easings of src/easing/
  LinearBuilder& easing(int ease_kind, int n = 0, R param = R(0)) {
    this->c_.d.easing = ease_kind;
    this->c_.d.easing_n = n;
    this->c_.d.easing_param = param;
    return *this;
  }
  LinearBuilder& domain(R a, R b) { return form(ENTERP_EQ_DOMAIN, a, b); }      // 427-453
  LinearBuilder& normalized() { return form(ENTERP_EQ_NORMALIZED, 0, 1); }
  LinearBuilder& distance(R start, R step) { return form(ENTERP_EQ_DISTANCE, start, step); }

 private:
  LinearBuilder& form(int f, R a, R b) {
    this->c_.d.eq_form = f;
    this->c_.d.eq_a = a;
    this->c_.d.eq_b = b;
    this->c_.validate(1);
    return *this;
  }
};

template <class R = double>
class BezierBuilder : public BuilderBase<BezierBuilder<R>, R> {
 public:
  explicit BezierBuilder(bool imm = false) {
    this->c_.d.kind = ENTERP_BEZIER;
    this->c_.d.space = ENTERP_SPACE_CONSTANT;
    this->c_.immediate = imm;
  }
  BezierBuilder& normalized() {  // src/bezier/builder.rs:273-289
    this->c_.d.input_form = ENTERP_INPUT_NORMALIZED;
    return *this;
  }
  BezierBuilder& domain(R a, R b) {
    this->c_.d.input_form = ENTERP_INPUT_DOMAIN;
    this->c_.d.in_a = a;
    this->c_.d.in_b = b;
    return *this;
  }
  BezierBuilder& dynamic() { return space(ENTERP_SPACE_DYNAMIC, 0); }          // 317-324
  BezierBuilder& constant() { return space(ENTERP_SPACE_CONSTANT, 0); }        // 330-340
  BezierBuilder& workspace(uint64_t n) { return space(ENTERP_SPACE_WORKSPACE, n); }  // 354-367

 private:
  BezierBuilder& space(int s, uint64_t n) {
    this->c_.d.space = s;
    this->c_.d.workspace_n = n;
    this->c_.validate(2);
    return *this;
  }
};

template <class R = double>
class BSplineBuilder : public BuilderBase<BSplineBuilder<R>, R> {
 public:
  explicit BSplineBuilder(bool imm = false) {
    this->c_.d.kind = ENTERP_BSPLINE;
    this->c_.d.mode = ENTERP_OPEN;
    this->c_.d.space = ENTERP_SPACE_DYNAMIC;
    this->c_.immediate = imm;
  }
  BSplineBuilder& open() { return mode(ENTERP_OPEN); }        // src/bspline/builder.rs:177-209
  BSplineBuilder& clamped() { return mode(ENTERP_CLAMPED); }
  BSplineBuilder& legacy() { return mode(ENTERP_LEGACY); }
  BSplineBuilder& knots(std::vector<R> k) {                   // 377-401 / 447-466 / 512-534
    this->c_.d.knot_base = ENTERP_KNOTS_SORTED;
    this->c_.knots = std::move(k);
    this->c_.validate(1);
    return *this;
  }
  // BSpline::new(elements, ConstEquidistant::<R, N>::new(), space) (src/bspline/mod.rs:212-235 over
  // src/base/list.rs:566-721): N knots i/(N-1) on [0,1]; open mode only, the constructor's own checks apply
  BSplineBuilder& const_equidistant(uint64_t n_knots) {
    this->c_.d.knot_base = ENTERP_KNOTS_CONST_EQUIDISTANT;
    this->c_.const_knots = n_knots;
    this->c_.validate(1);
    return *this;
  }
  BSplineBuilder& equidistant() {                             // 576-583
    this->c_.d.knot_base = ENTERP_KNOTS_EQUIDISTANT;
    this->c_.d.eq_form = ENTERP_EQ_NORMALIZED;
    return *this;
  }
  BSplineBuilder& degree(uint64_t deg) { return by(ENTERP_EQ_BY_DEGREE, deg); }      // 633-655 / 780-798
  BSplineBuilder& quantity(uint64_t q) { return by(ENTERP_EQ_BY_QUANTITY, q); }      // 673-692 / 820-836
  BSplineBuilder& domain(R a, R b) { return form(ENTERP_EQ_DOMAIN, a, b); }          // 903-933 / 972-1017
  BSplineBuilder& normalized() { return form(ENTERP_EQ_NORMALIZED, 0, 1); }
  BSplineBuilder& distance(R start, R step) { return form(ENTERP_EQ_DISTANCE, start, step); }
  BSplineBuilder& dynamic() { return space(ENTERP_SPACE_DYNAMIC, 0); }               // 1071-1078
  BSplineBuilder& constant(uint64_t n) { return space(ENTERP_SPACE_CONSTANT, n); }   // 1087-1103
  BSplineBuilder& workspace(uint64_t n) { return space(ENTERP_SPACE_WORKSPACE, n); } // 1114-1130

 private:
  BSplineBuilder& mode(int m) {
    this->c_.d.mode = m;
    return *this;
  }
  BSplineBuilder& by(int b, uint64_t n) {
    this->c_.d.eq_by = b;
    this->c_.d.eq_count = n;
    this->c_.validate(1);
    return *this;
  }
  BSplineBuilder& form(int f, R a, R b) {
    this->c_.d.eq_form = f;
    this->c_.d.eq_a = a;
    this->c_.d.eq_b = b;
    this->c_.validate(1);
    return *this;
  }
  BSplineBuilder& space(int s, uint64_t n) {
    this->c_.d.space = s;
    this->c_.d.workspace_n = n;
    this->c_.validate(2);
    return *this;
  }
};

struct Linear {
  template <class R = double>
  static LinearBuilder<R> builder() { return LinearBuilder<R>(false); }
  template <class R = double>
  static LinearBuilder<R> director() { return LinearBuilder<R>(true); }
};
struct Bezier {
  template <class R = double>
  static BezierBuilder<R> builder() { return BezierBuilder<R>(false); }
  template <class R = double>
  static BezierBuilder<R> director() { return BezierBuilder<R>(true); }
};
struct BSpline {
  template <class R = double>
  static BSplineBuilder<R> builder() { return BSplineBuilder<R>(false); }
  template <class R = double>
  static BSplineBuilder<R> director() { return BSplineBuilder<R>(true); }
};

}  // namespace enterp
