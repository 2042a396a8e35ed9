// C++ mirror smoke test: builder errors on any box; evaluation when a CUDA device is present.
// Mirrors reference tests: bspline/builder.rs clamped_errors, linear/mod.rs extrapolation, README.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/enterpolation.hpp"

#define EXPECT(cond)                                                  \
  do {                                                                \
    if (!(cond)) {                                                    \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond);      \
      return 1;                                                       \
    }                                                                 \
  } while (0)

template <class F>
static int status_of(F&& f) {
  try {
    f();
  } catch (const enterp::Error& e) {
    return (int)e.status;
  }
  return 0;
}

int main() {
  using namespace enterp;
  // Director semantics: fails AT the offending call (bspline/builder.rs:1423-1509)
  EXPECT(status_of([] { BSpline::director<float>().clamped().elements({0.0f}); }) == ENTERP_TOO_FEW_ELEMENTS);
  EXPECT(status_of([] { BSpline::director<float>().clamped().elements({0.f, 1.f, 2.f, 3.f}).knots({0.f}); }) ==
         ENTERP_TOO_FEW_KNOTS);
  EXPECT(status_of([] { BSpline::director<float>().clamped().elements({0.f, 1.f, 2.f, 3.f}).equidistant().degree(0); }) ==
         ENTERP_INVALID_DEGREE);
  EXPECT(status_of([] {
           BSpline::director<float>().clamped().elements({0.f, 1.f, 2.f, 3.f}).equidistant().degree(2).domain(0.f, 1.f).constant(2);
         }) == ENTERP_TOO_SMALL_WORKSPACE);
  EXPECT(status_of([] { Linear::director<double>().elements({1.0, 2.0}).knots({1.0, 2.0, 3.0}); }) ==
         ENTERP_KNOT_ELEMENT_INEQUALITY);
  EXPECT(status_of([] { Bezier::director<double>().elements(std::vector<double>{}, 1); }) == ENTERP_EMPTY);
  // Builder semantics: first error surfaces at build()
  {
    auto b = BSpline::builder<double>().elements({0.0}).knots({0.0}).constant(0);
    EXPECT(status_of([&] { b.build(); }) == ENTERP_TOO_FEW_ELEMENTS);
  }
  auto spline = BSpline::builder<double>().clamped().elements({0, 0, 0, 6, 0, 0, 0}).equidistant().degree(3).domain(-2.0, 2.0)
                    .constant(4).build();
  EXPECT(spline.degree() == 3 && spline.domain()[0] == -2.0 && spline.domain()[1] == 2.0);
  if (enterp_device_count() == 0) {
    bool loud = false;
    try {
      spline.take(4).collect();
    } catch (const DeviceError&) {
      loud = true;  // no CPU fallback
    }
    EXPECT(loud);
    std::printf("OK (no device: builder + loud-failure checks only)\n");
    return 0;
  }
  // README.md:52-79 three constructions agree; SURVEY appendix A values
  auto y = spline.take(10).collect();
  auto sorted = BSpline::builder<double>().elements({0, 0, 0, 6, 0, 0, 0}).knots({-2, -2, -2, -1, 0, 1, 2, 2, 2}).constant(4).build();
  auto y2 = sorted.take(10).collect();
  for (int i = 0; i < 10; ++i) EXPECT(std::fabs(y[i] - y2[i]) <= 4 * 2.3e-16 * (1 + std::fabs(y[i])));
  EXPECT(y[3] == 2.2222222222222223 && y[0] == 0.0 && y[9] == 0.0);
  // linear/mod.rs:256-267 extrapolation
  auto lin = Linear::builder<double>().elements({20.0, 100.0, 0.0, 200.0}).knots({1.0, 2.0, 3.0, 4.0}).build();
  auto v = lin.sample({1.5, 2.5, -1.0, 5.0});
  EXPECT(v[0] == 60.0 && v[1] == 50.0 && v[2] == -140.0 && v[3] == 400.0);
  // signal.rs:271-279 clamp doctest and 241-249 slice doctest
  auto l2 = Linear::builder<double>().elements({0.0, 3.0}).knots({0.0, 1.0}).build();
  auto cl = l2.clamp().sample({-1.0, 0.0, 0.5, 1.0, 2.0});
  EXPECT(cl[0] == 0.0 && cl[2] == 1.5 && cl[4] == 3.0);
  auto l3 = Linear::builder<double>().elements({0.0, 5.0, 3.0}).knots({0.0, 1.0, 2.0}).build();
  auto sl = l3.slice(0.5, 1.5).take(3).collect();
  EXPECT(sl[0] == 2.5 && sl[1] == 5.0 && sl[2] == 4.0);
  // easing + ConstEquidistantLinear (linear/mod.rs:281-291)
  auto ce = Linear::builder<double>().elements({20.0, 100.0, 0.0, 200.0}).const_equidistant().build().take(7).collect();
  EXPECT(ce[0] == 20.0 && ce[2] == 100.0 && ce[6] == 200.0);
  auto sm = Linear::builder<double>().elements({0.0, 1.0}).knots({0.0, 1.0}).easing(ENTERP_EASE_SMOOTHSTEP).build().sample({0.5});
  EXPECT(sm[0] == 0.5);
  // shards of one take agree with the whole
  auto whole = spline.take(1000).collect();
  auto part = spline.take(1000).shard(1, 3).collect();
  EXPECT(std::memcmp(part.data(), whole.data() + 333, part.size() * sizeof(double)) == 0);
  std::printf("OK (device path)\n");
  return 0;
}
