"""Known-answer tests transcribed from the reference's own unit tests, doctests and examples
(SURVEY.md §8c). Every entry cites its source under /root/reference. The same catalogue is replayed
through the CPU oracle (tests/test_oracle_kat.py, no GPU) and through the CUDA path
(tests/test_gpu_kat.py, -m gpu).

An entry is a dict:
  name     unique id
  src      reference file:line
  make     () -> builder chain (enterpolation_b200 builders; `.to_desc()` feeds the oracle)
  take     n          -> compare curve.take(n) with `expect`
  sample   [t...]     -> compare curve.sample(t) with `expect`
  expect   values (or rows)
  ulps     tolerance in representable steps (4 = assert_f*_near! default, 0 = assert_eq!)
  abs      optional absolute tolerance instead (assert_float_absolute_eq!)
"""
import math

import numpy as np

from enterpolation_b200 import Bezier, BSpline, Linear

F32 = np.float32
INF = float("inf")


def _f32(xs):
    return np.asarray(xs, dtype=np.float32)


KATS = []


def kat(**kw):
    KATS.append(kw)


# ---- B-spline values -------------------------------------------------------------------------
kat(name="bspline_linear_f32", src="src/bspline/mod.rs:269-290",
    make=lambda: BSpline.builder().elements(_f32([0.0, 1.0])).knots(_f32([0.0, 1.0])).constant(2),
    sample=_f32([-1.0, 0.0, 0.2, 0.4, 0.6, 0.8, 1.0, 2.0]),
    expect=_f32([-1.0, 0.0, 0.2, 0.4, 0.6, 0.8, 1.0, 2.0]), ulps=4)
kat(name="bspline_quadratic_f32", src="src/bspline/mod.rs:293-316",
    make=lambda: BSpline.builder().elements(_f32([0.0, 0.0, 1.0, 0.0, 0.0]))
    .knots(_f32([0.0, 0.0, 1.0, 2.0, 3.0, 3.0])).constant(3),
    sample=_f32([0.0, 0.5, 1.0, 1.4, 1.5, 1.6, 2.0, 2.5, 3.0]),
    expect=_f32([0.0, 0.125, 0.5, 0.74, 0.75, 0.74, 0.5, 0.125, 0.0]), ulps=4)
kat(name="bspline_cubic_f32", src="src/bspline/mod.rs:319-343",
    make=lambda: BSpline.builder().elements(_f32([0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0]))
    .knots(_f32([-2.0, -2.0, -2.0, -1.0, 0.0, 1.0, 2.0, 2.0, 2.0])).constant(4),
    sample=_f32([-2.0, -1.5, -1.0, -0.6, 0.0, 0.5, 1.5, 2.0]),
    expect=_f32([0.0, 0.125, 1.0, 2.488, 4.0, 2.875, 0.12500001, 0.0]), ulps=4)
_QUARTIC_PTS = [0.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0]
_QUARTIC_KNOTS = [0.0, 0.0, 0.0, 0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 5.0, 5.0, 5.0]
_QUARTIC_T = [0.0, 0.4, 1.0, 1.5, 2.0, 2.5, 3.0, 3.2, 4.1, 4.5, 5.0]
kat(name="bspline_quartic_f32", src="src/bspline/mod.rs:346-370",
    make=lambda: BSpline.builder().elements(_f32(_QUARTIC_PTS)).knots(_f32(_QUARTIC_KNOTS)).constant(5),
    sample=_f32(_QUARTIC_T),
    expect=_f32([0.0, 0.0010666668, 0.041666668, 0.19791667, 0.4583333, 0.5989583, 0.4583333, 0.35206667,
                 0.02733751, 0.002604167, 0.0]), ulps=4)
kat(name="bspline_quartic_f64", src="src/bspline/mod.rs:373-398",
    make=lambda: BSpline.builder().elements(_QUARTIC_PTS).knots(_QUARTIC_KNOTS).constant(5),
    sample=_QUARTIC_T,
    expect=[0.0, 0.001066666666666667, 0.041666666666666664, 0.19791666666666666, 0.45833333333333337,
            0.5989583333333334, 0.4583333333333333, 0.3520666666666666, 0.027337500000000046,
            0.002604166666666666, 0.0], ulps=4)
kat(name="bspline_clamped_equidistant_doc", src="src/bspline/mod.rs:13-24",
    make=lambda: BSpline.builder().clamped().elements([0.0, 5.0, 3.0, 10.0, 7.0]).equidistant("f64").degree(3)
    .normalized().constant(4),
    take=11, expect=[0.0, 2.346, 3.648, 4.302, 4.704, 5.25, 6.2, 7.27, 8.04, 8.09, 7.0], ulps=4)
kat(name="bspline_builder_doc", src="src/bspline/mod.rs:86-99",
    make=lambda: BSpline.builder().clamped().elements([20.0, 100.0, 0.0, 200.0]).equidistant("f64").degree(3)
    .normalized().constant(4),
    take=5, expect=[20.0, 53.75, 65.0, 98.75, 200.0], ulps=4)

# ---- Linear values ---------------------------------------------------------------------------
_LIN7 = [20.0, 60.0, 100.0, 50.0, 0.0, 100.0, 200.0]
kat(name="linear_equidistant", src="src/linear/mod.rs:226-240",
    make=lambda: Linear.builder().elements([20.0, 100.0, 0.0, 200.0]).equidistant("f64").normalized(),
    take=7, expect=_LIN7, ulps=4)
kat(name="linear_sorted", src="src/linear/mod.rs:242-254",
    make=lambda: Linear.builder().elements([20.0, 100.0, 0.0, 200.0]).knots([0.0, 1.0 / 3.0, 2.0 / 3.0, 1.0]),
    take=7, expect=_LIN7, ulps=4)
kat(name="linear_extrapolation", src="src/linear/mod.rs:256-267",
    make=lambda: Linear.builder().elements([20.0, 100.0, 0.0, 200.0]).knots([1.0, 2.0, 3.0, 4.0]),
    sample=[1.5, 2.5, -1.0, 5.0], expect=[60.0, 50.0, -140.0, 400.0], ulps=4)
kat(name="linear_weights", src="src/linear/mod.rs:269-278",
    make=lambda: Linear.builder().elements_with_weights([(0.0, 9.0), (1.0, 1.0)]).equidistant("f64").normalized(),
    sample=[0.5], expect=[0.1], ulps=4)
kat(name="linear_borrow_creation", src="src/linear/mod.rs:294-307",
    make=lambda: Linear.builder().elements([20.0, 100.0, 0.0, 200.0]).knots([0.0, 1.0, 2.0, 3.0]),
    sample=[0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0], expect=_LIN7, ulps=4)
kat(name="linear_module_doc", src="src/linear/mod.rs:10-16 / 87-95",
    make=lambda: Linear.builder().elements([0.0, 5.0, 3.0]).equidistant("f64").normalized(),
    take=5, expect=[0.0, 2.5, 5.0, 4.0, 3.0], ulps=4)
kat(name="linear_director_doc", src="src/linear/builder.rs:31-38",
    make=lambda: Linear.director().elements([1.0, 5.0, 100.0]).equidistant("f64").normalized(),
    take=5, expect=[1.0, 3.0, 5.0, 52.5, 100.0], ulps=4)
kat(name="linear_weighted_builder_doc", src="src/linear/builder.rs:200-208",
    make=lambda: Linear.builder().elements_with_weights([(1.0, 1.0), (2.0, 4.0), (3.0, 0.0)])
    .equidistant("f64").normalized(),
    take=5, expect=[1.0, 1.8, 2.0, 2.75, INF], ulps=4)
kat(name="signal_extract_doc", src="src/base/signal.rs:33-41 / 150-158",
    make=lambda: Linear.builder().elements([0.0, 3.0]).knots([0.0, 1.0]),
    sample=[0.0, 0.2, 0.4, 0.5, 0.55, 1.0], expect=[0.0, 0.6, 1.2, 1.5, 1.65, 3.0], ulps=4)
kat(name="signal_stack_weights_doc", src="src/base/signal.rs:71-78",
    make=lambda: Linear.builder().elements_with_weights([1.0, 5.0, 3.0], [1.0, 3.0, 2.0]).knots([0.0, 1.0, 2.0]),
    sample=[0.5], expect=[4.0], ulps=4)
kat(name="curve_take_doc", src="src/base/signal.rs:205-212",
    make=lambda: Linear.builder().elements([0.0, 5.0, 3.0]).knots([0.0, 1.0, 2.0]),
    take=11, expect=[0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 4.6, 4.2, 3.8, 3.4, 3.0], ulps=4)
kat(name="readme_linear", src="README.md:27-43 (printed values not asserted upstream; endpoints only)",
    make=lambda: Linear.builder().elements([0.0, 5.0, -5.0, 0.0]).knots([0.0, 0.2, 0.8, 1.0]),
    sample=[0.0, 0.2, 0.8, 1.0], expect=[0.0, 5.0, -5.0, 0.0], ulps=4)
# examples/linear.rs:6-21 — R's approx() cross-check, 17-digit printed output
kat(name="example_linear_approx", src="examples/linear.rs:9-21",
    make=lambda: Linear.builder().elements([0.0, 1.0, 2.0, 3.0]).knots([0.1, 0.2, 1.0, 10.0]),
    sample=[float(x) for x in range(1, 11)],
    expect=[2.0, 2.111111111111111, 2.2222222222222223, 2.3333333333333335, 2.4444444444444446,
            2.5555555555555554, 2.666666666666667, 2.7777777777777777, 2.888888888888889, 3.0], ulps=0)

kat(name="linear_const_creation", src="src/linear/mod.rs:281-291 (ConstEquidistantLinear)",
    make=lambda: Linear.equidistant_unchecked([20.0, 100.0, 0.0, 200.0], "f64"),
    take=7, expect=_LIN7, ulps=4)

# ---- Bezier values ---------------------------------------------------------------------------
kat(name="bezier_module_doc", src="src/bezier/mod.rs:11-19",
    make=lambda: Bezier.builder().elements([0.0, 5.0, 3.0]).normalized("f64").constant(3),
    take=3, expect=[0.0, 3.25, 3.0], ulps=4)
kat(name="bezier_builder_doc", src="src/bezier/mod.rs:184-194",
    make=lambda: Bezier.builder().elements([20.0, 100.0, 0.0, 200.0]).normalized("f64").constant(),
    take=5, expect=[20.0, 53.75, 65.0, 98.75, 200.0], ulps=4)
kat(name="bezier_extrapolation", src="src/bezier/mod.rs:333-343",
    make=lambda: Bezier.builder().elements([20.0, 0.0, 200.0]).normalized("f64").constant(),
    sample=[2.0, -1.0], expect=[820.0, 280.0], ulps=4)
kat(name="bezier_elements_doc_exact", src="src/bezier/builder.rs:198-206",
    make=lambda: Bezier.builder().elements([1.0, 2.0]).normalized("f64").constant(),
    take=3, expect=[1.0, 1.5, 2.0], ulps=0)
kat(name="bezier_weighted_doc_exact", src="src/bezier/builder.rs:237-245",
    make=lambda: Bezier.builder().elements_with_weights([(1.0, 1.0), (2.0, 4.0), (3.0, 0.0)])
    .normalized("f64").constant(),
    take=5, expect=[1.0, 15.0 / 8.25, 10.0 / 4.5, 19.0 / 6.25, INF], ulps=0)

# ---- NURBS unit circle (examples/nurbs.rs:87-118) ----------------------------------------------
_W = math.sqrt(2.0) / 2.0
_CIRCLE_PTS = [(1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (-1.0, 1.0), (-1.0, 0.0), (-1.0, -1.0), (0.0, -1.0), (1.0, -1.0),
               (1.0, 0.0)]
_CIRCLE_W = [1.0, _W, 1.0, _W, 1.0, _W, 1.0, _W, 1.0]
_CIRCLE_KNOTS = [0.0, 0.0, 1.0, 1.0, 2.0, 2.0, 3.0, 3.0, 4.0, 4.0]


def nurbs_circle():
    return BSpline.builder().elements_with_weights(_CIRCLE_PTS, _CIRCLE_W).knots(_CIRCLE_KNOTS).constant(3)


kat(name="nurbs_circle_quadrants", src="examples/nurbs.rs:112-117 (assert_float_absolute_eq!, 1e-6)",
    make=nurbs_circle, sample=[0.0, 1.0, 2.0, 3.0, 4.0],
    expect=[[math.cos(v * 0.5 * math.pi), math.sin(v * 0.5 * math.pi)] for v in range(5)], ulps=None, abs=1e-6)


# ---- adaptors and easings ------------------------------------------------
def _lin_012():
    return Linear.builder().elements([0.0, 5.0, 3.0]).knots([0.0, 1.0, 2.0])


kat(name="curve_slice_doc", src="src/base/signal.rs:241-249",
    make=_lin_012, adapt=lambda c: c.slice(0.5, 1.5), take=3, expect=[2.5, 5.0, 4.0], ulps=4)
kat(name="curve_clamp_doc", src="src/base/signal.rs:271-279",
    make=lambda: Linear.builder().elements([0.0, 3.0]).knots([0.0, 1.0]), adapt=lambda c: c.clamp(),
    sample=[-1.0, 0.0, 0.5, 1.0, 2.0], expect=[0.0, 0.0, 1.5, 3.0, 3.0], ulps=4)
# adaptors.rs:359-368 slices the Identity curve over [0,1] to 10..100; Identity is the degree-1 curve 0 -> 1
kat(name="adaptors_slice_test", src="src/base/adaptors.rs:359-368",
    make=lambda: Linear.builder().elements([0.0, 1.0]).knots([0.0, 1.0]), adapt=lambda c: c.slice(10.0, 100.0),
    sample=[0.0, 1.0], expect=[10.0, 100.0], ulps=4)


# adaptors.rs:338-356 wraps the Identity curve in TransformInput::new(identity, 0.0, 2.0); over [0,1] (and beyond: both
# extrapolate linearly) Identity is the two-point Bezier 0 -> 1, and a stored TransformInput is what the serde layout
# hands to the library verbatim (ENTERP_INPUT_TRANSFORM). The take() half pins TransformInput::domain (adaptors.rs:158-167)
def _identity_times_two():
    from enterpolation_b200 import serde
    return serde.loads("bezier", '{"addition":0.0,"multiplication":2.0,"inner":'
                       '{"elements":[0.0,1.0],"space":{"_phantom":null},"_input":null}}', "f64")


kat(name="adaptors_input_transform_extract", src="src/base/adaptors.rs:338-347", make=_identity_times_two,
    sample=[1.0, 0.0, 0.5, 1.0], expect=[2.0, 0.0, 1.0, 2.0], ulps=4)
kat(name="adaptors_input_transform_take", src="src/base/adaptors.rs:348-355", make=_identity_times_two,
    take=3, expect=[0.0, 0.5, 1.0], ulps=4)


def all_kats():
    return KATS
