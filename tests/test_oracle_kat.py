"""Pins the CPU oracle (oracle/enterp_oracle.hpp) against every literal known-answer test the
reference holds for the hot path. No GPU involved."""
import math

import numpy as np
import pytest

from enterpolation_b200 import Bezier, BSpline, Linear
from oracle import oracle
from oracle.oracle import Chain, BORDER_BUFFER, BORDER_DELETION

from kats import all_kats, nurbs_circle
from util import assert_near, ulp_distance


def lerp(a, b, f):  # src/utils.rs:6-12
    return a * (1.0 - f) + b * f


@pytest.mark.parametrize("k", all_kats(), ids=lambda k: k["name"])
def test_reference_kat(k):
    c = oracle.OracleCurve(k["make"]().to_desc())
    if "adapt" in k:
        c = k["adapt"](c)
    got = c.take(k["take"]) if "take" in k else c.eval(k["sample"])
    want = np.asarray(k["expect"], dtype=c.dtype)
    if k.get("abs") is not None:
        assert np.max(np.abs(got - want)) <= k["abs"], (got, want)
    else:
        assert_near(got, want, k["ulps"], k["name"])


# ---- span indices: exact assert_eq! in the reference -------------------------------------------
ARR = [0.0, 0.1, 0.2, 0.7, 0.7, 0.7, 0.8, 1.0]


def test_sorted_strict_upper_bound_clamped():  # src/base/list.rs:63-67
    a = Chain.sorted(ARR)
    assert a.strict_upper_bound_clamped(-1.0, 1, 5) == 1
    assert a.strict_upper_bound_clamped(0.15, 1, 5) == 2
    assert a.strict_upper_bound_clamped(0.7, 1, 5) == 5
    assert a.strict_upper_bound_clamped(20.0, 1, 5) == 5


def test_sorted_strict_upper_bound():  # src/base/list.rs:98-102
    a = Chain.sorted(ARR)
    assert [a.strict_upper_bound(x) for x in (-1.0, 0.15, 0.7, 20.0)] == [0, 2, 6, 8]


@pytest.mark.parametrize("make", [Chain.equidistant_normalized, Chain.const_equidistant],
                         ids=["Equidistant", "ConstEquidistant"])
def test_equidistant_bounds(make):  # list.rs:441-444, 471-474 / 617-620, 647-650
    e = make(11)
    assert [e.strict_upper_bound(x) for x in (-1.0, 0.15, 20.0)] == [0, 2, 11]
    assert [e.strict_upper_bound_clamped(x, 1, 3) for x in (-1.0, 0.15, 20.0)] == [1, 2, 3]


def test_border_deletion():  # src/bspline/adaptors.rs:154-165
    d = Chain.equidistant_normalized(11, adaptor=BORDER_DELETION)
    assert len(d) == 9
    assert d.strict_upper_bound_clamped(0.45, 0, 9) == 4
    assert d.strict_upper_bound_clamped(-1.0, 0, 9) == 0
    assert d.strict_upper_bound_clamped(10.0, 0, 9) == 9
    assert d.strict_upper_bound(0.45) == 4
    assert d.strict_upper_bound(-1.0) == 0
    assert d.strict_upper_bound(10.0) == 9
    assert d.strict_upper_bound_clamped(0.8, 1, 5) == 5
    assert d.strict_upper_bound_clamped(0.45, 3, 7) == 4


def test_border_buffer():  # src/bspline/adaptors.rs:168-179
    b = Chain.equidistant_normalized(11, adaptor=BORDER_BUFFER, border_n=3)
    assert len(b) == 17
    assert b.strict_upper_bound_clamped(0.45, 0, 17) == 8
    assert b.strict_upper_bound_clamped(-1.0, 0, 17) == 0
    assert b.strict_upper_bound_clamped(10.0, 0, 17) == 17
    assert b.strict_upper_bound(0.45) == 8
    assert b.strict_upper_bound(-1.0) == 0
    assert b.strict_upper_bound(10.0) == 17
    assert b.strict_upper_bound_clamped(0.8, 1, 5) == 5
    assert b.strict_upper_bound_clamped(0.45, 3, 9) == 8


# ---- span + factor round trips -------------------------------------------------------------------
def test_sorted_upper_border_roundtrip():  # list.rs:144-151
    a = Chain.sorted(ARR)
    for v in (-1.0, 0.0, 0.15, 0.7, 1.0, 20.0):
        lo, hi, f = a.upper_border(v)
        assert_near(np.float64(lerp(a.eval(lo), a.eval(hi), f)), v, 4)


def test_sorted_upper_border_repeated_edges():  # list.rs:158-167
    a = Chain.sorted([0.0, 0.0, 5.0, 5.0, 5.0])
    for v, r in ((-1.0, 0.0), (20.0, 5.0)):
        lo, hi, f = a.upper_border(v)
        assert_near(np.float64(lerp(a.eval(lo), a.eval(hi), f)), r, 4)


@pytest.mark.parametrize("make", [Chain.equidistant_normalized, Chain.const_equidistant],
                         ids=["Equidistant", "ConstEquidistant"])
def test_equidistant_upper_border_roundtrip(make):  # list.rs:523-530 / 693-700
    e = make(6)
    for v in (-1.0, 0.0, 0.15, 0.6, 1.0, 20.0):
        lo, hi, f = e.upper_border(v)
        assert_near(np.float64(lerp(e.eval(lo), e.eval(hi), f)), v, 4)


def test_equidistant_panics_on_nan_and_inf():  # list.rs:456,486,539-540 (to_usize().unwrap())
    e = Chain.equidistant_normalized(11)
    for bad in (float("nan"), float("inf")):
        with pytest.raises(RuntimeError):
            e.strict_upper_bound(bad)
        with pytest.raises(RuntimeError):
            e.upper_border(bad)
    assert e.strict_upper_bound(float("-inf")) == 0


# ---- stepper ---------------------------------------------------------------------------------
def test_stepper():  # src/base/signal.rs:652-664
    from enterpolation_b200 import _abi
    assert_near(oracle.stepper(_abi.F64, 0.0, 1.0, 11), [0.0, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0], 4)
    assert_near(oracle.stepper(_abi.F64, 3.0, 5.0, 5), [3.0, 3.5, 4.0, 4.5, 5.0], 4)


# ---- cross-curve equivalences ------------------------------------------------------------------
def test_readme_three_constructions():  # README.md:52-79
    el = [0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0]
    kn = [-2.0, -2.0, -2.0, -1.0, 0.0, 1.0, 2.0, 2.0, 2.0]
    a = BSpline.builder().clamped().elements(el).equidistant("f64").degree(3).domain(-2.0, 2.0).constant(4)
    b = BSpline.builder().elements(el).knots(kn).constant(4)
    c = BSpline.builder().elements(el).knots(kn).dynamic()
    ya, yb, yc = (oracle.OracleCurve(x.to_desc()).take(10) for x in (a, b, c))
    assert_near(ya, yb, 4)
    assert_near(yb, yc, 4)
    # SURVEY appendix A probe values (first take(10) of the README curve)
    assert_near(ya, [0.0, 0.0877914951989026, 0.7023319615912208, 2.2222222222222223, 3.7366255144032916,
                     3.7366255144032916, 2.2222222222222228, 0.7023319615912217, 0.08779149519890274, 0.0], 0)


def test_mode_equality():  # src/bspline/builder.rs:1346-1377
    el = [1.0, 3.0, 7.0]
    o = BSpline.builder().elements(el).knots([0.0, 0.0, 1.0, 1.0]).constant(3)
    c = BSpline.builder().clamped().elements(el).knots([0.0, 1.0]).constant(3)
    l = BSpline.builder().legacy().elements(el).knots([0.0, 0.0, 0.0, 1.0, 1.0, 1.0]).constant(3)
    yo, yc, yl = (oracle.OracleCurve(x.to_desc()).take(10) for x in (o, c, l))
    assert_near(yo, yc, 4)
    assert_near(yc, yl, 4)


def test_bspline_reasoning():  # examples/bspline_reasoning.rs:9-44
    el = [1.0, 3.0, 10.0, 100.0]
    lin = Linear.builder().elements(el).equidistant("f64").normalized()
    lin_bs = BSpline.builder().elements(el).equidistant("f64").degree(1).normalized().constant(2)
    assert_near(oracle.OracleCurve(lin.to_desc()).take(10), oracle.OracleCurve(lin_bs.to_desc()).take(10), 4)
    bez = Bezier.builder().elements(el).normalized("f64").constant()
    bez_bs = BSpline.builder().clamped().elements(el).knots([0.0, 1.0]).constant(4)
    assert_near(oracle.OracleCurve(bez.to_desc()).take(10), oracle.OracleCurve(bez_bs.to_desc()).take(10), 4)


def test_nurbs_unit_circle():  # examples/nurbs.rs:108-111: ||p|| within 4 ulp of 1 at take(32)
    c = oracle.OracleCurve(nurbs_circle().to_desc())
    p = c.take(32)
    norm = np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1])
    assert_near(norm, np.ones(32), 4)


def test_noise_example_period():  # examples/noise.rs:32-48: values at x and x+10 within 10 ulp
    n = 10_001
    vals = np.array([float(i % 10) for i in range(n)])  # ValueSignal: (input % 10) as f64  (noise.rs:17-22)
    b = BSpline.builder().clamped().elements(vals).equidistant("f64").degree(3).distance(0.0, 1.0).constant(4)
    c = oracle.OracleCurve(b.to_desc())
    for s in (3.3, 157.23, 989.98):
        assert_near(c.eval([s]), c.eval([s + 10.0]), 10)


def test_weighted_equals_manual_homogeneous():
    """Weights lift at every fetch == lifting once (product's table design), incl. w == 0."""
    el = np.array([[1.0, 2.0], [3.0, -1.0], [0.5, 0.25], [2.0, 2.0]])
    w = np.array([1.0, 0.0, 2.5, 0.75])
    b = BSpline.builder().elements_with_weights(el, w).knots([0.0, 0.0, 1.0, 2.0, 2.0]).constant(3)
    c = oracle.OracleCurve(b.to_desc())
    lifted = np.concatenate([np.where(w[:, None] == 0, el, el * w[:, None]), w[:, None]], axis=1)
    h = oracle.OracleCurve(BSpline.builder().elements(lifted).knots([0.0, 0.0, 1.0, 2.0, 2.0]).constant(3).to_desc())
    t = np.linspace(-0.5, 2.5, 41)
    hv = h.eval(t)
    with np.errstate(divide="ignore", invalid="ignore"):
        manual = hv[:, :2] / hv[:, 2:3]
    assert_near(c.eval(t), manual, 0)


def test_take_shards_agree():
    """take(n) evaluated in shards with the GLOBAL index equals the whole."""
    b = BSpline.builder().clamped().elements(np.arange(12, dtype=np.float32).reshape(6, 2)).equidistant("f32") \
        .degree(3).normalized().dynamic()
    c = oracle.OracleCurve(b.to_desc())
    whole = c.take(1000)
    parts = [c.take(1000, lo, hi - lo) for lo, hi in ((0, 333), (333, 700), (700, 1000))]
    assert np.array_equal(np.concatenate(parts).view(np.uint32), whole.view(np.uint32))
    assert np.array_equal(c.take(1000, threads=3).view(np.uint32), whole.view(np.uint32))


def test_bezier_tangent_and_derivatives():  # src/bezier/mod.rs:345-390
    c = oracle.OracleCurve(Bezier.builder().elements([1.0, 2.0, 3.0]).normalized("f64").constant().to_desc())
    assert_near(c.gen_with_tangent([0.5]).ravel(), [2.0, 2.0], 4)
    assert_near(c.gen_with_derivatives([0.5], 3).ravel(), [2.0, 2.0, 0.0], 4)
    assert_near(c.gen_with_derivatives([0.5], 5).ravel(), [2.0, 2.0, 0.0, 0.0, 0.0], 4)
    for els in ([5.0],):  # bigger_workspace / constant
        one = oracle.OracleCurve(Bezier.builder().elements(els).normalized("f64").constant().to_desc())
        assert_near(one.gen_with_tangent([0.5, 0.25, 0.75]).ravel(), [5.0, 0.0, 5.0, 0.0, 5.0, 0.0], 4)


def test_bspline_over_const_equidistant_equals_tabulated_knots():
    """BSpline::new over ConstEquidistant<R, N> (bspline/mod.rs:212-235, base/list.rs:566-721): the knots are
    fl(i)/fl(N-1), so away from the knots themselves (where `floor(x * (N-1))` and the comparison search may round to
    different sides) it is the same curve as the open B-spline over those knots given explicitly."""
    from enterpolation_b200 import BSpline
    for dtype, scalar in ((np.float64, "f64"), (np.float32, "f32")):
        for n_el, n_kn in ((6, 8), (9, 9), (12, 15)):
            el = (np.arange(n_el * 2, dtype=dtype).reshape(n_el, 2) % 7) * dtype(0.5) - dtype(1.0)
            ce = oracle.OracleCurve(BSpline.new_const_equidistant(el, n_kn, scalar).to_desc())
            kn = (np.arange(n_kn, dtype=dtype) / dtype(n_kn - 1)).astype(dtype)
            ex = oracle.OracleCurve(BSpline.builder().elements(el).knots(kn).dynamic().to_desc())
            assert ce.domain() == ex.domain()
            d0, d1 = float(ce.domain()[0]), float(ce.domain()[1])
            t = (d0 - 0.2 + (d1 - d0 + 0.4) * (np.arange(4001, dtype=np.float64) + 0.37) / 4001).astype(dtype)
            ia, _ = ce.span(t)
            ib, _ = ex.span(t)
            assert np.array_equal(ia, ib)
            assert np.array_equal(ce.eval(t).view(np.uint8), ex.eval(t).view(np.uint8))
    # ConstEquidistant's own doctests (list.rs:617-650) on the chain hook
    ch = oracle.Chain.const_equidistant(11)
    assert [ch.strict_upper_bound_clamped(x, 1, 3) for x in (-1.0, 0.15, 20.0)] == [1, 2, 3]
