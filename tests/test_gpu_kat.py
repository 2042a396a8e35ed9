"""The reference's known-answer tests replayed through the CUDA path (C ABI -> sm_100a kernels), and
each result additionally required to equal the CPU oracle bit for bit."""
import numpy as np
import pytest
import torch

from enterpolation_b200 import Bezier, BSpline, Linear
from oracle import oracle

from kats import all_kats, nurbs_circle
from util import assert_bits_equal, assert_near

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", all_kats(), ids=lambda k: k["name"])
def test_reference_kat_on_gpu(k):
    b = k["make"]()
    curve = b.build()
    ref = oracle.OracleCurve(b.to_desc())
    if "adapt" in k:
        curve, ref = k["adapt"](curve), k["adapt"](ref)
    if "take" in k:
        got = np.asarray(curve.take(k["take"]).collect())
        want_oracle = ref.take(k["take"])
    else:
        got = np.asarray(curve.sample(k["sample"]))
        want_oracle = ref.eval(k["sample"])
    want = np.asarray(k["expect"], dtype=curve.dtype)
    if k.get("abs") is not None:
        assert np.max(np.abs(got - want)) <= k["abs"], (got, want)
    else:
        assert_near(got, want, k["ulps"], k["name"])
    assert_bits_equal(got, want_oracle, k["name"] + " vs oracle")


def test_eval_single_and_iteration():
    lin = Linear.builder().elements([0.0, 3.0]).knots([0.0, 1.0]).build()
    assert lin.eval(0.5) == 1.5
    assert [float(v) for v in lin.take(3)] == [0.0, 1.5, 3.0]
    assert lin.domain() == [0.0, 1.0]


def test_readme_three_constructions_gpu():  # README.md:52-79
    el = [0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0]
    kn = [-2.0, -2.0, -2.0, -1.0, 0.0, 1.0, 2.0, 2.0, 2.0]
    a = BSpline.builder().clamped().elements(el).equidistant("f64").degree(3).domain(-2.0, 2.0).constant(4).build()
    b = BSpline.builder().elements(el).knots(kn).constant(4).build()
    c = BSpline.builder().elements(el).knots(kn).dynamic().build()
    ya, yb, yc = (x.take(10).collect() for x in (a, b, c))
    assert_near(ya, yb, 4)
    assert_near(yb, yc, 4)


def test_mode_equality_gpu():  # src/bspline/builder.rs:1346-1377
    el = [1.0, 3.0, 7.0]
    o = BSpline.builder().elements(el).knots([0.0, 0.0, 1.0, 1.0]).constant(3).build()
    c = BSpline.builder().clamped().elements(el).knots([0.0, 1.0]).constant(3).build()
    l = BSpline.builder().legacy().elements(el).knots([0.0, 0.0, 0.0, 1.0, 1.0, 1.0]).constant(3).build()
    yo, yc, yl = (x.take(10).collect() for x in (o, c, l))
    assert_near(yo, yc, 4)
    assert_near(yc, yl, 4)


def test_bspline_reasoning_gpu():  # examples/bspline_reasoning.rs:9-44
    el = [1.0, 3.0, 10.0, 100.0]
    lin = Linear.builder().elements(el).equidistant("f64").normalized().build()
    lin_bs = BSpline.builder().elements(el).equidistant("f64").degree(1).normalized().constant(2).build()
    assert_near(lin.take(10).collect(), lin_bs.take(10).collect(), 4)
    bez = Bezier.builder().elements(el).normalized("f64").constant().build()
    bez_bs = BSpline.builder().clamped().elements(el).knots(bez.domain()).constant(4).build()
    assert_near(bez.take(10).collect(), bez_bs.take(10).collect(), 4)


def test_nurbs_unit_circle_gpu():  # examples/nurbs.rs:108-111
    p = nurbs_circle().build().take(32).collect()
    assert_near(np.sqrt(p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1]), np.ones(32), 4)


def test_noise_example_gpu():  # examples/noise.rs:32-48
    vals = np.array([float(i % 10) for i in range(10_001)])
    c = BSpline.builder().clamped().elements(vals).equidistant("f64").degree(3).distance(0.0, 1.0).constant(4).build()
    s = np.array([3.3, 157.23, 989.98])
    assert_near(c.sample(s), c.sample(s + 10.0), 10)


def test_span_kats_on_gpu():
    """list.rs / adaptors.rs span KATs through enterp_span_batch (BSpline index = clamped search)."""
    # BorderBuffer(Equidistant::normalized(11), 3): a clamped equidistant quartic with 14 elements has exactly
    # this knot chain (inner len = e - d + 1 = 11, n = d - 1 = 3); search range [4, 13].
    b = BSpline.builder().clamped().elements(np.zeros(14)).equidistant("f64").degree(4).normalized().constant(5)
    c = b.build()
    ref = oracle.OracleCurve(b.to_desc())
    t = np.array([0.45, -1.0, 10.0, 0.8, 0.0, 1.0, 0.3, 0.30000000000000004, 0.29999999999999993])
    idx, _ = c.span_device(torch.from_numpy(t).cuda())
    widx, _ = ref.span(t)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx)
    assert int(idx[0]) == 8  # adaptors.rs:171: strict_upper_bound_clamped(0.45, 0, 17) == 8 lies inside [4, 13]
