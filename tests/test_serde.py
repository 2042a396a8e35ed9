"""enterpolation_b200.serde: the reference's serde layout of built curves.

No serde fixture exists upstream, so the layout is checked against values derived by hand from the derives
(struct field names, src/bspline/mod.rs:61-67 etc.) and the reference's constructors, and by round trips:
a curve written and read back must evaluate bit-identically in the CPU oracle."""
import json

import numpy as np
import pytest

from enterpolation_b200 import Bezier, BSpline, Linear, easing, serde
from enterpolation_b200.errors import InvalidArgument
from oracle import oracle

from util import assert_bits_equal, u01


def test_readme_cubic_layout():
    """README.md:52-62 curve: BSpline<BorderBuffer<Equidistant<f64>>, [f64; 7], ConstSpace<f64, 4>>.
    Equidistant::new(5, -2, 2): step = (2 - -2) / 4 = 1, offset = -2 (list.rs:393-399); BorderBuffer n = degree - 1."""
    b = BSpline.builder().clamped().elements([0, 0, 0, 6, 0, 0, 0]).equidistant("f64").degree(3).domain(-2, 2).constant(4)
    assert serde.dumps(b) == ('{"elements":[0.0,0.0,0.0,6.0,0.0,0.0,0.0],"knots":{"inner":{"len":5,"step":1.0,"offset":-2.0},"n":2},'
                              '"space":{"_phantom":null},"degree":3}')


def test_linear_layouts():
    # examples/linear.rs:6-10 shape: Linear<Sorted<[f64;3]>, [f64;3], Identity>
    b = Linear.builder().elements([1.0, 5.0, 100.0]).knots([0.0, 1.0, 3.0])
    assert serde.to_value(b) == {"elements": [1.0, 5.0, 100.0], "knots": [0.0, 1.0, 3.0], "easing": {}}
    # linear/mod.rs:10-16: equidistant normalized: Equidistant::normalized(3) = {3, 1/(3-1), 0} (list.rs:380-386)
    b = Linear.builder().elements([0.0, 5.0, 3.0]).equidistant("f64").normalized()
    assert serde.to_value(b)["knots"] == {"len": 3, "step": 0.5, "offset": 0.0}
    # weighted (linear/builder.rs:200-208): Weighted<Linear<_, Weights<Vec<(f64, f64)>>, _>>
    b = Linear.builder().elements_with_weights([(1.0, 1.0), (2.0, 4.0), (3.0, 0.0)]).equidistant("f64").normalized()
    v = serde.to_value(b)
    assert set(v) == {"inner"} and v["inner"]["elements"] == {"signal": [[1.0, 1.0], [2.0, 4.0], [3.0, 0.0]]}
    # Plateau::new(0.5): min = 0 + 0.25, max = 1 - 0.25 (plateau.rs:19-25); ConstEquidistant is PhantomData
    b = Linear.equidistant_unchecked([[0.0, 1.0], [2.0, 3.0], [4.0, 5.0]], "f32").easing(easing.Plateau(0.5))
    assert serde.dumps(b) == '{"elements":[[0.0,1.0],[2.0,3.0],[4.0,5.0]],"knots":null,"easing":{"min":0.25,"max":0.75}}'
    with pytest.raises(InvalidArgument):
        serde.to_value(Linear.builder().elements([0.0, 1.0]).equidistant("f64").normalized().easing(easing.smoothstep))


def test_bezier_layouts():
    b = Bezier.builder().elements([0.0, 5.0, 3.0]).normalized("f64").constant()
    assert serde.dumps(b) == '{"elements":[0.0,5.0,3.0],"space":{"_phantom":null},"_input":null}'
    b = Bezier.builder().elements([[0.0, 1.0], [5.0, 2.0], [3.0, 3.0]]).normalized("f64").dynamic()
    assert serde.to_value(b)["space"] == {"len": 3, "_phantom": None}
    # TransformInput::normalized_to_domain(_, 1, 3): addition = -1, multiplication = recip(2) (adaptors.rs:137-139)
    b = Bezier.builder().elements_with_weights([(1.0, 1.0), (2.0, 0.5)]).domain(1.0, 3.0, "f64").constant()
    v = serde.to_value(b)
    assert (v["addition"], v["multiplication"]) == (-1.0, 0.5) and set(v["inner"]) == {"inner"}
    assert v["inner"]["inner"]["elements"] == {"signal": [[1.0, 1.0], [2.0, 0.5]]}


def test_f32_numbers_are_shortest_round_trip():
    b = Linear.builder().elements(np.array([0.1, 0.2, 1e-30], dtype=np.float32)).equidistant("f32").domain(0.0, 0.7)
    text = serde.dumps(b)
    assert text.startswith('{"elements":[0.1,0.2,1e-30],')
    step = json.loads(text)["knots"]["step"]
    assert np.float32(step) == np.float32(np.float32(0.7) - np.float32(0.0)) / np.float32(2.0)


def _cases():
    el = lambda n, d, dt, s: (u01(n * d, s, dt) * dt(2) - dt(1)).reshape(n, d) if d > 1 else (u01(n, s, dt) * dt(2) - dt(1))
    w = lambda n, dt: (0.5 + u01(n, 0x33, dt)).astype(dt)
    kn = lambda n, dt: np.cumsum(0.5 + u01(n, 0x44, np.float64)).astype(dt)
    for sc, dt in (("f32", np.float32), ("f64", np.float64)):
        yield "linear", sc, Linear.builder().elements(el(9, 3, dt, 1)).knots(kn(9, dt))
        yield "linear", sc, Linear.builder().elements_with_weights(el(9, 2, dt, 2), w(9, dt)).equidistant(sc).domain(-1.0, 0.7)
        yield "linear", sc, Linear.builder().elements(el(6, 1, dt, 3)).equidistant(sc).distance(0.3, 0.1).easing(easing.Plateau(0.3))
        yield "linear", sc, Linear.equidistant_unchecked(el(7, 4, dt, 4), sc)
        yield "bezier", sc, Bezier.builder().elements(el(5, 3, dt, 5)).normalized(sc).constant()
        yield "bezier", sc, Bezier.builder().elements(el(6, 1, dt, 6)).domain(0.3, 1.7, sc).dynamic()
        yield "bezier", sc, Bezier.builder().elements_with_weights(el(4, 2, dt, 7), w(4, dt)).domain(-2.0, 5.0, sc).workspace(9)
        yield "bspline", sc, BSpline.builder().elements(el(9, 2, dt, 8)).knots(kn(11, dt)).constant(4)
        yield "bspline", sc, BSpline.builder().clamped().elements(el(9, 1, dt, 9)).knots(kn(7, dt)).dynamic()
        yield "bspline", sc, BSpline.builder().legacy().elements(el(9, 3, dt, 10)).knots(kn(13, dt)).constant(5)
        yield "bspline", sc, BSpline.builder().elements(el(8, 1, dt, 11)).equidistant(sc).degree(3).domain(-1.0, 2.3).dynamic()
        yield "bspline", sc, BSpline.builder().clamped().elements_with_weights(el(8, 3, dt, 12), w(8, dt)).equidistant(sc).degree(2).domain(0.1, 0.9).workspace(7)
        yield "bspline", sc, BSpline.builder().clamped().elements(el(12, 4, dt, 13)).equidistant(sc).quantity(9).normalized().constant(5)


@pytest.mark.parametrize("kind,scalar,builder", list(_cases()), ids=lambda v: v if isinstance(v, str) else "b")
def test_round_trip_is_bit_exact(kind, scalar, builder):
    text = serde.dumps(builder)
    back = serde.loads(kind, text, scalar)
    assert serde.dumps(back) == text
    a, b = oracle.OracleCurve(builder.to_desc()), oracle.OracleCurve(back.to_desc())
    da, db = a.domain(), b.domain()
    assert_bits_equal(np.asarray(da), np.asarray(db), "domain")
    dt = np.float32 if scalar == "f32" else np.float64
    span = float(da[1] - da[0])
    t = (u01(4000, 0x55, dt) * dt(1.4 * span) + dt(float(da[0]) - 0.2 * span)).astype(dt)
    assert_bits_equal(a.eval(t), b.eval(t), "eval")
    assert_bits_equal(a.take(257), b.take(257), "take")


def test_rejects_foreign_shapes():
    with pytest.raises(InvalidArgument):
        serde.loads("bspline", '{"elements":[0.0,1.0],"knots":[0.0,1.0]}', "f64")
    with pytest.raises(InvalidArgument):
        serde.loads("linear", '{"elements":[0.0,1.0],"knots":{"len":3,"step":0.5,"offset":0.0},"easing":{}}', "f64")
    with pytest.raises(InvalidArgument):
        serde.loads("linear", '{"elements":[0.0,1.0],"knots":[0.0,1.0],"easing":{"min":0.25,"max":0.5}}', "f64")
    with pytest.raises(InvalidArgument):
        serde.loads("bspline", '{"elements":[0.0,1.0,2.0,3.0],"knots":[0.0,1.0,2.0,3.0,4.0],"space":{"_phantom":null},"degree":3}', "f64")
    with pytest.raises(InvalidArgument):
        serde.loads("curve", "{}", "f64")


def test_non_finite_values_are_refused():
    """serde_json writes NaN/inf as null and cannot read null back into a float: dumps refuses them by default."""
    b = Linear.builder().elements([1.0, float("nan"), 3.0]).knots([0.0, 1.0, 2.0])
    with pytest.raises(InvalidArgument):
        serde.dumps(b)
    text = serde.dumps(b, allow_nonfinite=True)
    assert "null" in text
    ok = Linear.equidistant_unchecked([0.0, 1.0], "f64")  # "knots": null of ConstEquidistant is not a number
    assert '"knots":null' in serde.dumps(ok)
