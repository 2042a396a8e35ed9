"""Builder validation through the C library (host side, no GPU): the reference's error matrices
(src/bspline/builder.rs:1325-1685, src/linear/builder.rs:600-679, src/bezier/builder.rs:554-597,
src/linear/mod.rs, src/bezier/mod.rs) replayed against the product's `enterp_curve_validate` /
`enterp_curve_create`, and cross-checked against the oracle's independent restatement."""
import ctypes as C
import itertools

import numpy as np
import pytest

from enterpolation_b200 import (Bezier, BSpline, Empty, IncongruousElementsDegree, IncongruousElementsKnots,
                                InvalidDegree, KnotElementInequality, Linear, NotSorted, TooFewElements, TooFewKnots,
                                TooSmallWorkspace, Unsupported, BSplineError, LinearError, BezierError, _abi)
from oracle import oracle


def D():
    return BSpline.director()


def test_bspline_degenerate_creations():  # builder.rs:1334-1353
    with pytest.raises(BSplineError):
        BSpline.builder().elements([]).knots([]).constant(1).build()
    with pytest.raises(BSplineError):
        BSpline.builder().elements([1.0]).knots([1.0]).constant(2).build()


def test_bspline_elements_with_weights_build():  # builder.rs:1380-1420 (construction succeeds)
    BSpline.builder().elements_with_weights([(1.0, 1.0), (2.0, 2.0), (3.0, 0.0)]).equidistant("f64").degree(2) \
        .domain(0.0, 5.0).constant(3).build()
    BSpline.builder().elements_with_weights([1.0, 2.0, 3.0], [1.0, 2.0, 0.0]).equidistant("f64").degree(1) \
        .normalized().constant(2).build()
    BSpline.builder().elements([0.1, 0.2, 0.3]).equidistant("f64").degree(1).normalized().constant(2).build()


def test_clamped_errors():  # builder.rs:1423-1509
    with pytest.raises(TooFewElements):
        D().clamped().elements([0.0])
    with pytest.raises(TooFewKnots):
        D().clamped().elements([0.0, 1.0, 2.0, 3.0]).knots([0.0])
    with pytest.raises(TooFewKnots):
        D().clamped().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").quantity(1)
    with pytest.raises(InvalidDegree):
        D().clamped().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").degree(0)
    with pytest.raises(TooSmallWorkspace) as e:
        D().clamped().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").degree(2).domain(0.0, 1.0).constant(2)
    assert (e.value.found, e.value.necessary) == (2, 2)
    with pytest.raises(IncongruousElementsDegree):
        D().clamped().elements([0.0, 1.0, 2.0]).equidistant("f32").degree(3)
    D().clamped().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").degree(3)
    with pytest.raises(IncongruousElementsKnots):
        D().clamped().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(4)
    D().clamped().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(3)


def test_open_errors():  # builder.rs:1512-1624
    with pytest.raises(TooFewElements):
        D().open().elements([0.0])
    with pytest.raises(TooFewKnots):
        D().open().elements([0.0, 1.0, 2.0, 3.0]).knots([0.0])
    with pytest.raises(TooFewKnots):
        D().open().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").quantity(1)
    with pytest.raises(InvalidDegree):
        D().open().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").degree(0)
    with pytest.raises(TooSmallWorkspace):
        D().open().elements([0.0, 1.0, 2.0, 3.0]).equidistant("f32").degree(2).domain(0.0, 1.0).constant(2)
    with pytest.raises(IncongruousElementsKnots) as e:
        D().open().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(2)
    assert e.value.mode == _abi.LEGACY  # builder.rs:680 reports a legacy-flavoured error (upstream quirk)
    D().open().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(3)
    D().open().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(4)
    with pytest.raises(IncongruousElementsKnots) as e:
        D().open().elements([0.0, 1.0, 2.0]).equidistant("f32").quantity(5)
    assert e.value.mode == _abi.OPEN
    D().open().elements([0.0, 1.0]).equidistant("f32").degree(1)
    with pytest.raises(IncongruousElementsDegree):
        D().open().elements([0.0, 1.0]).equidistant("f32").degree(2)


def test_legacy_errors():  # builder.rs:1627-1684
    with pytest.raises(TooFewElements):
        D().legacy().elements([0.0])
    with pytest.raises(TooFewKnots):
        D().legacy().elements([0.0, 1.0, 2.0, 3.0]).knots([0.0, 1.0, 2.0])
    with pytest.raises(TooSmallWorkspace):
        D().legacy().elements([0.0, 1.0, 2.0, 3.0]).knots([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0, 7.0]).constant(2)
    with pytest.raises(IncongruousElementsKnots):
        D().legacy().elements([0.0, 1.0, 3.0]).knots([0.0, 1.0, 2.0, 3.0])
    D().legacy().elements([0.0, 1.0, 3.0]).knots([0.0, 1.0, 2.0, 3.0, 4.0])
    D().legacy().elements([0.0, 1.0, 3.0]).knots([0.0, 1.0, 2.0, 3.0, 4.0, 5.0])
    with pytest.raises(IncongruousElementsKnots):
        D().legacy().elements([0.0, 1.0, 3.0]).knots([0.0, 1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    with pytest.raises(Unsupported):  # no impl block upstream for legacy + equidistant
        D().legacy().elements([0.0, 1.0, 3.0]).equidistant("f64").degree(2)


def test_linear_errors():  # linear/builder.rs:634-679
    with pytest.raises(LinearError):
        Linear.builder().elements([]).knots([]).build()
    with pytest.raises(LinearError):
        Linear.builder().elements([1.0]).knots([1.0]).build()
    with pytest.raises(KnotElementInequality):
        Linear.builder().elements([1.0, 2.0]).knots([1.0, 2.0, 3.0]).build()
    with pytest.raises(TooFewElements):
        Linear.director().elements([0.0])
    with pytest.raises(KnotElementInequality):
        Linear.director().elements([0.0, 1.0]).knots([1.0])
    with pytest.raises(KnotElementInequality) as e:
        Linear.director().elements([1.0, 2.0]).knots([1.0, 2.0, 3.0])
    assert (e.value.found, e.value.necessary) == (2, 3)
    Linear.director().elements([1.0, 2.0]).knots([1.0, 2.0])


def test_not_sorted_index():  # list.rs:263-278: index of the first element smaller than its predecessor
    with pytest.raises(NotSorted) as e:
        Linear.director().elements([1.0, 2.0, 3.0, 4.0]).knots([0.0, 1.0, 0.5, 2.0])
    assert e.value.index == 2
    with pytest.raises(NotSorted) as e:
        Linear.director().elements([1.0, 2.0, 3.0]).knots([0.0, float("nan"), 2.0])
    assert e.value.index == 1
    Linear.director().elements([1.0, 2.0, 3.0]).knots([0.0, 0.0, 0.0])  # non-strictly increasing is fine


def test_bezier_errors():  # bezier/builder.rs:554-597; bezier/mod.rs:296-308
    with pytest.raises(Empty):
        Bezier.director().elements([])
    Bezier.director().elements([1.0])
    with pytest.raises(BezierError):
        Bezier.builder().elements([]).normalized("f64").constant().build()
    with pytest.raises(TooSmallWorkspace) as e:
        Bezier.director().elements([1.0, 2.0, 3.0]).normalized("f64").workspace(2)
    assert (e.value.found, e.value.necessary) == (3, 2)  # builder.rs:359 swaps found/necessary (upstream quirk)
    Bezier.director().elements([1.0, 2.0, 3.0]).normalized("f64").workspace(5)


def test_builder_defers_first_error_to_build():  # bspline/builder.rs:138-140
    b = BSpline.builder().elements([0.0]).knots([0.0]).constant(0)  # three invalid steps, no raise yet
    with pytest.raises(TooFewElements):
        b.build()


def test_resolved_shapes():
    c = BSpline.builder().clamped().elements([0.0] * 7).equidistant("f64").degree(3).domain(-2.0, 2.0).constant(4).build()
    assert (c.info.n_knots, c.degree, c.info.adaptor, c.info.border) == (9, 3, 1, 2)
    assert c.domain() == [-2.0, 2.0]
    c = BSpline.builder().legacy().elements([0.0] * 3).knots([0.0, 0.0, 0.0, 1.0, 1.0, 1.0]).constant(3).build()
    assert (c.info.n_knots, c.degree, c.info.adaptor) == (4, 2, 2)
    c = Bezier.builder().elements([1.0, 2.0, 3.0]).domain(0.0, 2.0, "f64").constant().build()
    assert c.domain() == [0.0, 2.0]
    c = Bezier.builder().elements([1.0, 2.0, 3.0]).domain(1.0, 3.0, "f64").constant().build()
    assert c.domain() == [2.0, 4.0]  # base/adaptors.rs:137-139,161-166: t/(end-start) - start (upstream quirk)


def _status(lib_call, desc):
    info = _abi.ErrorInfo()
    st = lib_call(desc, info)
    return st, info.found, info.necessary, info.index, info.mode


@pytest.mark.parametrize("kind", [_abi.LINEAR, _abi.BEZIER, _abi.BSPLINE])
def test_product_and_oracle_validation_agree_exhaustively(kind):
    """Independent restatements (product resolve.cpp vs oracle capi) over a grid of small shapes."""
    lib = _abi.load_library()
    ol = oracle.lib()
    checked = 0
    for e, k, mode, base, by, cnt, space, ws in itertools.product(
            range(0, 6), range(0, 9), (_abi.OPEN, _abi.CLAMPED, _abi.LEGACY), (_abi.KNOTS_SORTED, _abi.KNOTS_EQUIDISTANT),
            (_abi.EQ_BY_DEGREE, _abi.EQ_BY_QUANTITY), range(0, 7), (_abi.SPACE_CONSTANT, _abi.SPACE_DYNAMIC, _abi.SPACE_WORKSPACE),
            (0, 2, 4)):
        if kind != _abi.BSPLINE and (mode != _abi.OPEN or by != _abi.EQ_BY_DEGREE or cnt > 0):
            continue
        if base == _abi.KNOTS_EQUIDISTANT and k > 0:
            continue
        if base == _abi.KNOTS_SORTED and cnt > 0:
            continue
        if kind == _abi.BEZIER and (base != _abi.KNOTS_SORTED or k > 0):
            continue
        knots = None if base == _abi.KNOTS_EQUIDISTANT or kind == _abi.BEZIER else np.arange(k, dtype=np.float64)
        eq = (by, cnt, _abi.EQ_NORMALIZED, 0.0, 1.0) if base == _abi.KNOTS_EQUIDISTANT else None
        da = _abi.make_desc(kind, _abi.F64, np.arange(e, dtype=np.float64), mode=mode, knots=knots, equidistant=eq,
                            space=(space, ws))
        info_p, info_o = _abi.ErrorInfo(), _abi.ErrorInfo()
        sp = lib.enterp_curve_validate(C.byref(da.desc), 2, C.byref(info_p))
        h = C.c_void_p()
        so = ol.ora_create(C.byref(da.desc), C.byref(h), C.byref(info_o))
        if h:
            ol.ora_destroy(h)
        assert sp == so, (kind, e, k, mode, base, by, cnt, space, ws, sp, so)
        if sp:
            assert (info_p.found, info_p.necessary, info_p.index, info_p.mode) == \
                   (info_o.found, info_o.necessary, info_o.index, info_o.mode), (kind, e, k, mode, base, by, cnt, space, ws)
        checked += 1
    assert checked > 50


def test_bspline_over_const_equidistant_knots_validation():
    """BSpline::new(elements, ConstEquidistant::<R, N>::new(), space) (bspline/mod.rs:212-235): the constructor's
    checks in its order — TooFewElements, IncongruousElementsKnots::open (twice), TooSmallWorkspace — and agreement
    of the product's validation with the oracle's over a grid."""
    with pytest.raises(TooFewElements):
        BSpline.new_const_equidistant([1.0], 5).build()
    with pytest.raises(IncongruousElementsKnots) as e:
        BSpline.new_const_equidistant([1.0, 2.0, 3.0], 2).build()          # knots.len() < elements.len()
    assert (e.value.found, e.value.necessary, e.value.mode) == (3, 2, _abi.OPEN)
    with pytest.raises(IncongruousElementsKnots):
        BSpline.new_const_equidistant([1.0, 2.0, 3.0], 5).build()          # elements.len() <= degree
    with pytest.raises(TooSmallWorkspace) as e:
        BSpline.new_const_equidistant([1.0, 2.0, 3.0, 4.0], 5, workspace=2).build()
    assert (e.value.found, e.value.necessary) == (2, 2)
    c = BSpline.new_const_equidistant([1.0, 2.0, 3.0, 4.0], 5, workspace=3).build()
    assert (c.degree, c.info.n_knots, c.info.adaptor) == (2, 5, 0)
    assert c.domain() == [0.25, 0.75]  # knots[p-1], knots[len-p] of i/(N-1) (bspline/mod.rs:180-185)
    lib, ol = _abi.load_library(), oracle.lib()
    for e_n, n, mode, space, ws in itertools.product(range(0, 7), range(0, 12), (_abi.OPEN, _abi.CLAMPED, _abi.LEGACY),
                                                     (_abi.SPACE_CONSTANT, _abi.SPACE_DYNAMIC), (0, 2, 3, 5)):
        da = _abi.make_desc(_abi.BSPLINE, _abi.F32, np.arange(e_n, dtype=np.float32), mode=mode, const_equidistant=max(n, 0) or True,
                            space=(space, ws))
        da.desc.n_knots = n
        info_p, info_o = _abi.ErrorInfo(), _abi.ErrorInfo()
        sp = lib.enterp_curve_validate(C.byref(da.desc), 2, C.byref(info_p))
        h = C.c_void_p()
        so = ol.ora_create(C.byref(da.desc), C.byref(h), C.byref(info_o))
        if h:
            ol.ora_destroy(h)
        assert sp == so, (e_n, n, mode, space, ws, sp, so)
        if sp:
            assert (info_p.found, info_p.necessary, info_p.mode) == (info_o.found, info_o.necessary, info_o.mode)
