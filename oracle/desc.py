"""TEST INFRASTRUCTURE — the checker's own ctypes view of include/enterp.h's descriptor structs, so that the oracle
(and `bench.py --impl reference`) never has to import the product package. Layouts are checked against the
product's view in tests/test_abi.py.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

LINEAR, BEZIER, BSPLINE = 0, 1, 2
F32, F64 = 0, 1
OPEN, CLAMPED, LEGACY = 0, 1, 2
KNOTS_SORTED, KNOTS_EQUIDISTANT, KNOTS_CONST_EQUIDISTANT = 0, 1, 2
EQ_BY_DEGREE, EQ_BY_QUANTITY = 0, 1
EQ_DOMAIN, EQ_NORMALIZED, EQ_DISTANCE = 0, 1, 2
SPACE_CONSTANT, SPACE_DYNAMIC, SPACE_WORKSPACE = 0, 1, 2
NP_DTYPE = {F32: np.float32, F64: np.float64}


class CurveDesc(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scalar", C.c_int32), ("dim", C.c_int32), ("weighted", C.c_int32),
        ("n_elements", C.c_uint64), ("elements", C.c_void_p), ("weights", C.c_void_p),
        ("mode", C.c_int32), ("knot_base", C.c_int32),
        ("n_knots", C.c_uint64), ("knots", C.c_void_p),
        ("eq_by", C.c_int32), ("eq_form", C.c_int32),
        ("eq_count", C.c_uint64), ("eq_a", C.c_double), ("eq_b", C.c_double),
        ("space", C.c_int32), ("input_form", C.c_int32),
        ("workspace_n", C.c_uint64), ("in_a", C.c_double), ("in_b", C.c_double),
        ("easing", C.c_int32), ("easing_n", C.c_int32), ("easing_param", C.c_double),
    ]


class ErrorInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("mode", C.c_int32),
                ("found", C.c_uint64), ("necessary", C.c_uint64), ("index", C.c_uint64)]


class CurveInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("scalar", C.c_int32), ("dim", C.c_int32), ("weighted", C.c_int32),
        ("knot_base", C.c_int32), ("adaptor", C.c_int32),
        ("n_elements", C.c_uint64), ("n_knots", C.c_uint64), ("degree", C.c_uint64),
        ("border", C.c_uint64), ("domain", C.c_double * 2),
    ]


class Desc:
    """A descriptor plus the numpy arrays it points into."""

    def __init__(self, desc: CurveDesc, keep):
        self.desc, self.keep = desc, keep


def bspline(elements, scalar, *, mode=OPEN, knots=None, weights=None, equidistant=None, space=(SPACE_DYNAMIC, 0)) -> Desc:
    """BSpline::builder()[.clamped()].elements(..)[_with_weights].knots(K) | .equidistant().degree|quantity(..)
    .domain|normalized|distance(..) .constant::<N>() | .dynamic(): equidistant = (eq_by, count, eq_form, a, b)."""
    dt = NP_DTYPE[scalar]
    el = np.ascontiguousarray(np.asarray(elements, dtype=dt))
    if el.ndim == 1:
        el = el.reshape(-1, 1)
    d = CurveDesc()
    keep = [el]
    d.kind, d.scalar, d.dim, d.n_elements = BSPLINE, scalar, el.shape[1], el.shape[0]
    d.elements = el.ctypes.data
    if weights is not None:
        w = np.ascontiguousarray(np.asarray(weights, dtype=dt)).reshape(-1)
        keep.append(w)
        d.weighted, d.weights = 1, w.ctypes.data
    d.mode = mode
    if knots is not None:
        k = np.ascontiguousarray(np.asarray(knots, dtype=dt)).reshape(-1)
        keep.append(k)
        d.knot_base, d.n_knots, d.knots = KNOTS_SORTED, k.shape[0], k.ctypes.data
    else:
        d.knot_base = KNOTS_EQUIDISTANT
        d.eq_by, d.eq_count, d.eq_form, a, b = equidistant
        d.eq_a, d.eq_b = float(dt(a)), float(dt(b))
    d.space, d.workspace_n = space
    return Desc(d, keep)
