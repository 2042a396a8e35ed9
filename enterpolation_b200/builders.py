"""Fluent builders mirroring the reference's typestate builders (same method names and order).

    BSpline.builder()  src/bspline/mod.rs:115   -> src/bspline/builder.rs
    Linear.builder()   src/linear/mod.rs:105    -> src/linear/builder.rs
    Bezier.builder()   src/bezier/mod.rs:207    -> src/bezier/builder.rs

`*.builder()` defers the first error to `build()` (the `*Builder` types, e.g. bspline/builder.rs:
138-140); `*.director()` raises at the offending call (the `*Director` types). The rules themselves
are NOT restated here: every check is done by the C library (`enterp_curve_validate` /
`enterp_curve_create`), this file only records the call chain into an `enterp_curve_desc`.

Python has no type inference for `R`, so it is named where Rust would infer it:
`.equidistant("f64")`, `.normalized("f32")` (Bezier), or the dtype of the `knots` array.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .curve import Curve
from .errors import EnterpolationError, InvalidArgument, from_info

_STAGE_ELEMENTS, _STAGE_KNOTS, _STAGE_SPACE = 0, 1, 2


def _scalar_of(x, default=None):
    if x is None:
        return default
    if isinstance(x, str):
        s = x.lower()
        if s in ("f32", "float32", "single"):
            return _abi.F32
        if s in ("f64", "float64", "double"):
            return _abi.F64
        raise InvalidArgument(f"unknown scalar type {x!r}")
    if x in (np.float32,) or getattr(x, "name", None) == "float32" or str(x) == "torch.float32":
        return _abi.F32
    if x in (np.float64, float) or getattr(x, "name", None) == "float64" or str(x) == "torch.float64":
        return _abi.F64
    raise InvalidArgument(f"unknown scalar type {x!r}")


def _dtype_scalar(arr):
    a = np.asarray(arr)
    return _abi.F32 if a.dtype == np.float32 else _abi.F64


class _Chain:
    """Shared recording machinery."""

    _kind = None

    def __init__(self, immediate: bool):
        self._immediate = immediate
        self._error: EnterpolationError | None = None
        self._elements = None
        self._weights = None
        self._knots = None
        self._scalar = None
        self._mode = _abi.OPEN
        self._eq = None           # [eq_by, eq_count, eq_form, a, b]
        self._space = (_abi.SPACE_DYNAMIC, 0)
        self._bezier_domain = None
        self._bezier_transform = None  # stored TransformInput fields of a deserialised curve (serde.py)
        self._easing = (_abi.EASE_IDENTITY, 0, 0.0)
        self._const_equidistant = False
        self._device = 0

    # -- helpers ---------------------------------------------------------------------------------
    def _resolved_scalar(self):
        if self._scalar is not None:
            return self._scalar
        if self._knots is not None:
            return _dtype_scalar(self._knots)
        if self._elements is not None and np.asarray(self._elements).dtype == np.float32:
            return _abi.F32
        return _abi.F64

    def _desc(self) -> _abi.DescArrays:
        eq = None
        if self._eq is not None:
            eq = tuple(self._eq)
        return _abi.make_desc(self._kind, self._resolved_scalar(),
                              [] if self._elements is None else self._elements, self._weights,
                              mode=self._mode, knots=self._knots, equidistant=eq, space=self._space,
                              bezier_domain=self._bezier_domain, easing=self._easing,
                              const_equidistant=self._const_equidistant, bezier_transform=self._bezier_transform)

    def _validate(self, stage: int):
        """Ask the C library whether the chain so far is valid (Director semantics)."""
        if self._error is not None:
            return self
        lib = _abi.load_library()
        da = self._desc()
        info = _abi.ErrorInfo()
        st = lib.enterp_curve_validate(C.byref(da.desc), stage, C.byref(info))
        if st != _abi.OK:
            info.status = st
            e = from_info(info)
            if self._immediate:
                raise e
            self._error = e
        return self

    def on_device(self, device: int):
        """Not in the reference: the GPU the built curve evaluates on by default."""
        self._device = int(device)
        return self

    def to_desc(self) -> _abi.DescArrays:
        """The C descriptor of this chain (used by the parity tests to feed the oracle)."""
        return self._desc()

    def build(self) -> Curve:
        if self._error is not None:
            raise self._error
        return Curve(self._desc(), device=self._device)


class _ElementsMixin:
    def elements(self, elements):
        """`.elements(E)`: [n] scalars or [n][2..4] vectors."""
        self._elements = np.asarray(elements)
        self._weights = None
        return self._validate(_STAGE_ELEMENTS)

    def elements_with_weights(self, elements, weights=None):
        """`.elements_with_weights(G)`: either an iterable of (element, weight) pairs or two arrays
        (the `elements.stack(weights)` form, signal.rs:71-78)."""
        if weights is None:
            pairs = list(elements)
            self._elements = np.asarray([p[0] for p in pairs])
            self._weights = np.asarray([p[1] for p in pairs])
        else:
            self._elements = np.asarray(elements)
            self._weights = np.asarray(weights)
        return self._validate(_STAGE_ELEMENTS)


class _EquidistantMixin:
    def domain(self, start, end):
        self._eq[2:] = [_abi.EQ_DOMAIN, start, end]
        return self._validate(_STAGE_KNOTS)

    def normalized(self):
        self._eq[2:] = [_abi.EQ_NORMALIZED, 0.0, 1.0]
        return self._validate(_STAGE_KNOTS)

    def distance(self, start, step):
        self._eq[2:] = [_abi.EQ_DISTANCE, start, step]
        return self._validate(_STAGE_KNOTS)


# ------------------------------------------------------------------------------------------------
class LinearBuilder(_Chain, _ElementsMixin, _EquidistantMixin):
    """src/linear/builder.rs — elements -> knots | equidistant.{domain,normalized,distance} -> build."""

    _kind = _abi.LINEAR

    def knots(self, knots):
        self._knots = np.asarray(knots)
        if self._knots.dtype not in (np.float32, np.float64):
            self._knots = self._knots.astype(np.float64)
        self._eq = None
        return self._validate(_STAGE_KNOTS)

    def equidistant(self, scalar="f64"):
        self._scalar = _scalar_of(scalar)
        self._knots = None
        self._eq = [0, 0, _abi.EQ_NORMALIZED, 0.0, 1.0]
        return self

    def easing(self, easing):
        """`.easing(F)` (linear/builder.rs:496-503): one of the built-ins in enterpolation_b200.easing
        (Identity, flip, smoothstart(n), smoothend(n), smoothstep, smootherstep, Plateau(strength))."""
        from .easing import Easing
        if easing is None:
            self._easing = (_abi.EASE_IDENTITY, 0, 0.0)
        elif isinstance(easing, Easing):
            self._easing = easing.triple()
        else:
            raise NotImplementedError("only the built-in easings cross to the device; closures (FuncEase) cannot")
        return self


class BezierBuilder(_Chain, _ElementsMixin):
    """src/bezier/builder.rs — elements -> normalized::<R>() | domain(a,b) -> space -> build."""

    _kind = _abi.BEZIER

    def __init__(self, immediate):
        super().__init__(immediate)
        self._space = (_abi.SPACE_CONSTANT, 0)

    def normalized(self, scalar="f64"):
        self._scalar = _scalar_of(scalar)
        self._bezier_domain = None
        return self

    def domain(self, start, end, scalar=None):
        self._scalar = _scalar_of(scalar, self._scalar if self._scalar is not None else _abi.F64)
        self._bezier_domain = (start, end)
        return self

    def dynamic(self):
        self._space = (_abi.SPACE_DYNAMIC, 0)
        return self._validate(_STAGE_SPACE)

    def constant(self, n: int | None = None):
        """`constant::<N>()`; Rust pins N to the array length at compile time (builder.rs:330-340)."""
        self._space = (_abi.SPACE_CONSTANT, 0 if n is None else int(n))
        return self._validate(_STAGE_SPACE)

    def workspace(self, n: int):
        """`workspace(S)` with `S.len() == n`."""
        self._space = (_abi.SPACE_WORKSPACE, int(n))
        return self._validate(_STAGE_SPACE)


class BSplineBuilder(_Chain, _ElementsMixin, _EquidistantMixin):
    """src/bspline/builder.rs — [open|clamped|legacy] -> elements -> knots | equidistant.(degree|quantity)
    .(domain|normalized|distance) -> dynamic|constant|workspace -> build."""

    _kind = _abi.BSPLINE

    def open(self):
        self._mode = _abi.OPEN
        return self

    def clamped(self):
        self._mode = _abi.CLAMPED
        return self

    def legacy(self):
        self._mode = _abi.LEGACY
        return self

    def knots(self, knots):
        self._knots = np.asarray(knots)
        if self._knots.dtype not in (np.float32, np.float64):
            self._knots = self._knots.astype(np.float64)
        self._eq = None
        return self._validate(_STAGE_KNOTS)

    def equidistant(self, scalar="f64"):
        self._scalar = _scalar_of(scalar)
        self._knots = None
        self._eq = [_abi.EQ_BY_DEGREE, 0, _abi.EQ_NORMALIZED, 0.0, 1.0]
        self._eq_pending = True
        return self

    def degree(self, degree: int):
        self._eq[0:2] = [_abi.EQ_BY_DEGREE, int(degree)]
        return self._validate(_STAGE_KNOTS)

    def quantity(self, quantity: int):
        self._eq[0:2] = [_abi.EQ_BY_QUANTITY, int(quantity)]
        return self._validate(_STAGE_KNOTS)

    def dynamic(self):
        self._space = (_abi.SPACE_DYNAMIC, 0)
        return self._validate(_STAGE_SPACE)

    def constant(self, n: int):
        self._space = (_abi.SPACE_CONSTANT, int(n))
        return self._validate(_STAGE_SPACE)

    def workspace(self, n: int):
        self._space = (_abi.SPACE_WORKSPACE, int(n))
        return self._validate(_STAGE_SPACE)


class Linear:
    @staticmethod
    def builder() -> LinearBuilder:
        return LinearBuilder(immediate=False)

    @staticmethod
    def equidistant_unchecked(elements, scalar="f64") -> LinearBuilder:
        """`Linear::equidistant_unchecked(elements)` / `ConstEquidistantLinear` (linear/mod.rs:193-216): knots
        i/(N-1) on [0,1] with ConstEquidistant's own arithmetic (list.rs:566-721). Returns a chain whose
        `.build()` yields the curve (and whose `.to_desc()` feeds the oracle)."""
        b = LinearBuilder(immediate=False)
        b._scalar = _scalar_of(scalar)
        b._const_equidistant = True
        return b.elements(elements)

    @staticmethod
    def director() -> LinearBuilder:
        return LinearBuilder(immediate=True)


class Bezier:
    @staticmethod
    def builder() -> BezierBuilder:
        return BezierBuilder(immediate=False)

    @staticmethod
    def director() -> BezierBuilder:
        return BezierBuilder(immediate=True)


class BSpline:
    @staticmethod
    def builder() -> BSplineBuilder:
        return BSplineBuilder(immediate=False)

    @staticmethod
    def director() -> BSplineBuilder:
        return BSplineBuilder(immediate=True)

    @staticmethod
    def new_const_equidistant(elements, n_knots: int, scalar="f64", workspace: int | None = None, weights=None) -> BSplineBuilder:
        """`BSpline::new(elements, ConstEquidistant::<R, N>::new(), space)` (bspline/mod.rs:212-235 over
        base/list.rs:566-721): N = n_knots knots i/(N-1) on [0,1], span search by ConstEquidistant's own
        `x * (N-1)` arithmetic, degree = N - len(elements) + 1. No builder chain reaches this type upstream; the
        constructor's checks apply. workspace=None is `DynSpace::new(degree + 1)`, an int is `ConstSpace::<N>`.
        Returns a chain whose `.build()` yields the curve (and whose `.to_desc()` feeds the oracle)."""
        b = BSplineBuilder(immediate=False)
        b._scalar = _scalar_of(scalar)
        b._const_equidistant = int(n_knots)
        b._space = (_abi.SPACE_DYNAMIC, 0) if workspace is None else (_abi.SPACE_CONSTANT, int(workspace))
        if weights is not None:
            b._elements, b._weights = np.asarray(elements), np.asarray(weights)
        else:
            b._elements = np.asarray(elements)
        return b._validate(_STAGE_SPACE)
