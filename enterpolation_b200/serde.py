"""The reference's serde layout of built curves, as the on-disk format.

UNVERIFIED AGAINST UPSTREAM: the reference ships no serde test or fixture and no Rust toolchain exists in this
image, so nothing here has been checked against real `serde_json` output. The layout below is DERIVED from the
`#[derive(Serialize, Deserialize)]` attributes and serde's documented data model; tests/test_serde.py pins it to
values derived by hand from those rules, not to upstream bytes. Treat it as this package's own stable format that
is meant to coincide with the reference's. Non-finite numbers are NOT interchangeable: serde_json writes NaN/inf
as `null` and cannot read `null` back into an f32/f64, so `dumps` refuses them unless told otherwise.

With its `serde` feature the reference derives `Serialize`/`Deserialize` on the curve types themselves
(`src/bspline/mod.rs:61-67`, `src/linear/mod.rs:65-70`, `src/bezier/mod.rs:161-166`) and on everything they
are generic over, so a serialised curve is the *resolved* structure the builders produced, not the builder
calls. serde's data model maps those derives onto JSON mechanically (struct -> object with the field names,
newtype struct -> its content, tuple / array / Vec -> array, `PhantomData` -> null, fieldless braced struct ->
`{}`), which is what this module writes and reads:

    BSpline<K, E, S>          {"elements": E, "knots": K, "space": S, "degree": usize}
    Linear<K, E, F>           {"elements": E, "knots": K, "easing": F}
    Bezier<R, E, S>           {"elements": E, "space": S, "_input": null}
    Weighted<G>               {"inner": G}                                  weights/weighted.rs:12-15
    Weights<G>                {"signal": [[element, weight], ...]}          weights/mod.rs:19-22 (G yields (T, R) pairs)
    TransformInput<G, R, R>   {"addition": R, "multiplication": R, "inner": G}   base/adaptors.rs:112-117
    Sorted<C>                 C (newtype)                                    base/list.rs:254-255
    Equidistant<R>            {"len": usize, "step": R, "offset": R}         base/list.rs:353-358
    ConstEquidistant<R, N>    null (newtype over PhantomData)                base/list.rs:565-566
    BorderBuffer<G>           {"inner": G, "n": usize}                       bspline/adaptors.rs:6-10
    BorderDeletion<G>         {"inner": G}                                   bspline/adaptors.rs:93-96
    ConstSpace<T, N>          {"_phantom": null}                             base/space.rs:28-31
    DynSpace<T>               {"len": usize, "_phantom": null}               base/space.rs:69-73
    Identity                  {}                                             easing/mod.rs:51-52
    Plateau<R>                {"min": R, "max": R}                           easing/plateau.rs:8-13

The constants inside (`Equidistant::step/offset`, `TransformInput::addition/multiplication`, `Plateau::min/max`)
are the values the reference's constructors compute in R arithmetic; they come from the C library
(`enterp_curve_get_resolved`), which computes them with the same operations. Reading goes the other way: the
stored constants are handed to the library verbatim (`distance(offset, step)`, `ENTERP_INPUT_TRANSFORM`), so a
round trip reproduces the curve bit for bit.

What JSON cannot carry is Rust's type information: the scalar type `R`, `N` of `ConstSpace`/`ConstEquidistant`
and the element container type live in the Rust type, so `from_value` takes the scalar as an argument, as
`serde_json::from_str::<BSpline<..f32..>>` does through its type parameter. `FuncEase` easings hold a closure
or fn item and are not serialisable upstream either.

Status: the reference has no serde tests or fixtures, and no Rust toolchain exists here, so this layout is derived
from the derives and serde's documented data model, not checked against serde_json output ("parity unpinned" for
the textual form; numbers are written as shortest round-trip decimals of their scalar type).
"""
from __future__ import annotations

import json
import math

import numpy as np

from . import _abi
from .builders import Bezier, BSpline, Linear, _Chain
from .easing import Plateau
from .errors import InvalidArgument

_KIND_NAMES = {"linear": _abi.LINEAR, "bezier": _abi.BEZIER, "bspline": _abi.BSPLINE}


def _num(x):
    """serde_json writes non-finite floats as null."""
    x = float(x)
    return x if math.isfinite(x) else None


def _elements_value(b: _Chain, dt):
    el = np.asarray(b._elements, dtype=dt)
    rows = [_num(v) for v in el] if el.ndim == 1 else [[_num(v) for v in row] for row in el]
    if b._weights is None:
        return rows
    w = np.asarray(b._weights, dtype=dt).reshape(-1)
    return {"signal": [[rows[i], _num(w[i])] for i in range(len(rows))]}


def _space_value(kind, n):
    if kind == _abi.SPACE_CONSTANT:
        return {"_phantom": None}
    return {"len": int(n), "_phantom": None}


def to_value(builder: _Chain):
    """The serde data-model value (JSON-compatible Python objects) of the curve `builder.build()` yields (layout
    derived from the reference's derives, unverified against upstream output; non-finite numbers become None)."""
    curve = builder.build()
    info, res = curve.info, curve.resolved()
    dt = curve.dtype
    elements = _elements_value(builder, dt)
    if info.kind == _abi.BEZIER:
        v = {"elements": elements, "space": _space_value(builder._space[0], res.space_len), "_input": None}
    else:
        if info.knot_base == _abi.KNOTS_SORTED:
            inner = [_num(k) for k in np.asarray(builder._knots, dtype=dt).reshape(-1)]
        elif info.knot_base == _abi.KNOTS_CONST_EQUIDISTANT:
            inner = None
        else:
            inner = {"len": int(res.eq_len), "step": _num(dt(res.eq_step)), "offset": _num(dt(res.eq_offset))}
        if info.kind == _abi.LINEAR:
            if res.easing == _abi.EASE_IDENTITY:
                easing = {}
            elif res.easing == _abi.EASE_PLATEAU:
                easing = {"min": _num(dt(res.ease_min)), "max": _num(dt(res.ease_max))}
            else:
                raise InvalidArgument("FuncEase easings hold a function and have no serialised form (easing/mod.rs:17-20)")
            v = {"elements": elements, "knots": inner, "easing": easing}
        else:
            knots = inner
            if info.adaptor == 1:
                knots = {"inner": inner, "n": int(info.border)}
            elif info.adaptor == 2:
                knots = {"inner": inner}
            v = {"elements": elements, "knots": knots, "space": _space_value(builder._space[0], res.space_len),
                 "degree": int(info.degree)}
    if info.weighted:
        v = {"inner": v}
    if res.has_transform:
        v = {"addition": _num(dt(res.in_addition)), "multiplication": _num(dt(res.in_multiplication)), "inner": v}
    return v


# ---- reading ------------------------------------------------------------------------------------------------
def _need(cond, what):
    if not cond:
        raise InvalidArgument(f"not the serialised form of this curve type: {what}")


def _float(x):
    return float("nan") if x is None else float(x)  # null is how non-finite values were written


def _elements_from(b, value):
    if isinstance(value, dict):
        _need(set(value) == {"signal"}, "Weights { signal }")
        pairs = value["signal"]
        _need(all(isinstance(p, list) and len(p) == 2 for p in pairs), "(element, weight) pairs")
        el = [[_float(c) for c in p[0]] if isinstance(p[0], list) else _float(p[0]) for p in pairs]
        return b.elements_with_weights(el, [_float(p[1]) for p in pairs]), True
    el = [[_float(c) for c in e] if isinstance(e, list) else _float(e) for e in value]
    return b.elements(el), False


def _apply_space(b, value, natural_len):
    _need(isinstance(value, dict) and "_phantom" in value, "ConstSpace / DynSpace")
    if "len" not in value:
        return b.constant(natural_len)
    n = int(value["len"])
    return b.dynamic() if n == natural_len else b.workspace(n)


def from_value(kind: str, value, scalar: str = "f64") -> _Chain:
    """Rebuild the builder chain of a serialised curve. `kind` in {"linear", "bezier", "bspline"}; `scalar` names R."""
    k = _KIND_NAMES.get(kind.lower())
    if k is None:
        raise InvalidArgument(f"unknown curve kind {kind!r}")
    sc = "f32" if scalar.lower() in ("f32", "float32") else "f64"
    dt = np.float32 if sc == "f32" else np.float64
    transform = None
    if isinstance(value, dict) and "addition" in value:
        _need(k == _abi.BEZIER and set(value) == {"addition", "multiplication", "inner"}, "TransformInput")
        transform = (_float(value["addition"]), _float(value["multiplication"]))
        value = value["inner"]
    weighted_wrapper = isinstance(value, dict) and set(value) == {"inner"}
    if weighted_wrapper:
        value = value["inner"]
    _need(isinstance(value, dict) and "elements" in value, "curve object")

    if k == _abi.BEZIER:
        _need(set(value) == {"elements", "space", "_input"}, "Bezier { elements, space, _input }")
        b, w = _elements_from(Bezier.builder(), value["elements"])
        b = b.normalized(sc)
        b._bezier_transform = transform
        b = _apply_space(b, value["space"], len(b._elements))
    elif k == _abi.LINEAR:
        _need(set(value) == {"elements", "knots", "easing"}, "Linear { elements, knots, easing }")
        kn, ez = value["knots"], value["easing"]
        if kn is None:  # ConstEquidistant
            _need(not isinstance(value["elements"], dict), "ConstEquidistantLinear has no weighted form")
            el = [[_float(c) for c in e] if isinstance(e, list) else _float(e) for e in value["elements"]]
            b, w = Linear.equidistant_unchecked(el, sc), False
        else:
            b, w = _elements_from(Linear.builder(), value["elements"])
            if isinstance(kn, list):
                b = b.knots(np.asarray([_float(x) for x in kn], dtype=dt))
            else:
                _need(set(kn) == {"len", "step", "offset"} and int(kn["len"]) == len(b._elements), "Equidistant { len, step, offset }")
                b = b.equidistant(sc).distance(_float(kn["offset"]), _float(kn["step"]))
        _need(isinstance(ez, dict), "easing")
        if ez:
            _need(set(ez) == {"min", "max"}, "Plateau { min, max }")
            lo, hi = dt(_float(ez["min"])), dt(_float(ez["max"]))
            strength = dt(2) * lo  # Plateau::new: min = 0 + strength / 2 (exact), max = 1 - strength / 2
            _need(dt(1) - strength / dt(2) == hi, "Plateau built by Plateau::new")
            b = b.easing(Plateau(float(strength)))
    else:
        _need(set(value) == {"elements", "knots", "space", "degree"}, "BSpline { elements, knots, space, degree }")
        kn = value["knots"]
        b = BSpline.builder()
        border = None
        if isinstance(kn, dict) and "inner" in kn:
            if "n" in kn:
                b, border = b.clamped(), int(kn["n"])
            else:
                b = b.legacy()
            kn = kn["inner"]
        b, w = _elements_from(b, value["elements"])
        if isinstance(kn, list):
            b = b.knots(np.asarray([_float(x) for x in kn], dtype=dt))
        else:
            _need(isinstance(kn, dict) and set(kn) == {"len", "step", "offset"}, "Equidistant { len, step, offset }")
            b = b.equidistant(sc).quantity(int(kn["len"])).distance(_float(kn["offset"]), _float(kn["step"]))
        b = _apply_space(b, value["space"], int(value["degree"]) + 1)
        info = b.build().info  # the stored redundancy must agree with what the builders derive
        _need(int(info.degree) == int(value["degree"]), "degree")
        _need(border is None or int(info.border) == border, "BorderBuffer n")
    _need(w == weighted_wrapper, "Weighted wrapper and Weights elements go together")
    return b


# ---- text ---------------------------------------------------------------------------------------------------
def _fmt(x, f32):
    if x is None:
        return "null"
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    if isinstance(x, float):
        s = str(np.float32(x)) if f32 else repr(x)  # shortest decimal that round-trips in R
        return s if any(c in s for c in ".en") else s + ".0"
    if isinstance(x, list):
        return "[" + ",".join(_fmt(v, f32) for v in x) + "]"
    if isinstance(x, dict):
        return "{" + ",".join(json.dumps(k) + ":" + _fmt(v, f32) for k, v in x.items()) + "}"
    raise TypeError(type(x))


def _has_null_number(v) -> bool:
    if isinstance(v, dict):
        return any(_has_null_number(x) for k, x in v.items() if k not in ("_phantom", "_input", "knots") or isinstance(x, (dict, list)))
    if isinstance(v, list):
        return any(x is None or _has_null_number(x) for x in v)
    return False


def dumps(builder: _Chain, allow_nonfinite: bool = False) -> str:
    """`serde_json::to_string(&curve)`-shaped text of the curve this chain builds. Layout derived from the
    reference's serde derives, unverified against upstream output (module docstring). NaN/inf elements, knots or
    weights would be written as `null`, which `serde_json::from_str` cannot read back as a float: InvalidArgument
    unless `allow_nonfinite=True` (then this package's own `loads` reads them back as NaN)."""
    value = to_value(builder)
    if not allow_nonfinite and _has_null_number(value):
        raise InvalidArgument("non-finite number: serde_json would write null and could not read it back")
    return _fmt(value, builder.build().dtype == np.float32)


def loads(kind: str, text: str, scalar: str = "f64") -> _Chain:
    """`serde_json::from_str::<Curve<..R..>>(text)`: the builder chain whose `.build()` is that curve. Layout
    derived from the reference's serde derives, unverified against upstream output (module docstring)."""
    return from_value(kind, json.loads(text), scalar)
