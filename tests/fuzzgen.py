"""Seeded random curve generators of the fuzz tests (tests/test_gpu_fuzz.py: CUDA path vs oracle; tests/
test_oracle_sanitized.py: the oracle alone under AddressSanitizer / UBSan). No torch, no GPU."""
import numpy as np

from enterpolation_b200 import Bezier, BSpline, Linear, easing as E

DT = {"f32": np.float32, "f64": np.float64}


def _elements(rng, b, n, dim, dtype, weighted):
    el = (rng.random((n, dim)) * 2 - 1).astype(dtype)
    el = el if dim > 1 else el[:, 0]
    if weighted:
        w = (0.25 + rng.random(n)).astype(dtype)
        if rng.random() < 0.1:
            w[rng.integers(0, n)] = 0  # a zero weight is a direction (homogeneous.rs:83-88)
        return b.elements_with_weights(el, w)
    return b.elements(el)


def _sorted_knots(rng, n, dtype):
    k = np.cumsum(0.05 + rng.random(n) * rng.choice([1.0, 1.0, 40.0]))
    if rng.random() < 0.3 and n > 6:
        for i in rng.integers(1, n - 1, size=max(1, n // 10)):
            k[i] = k[i - 1]  # repeated knots
        k = np.sort(k)
    return (k * rng.choice([1.0, 1e-3, 1e4]) + rng.choice([0.0, -3.0, 1e3])).astype(dtype)


def _random_curve(rng):
    scalar = rng.choice(["f32", "f64"])
    dtype = DT[scalar]
    dim = int(rng.integers(1, 5))
    weighted = bool(rng.random() < 0.25)
    kind = rng.choice(["linear", "bezier", "bspline"], p=[0.35, 0.15, 0.5])
    size = int(rng.choice([3, 5, 9, 33, 200, 600, 5000, 9000], p=[0.15, 0.2, 0.2, 0.15, 0.1, 0.08, 0.07, 0.05]))
    if kind == "linear":
        n = max(2, size)
        b = _elements(rng, Linear.builder(), n, dim, dtype, weighted)
        how = rng.choice(["sorted", "domain", "normalized", "distance"])
        if how == "sorted":
            b = b.knots(_sorted_knots(rng, n, dtype))
        else:
            b = b.equidistant(scalar)
            b = {"domain": lambda: b.domain(float(rng.normal()), float(rng.normal()) + 3.5),
                 "normalized": lambda: b.normalized(),
                 "distance": lambda: b.distance(float(rng.normal()), float(rng.choice([0.125, 0.1, 3.0, 1e-4])))}[how]()
        if rng.random() < 0.3:
            b = b.easing(rng.choice([E.flip, E.smoothstep, E.smootherstep, E.smoothstart(3), E.smoothend(2), E.Plateau(0.35)]))
    elif kind == "bezier":
        n = int(rng.choice([1, 2, 3, 4, 6, 8, 9, 16, 17, 30]))
        b = _elements(rng, Bezier.builder(), n, dim, dtype, weighted)
        b = b.domain(float(rng.normal()), float(rng.normal()) + 4.0, scalar) if rng.random() < 0.4 else b.normalized(scalar)
        b = b.constant() if rng.random() < 0.5 else b.dynamic()
    else:
        degree = int(rng.choice([1, 2, 3, 3, 3, 4, 5, 7, 8]))
        n = max(degree + 2, size)
        mode = rng.choice(["open", "clamped", "legacy"], p=[0.3, 0.5, 0.2])
        b = BSpline.builder()
        b = {"open": b.open, "clamped": b.clamped, "legacy": b.legacy}[mode]()
        b = _elements(rng, b, n, dim, dtype, weighted)
        if mode == "legacy" or rng.random() < 0.5:
            nk = {"open": n + degree - 1, "clamped": n - degree + 1, "legacy": n + degree + 1}[mode]
            b = b.knots(_sorted_knots(rng, nk, dtype))
        else:
            b = b.equidistant(scalar).degree(degree)
            b = b.normalized() if rng.random() < 0.4 else b.domain(float(rng.normal()), float(rng.normal()) + 3.5)
        b = b.dynamic() if rng.random() < 0.6 else b.constant(degree + 1)
    return scalar, b


def _weird(rng, dtype):
    """Values the builders accept without a check: zeros, denormals, huge magnitudes, infinities."""
    fi = np.finfo(dtype)
    return float(rng.choice([0.0, -0.0, fi.tiny, -fi.tiny, fi.tiny / 4, fi.max / 4, -fi.max / 4, fi.max, np.inf, -np.inf, 1.0, -1.0,
                             fi.eps, 1e-3, 1e4, -7.25]))


def _degenerate_curve(rng):
    """Curves whose numbers make the reference's arithmetic degenerate (descending or zero-width knot vectors, infinite and
    huge knots, all-equal knots, denormal spacings, zero / negative / infinite weights, non-finite elements): the
    builders upstream accept all of them, evaluation is whatever IEEE makes of it or a panic (-> NaN + invalid count)."""
    scalar = rng.choice(["f32", "f64"])
    dtype = DT[scalar]
    dim = int(rng.integers(1, 5))
    kind = rng.choice(["linear", "bezier", "bspline"], p=[0.4, 0.1, 0.5])
    n = int(rng.choice([2, 3, 5, 9, 33, 40, 200, 5000]))
    el = (rng.random((n, dim)) * 2 - 1).astype(dtype)
    if rng.random() < 0.3:
        el[rng.integers(0, n), rng.integers(0, dim)] = _weird(rng, dtype)
    el = el if dim > 1 else el[:, 0]
    w = None
    if rng.random() < 0.3:
        w = (0.25 + rng.random(n)).astype(dtype)
        for _ in range(int(rng.integers(0, 3))):
            w[rng.integers(0, n)] = dtype(_weird(rng, dtype))

    def with_elements(b):
        return b.elements_with_weights(el, w) if w is not None else b.elements(el)

    def weird_sorted(m):
        how = rng.choice(["equal", "inf_ends", "huge", "denormal", "clustered", "two_values"])
        if how == "equal":
            k = np.full(m, rng.normal())
        elif how == "inf_ends":
            k = np.cumsum(0.05 + rng.random(m))
            k[0] = -np.inf if rng.random() < 0.6 else k[0]
            k[-1] = np.inf if rng.random() < 0.6 else k[-1]
        elif how == "huge":
            k = np.linspace(-1.0, 1.0, m) * float(np.finfo(dtype).max) * float(rng.choice([1.0, 0.6, 0.25]))
        elif how == "denormal":
            k = np.arange(m) * float(np.finfo(dtype).tiny) * float(rng.choice([0.25, 1.0, 3.0]))
        elif how == "clustered":
            k = np.sort(np.concatenate([np.full(m - m // 2, 1.0) + rng.random(m - m // 2) * 1e-6, rng.random(m // 2) * 100]))
        else:
            k = np.sort(rng.choice([0.0, 1.0], size=m))
        return np.asarray(k, dtype=np.float64).astype(dtype)

    def weird_equidistant(b):
        how = rng.choice(["domain", "distance"])
        if how == "domain":
            a = float(rng.normal())
            return b.domain(a, float(rng.choice([a, a - 1.5, a + _weird(rng, dtype), _weird(rng, dtype)])))
        return b.distance(float(rng.choice([0.0, rng.normal(), _weird(rng, dtype)])), _weird(rng, dtype))

    if kind == "linear":
        b = with_elements(Linear.builder())
        b = b.knots(weird_sorted(n)) if rng.random() < 0.5 else weird_equidistant(b.equidistant(scalar))
        if rng.random() < 0.2:
            b = b.easing(rng.choice([E.smoothstep, E.smoothstart(3), E.Plateau(0.35)]))
    elif kind == "bezier":
        n = min(n, 17)
        el = el[:n]
        w = None if w is None else w[:n]
        b = with_elements(Bezier.builder())
        a = float(rng.normal())
        b = b.domain(a, float(rng.choice([a, a - 2.0, _weird(rng, dtype)])), scalar).dynamic()
    else:
        degree = int(rng.choice([1, 2, 3, 3, 4, 5, 8]))
        if n <= degree + 1:
            n = degree + 2
            el = (rng.random((n, dim)) * 2 - 1).astype(dtype)
            el = el if dim > 1 else el[:, 0]
            w = None
        mode = rng.choice(["open", "clamped", "legacy"], p=[0.4, 0.4, 0.2])
        b = BSpline.builder()
        b = with_elements({"open": b.open, "clamped": b.clamped, "legacy": b.legacy}[mode]())
        if mode == "legacy" or rng.random() < 0.5:
            nk = {"open": n + degree - 1, "clamped": n - degree + 1, "legacy": n + degree + 1}[mode]
            b = b.knots(weird_sorted(nk))
        else:
            b = weird_equidistant(b.equidistant(scalar).degree(degree))
        b = b.dynamic()
    return scalar, b


def _degenerate_inputs(rng, dtype, lo, hi, m):
    fi = np.finfo(dtype)
    centre = lo if np.isfinite(lo) else (hi if np.isfinite(hi) else 0.0)
    width = hi - lo if np.isfinite(hi - lo) and hi != lo else 1.0
    with np.errstate(all="ignore"):
        t = (centre + (rng.random(m) * 1.6 - 0.3) * width).astype(dtype)
    special = np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, lo, hi, fi.max, -fi.max, fi.tiny, -fi.tiny / 4, 1.0], dtype=dtype)
    t[rng.integers(0, m, size=min(m, 24))] = rng.choice(special, size=min(m, 24))
    return t


def _random_const_equidistant(rng):
    """ConstEquidistant knots (list.rs:566-721): no builder chain reaches them, so _random_curve never does."""
    scalar = rng.choice(["f32", "f64"])
    dtype = DT[scalar]
    dim = int(rng.integers(1, 5))
    n = int(rng.choice([2, 3, 9, 33, 200, 5000]))
    el = (rng.random((n, dim)) * 2 - 1).astype(dtype)
    el = el if dim > 1 else el[:, 0]
    if rng.random() < 0.5:
        return scalar, Linear.equidistant_unchecked(el, scalar)
    degree = int(rng.choice([1, 2, 3, 4, 6]))
    if n <= degree:
        return scalar, Linear.equidistant_unchecked(el, scalar)
    w = (0.25 + rng.random(n)).astype(dtype) if rng.random() < 0.3 else None
    return scalar, BSpline.new_const_equidistant(el, n + degree - 1, scalar, workspace=None if rng.random() < 0.5 else degree + 1, weights=w)
