"""Seeded random differential test: random curve configurations x random evaluation calls, CUDA path vs CPU oracle,
bit for bit. Complements the structured cases of test_gpu_parity.py by mixing features the kernels dispatch on
(table sizes around the shared-memory / bucket-index / lean-kernel thresholds, adaptors, easings, weights, windows)."""
import os

import numpy as np
import pytest
import torch

from enterpolation_b200 import Bezier, BSpline, Linear, easing as E
from oracle import oracle

from util import assert_bits_equal, u01

pytestmark = pytest.mark.gpu

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


def _check(rng, scalar, b, tag):
    dtype = DT[scalar]
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    d = r.domain()
    lo, hi = float(d[0]), float(d[1])
    span = hi - lo if np.isfinite(hi - lo) and hi > lo else 1.0
    variants = [(c, r, "plain")]
    if rng.random() < 0.3:
        variants.append((c.clamp(), r.clamp(), "clamp"))
    if rng.random() < 0.3:
        a, bb = lo + span * float(rng.random()) * 0.5, hi + span * float(rng.normal()) * 0.3
        variants.append((c.slice(a, bb), r.slice(a, bb), "slice"))
    for cv, rv, what in variants:
        n = int(rng.choice([1, 2, 33, 257, 4099, 70001]))
        total = n + int(rng.choice([0, 0, 12345, (1 << 32) + 77]))
        first = int(rng.integers(0, total - n + 1))
        before = c.invalid_count()
        got = cv.take_device(total, first, n).cpu().numpy()
        want, inv = rv.take(total, first, n, return_invalid=True)
        assert_bits_equal(got, want, f"{tag} {what} take({total})[{first}:+{n}]")
        assert c.invalid_count() - before == inv, f"{tag} {what} take invalid count"
        m = int(rng.choice([1, 31, 1000, 20011]))
        t = (rng.random(m) * 1.6 * span + (lo - 0.3 * span)).astype(dtype)
        if rng.random() < 0.3:
            t[rng.integers(0, m, size=min(m, 5))] = rng.choice([np.nan, np.inf, -np.inf, 0.0, lo, hi])
        before = c.invalid_count()
        got = cv.sample_device(torch.from_numpy(t).cuda()).cpu().numpy()
        want, inv = rv.eval(t, return_invalid=True)
        assert_bits_equal(got, want, f"{tag} {what} batch({m})")
        assert c.invalid_count() - before == inv, f"{tag} {what} batch invalid count"


# ENTERP_FUZZ_SEEDS="first:count" widens the sweep (e.g. "24:400" on a spare GPU minute); the default 24 seeds x 12
# curves are what every `pytest -m gpu` run covers.
_FIRST, _COUNT = (int(v) for v in os.environ.get("ENTERP_FUZZ_SEEDS", "0:24").split(":"))


@pytest.mark.parametrize("seed", range(_FIRST, _FIRST + _COUNT))
def test_random_curves(seed):
    rng = np.random.default_rng(0xE17E + seed)
    done = 0
    while done < 12:
        scalar, b = _random_curve(rng)
        try:
            b.build()
        except Exception:
            continue  # the random shape violated a builder rule (covered by test_builder_errors.py)
        _check(rng, scalar, b, f"seed {seed} case {done}")
        done += 1


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


# Windows of long takes laid over the indices where an f32 stepper changes its run length (2^24 .. 2^32, rows.cuh RUNS,
# capi.cu generic_runs_take / run-length host transfer), for random curves, through the device AND the host-buffer entry
# points (the latter decide per chunk between the plain and the run-length pipeline).
@pytest.mark.parametrize("seed", range(_FIRST // 2, _FIRST // 2 + max(1, _COUNT // 2)))
def test_random_windows_over_run_boundaries(seed):
    from enterpolation_b200.curve import Take
    rng = np.random.default_rng(0xB0DA + seed)
    done = 0
    while done < 6:
        scalar, b = _random_const_equidistant(rng) if rng.random() < 0.2 else _random_curve(rng)
        try:
            c = b.build()
        except Exception:
            continue
        r = oracle.OracleCurve(b.to_desc())
        variants = [(c, r, "plain")]
        if rng.random() < 0.25:
            variants.append((c.clamp(), r.clamp(), "clamp"))
        for cv, rv, what in variants:
            k = int(rng.integers(24, 33))
            n = int(rng.choice([4099, 70001, 300000]))
            first = (1 << k) - int(rng.integers(0, n + 1))
            total = first + n + int(rng.choice([0, 1, 12345, 1 << k]))
            tag = f"seed {seed} case {done} {what} take({total})[{first}:+{n}]"
            want, inv = rv.take(total, first, n, return_invalid=True)
            before = c.invalid_count()
            got = cv.take_device(total, first, n).cpu().numpy()
            assert_bits_equal(got, want, tag + " device")
            assert c.invalid_count() - before == inv, tag + " invalid count"
            got = Take(cv, total, first, n).collect()
            assert_bits_equal(got, want, tag + " host")
            m = int(rng.choice([1, 1000, 20011]))
            d = rv.domain()
            lo, hi = float(d[0]), float(d[1])
            span = hi - lo if np.isfinite(hi - lo) and hi > lo else 1.0
            t = (rng.random(m) * 1.6 * span + (lo - 0.3 * span)).astype(DT[scalar])
            want, inv = rv.eval(t, return_invalid=True)
            assert_bits_equal(np.asarray(cv.sample(t)), want, tag + f" host batch({m})")
        done += 1
