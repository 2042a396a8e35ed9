"""Differential parity: CUDA path (through the C ABI) vs the CPU oracle on seeded random and
adversarial inputs. Bar (BASELINE.json north_star): span indices bit-exact, f64 within 4 ulp, f32
within 1e-6 relative. The kernels are built from unfused IEEE operations in the reference's order,
so the tests demand the stronger property: bit-for-bit equality (any NaN == any NaN)."""
import itertools

import numpy as np
import pytest
import torch

from enterpolation_b200 import Bezier, BSpline, Linear, _abi, bezier_many_take, stepper_device
from oracle import oracle

from util import assert_bits_equal, u01

pytestmark = pytest.mark.gpu

DT = {"f32": np.float32, "f64": np.float64}


def adversarial(knots, dtype, rng, n_random=3000):
    """Exact knots, their nextafter neighbours, out-of-domain values, +-0, denormals, huge values."""
    k = np.unique(np.asarray(knots, dtype=dtype))
    lo, hi = k[0], k[-1]
    span = max(float(hi - lo), 1.0)
    parts = [k, np.nextafter(k, dtype(np.inf)), np.nextafter(k, dtype(-np.inf)),
             np.array([0.0, -0.0, np.finfo(dtype).tiny, -np.finfo(dtype).tiny, np.finfo(dtype).tiny / 4,
                       lo - span, hi + span, lo - 1e6, hi + 1e6, np.finfo(dtype).max / 8, -np.finfo(dtype).max / 8],
                      dtype=dtype),
             rng.uniform(float(lo) - 0.25 * span, float(hi) + 0.25 * span, n_random).astype(dtype)]
    return np.concatenate(parts).astype(dtype)


def check_curve(builder, t, take_n=777, what=""):
    curve = builder.build()
    ref = oracle.OracleCurve(builder.to_desc())
    tt = torch.from_numpy(np.ascontiguousarray(t)).cuda()
    before = curve.invalid_count()
    got = curve.sample_device(tt).cpu().numpy()
    want, inv = ref.eval(t, return_invalid=True)
    assert_bits_equal(got, want, what + " batch")
    assert curve.invalid_count() - before == inv, what + " invalid counter"
    idx, fac = curve.span_device(tt)
    widx, wfac = ref.span(t)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx), what + " span indices"
    assert_bits_equal(fac.cpu().numpy(), wfac, what + " span factor")
    # take(): whole, and a shard with a non-zero global offset
    got = curve.take_device(take_n).cpu().numpy()
    assert_bits_equal(got, ref.take(take_n), what + " take")
    lo, cnt = take_n // 3, take_n // 2
    assert_bits_equal(curve.take_device(take_n, lo, cnt).cpu().numpy(), ref.take(take_n, lo, cnt), what + " take shard")
    # host-buffer entry points (the call a user of the reference makes)
    assert_bits_equal(np.asarray(curve.sample(t)), want, what + " sample_host")
    assert_bits_equal(np.asarray(curve.take(take_n).collect()), ref.take(take_n), what + " take_host")
    return curve, ref


def rand_elements(n, dim, dtype, seed):
    e = (u01(n * dim, seed, dtype) * dtype(2) - dtype(1)).astype(dtype)
    return e.reshape(n, dim) if dim > 1 else e


def sorted_knots(n, dtype, seed, repeat_every=0):
    k = np.cumsum(0.5 + u01(n, seed, np.float64)).astype(dtype)
    if repeat_every:
        for i in range(repeat_every, n, repeat_every):
            k[i] = k[i - 1]
    return k


@pytest.mark.parametrize("scalar,dim,weighted", [(s, d, w) for s in ("f32", "f64") for d in (1, 2, 3, 4) for w in (0, 1)])
@pytest.mark.parametrize("degree", [1, 2, 3, 5, 7, 9])
def test_bspline_sorted_open(scalar, dim, weighted, degree):
    dtype = DT[scalar]
    rng = np.random.default_rng(degree * 100 + dim * 10 + weighted)
    e = degree + 12
    k = e + degree - 1
    knots = sorted_knots(k, dtype, 0xE2 + degree, repeat_every=5)
    el = rand_elements(e, dim, dtype, 0xE1 + dim)
    b = BSpline.builder()
    b = b.elements_with_weights(el, (0.5 + u01(e, 0xE3, dtype)).astype(dtype)) if weighted else b.elements(el)
    b = b.knots(knots)
    b = b.constant(degree + 1) if degree % 2 else b.dynamic()
    check_curve(b, adversarial(knots, dtype, rng), what=f"bspline sorted {scalar} d{dim} w{weighted} p{degree}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["clamped", "legacy"])
@pytest.mark.parametrize("degree", [1, 2, 3, 4])
def test_bspline_sorted_modes(scalar, mode, degree):
    dtype = DT[scalar]
    rng = np.random.default_rng(degree)
    e = degree + 9
    if mode == "clamped":
        k = e - degree + 1  # BorderBuffer n = e - k = degree - 1
        b = BSpline.builder().clamped()
    else:
        k = e + degree + 1  # legacy: textbook knot vector
        b = BSpline.builder().legacy()
    knots = sorted_knots(k, dtype, 0xA0 + degree)
    b = b.elements(rand_elements(e, 3, dtype, 0xA1)).knots(knots).constant(degree + 1)
    curve, _ = check_curve(b, adversarial(knots, dtype, rng), what=f"bspline {mode} {scalar} p{degree}")
    assert curve.degree == degree


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("mode", ["open", "clamped"])
@pytest.mark.parametrize("form", ["domain", "normalized", "distance"])
@pytest.mark.parametrize("degree,dim,weighted", [(1, 1, 0), (2, 2, 1), (3, 4, 0), (3, 3, 1), (4, 1, 1), (6, 2, 0)])
def test_bspline_equidistant(scalar, mode, form, degree, dim, weighted):
    dtype = DT[scalar]
    rng = np.random.default_rng(degree * 7 + dim)
    e = degree + 8
    el = rand_elements(e, dim, dtype, 0xB1 + degree)
    b = BSpline.builder()
    b = b.clamped() if mode == "clamped" else b.open()
    b = b.elements_with_weights(el, (0.5 + u01(e, 0xB3, dtype)).astype(dtype)) if weighted else b.elements(el)
    b = b.equidistant(scalar).degree(degree)
    if form == "domain":
        b = b.domain(-2.0, 3.0)
    elif form == "normalized":
        b = b.normalized()
    else:
        b = b.distance(0.25, 0.3)
    b = b.constant(degree + 1) if degree != 2 else b.dynamic()
    ref = oracle.OracleCurve(b.to_desc())
    d = ref.domain()
    n_inner = e - 1 + degree if mode == "open" else e - degree + 1
    approx_knots = np.linspace(float(d[0]), float(d[1]), 4 * n_inner + 1)
    t = adversarial(approx_knots, dtype, rng)
    t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype)])  # reference panics on NaN/+inf
    check_curve(b, t, what=f"bspline equi {scalar} {mode} {form} p{degree} d{dim} w{weighted}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("dim,weighted", [(1, 0), (3, 0), (2, 1), (4, 1)])
@pytest.mark.parametrize("knots_kind", ["sorted", "sorted_repeats", "domain", "normalized", "distance", "const"])
def test_linear(scalar, dim, weighted, knots_kind):
    dtype = DT[scalar]
    rng = np.random.default_rng(dim)
    e = 37
    el = rand_elements(e, dim, dtype, 0xC1)
    b = Linear.builder()
    if knots_kind == "const":  # ConstEquidistantLinear (linear/mod.rs:193-216) has no weighted constructor
        if weighted:
            pytest.skip("equidistant_unchecked takes plain elements")
        b = Linear.equidistant_unchecked(el, scalar)
        t = adversarial(np.linspace(0.0, 1.0, 4 * (e - 1) + 1), dtype, rng)
        t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
        check_curve(b, t, what=f"linear {scalar} d{dim} const-equidistant")
        return
    b = b.elements_with_weights(el, (0.5 + u01(e, 0xC3, dtype)).astype(dtype)) if weighted else b.elements(el)
    if knots_kind.startswith("sorted"):
        knots = sorted_knots(e, dtype, 0xC2, repeat_every=4 if knots_kind == "sorted_repeats" else 0)
        if knots_kind == "sorted_repeats":
            knots[1] = knots[0]
            knots[-2] = knots[-1]  # coincident edge knots: linear_factor returns the zero difference
        b = b.knots(knots)
        t = adversarial(knots, dtype, rng)
        t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
    else:
        b = b.equidistant(scalar)
        b = {"domain": lambda: b.domain(-1.0, 4.0), "normalized": lambda: b.normalized(),
             "distance": lambda: b.distance(2.0, 0.125)}[knots_kind]()
        d = oracle.OracleCurve(b.to_desc()).domain()
        t = adversarial(np.linspace(float(d[0]), float(d[1]), 4 * (e - 1) + 1), dtype, rng)
        t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
    check_curve(b, t, what=f"linear {scalar} d{dim} w{weighted} {knots_kind}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 8, 16, 17, 40])
@pytest.mark.parametrize("dim,weighted,domain", [(1, 0, None), (3, 0, None), (2, 1, None), (4, 1, (0.0, 2.0)), (3, 0, (1.0, 3.0))])
def test_bezier(scalar, n, dim, weighted, domain):
    dtype = DT[scalar]
    rng = np.random.default_rng(n)
    el = rand_elements(n, dim, dtype, 0xD1 + n)
    b = Bezier.builder()
    b = b.elements_with_weights(el, (0.5 + u01(n, 0xD3, dtype)).astype(dtype)) if weighted else b.elements(el)
    b = b.domain(domain[0], domain[1], scalar) if domain else b.normalized(scalar)
    b = b.constant() if n % 2 else b.dynamic()
    t = adversarial([0.0, 0.5, 1.0, 2.0, 3.0], dtype, rng, n_random=1500)
    check_curve(b, t, what=f"bezier {scalar} n{n} d{dim} w{weighted} dom{domain}")


def test_negative_zero_weight_becomes_zero():
    """`weighted_or_infinite` (homogeneous.rs:83-88) sends a weight that `is_zero()` — -0.0 included — to
    `Homogeneous::infinity`, whose rational is `R::zero()` = +0.0: the sign of the resulting infinities follows the element,
    not the weight's sign bit (found by the degenerate fuzz sweep: -inf upstream, +inf here)."""
    for scalar in ("f32", "f64"):
        dtype = DT[scalar]
        el = np.array([[-0.93, -0.28], [-0.95, -0.76], [0.72, 0.38]], dtype=dtype)
        w = np.array([-0.0, 0.78, 0.38], dtype=dtype)
        t = np.array([-0.0, 0.0, -1e-39, 0.25, 1.0], dtype=dtype)
        for b in (Linear.builder().elements_with_weights(el, w).equidistant(scalar).normalized(),
                  Bezier.builder().elements_with_weights(el, w).normalized(scalar).constant(),
                  BSpline.builder().clamped().elements_with_weights(el, w).equidistant(scalar).degree(2).normalized().dynamic()):
            curve, ref = check_curve(b, t, take_n=5, what=f"-0.0 weight {type(b).__name__} {scalar}")
            assert np.all(np.asarray(curve.sample(t[:1])) == -np.inf)


def test_weight_zero_is_a_direction():  # homogeneous.rs:83-88; bezier/builder.rs:237-245 ends in inf
    b = Bezier.builder().elements_with_weights([(1.0, 1.0), (2.0, 4.0), (3.0, 0.0)]).normalized("f64").constant()
    check_curve(b, np.linspace(-0.5, 1.5, 201), what="weight zero")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
def test_stepper_matches_oracle(scalar):
    sc = _abi.F32 if scalar == "f32" else _abi.F64
    for (a, b, n, first, count) in [(0.0, 1.0, 11, 0, 11), (3.0, 5.0, 5, 0, 5), (-2.0, 2.0, 1 << 24, (1 << 24) - 5000, 5000),
                                    (0.0, 1.0, 1 << 32, (1 << 32) - 4096, 4096), (0.0, 1.0, 1 << 32, (1 << 31) + 12345, 4096),
                                    (0.0, 1.0, 1, 0, 1)]:
        got = stepper_device(sc, a, b, n, first, count).cpu().numpy()
        assert_bits_equal(got, oracle.stepper(sc, a, b, n, first, count), f"stepper {scalar} n={n}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n_ctrl,dim", [(16, 3), (4, 2), (2, 1), (1, 4), (20, 3)])
def test_bezier_many(scalar, n_ctrl, dim):
    dtype = DT[scalar]
    n_curves, samples = 257, 100
    ctrl = rand_elements(n_curves * n_ctrl, dim, dtype, 0xF1).reshape(n_curves, n_ctrl, dim)
    got = bezier_many_take(torch.from_numpy(ctrl).cuda(), samples).cpu().numpy()
    assert_bits_equal(got, oracle.bezier_many_take(ctrl, samples), f"bezier_many {scalar} n{n_ctrl} d{dim}")
    # each curve equals the single-curve builder path
    c0 = Bezier.builder().elements(ctrl[5] if dim > 1 else ctrl[5, :, 0]).normalized(scalar).constant().build()
    assert_bits_equal(np.asarray(c0.take(samples).collect()).reshape(samples, dim), got[5], "many == single")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n_ctrl,dim", [(1, 3), (2, 2), (3, 4), (4, 3), (4, 1), (8, 2), (9, 3), (16, 3)])
@pytest.mark.parametrize("samples", [2, 3, 4, 7, 31, 32, 33, 64, 127, 128, 300])
def test_bezier_many_small_sample_counts(scalar, n_ctrl, dim, samples):
    """Few samples per curve take the flattened (curve, sample) kernel with its multiply-shift division; larger counts
    the warp-per-curve kernel (control values hoisted for small curves). Odd curve counts leave ragged tails."""
    dtype = DT[scalar]
    n_curves = 1031
    ctrl = rand_elements(n_curves * n_ctrl, dim, dtype, 0xF7 + samples).reshape(n_curves, n_ctrl, dim)
    got = bezier_many_take(torch.from_numpy(ctrl).cuda(), samples).cpu().numpy()
    assert_bits_equal(got, oracle.bezier_many_take(ctrl, samples), f"bezier_many {scalar} n{n_ctrl} d{dim} s{samples}")
    # a misaligned control array keeps the scalar-load kernel
    buf = torch.empty(ctrl.size + 1, dtype=torch.from_numpy(ctrl).dtype, device="cuda")
    view = buf[1:].view(n_curves, n_ctrl, dim)
    view.copy_(torch.from_numpy(ctrl))
    assert_bits_equal(bezier_many_take(view, samples).cpu().numpy(), got, "misaligned ctrl")


def test_large_sorted_search_levels():
    """4-level pivot pyramid (70 000 knots): exact knots + neighbours + random, f32 Linear [f32;3]."""
    dtype = np.float32
    n = 70_000
    knots = sorted_knots(n, dtype, 0xE2)
    el = rand_elements(n, 3, dtype, 0xE1)
    b = Linear.builder().elements(el).knots(knots)
    rng = np.random.default_rng(0)
    pick = rng.integers(0, n, 4000)
    t = np.concatenate([knots[pick], np.nextafter(knots[pick], dtype(np.inf)), np.nextafter(knots[pick], dtype(-np.inf)),
                        rng.uniform(knots[0] - 10, knots[-1] + 10, 50_000).astype(dtype),
                        knots[:40], knots[-40:], np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
    check_curve(b, t.astype(dtype), take_n=5000, what="large sorted linear")


def test_config_shapes_small():
    """BASELINE.json configs at sizes the oracle finishes in seconds."""
    # C1: README cardinal cubic, f64, take
    b = BSpline.builder().clamped().elements([0.0, 0.0, 0.0, 6.0, 0.0, 0.0, 0.0]).equidistant("f64").degree(3) \
        .domain(-2.0, 2.0).constant(4)
    c = b.build()
    ref = oracle.OracleCurve(b.to_desc())
    n = 1 << 18
    assert_bits_equal(c.take_device(n).cpu().numpy(), ref.take(n, threads=8), "C1")
    idx, _ = c.span_take_device(n)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), ref.span_take(n)[0])
    # C5: dynamic clamped equidistant cubic [f32;4], global n = 2^32, windows at both ends and the middle
    for stops in (5, 256):
        el = u01(stops * 4, 0xE1, np.float32).reshape(stops, 4)
        b = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic()
        c = b.build()
        ref = oracle.OracleCurve(b.to_desc())
        for first in (0, (1 << 31) - 50_000, (1 << 32) - 100_000):
            assert_bits_equal(c.take_device(1 << 32, first, 100_000).cpu().numpy(), ref.take(1 << 32, first, 100_000, threads=8),
                              f"C5 stops={stops} first={first}")


# ---- input adaptors and built-in easings ---------------------------------------
def check_adapted(curve, ref, t, what, take_n=333):
    tt = torch.from_numpy(np.ascontiguousarray(t)).cuda()
    assert_bits_equal(curve.sample_device(tt).cpu().numpy(), ref.eval(t), what + " batch")
    assert_bits_equal(curve.take_device(take_n).cpu().numpy(), ref.take(take_n), what + " take")
    assert_bits_equal(np.asarray(curve.take(take_n).collect()), ref.take(take_n), what + " take_host")
    idx, fac = curve.span_device(tt)
    widx, wfac = ref.span(t)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx), what + " span"
    assert [float(x) for x in curve.domain()] == [float(x) for x in ref.domain()], what + " domain"


@pytest.mark.parametrize("scalar", ["f32", "f64"])
def test_clamp_and_slice_adaptors(scalar):
    from enterpolation_b200 import easing as E  # noqa: F401
    dtype = DT[scalar]
    rng = np.random.default_rng(3)
    el = rand_elements(9, 3, dtype, 0x51)
    builders = {
        "bspline_equi": BSpline.builder().clamped().elements(el).equidistant(scalar).degree(3).domain(-1.0, 2.0).constant(4),
        "bspline_sorted": BSpline.builder().elements(el).knots(sorted_knots(10, dtype, 0x52)).constant(3),
        "linear": Linear.builder().elements(el).knots(sorted_knots(9, dtype, 0x53)),
        "bezier_dom": Bezier.builder().elements(el).domain(0.0, 2.0, scalar).constant(),
        "nurbs": BSpline.builder().elements_with_weights(el, (0.5 + u01(9, 0x54, dtype)).astype(dtype)).equidistant(scalar)
        .degree(2).normalized().dynamic(),
    }
    for name, b in builders.items():
        base, rbase = b.build(), oracle.OracleCurve(b.to_desc())
        d = rbase.domain()
        lo, hi = float(d[0]), float(d[1])
        t = adversarial(np.linspace(lo, hi, 9), dtype, rng, n_random=800)
        mid0, mid1 = lo + 0.25 * (hi - lo), lo + 0.6 * (hi - lo)
        check_adapted(base.clamp(), rbase.clamp(), t, f"{name} clamp {scalar}")
        check_adapted(base.slice(mid0, mid1), rbase.slice(mid0, mid1), t, f"{name} slice {scalar}")
        check_adapted(base.slice(None, mid1), rbase.slice(None, mid1), t, f"{name} slice-open {scalar}")
        # nesting: clamp of a slice, slice of a clamp
        check_adapted(base.slice(mid0, mid1).clamp(), rbase.slice(mid0, mid1).clamp(), t, f"{name} slice.clamp {scalar}")
        check_adapted(base.clamp().slice(mid0, mid1), rbase.clamp().slice(mid0, mid1), t, f"{name} clamp.slice {scalar}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("knots_kind", ["sorted", "equidistant"])
def test_linear_easings(scalar, knots_kind):
    from enterpolation_b200 import easing as E
    dtype = DT[scalar]
    rng = np.random.default_rng(5)
    el = rand_elements(6, 2, dtype, 0x61)
    easings = [E.Identity(), E.flip, E.smoothstart(1), E.smoothstart(2), E.smoothstart(5), E.smoothend(2), E.smoothend(3),
               E.smoothstep, E.smootherstep, E.Plateau(0.7), E.Plateau(0.0), E.Plateau(1.0)]
    for ez in easings:
        b = Linear.builder().elements(el)
        if knots_kind == "sorted":
            kn = sorted_knots(6, dtype, 0x62)
            b = b.knots(kn)
            t = adversarial(kn, dtype, rng, n_random=1500)
        else:
            b = b.equidistant(scalar).domain(-1.0, 1.0)  # examples/plateaus.rs:35-41
            t = adversarial(np.linspace(-1.0, 1.0, 21), dtype, rng, n_random=1500)
        b = b.easing(ez)
        check_adapted(b.build(), oracle.OracleCurve(b.to_desc()), t, f"easing {ez} {knots_kind} {scalar}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 8, 9, 16, 17, 31, 32, 33, 63, 64, 65, 255, 256, 257, 1025, 4100])
def test_pivot_tree_shapes(scalar, n):
    """Every node-count corner of the pivot tree (full / partial last nodes, 1..5 levels), with x below,
    at and beyond every knot — in particular x >= all knots, where every upper-level key is <= x."""
    dtype = DT[scalar]
    knots = sorted_knots(n, dtype, 0x71 + n)
    el = rand_elements(n, 1, dtype, 0x72)
    b = Linear.builder().elements(el).knots(knots)
    t = np.concatenate([knots, np.nextafter(knots, dtype(np.inf)), np.nextafter(knots, dtype(-np.inf)),
                        np.array([knots[0] - 1, knots[-1] + 1, knots[-1] * 4, np.inf, -np.inf, np.nan], dtype=dtype)]).astype(dtype)
    curve, ref = b.build(), oracle.OracleCurve(b.to_desc())
    tt = torch.from_numpy(t).cuda()
    idx, fac = curve.span_device(tt)
    widx, wfac = ref.span(t)
    assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx), f"linear span n={n}"
    assert_bits_equal(curve.sample_device(tt).cpu().numpy(), ref.eval(t), f"linear eval n={n}")
    if n >= 4:  # the same knots as an open quadratic B-spline (k knots, e = k - 1 elements)
        b2 = BSpline.builder().elements(rand_elements(n - 1, 2, dtype, 0x73)).knots(knots).dynamic()
        c2, r2 = b2.build(), oracle.OracleCurve(b2.to_desc())
        idx, _ = c2.span_device(tt)
        assert np.array_equal(idx.cpu().numpy().view(np.uint32), r2.span(t)[0]), f"bspline span n={n}"
        assert_bits_equal(c2.sample_device(tt).cpu().numpy(), r2.eval(t), f"bspline eval n={n}")


# ---- edge cases: empty / ragged / misaligned / huge index ranges / concurrency --------------------------
def _c5_like(scalar="f32", stops=7, dim=4):
    dtype = DT[scalar]
    el = rand_elements(stops, dim, dtype, 0x81)
    return BSpline.builder().clamped().elements(el).equidistant(scalar).degree(3).normalized().dynamic()


def test_empty_and_single_sample():
    b = _c5_like()
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    assert c.sample(np.zeros(0, np.float32)).shape == (0, 4)
    assert c.take_device(10, 3, 0).shape == (0, 4)
    assert_bits_equal(np.asarray(c.sample(np.array([0.3], np.float32))), r.eval(np.array([0.3], np.float32)))
    # take(1): step = (end-start)/fl(0) -> the single parameter is NaN; the reference panics on it
    before = c.invalid_count()
    got = c.take_device(1).cpu().numpy()
    assert np.all(np.isnan(got)) and c.invalid_count() == before + 1
    assert np.all(np.isnan(r.take(1)))


@pytest.mark.parametrize("n", [1, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 100003])
@pytest.mark.parametrize("scalar,dim", [("f32", 4), ("f32", 3), ("f64", 1), ("f64", 2)])
def test_ragged_sizes_and_offsets(n, scalar, dim):
    """Tile tails of the 2-samples-per-trip kernel, for take (with a global offset) and batch."""
    b = _c5_like(scalar, 9, dim)
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    assert_bits_equal(c.take_device(n).cpu().numpy(), r.take(n), f"take {n}")
    total = n + 12345
    assert_bits_equal(c.take_device(total, 777, n).cpu().numpy(), r.take(total, 777, n), f"take window {n}")
    t = (u01(n, 0x82, DT[scalar]) * DT[scalar](1.5) - DT[scalar](0.25)).astype(DT[scalar])
    assert_bits_equal(c.sample_device(torch.from_numpy(t).cuda()).cpu().numpy(), r.eval(t), f"batch {n}")


def test_misaligned_output_pointer_uses_scalar_stores():
    b = _c5_like("f32", 6, 4)
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    n = 1000
    buf = torch.empty(n * 4 + 1, dtype=torch.float32, device="cuda")
    out = buf[1:].view(n, 4)  # 4-byte aligned only
    assert out.data_ptr() % 16 != 0
    c.take_device(n, out=out)
    assert_bits_equal(out.cpu().numpy(), r.take(n), "misaligned take")
    t = u01(n, 0x83, np.float32)
    c.sample_device(torch.from_numpy(t).cuda(), out=out)
    assert_bits_equal(out.cpu().numpy(), r.eval(t), "misaligned batch")


def test_f64x4_output_alignment_classes():
    """[f64;4] samples are written with one 256-bit store when the output is 32-byte aligned, two 128-bit stores at
    16-byte alignment, scalar stores below that (kernels.cuh store_elem)."""
    el = rand_elements(12, 4, np.float64, 0x84)
    for b in (Linear.builder().elements(el).equidistant("f64").normalized(),
              Bezier.builder().elements(el[:5]).normalized("f64").constant(),
              BSpline.builder().clamped().elements(el).equidistant("f64").degree(2).normalized().dynamic()):
        c, r = b.build(), oracle.OracleCurve(b.to_desc())
        n = 1000
        want = r.take(n)
        t = u01(n, 0x85, np.float64)
        for offset in (0, 2, 1):  # doubles: 32-, 16-, 8-byte aligned
            buf = torch.zeros(n * 4 + 4, dtype=torch.float64, device="cuda")
            out = buf[offset:offset + n * 4].view(n, 4)
            assert out.data_ptr() % 32 == (offset * 8) % 32
            c.take_device(n, out=out)
            assert_bits_equal(out.cpu().numpy(), want, f"take, offset {offset}")
            c.sample_device(torch.from_numpy(t).cuda(), out=out)
            assert_bits_equal(out.cpu().numpy(), r.eval(t), f"batch, offset {offset}")
            assert float(buf[offset + n * 4:].abs().sum()) == 0.0 and float(buf[:offset].abs().sum()) == 0.0


@pytest.mark.parametrize("scalar", ["f32", "f64"])
def test_huge_global_index_range(scalar):
    """take(2^40): 64-bit global indices and fl(i) beyond 2^32 (shards of a very long stepper)."""
    b = _c5_like(scalar, 8, 2)
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    n_total = 1 << 40
    for first in (0, (1 << 32) - 1000, (1 << 39) + 12345, n_total - 5000):
        assert_bits_equal(c.take_device(n_total, first, 5000).cpu().numpy(), r.take(n_total, first, 5000), f"{scalar} first={first}")


def test_concurrent_host_threads_share_a_curve():
    """`eval(&self)` re-entrancy (space.rs:63-66): several host threads drive one handle on their own streams."""
    import threading
    b = _c5_like("f32", 11, 4)
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    c.upload(0)
    want = {k: r.take(5000 + k, k, 4000) for k in range(4)}
    got, errs = {}, []

    def work(k):
        try:
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                for _ in range(20):
                    o = c.take_device(5000 + k, k, 4000, stream=s)
                s.synchronize()
                got[k] = o.cpu().numpy()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    ts = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for k in range(4):
        assert_bits_equal(got[k], want[k], f"thread {k}")


def test_degree_limits_and_unsupported():
    from enterpolation_b200 import Unsupported
    # runtime-degree path: degree 20 (DynSpace semantics), and the documented limit (degree > 63)
    for deg in (8, 20, 63):
        e = deg + 5
        knots = sorted_knots(e + deg - 1, np.float64, 0x90 + deg)
        b = BSpline.builder().elements(rand_elements(e, 2, np.float64, 0x91)).knots(knots).dynamic()
        c, r = b.build(), oracle.OracleCurve(b.to_desc())
        t = np.linspace(float(r.domain()[0]) - 0.1, float(r.domain()[1]) + 0.1, 301)
        assert_bits_equal(np.asarray(c.sample(t)), r.eval(t), f"degree {deg}")
    e = 70
    b = BSpline.builder().elements(np.zeros(e)).knots(np.arange(e + 63, dtype=np.float64)).dynamic()  # degree 64
    with pytest.raises(Unsupported):
        b.build().sample(np.array([1.0]))


# ---- Bezier tangent / derivatives ---------------------------------------------------
def test_bezier_tangent_reference_kats():  # bezier/mod.rs:345-390
    c = Bezier.builder().elements([1.0, 2.0, 3.0]).normalized("f64").constant().build()
    assert c.gen_with_tangent([0.5]).ravel().tolist() == [2.0, 2.0]
    assert c.gen_with_derivatives([0.5], 3).ravel().tolist() == [2.0, 2.0, 0.0]
    assert c.gen_with_derivatives([0.5], 5).ravel().tolist() == [2.0, 2.0, 0.0, 0.0, 0.0]
    one = Bezier.builder().elements([5.0]).normalized("f64").constant().build()
    assert one.gen_with_tangent([0.25, 0.75]).ravel().tolist() == [5.0, 0.0, 5.0, 0.0]
    from enterpolation_b200 import InvalidArgument, Unsupported
    with pytest.raises(InvalidArgument):
        c.gen_with_derivatives([0.5], 2)  # K <= len - 1 panics upstream
    with pytest.raises(Unsupported):
        Linear.builder().elements([0.0, 1.0]).knots([0.0, 1.0]).build().gen_with_tangent([0.5])


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n,dim", [(1, 1), (2, 3), (3, 2), (5, 4), (16, 3), (33, 1)])
def test_bezier_derivatives_vs_oracle(scalar, n, dim):
    dtype = DT[scalar]
    el = rand_elements(n, dim, dtype, 0xA1 + n)
    b = Bezier.builder().elements(el).normalized(scalar).constant()
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    t = np.concatenate([np.linspace(-0.5, 1.5, 101), [0.0, 1.0, -0.0]]).astype(dtype)
    assert_bits_equal(c.gen_with_tangent(t), r.gen_with_tangent(t), f"tangent n={n}")
    for k in {n, n + 1, n + 3}:
        assert_bits_equal(c.gen_with_derivatives(t, k), r.gen_with_derivatives(t, k), f"derivatives n={n} K={k}")


# ---- lean take kernels (csrc/lean.cuh): host-verified preconditions, else the generic kernels ---------------
@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("dim,weighted", [(1, 0), (2, 0), (3, 0), (4, 0), (1, 1), (3, 1), (4, 1)])
@pytest.mark.parametrize("knots_kind", ["domain", "normalized_pow2", "normalized", "distance", "tiny_step", "const"])
def test_lean_linear_take(scalar, dim, weighted, knots_kind):
    """take() of Linear over equidistant knots: every division mode of the lean kernel, whole takes, windows with
    a global offset, and windows past the 32-bit index limit (which must fall back to the generic kernel)."""
    dtype = DT[scalar]
    e = 33 if knots_kind == "normalized_pow2" else 37
    el = rand_elements(e, dim, dtype, 0xE1 + dim)
    if knots_kind == "const":
        if weighted:
            pytest.skip("equidistant_unchecked takes plain elements")
        b = Linear.equidistant_unchecked(el, scalar)
    else:
        b = Linear.builder()
        b = b.elements_with_weights(el, (0.5 + u01(e, 0xE3, dtype)).astype(dtype)) if weighted else b.elements(el)
        b = b.equidistant(scalar)
        b = {"domain": lambda: b.domain(-1.0, 4.0), "normalized": lambda: b.normalized(),
             "normalized_pow2": lambda: b.normalized(), "distance": lambda: b.distance(2.0, 0.125),
             "tiny_step": lambda: b.distance(0.0, 3e-30 if scalar == "f32" else 3e-250)}[knots_kind]()
    if (dim + weighted) % 2 == 0:  # half of the cases carry an easing (EASE instantiations)
        from enterpolation_b200 import easing as E
        b = b.easing([E.smoothstep, E.smootherstep, E.Plateau(0.6), E.smoothend(3)][dim - 1])
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    for n in (2, 3, 36, 37, 255, 257, 100003):
        assert_bits_equal(c.take_device(n).cpu().numpy(), r.take(n), f"{knots_kind} take {n}")
    # batch (Signal::sample): arbitrary parameters incl. out-of-domain, NaN, inf -> panic counter
    d = r.domain()
    t = adversarial(np.linspace(float(d[0]), float(d[1]), 4 * (e - 1) + 1), dtype, np.random.default_rng(dim))
    t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype), u01(100003, 0xE5, dtype) * dtype(d[1] - d[0]) + dtype(d[0])])
    before = c.invalid_count()
    got = c.sample_device(torch.from_numpy(t).cuda()).cpu().numpy()
    want, inv = r.eval(t, return_invalid=True)
    assert_bits_equal(got, want, f"{knots_kind} batch")
    assert c.invalid_count() - before == inv and inv >= 1
    total = 100003 + 54321
    assert_bits_equal(c.take_device(total, 777, 100003).cpu().numpy(), r.take(total, 777, 100003), "take window")
    n_total = (1 << 32) + 12345
    for first in (0, (1 << 31) + 17, (1 << 32) - (1 << 24) - 3000, (1 << 32) - 1500, n_total - 3000):
        assert_bits_equal(c.take_device(n_total, first, 3000).cpu().numpy(), r.take(n_total, first, 3000),
                          f"{knots_kind} first={first}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("dim,weighted,domain", [(1, 0, None), (2, 0, (-3.0, 5.0)), (3, 0, None), (4, 0, None),
                                                 (1, 1, None), (3, 1, (1.0, 3.0)), (4, 1, None)])
def test_lean_bezier_take(scalar, n, dim, weighted, domain):
    dtype = DT[scalar]
    el = rand_elements(n, dim, dtype, 0xF1 + n)
    b = Bezier.builder()
    b = b.elements_with_weights(el, (0.5 + u01(n, 0xF3, dtype)).astype(dtype)) if weighted else b.elements(el)
    b = b.domain(domain[0], domain[1], scalar) if domain else b.normalized(scalar)
    b = b.constant()
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    for cnt in (2, 33, 256, 100003):
        assert_bits_equal(c.take_device(cnt).cpu().numpy(), r.take(cnt), f"bezier take {cnt}")
    t = np.concatenate([adversarial([0.0, 0.5, 1.0, 2.0, 3.0], dtype, np.random.default_rng(n), n_random=500),
                        u01(100003, 0xF5, dtype) * dtype(1.5) - dtype(0.25)]).astype(dtype)
    assert_bits_equal(c.sample_device(torch.from_numpy(t).cuda()).cpu().numpy(), r.eval(t), "bezier batch")
    n_total = (1 << 32) + 12345
    for first in (5, (1 << 32) - (1 << 24) - 2000, (1 << 32) - 1000):
        assert_bits_equal(c.take_device(n_total, first, 2500).cpu().numpy(), r.take(n_total, first, 2500), f"first={first}")
    # a sliced Bezier carries two affine maps: generic kernel
    lo, hi = (0.25, 0.75) if domain is None else (domain[0] + 0.5, domain[1] - 0.5)
    s, rs = c.slice(lo, hi), r.slice(lo, hi)
    assert_bits_equal(s.take_device(1001).cpu().numpy(), rs.take(1001), "sliced bezier take")


def test_serde_round_trip_on_device():
    """A curve written in the reference's serde layout and read back evaluates bit-identically on the GPU."""
    from enterpolation_b200 import serde
    el = rand_elements(12, 4, np.float32, 0x91)
    for kind, b in (("bspline", BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).domain(-1.0, 2.0).dynamic()),
                    ("linear", Linear.builder().elements(el).knots(sorted_knots(12, np.float32, 0x92))),
                    ("bezier", Bezier.builder().elements(el[:5]).domain(0.5, 2.5, "f32").constant())):
        back = serde.loads(kind, serde.dumps(b), "f32")
        assert_bits_equal(back.build().take_device(10007).cpu().numpy(), b.build().take_device(10007).cpu().numpy(), kind)
        assert_bits_equal(back.build().take_device(10007).cpu().numpy(), oracle.OracleCurve(b.to_desc()).take(10007), kind)


# ---- chunked global-rows take variant (rows.cuh VARIANT 3) ------------------------------------------------------
@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("knots_kind", ["sorted", "sorted_repeats", "equidistant", "equidistant_pow2"])
@pytest.mark.parametrize("stops,degree,dim,weighted", [(9, 2, 2, 1), (40, 3, 4, 0), (700, 3, 4, 0), (3000, 3, 1, 1), (5000, 5, 3, 0)])
def test_rows_take_chunked(scalar, knots_kind, stops, degree, dim, weighted):
    """take() of B-splines through the span-row table in global memory: tables beyond the 64 KB shared-memory
    budget (equidistant and sorted) and, for sorted knots of any size, the search-skipping hold test. Sample counts
    put anything from many samples per span to several spans per sample step; windows start mid-take."""
    dtype = DT[scalar]
    el = rand_elements(stops, dim, dtype, 0xA1 + stops)
    b = BSpline.builder().clamped()
    b = b.elements_with_weights(el, (0.5 + u01(stops, 0xA3, dtype)).astype(dtype)) if weighted else b.elements(el)
    if knots_kind.startswith("sorted"):
        k = sorted_knots(stops - degree + 1, dtype, 0xA2, repeat_every=7 if knots_kind == "sorted_repeats" else 0)
        b = b.knots(k)
    elif knots_kind == "equidistant":
        b = b.equidistant(scalar).degree(degree).domain(-1.0, 2.0)
    else:
        if (stops - degree) & (stops - degree - 1):
            pytest.skip("inner knot count - 1 is not a power of two")
        b = b.equidistant(scalar).degree(degree).normalized()
    b = b.dynamic()
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    for n in (2, 513, 100003):
        assert_bits_equal(c.take_device(n).cpu().numpy(), r.take(n), f"take {n}")
    total = 1 << 22
    for first, cnt in ((0, 70001), (total // 3, 50000), (total - 1025, 1025)):
        assert_bits_equal(c.take_device(total, first, cnt).cpu().numpy(), r.take(total, first, cnt), f"window {first}+{cnt}")
    # batches of the same curves keep their own paths (shared-memory rows when they fit, generic otherwise)
    d = r.domain()
    t = (u01(20011, 0xA5, dtype) * dtype(1.2 * float(d[1] - d[0])) + dtype(float(d[0]) - 0.1 * float(d[1] - d[0]))).astype(dtype)
    assert_bits_equal(c.sample_device(torch.from_numpy(t).cuda()).cpu().numpy(), r.eval(t), "batch")
    assert_bits_equal(c.clamp().take_device(5000).cpu().numpy(), r.clamp().take(5000), "clamp().take")
    # a slice that runs far out of the domain: decreasing / huge parameters; with equidistant knots the far end
    # makes the reference panic (floor(scaled) >= 2^64) -> NaN samples and the invalid counter, mid-take
    for lo_, hi_ in ((float(d[1]), float(d[0])), (float(d[0]), 1e30)):
        cs, rs = c.slice(lo_, hi_), r.slice(lo_, hi_)
        before = c.invalid_count()
        got = cs.take_device(40001).cpu().numpy()
        want, inv = rs.take(40001, return_invalid=True)
        assert_bits_equal(got, want, f"slice({lo_}, {hi_}).take")
        assert c.invalid_count() - before == inv
        assert (inv > 0) == (hi_ == 1e30 and knots_kind.startswith("equidistant"))


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("dim,weighted", [(1, 0), (2, 0), (3, 0), (4, 0), (1, 1), (3, 1), (4, 1)])
@pytest.mark.parametrize("n_knots,repeats", [(2, 0), (3, 0), (37, 0), (37, 5), (5000, 0), (5000, 3)])
def test_lean_linear_sorted_take(scalar, dim, weighted, n_knots, repeats):
    """take() of Linear over sorted knots (chunked walk, the search is skipped inside the held knot interval):
    few and many samples per interval, repeated knots incl. coincident edge knots, easings, windows."""
    dtype = DT[scalar]
    el = rand_elements(n_knots, dim, dtype, 0xB1 + n_knots)
    knots = sorted_knots(n_knots, dtype, 0xB2, repeat_every=repeats)
    if repeats and n_knots > 4:
        knots[1] = knots[0]
        knots[-2] = knots[-1]
    b = Linear.builder()
    b = b.elements_with_weights(el, (0.5 + u01(n_knots, 0xB3, dtype)).astype(dtype)) if weighted else b.elements(el)
    b = b.knots(knots)
    if (dim + weighted) % 2 == 0:
        from enterpolation_b200 import easing as E
        b = b.easing([E.smoothstep, E.Plateau(0.4), E.smoothstart(3), E.flip][dim - 1])
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    for n in (2, 3, 255, 257, 100003):
        assert_bits_equal(c.take_device(n).cpu().numpy(), r.take(n), f"take {n}")
    total = 1 << 22
    for first, cnt in ((0, 70001), (total // 3, 50000), (total - 1025, 1025)):
        assert_bits_equal(c.take_device(total, first, cnt).cpu().numpy(), r.take(total, first, cnt), f"window {first}+{cnt}")
    n_total = (1 << 32) + 12345
    for first in ((1 << 32) - (1 << 24) - 3000, (1 << 32) - 1500):
        assert_bits_equal(c.take_device(n_total, first, 3000).cpu().numpy(), r.take(n_total, first, 3000), f"first={first}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("shape", ["uniform", "one_cluster", "heavy_clusters", "repeats", "huge_range", "tiny_range"])
def test_bucket_index_search(scalar, shape):
    """Sorted searches over >= 4096 knots go through the bucket index (kernels.cuh count_le): spread-out knots, a
    cluster that overflows its bucket (tree fallback at run time), clustering that disables the index at build time,
    repeated knots, extreme knot ranges. Parameters: every knot, its float neighbours, out-of-domain, NaN, inf."""
    dtype = DT[scalar]
    rng = np.random.default_rng(11)
    n = 8192
    k = np.cumsum(0.5 + u01(n, 0xD2, np.float64))
    if shape == "one_cluster":
        k[3000:3100] = k[3000] + np.arange(100) * 1e-3
    elif shape == "heavy_clusters":
        k = np.sort(np.concatenate([k[: n // 2], k[n // 2] + u01(n // 2, 0xD3, np.float64) * 1e-2]))
    elif shape == "repeats":
        k[1::3] = k[0::3][: len(k[1::3])]
        k = np.sort(k)
    elif shape == "huge_range":
        k = k * (1e30 if scalar == "f32" else 1e250)
    elif shape == "tiny_range":
        k = 1.0 + k * (2e-7 if scalar == "f32" else 3e-16)
    k = np.sort(k.astype(dtype))
    t = adversarial(k, dtype, rng, n_random=20000)
    t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=dtype)])
    check_curve(Linear.builder().elements(rand_elements(n, 3, dtype, 0xD4)).knots(k), t, what=f"linear {shape}")
    check_curve(BSpline.builder().elements(rand_elements(n - 2, 2, dtype, 0xD5)).knots(k).constant(4), t, what=f"bspline {shape}")


def test_launches_are_cuda_graph_capturable():
    """Small takes are launch-latency bound (C1 at 2^20 samples: ~16 us); after the one-time upload every eval call
    only enqueues kernels on the caller's stream, so a sequence of them can be captured into a CUDA graph and replayed."""
    curves = [_c5_like("f32", 7, 4),
              Linear.builder().elements(rand_elements(12, 3, np.float32, 0x95)).knots(sorted_knots(12, np.float32, 0x96)),
              Bezier.builder().elements(rand_elements(4, 2, np.float64, 0x97)).normalized("f64").constant(),
              BSpline.builder().elements(rand_elements(9, 1, np.float64, 0x98)).knots(sorted_knots(11, np.float64, 0x99)).constant(4)]
    built = [b.build() for b in curves]
    refs = [oracle.OracleCurve(b.to_desc()) for b in curves]
    n = 4099
    outs = [c.take_device(n) for c in built]  # warm-up: uploads the tables, outside the capture
    t = [torch.from_numpy(u01(n, 0x9A + i, c.dtype)).cuda() for i, c in enumerate(built)]
    bouts = [c.sample_device(t[i]) for i, c in enumerate(built)]
    for o in outs + bouts:
        o.zero_()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i, c in enumerate(built):
            c.take_device(n, out=outs[i])
            c.sample_device(t[i], out=bouts[i])
    for o in outs + bouts:
        o.zero_()  # the capture itself must not have run anything we rely on
    g.replay()
    torch.cuda.synchronize()
    for i, r in enumerate(refs):
        assert_bits_equal(outs[i].cpu().numpy(), r.take(n), f"graph take {i}")
        assert_bits_equal(bouts[i].cpu().numpy(), r.eval(t[i].cpu().numpy()), f"graph batch {i}")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("knots_kind", ["equidistant", "equidistant_pow2", "sorted"])
def test_hold_intervals_at_every_float_around_the_knots(scalar, knots_kind):
    """The chunked take variants skip the span computation while hold_lo <= x < hold_hi. Zooming a slice onto each knot
    (parameter spacing far below one ulp) walks through every representable parameter on both sides of every span
    boundary, where a hold interval that is off by a single float would pick the wrong row."""
    dtype = DT[scalar]
    stops, degree = 11, 3
    el = rand_elements(stops, 2, dtype, 0xC7)
    b = BSpline.builder().clamped().elements(el)
    if knots_kind == "sorted":
        b = b.knots(sorted_knots(stops - degree + 1, dtype, 0xC8))
    elif knots_kind == "equidistant":
        b = b.equidistant(scalar).degree(degree).domain(-1.0, 2.0)   # step 3/8... not a power of two after rounding
    else:
        b = b.equidistant(scalar).degree(degree).normalized()        # 8 intervals: step 2^-3
    b = b.dynamic()
    c, r = b.build(), oracle.OracleCurve(b.to_desc())
    d = r.domain()
    n_inner = stops - degree + 1
    if knots_kind == "sorted":
        knots = sorted_knots(n_inner, dtype, 0xC8)
    else:
        knots = np.linspace(float(d[0]), float(d[1]), n_inner).astype(dtype)
    eps = 64 * float(np.finfo(dtype).eps)
    for k in knots:
        lo, hi = float(k) - eps * max(1.0, abs(float(k))), float(k) + eps * max(1.0, abs(float(k)))
        cs, rs = c.slice(lo, hi), r.slice(lo, hi)
        assert_bits_equal(cs.take_device(8192).cpu().numpy(), rs.take(8192), f"zoom on knot {k}")
    if scalar == "f32":  # plain take(): a stepper finer than one ulp, windows centred on every knot
        n_total = 1 << 31
        step = (float(d[1]) - float(d[0])) / (n_total - 1)
        for k in knots[1:-1]:
            first = int(round((float(k) - float(d[0])) / step)) - 2048
            assert_bits_equal(c.take_device(n_total, first, 4096).cpu().numpy(), r.take(n_total, first, 4096), f"window on {k}")
        lin = Linear.builder().elements(rand_elements(n_inner, 3, dtype, 0xC9)).knots(np.asarray(knots, dtype=dtype))
        lc, lr = lin.build(), oracle.OracleCurve(lin.to_desc())
        for k in knots[1:-1]:
            first = int(round((float(k) - float(d[0])) / step)) - 2048
            assert_bits_equal(lc.take_device(n_total, first, 4096).cpu().numpy(), lr.take(n_total, first, 4096), f"linear window on {k}")


# ---- f32 stepper runs (rows.cuh RUNS / VARIANT 4): from index 2^24 on, from_usize(i) repeats in runs ------------------
def _runs_builders():
    el4 = rand_elements(5, 4, np.float32, 0x51)
    yield "C5 5 stops", BSpline.builder().clamped().elements(el4).equidistant("f32").degree(3).normalized().dynamic()
    el256 = rand_elements(256, 4, np.float32, 0x52)
    yield "C5 256 stops", BSpline.builder().clamped().elements(el256).equidistant("f32").degree(3).normalized().dynamic()
    for dim in (1, 2, 3):
        el = rand_elements(23, dim, np.float32, 0x53 + dim)
        yield f"sorted quadratic dim {dim}", BSpline.builder().elements(el).knots(sorted_knots(24, np.float32, 0x54)).constant(3)
    pts = rand_elements(40, 3, np.float32, 0x58)
    w = (u01(40, 0x59, np.float32) + np.float32(0.5)).astype(np.float32)
    yield "nurbs", BSpline.builder().elements_with_weights(pts, w).knots(sorted_knots(42, np.float32, 0x5A)).constant(4)


@pytest.mark.parametrize("name,builder", list(_runs_builders()), ids=lambda v: v if isinstance(v, str) else "")
def test_take_f32_stepper_runs(name, builder):
    curve = builder.build()
    ref = oracle.OracleCurve(builder.to_desc())
    lib = _abi.load_library()
    cases = []
    n32 = 1 << 32
    for k in range(24, 32):  # run-length changes: 2^k - 300 .. 2^k + 300 (and odd, unaligned windows)
        cases.append((n32, (1 << k) - 300, 601))
        cases.append((n32, (1 << k) - 4097 - k, 9001))
    cases += [(n32, (1 << 24) - 70000, 200001),      # head below 2^24 + tail above in one call
              (n32, n32 - 100000, 100000),           # the end of the take (fl(i) rounds up to 2^32)
              (n32 - 12345, 3 * (1 << 29) + 17, 70001),
              ((1 << 25) + 5, 0, (1 << 25) + 5),     # a whole take that crosses 2^24
              ((1 << 40) + 3, (1 << 39) + 12345, 150000),   # runs longer than a tile
              ((1 << 40) + 3, (1 << 37) - 4000, 150000)]
    for n_total, first, count in cases:
        before = lib.enterp_launch_count()
        got = curve.take_device(n_total, first, count).cpu().numpy()
        assert lib.enterp_launch_count() - before in (1, 2)
        assert_bits_equal(got, ref.take(n_total, first, count), f"{name} take({n_total}) [{first}, +{count})")


# ---- host-buffer calls over several devices, scratch pool, counters, capture guard ------------------------------------
def test_take_and_sample_host_multi_device():
    """enterp_take_host_multi / enterp_sample_host_multi: ONE call from one host thread shards the range over every GPU
    of the box (contiguous shares, global stepper index): equal to the single-device call and to the oracle."""
    n_dev = torch.cuda.device_count()
    devices = list(range(n_dev))
    b = _c5_like("f32", 9, 4)
    curve = b.build()
    ref = oracle.OracleCurve(b.to_desc())
    n_total, first, count = (1 << 26) + 11, (1 << 24) - 1000, (1 << 23) + 12345  # several chunks per device, crosses 2^24
    from enterpolation_b200.curve import Take
    got = Take(curve, n_total, first, count).collect(devices=devices)
    assert_bits_equal(got, ref.take(n_total, first, count, threads=8), f"take_host_multi over {n_dev} device(s)")
    one = Take(curve, n_total, first, count).collect()
    assert_bits_equal(got, one, "multi == single")
    lin_b = Linear.builder().elements(rand_elements(300, 3, np.float64, 0xD1)).knots(sorted_knots(300, np.float64, 0xD2))
    lin = lin_b.build()
    t = (u01((1 << 22) + 77, 0xD3, np.float64) * 200.0 - 10.0)
    assert_bits_equal(lin.sample(t, devices=devices), oracle.OracleCurve(lin_b.to_desc()).eval(t, threads=8), "sample_host_multi")
    lib = _abi.load_library()
    import ctypes as C
    arr = (C.c_int * 2)(0, 0)
    out = np.empty((16, 4), dtype=np.float32)
    assert lib.enterp_take_host_multi(curve._h, arr, 2, 16, 0, 16, out.ctypes.data) == _abi.INVALID_ARGUMENT  # repeated ordinal
    assert lib.enterp_take_host_multi(curve._h, arr, 1, 16, 2 ** 64 - 1, 2, out.ctypes.data) == _abi.INVALID_ARGUMENT  # wraps
    assert lib.enterp_host_scratch_release(-1) >= 1  # the idle pipes of the calls above
    assert_bits_equal(Take(curve, 99).collect(devices=devices), ref.take(99), "after scratch release")


def test_range_checks_do_not_wrap():
    curve = _c5_like("f32", 5, 4).build()
    out = torch.empty((8, 4), dtype=torch.float32, device="cuda")
    lib = _abi.load_library()
    big = 2 ** 64 - 1
    assert lib.enterp_eval_take(curve._h, 0, 1000, big, 2, out.data_ptr(), None) == _abi.INVALID_ARGUMENT
    assert lib.enterp_eval_take(curve._h, 0, 1000, 999, 2, out.data_ptr(), None) == _abi.INVALID_ARGUMENT
    assert lib.enterp_span_take(curve._h, 0, 1000, big, 2, out.data_ptr(), None, None) == _abi.INVALID_ARGUMENT
    assert lib.enterp_stepper(_abi.F32, 0, 0.0, 1.0, 1000, big, 2, out.data_ptr(), None) == _abi.INVALID_ARGUMENT
    assert lib.enterp_eval_take(curve._h, 0, 1000, 999, 1, out.data_ptr(), None) == _abi.OK


def test_invalid_counter_reset_and_stream_scope():
    b = BSpline.builder().clamped().elements(rand_elements(7, 2, np.float32, 0xD5)).equidistant("f32").degree(3).domain(0.0, 3.0).dynamic()
    curve = b.build()
    s = torch.cuda.Stream()
    t = torch.tensor([0.5, float("nan"), 1.5, float("inf"), 2.0, float("nan")], dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    curve.sample_device(t, stream=s)
    assert curve.invalid_count(stream=s) == 3
    curve.invalid_reset(stream=s)
    assert curve.invalid_count(stream=s) == 0
    curve.sample_device(t[:3], stream=s)
    assert curve.invalid_count(stream=s) == 1
    assert curve.invalid_count() == 1


def test_first_use_inside_a_capture_is_reported():
    """The one-time upload cannot run while the caller's stream is being captured: ENTERP_NOT_UPLOADED, the capture
    survives, and after enterp_curve_upload the same call captures fine."""
    from enterpolation_b200 import DeviceError
    b = _c5_like("f32", 6, 4)
    curve = b.build()
    n = 1000
    out = torch.zeros((n, 4), dtype=torch.float32, device="cuda")
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        g.capture_begin()
        with pytest.raises(DeviceError):
            curve.take_device(n, out=out, stream=s)
        g.capture_end()
    curve.upload(0)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        g2.capture_begin()
        curve.take_device(n, out=out, stream=s)
        g2.capture_end()
    g2.replay()
    torch.cuda.synchronize()
    assert_bits_equal(out.cpu().numpy(), oracle.OracleCurve(b.to_desc()).take(n), "captured after upload")


@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("n_el,n_kn,dim,weighted,workspace", [(6, 8, 1, 0, None), (9, 9, 3, 0, 2), (12, 15, 4, 1, 8), (40, 47, 2, 1, None),
                                                               (5000, 5002, 3, 0, 4)])
def test_bspline_over_const_equidistant_knots(scalar, n_el, n_kn, dim, weighted, workspace):
    """BSpline::new(elements, ConstEquidistant::<R, N>, space) (bspline/mod.rs:212-235; span search by x * fl(N-1),
    base/list.rs:652-664): batch, spans, take, host paths against the oracle, incl. every knot and its float neighbours,
    NaN/inf (the reference panics: NaN output + counter)."""
    dtype = DT[scalar]
    rng = np.random.default_rng(n_el + dim)
    el = rand_elements(n_el, dim, dtype, 0xCE)
    w = (u01(n_el, 0xCF, dtype) + dtype(0.5)).astype(dtype) if weighted else None
    b = BSpline.new_const_equidistant(el, n_kn, scalar, workspace=workspace, weights=w)
    kn = (np.arange(n_kn, dtype=dtype) / dtype(n_kn - 1)).astype(dtype)
    t = adversarial(kn, dtype, rng, n_random=3000)
    t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf, -0.0, 1.0, 2.0], dtype=dtype)])
    check_curve(b, t, what=f"bspline over ConstEquidistant<{scalar},{n_kn}>")


@pytest.mark.parametrize("dim", [1, 2, 3, 4])
def test_take_host_run_length_transfer(dim):
    """Host-buffer takes of f32 curves beyond index 2^26 ship only the distinct values over PCIe and replicate them on
    the host (capi.cu, run-length transfer): same bytes as the oracle, pageable and pinned outputs, windows that cross
    binades and chunk boundaries, several devices; and the PCIe byte counter shows the saving."""
    import ctypes as C
    from enterpolation_b200.curve import Take
    lib = _abi.load_library()
    el = rand_elements(37, dim, np.float32, 0xA0 + dim)
    b = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic()
    curve = b.build()
    ref = oracle.OracleCurve(b.to_desc())
    n32 = 1 << 32
    devices = list(range(torch.cuda.device_count()))
    cases = [(n32, (1 << 31) - 3_000_000, 9_000_001),         # crosses 2^31, three chunks, unaligned
             (n32, (1 << 26) - 5000, (1 << 23) + 77),         # starts below the threshold
             (n32, n32 - (1 << 23) - 3, (1 << 23) + 3),       # the end of the take
             ((1 << 29) + 12345, (1 << 28) - 17, (1 << 23) + 1)]
    for n_total, first, count in cases:
        want = ref.take(n_total, first, count, threads=8)
        h2d0, d2h0 = C.c_ulonglong(0), C.c_ulonglong(0)
        lib.enterp_host_transfer_bytes(C.byref(h2d0), C.byref(d2h0))
        got = Take(curve, n_total, first, count).collect()      # pageable numpy output
        h2d1, d2h1 = C.c_ulonglong(0), C.c_ulonglong(0)
        lib.enterp_host_transfer_bytes(C.byref(h2d1), C.byref(d2h1))
        assert_bits_equal(got, want, f"rle take_host dim {dim} [{first}, +{count}) of {n_total}")
        moved = d2h1.value - d2h0.value
        assert moved <= count * dim * 4
        # (how much is saved depends on how the chunks were shared between the copy engine and the run-length path)
        pinned = torch.empty((count, dim) if dim > 1 else (count,), dtype=torch.float32, pin_memory=True)
        arr = (C.c_int * len(devices))(*devices)
        st = lib.enterp_take_host_multi(curve._h, arr, len(devices), n_total, first, count, pinned.data_ptr())
        assert st == 0
        assert_bits_equal(pinned.numpy(), want, f"rle take_host_multi dim {dim} pinned")
    # the other curve kinds go through the same transfer (their batch kernels evaluate the distinct values)
    others = [Linear.builder().elements(rand_elements(40, dim, np.float32, 0xA5)).equidistant("f32").normalized(),
              Linear.builder().elements(rand_elements(40, dim, np.float32, 0xA6)).knots(sorted_knots(40, np.float32, 0xA7)),
              Bezier.builder().elements(rand_elements(6, dim, np.float32, 0xA8)).normalized("f32").constant()]
    if dim > 1:
        w = (u01(30, 0xA9, np.float32) + np.float32(0.5)).astype(np.float32)
        others.append(BSpline.builder().elements_with_weights(rand_elements(30, dim, np.float32, 0xAA), w)
                      .knots(sorted_knots(32, np.float32, 0xAB)).constant(4))
    for ob in others:
        oc, oref = ob.build(), oracle.OracleCurve(ob.to_desc())
        n_total, first, count = n32, 3 * (1 << 30) - 70_000, (1 << 22) + 140_001
        assert_bits_equal(Take(oc, n_total, first, count).collect(), oref.take(n_total, first, count, threads=8),
                          f"rle take_host, other kinds, dim {dim}")
    # adaptors switch the run-length transfer off (a mapped parameter could leave the valid range): still exact
    cl = curve.clamp()
    n_total, first, count = n32, (1 << 30) + 5, 300_001
    d0 = C.c_ulonglong(0)
    lib.enterp_host_transfer_bytes(None, C.byref(d0))
    got = Take(cl, n_total, first, count).collect()
    d1 = C.c_ulonglong(0)
    lib.enterp_host_transfer_bytes(None, C.byref(d1))
    assert_bits_equal(got, ref.clamp().take(n_total, first, count, threads=8), "clamped curve, plain transfer")
    assert d1.value - d0.value == count * dim * 4


@pytest.mark.parametrize("degree", [1, 2, 3])
@pytest.mark.parametrize("mode", ["open", "clamped", "legacy"])
def test_bspline_knot_row_slots(degree, mode):
    """f32 B-splines over >= 4096 sorted knots, degrees 1..3: the generic kernel finds span and knot row through the
    grid-indexed knot-row slots (bspline_slots_kernel). Every knot, its float neighbours, out of domain, NaN/inf; knots
    with repeats and a cluster; all three builder modes (the span index a slot carries includes the adaptor's shift)."""
    rng = np.random.default_rng(degree * 7 + len(mode))
    n_k = 5000
    k = np.cumsum(0.5 + u01(n_k, 0xB1 + degree, np.float64))
    k[1000:1040] = k[1000] + np.arange(40) * 1e-3      # a cluster that overflows its cell
    k[2000:2003] = k[2000]                              # repeated interior knots
    k = np.sort(k).astype(np.float32)
    if mode == "open":
        e = n_k - degree + 1
        b = BSpline.builder().elements(rand_elements(e, 3, np.float32, 0xB5)).knots(k).constant(degree + 1)
    elif mode == "clamped":
        e = n_k + degree - 1
        b = BSpline.builder().clamped().elements(rand_elements(e, 3, np.float32, 0xB6)).knots(k).constant(degree + 1)
    else:
        e = n_k - degree - 1
        b = BSpline.builder().legacy().elements(rand_elements(e, 3, np.float32, 0xB7)).knots(k).constant(degree + 1)
    t = adversarial(k, np.float32, rng, n_random=20000)
    t = np.concatenate([t, np.array([np.nan, np.inf, -np.inf], dtype=np.float32)])
    check_curve(b, t, what=f"knot-row slots degree {degree} {mode}")


def test_linear_and_bezier_takes_across_the_lean_index_limit():
    """The lean Linear / Bezier take kernels index with 32 bits up to 2^32 - 2^24; a take that crosses that index is cut
    there (capi.cu dispatch_eval) instead of leaving the whole call to the generic kernels."""
    lim = (1 << 32) - (1 << 24)
    builders = [Linear.builder().elements(rand_elements(64, 4, np.float32, 0xC1)).equidistant("f32").normalized(),
                Linear.builder().elements(rand_elements(64, 3, np.float64, 0xC2)).knots(sorted_knots(64, np.float64, 0xC3)),
                Bezier.builder().elements(rand_elements(4, 4, np.float32, 0xC4)).normalized("f32").constant()]
    lib = _abi.load_library()
    for b in builders:
        c, ref = b.build(), oracle.OracleCurve(b.to_desc())
        for n_total, first, count in (((1 << 32), lim - 5000, 10001), ((1 << 33) + 5, lim - 1, 4097), ((1 << 32), lim, 3000)):
            before = lib.enterp_launch_count()
            got = c.take_device(n_total, first, count).cpu().numpy()
            assert lib.enterp_launch_count() - before >= (2 if first < lim else 1)  # beyond the limit: generic stepper runs
            assert_bits_equal(got, ref.take(n_total, first, count), f"take({n_total}) [{first}, +{count})")


def test_take_f32_stepper_runs_with_adaptors_and_panics():
    """Stepper runs (rows VARIANT 6): clamped / sliced f32 B-splines beyond index 2^24, incl. a slice that drives the
    parameter out of `to_usize` range (the reference panics there): NaN for every sample of such a run, and the
    invalid counter advances by the samples, not by the distinct values."""
    el = rand_elements(9, 4, np.float32, 0xE5)
    b = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).domain(0.0, 3.0).dynamic()
    base, rbase = b.build(), oracle.OracleCurve(b.to_desc())
    cases = [(base.clamp(), rbase.clamp()), (base.slice(0.5, 2.5), rbase.slice(0.5, 2.5)),
             (base.clamp().slice(-1.0, 4.0), rbase.clamp().slice(-1.0, 4.0)),
             (base.slice(0.0, 3.0e25), rbase.slice(0.0, 3.0e25))]  # scaled >= 2^64 from some index on: panics upstream
    n32 = 1 << 32
    windows = [(n32, (1 << 24) - 5000, 30001), (n32, (1 << 31) - 777, 5000), (n32, n32 - 40000, 40000), (n32, 3 << 29, 1 << 20)]
    lib = _abi.load_library()
    for c, r in cases:
        for n_total, first, count in windows:
            before = c.invalid_count()
            launches = lib.enterp_launch_count()
            got = c.take_device(n_total, first, count).cpu().numpy()
            assert lib.enterp_launch_count() - launches in (1, 2)
            want, inv = r.take(n_total, first, count, return_invalid=True, threads=4)
            assert_bits_equal(got, want, f"adaptor take({n_total}) [{first}, +{count})")
            assert c.invalid_count() - before == inv


def _generic_runs_builders():
    yield "bspline degree 5 sorted", (BSpline.builder().elements(rand_elements(30, 3, np.float32, 0xF1))
                                      .knots(sorted_knots(34, np.float32, 0xF2)).constant(6))
    yield "bspline runtime degree 9", (BSpline.builder().elements(rand_elements(40, 2, np.float32, 0xF3))
                                       .knots(sorted_knots(48, np.float32, 0xF4)).dynamic())
    yield "bspline over ConstEquidistant", BSpline.new_const_equidistant(rand_elements(20, 4, np.float32, 0xF5), 22, "f32")
    yield "bezier 12 control points", Bezier.builder().elements(rand_elements(12, 3, np.float32, 0xF6)).normalized("f32").constant()
    w = (u01(25, 0xF8, np.float32) + np.float32(0.5)).astype(np.float32)
    yield "weighted linear sorted", Linear.builder().elements_with_weights(rand_elements(25, 1, np.float32, 0xF7), w) \
        .knots(sorted_knots(25, np.float32, 0xF9))
    yield "linear plateau easing", Linear.builder().elements(rand_elements(33, 4, np.float32, 0xFA)).equidistant("f32") \
        .normalized().easing(__import__("enterpolation_b200").easing.Plateau(0.3))


@pytest.mark.parametrize("name,builder", list(_generic_runs_builders()), ids=lambda v: v if isinstance(v, str) else "")
def test_take_f32_generic_stepper_runs(name, builder):
    """f32 takes beyond index 2^26 of curves the span-row kernels do not serve: the distinct stepper values are
    evaluated once by the curve's batch kernel and expanded (capi.cu generic_runs_take). Windows across the 2^26
    threshold, binade boundaries, the lean kernels' index limit, the end of take(2^32), beyond 2^32."""
    curve = builder.build()
    ref = oracle.OracleCurve(builder.to_desc())
    n32 = 1 << 32
    cases = [(n32, (1 << 24) - 40000, 150001), (n32, (1 << 26) - 999, 100000), (n32, (1 << 31) - 70001, 140003),
             (n32, n32 - (1 << 24) - 5000, 70000), (n32, n32 - 100000, 100000), ((1 << 33) + 5, (1 << 32) - 33333, 100001),
             ((1 << 27) + 3, 0, (1 << 27) + 3)]
    for n_total, first, count in cases:
        got = curve.take_device(n_total, first, count).cpu().numpy()
        assert_bits_equal(got, ref.take(n_total, first, count, threads=8), f"{name} take({n_total}) [{first}, +{count})")


# The builders take `domain(a, b)` with a > b as it comes (no check upstream): the equidistant knots then DESCEND.
# Open B-splines: list.rs:476-488 returns `max.min(min_index + 1)`, which is below the lower cut for parameters between the
# first knot and knot[degree]; `index - degree` underflows and the element access panics (bspline/mod.rs:131) -> NaN +
# invalid counter here (found by the seeded fuzz sweep: the kernel used to gather in front of its tables instead).
# Parameters beyond the first knot fail `to_usize` (negative quotient); Linear: the same through upper_border.
@pytest.mark.parametrize("scalar", ["f32", "f64"])
@pytest.mark.parametrize("kind,degree,dim,weighted", [("open", 1, 1, 0), ("open", 3, 4, 0), ("open", 7, 1, 0), ("open", 9, 2, 1),
                                                      ("clamped", 3, 3, 1), ("clamped", 5, 2, 0), ("linear", 1, 3, 0)])
def test_descending_equidistant_knots(scalar, kind, degree, dim, weighted):
    dtype = DT[scalar]
    n = degree + 6
    el = rand_elements(n, dim, dtype, 0xD5C)
    w = (dtype(0.5) + u01(n, 0xD5D, dtype)).astype(dtype)
    if kind == "linear":
        b = Linear.builder()
        b = b.elements_with_weights(el, w) if weighted else b.elements(el)
        b = b.equidistant(scalar).domain(2.75, 1.5)
    else:
        b = BSpline.builder().open() if kind == "open" else BSpline.builder().clamped()
        b = b.elements_with_weights(el, w) if weighted else b.elements(el)
        b = b.equidistant(scalar).degree(degree).domain(2.75, 1.5).dynamic()
    rng = np.random.default_rng(degree * 10 + dim)
    t = np.concatenate([rng.uniform(1.0, 3.25, 4000), [2.75, 1.5, np.nextafter(2.75, 3), np.nextafter(2.75, 0)]]).astype(dtype)
    curve, ref = check_curve(b, t, what=f"descending {kind} degree {degree} {scalar}")
    want, inv = ref.eval(t, return_invalid=True)
    assert inv > 0 and np.isfinite(want).any()


# The stepper-run paths (generic runs on the device, run-length transfer to host buffers) evaluate distinct VALUES; a
# curve whose in-domain parameters make the reference panic (zero-width or descending equidistant knots: the quotient is
# NaN or negative) must keep the per-sample kernels, or the invalid counter counts values instead of samples (found by the
# degenerate fuzz sweep: take(2^32 + 78)[694938518:+1] counted 2).
@pytest.mark.parametrize("domain", [(0.0, 0.0), (2.75, 1.5), (1e4, 1e4), ("distance", 0.2992507, 2.9e-39), ("distance", 1e4, 2.0 ** -14)])
@pytest.mark.parametrize("kind", ["bspline", "linear"])
def test_run_paths_count_invalid_samples_not_values(domain, kind):
    from enterpolation_b200.curve import Take
    el = rand_elements(9, 4, np.float32, 0xBAD)
    b = BSpline.builder().open().elements(el).equidistant("f32").degree(5) if kind == "bspline" else Linear.builder().elements(el).equidistant("f32")
    # ("distance", offset, step): knots that collapse onto each other (a step far below the ulp of the offset)
    b = b.distance(domain[1], domain[2]) if domain[0] == "distance" else b.domain(*domain)
    if kind == "bspline":
        b = b.dynamic()
    curve, ref = b.build(), oracle.OracleCurve(b.to_desc())
    total = (1 << 32) + 78
    for first, n in ((694938518, 1), ((1 << 26) - 3000, 70001), ((1 << 31) - 5, 300000)):
        want, inv = ref.take(total, first, n, return_invalid=True)
        before = curve.invalid_count()
        assert_bits_equal(curve.take_device(total, first, n).cpu().numpy(), want, f"{kind} {domain} device take")
        assert curve.invalid_count() - before == inv, f"{kind} {domain} take({total})[{first}:+{n}] invalid count"
        before = curve.invalid_count()
        assert_bits_equal(Take(curve, total, first, n).collect(), want, f"{kind} {domain} host take")
        assert curve.invalid_count() - before == inv, f"{kind} {domain} host take invalid count"
