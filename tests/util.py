"""Comparison helpers shared by the tests (assert_float_eq semantics: 4 ULP steps by default)."""
import numpy as np


def _ordered(a):
    a = np.ascontiguousarray(a)
    if a.dtype == np.float32:
        i = a.view(np.int32).astype(np.int64)
        return np.where(i < 0, -(i & 0x7FFFFFFF), i)
    i = a.view(np.int64)
    return np.where(i < 0, -(i & 0x7FFFFFFFFFFFFFFF), i)


def ulp_distance(a, b):
    """Element-wise distance in representable steps (both NaN -> 0, one NaN -> huge)."""
    a = np.asarray(a)
    b = np.asarray(b, dtype=a.dtype)
    d = np.abs(_ordered(a).astype(np.float64) - _ordered(b).astype(np.float64))
    an, bn = np.isnan(a), np.isnan(b)
    d = np.where(an & bn, 0.0, d)
    d = np.where(an ^ bn, np.inf, d)
    return d


def assert_near(a, b, ulps=4, what=""):
    """assert_f64_near! / assert_f32_near!: at most `ulps` representable steps apart."""
    a = np.asarray(a)
    d = ulp_distance(a, np.asarray(b, dtype=a.dtype))
    assert np.all(d <= ulps), f"{what}: max ulp distance {d.max()} > {ulps}\n got {a}\n want {b}"


def assert_bits_equal(a, b, what=""):
    """Bit-for-bit equality, except that every NaN equals every NaN."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.dtype == b.dtype and a.shape == b.shape, f"{what}: {a.dtype}{a.shape} vs {b.dtype}{b.shape}"
    if a.dtype.kind == "f":
        u = np.uint32 if a.dtype == np.float32 else np.uint64
        same = (a.view(u) == b.view(u)) | (np.isnan(a) & np.isnan(b))
    else:
        same = a == b
    if not np.all(same):
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError(f"{what}: {bad.shape[0]} mismatches of {a.size}; first at {i}: got {a[i]!r} want {b[i]!r}")


def splitmix64(idx, seed):
    """SURVEY.md §8d deterministic generator: splitmix64(seed + index)."""
    z = (np.asarray(idx, dtype=np.uint64) + np.uint64(seed)) * np.uint64(1)
    with np.errstate(over="ignore"):
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def u01(n, seed, dtype):
    z = splitmix64(np.arange(n, dtype=np.uint64), seed)
    if dtype == np.float32:
        return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (z >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
