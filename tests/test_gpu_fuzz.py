"""Seeded random differential test: random curve configurations x random evaluation calls, CUDA path vs CPU oracle,
bit for bit. Complements the structured cases of test_gpu_parity.py by mixing features the kernels dispatch on
(table sizes around the shared-memory / bucket-index / lean-kernel thresholds, adaptors, easings, weights, windows)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

from fuzzgen import DT, _degenerate_curve, _degenerate_inputs, _random_const_equidistant, _random_curve, _weird
from util import assert_bits_equal

pytestmark = pytest.mark.gpu


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


@pytest.mark.parametrize("seed", range(_FIRST // 2, _FIRST // 2 + max(1, _COUNT // 2)))
def test_random_degenerate_curves(seed):
    from enterpolation_b200.curve import Take
    rng = np.random.default_rng(0xDE6E + seed)
    done = 0
    while done < 10:
        scalar, b = _degenerate_curve(rng)
        try:
            c = b.build()
        except Exception:
            continue
        r = oracle.OracleCurve(b.to_desc())
        base = c  # the invalid counter lives with the core curve
        pick = rng.random()
        if pick < 0.2:
            c, r = c.clamp(), r.clamp()
        elif pick < 0.4:  # Slice::new with bounds as they come: empty, reversed, non-finite (adaptors.rs:61-81)
            d0 = r.domain()
            a = float(rng.choice([float(d0[0]), 0.0, _weird(rng, DT[scalar]), rng.normal()]))
            bb = float(rng.choice([a, float(d0[1]), _weird(rng, DT[scalar]), rng.normal()]))
            c, r = c.slice(a, bb), r.slice(a, bb)
        d = r.domain()
        tag = f"seed {seed} case {done}"
        t = _degenerate_inputs(rng, DT[scalar], float(d[0]), float(d[1]), int(rng.choice([7, 1000, 20011])))
        want, inv = r.eval(t, return_invalid=True)
        before = base.invalid_count()
        got = c.sample_device(torch.from_numpy(t).cuda()).cpu().numpy()
        assert_bits_equal(got, want, tag + " batch")
        assert base.invalid_count() - before == inv, tag + " batch invalid count"
        assert_bits_equal(np.asarray(c.sample(t)), want, tag + " host batch")
        idx, fac = c.span_device(torch.from_numpy(t).cuda())
        widx, wfac = r.span(t)
        assert np.array_equal(idx.cpu().numpy().view(np.uint32), widx), tag + " span indices"
        assert_bits_equal(fac.cpu().numpy(), wfac, tag + " span factor")
        n = int(rng.choice([1, 2, 33, 4099, 70001]))
        total = n + int(rng.choice([0, 12345, (1 << 27) + 5, (1 << 32) + 77]))
        first = int(rng.integers(0, total - n + 1))
        want, inv = r.take(total, first, n, return_invalid=True)
        before = base.invalid_count()
        got = c.take_device(total, first, n).cpu().numpy()
        assert_bits_equal(got, want, tag + f" take({total})[{first}:+{n}]")
        assert base.invalid_count() - before == inv, tag + " take invalid count"
        assert_bits_equal(Take(c, total, first, n).collect(), want, tag + f" host take({total})[{first}:+{n}]")
        done += 1


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
