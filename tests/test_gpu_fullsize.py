"""BASELINE.json configs at (or near) their full sizes: bit-exact against the multi-threaded oracle where it
finishes in seconds, and size-independent properties (shards == whole on the device, convex hull, spans
monotone) where it does not. Workloads come from bench.py so that tests and benchmark cannot drift apart."""
import os

import numpy as np
import pytest
import torch

import bench
from oracle import oracle

from util import assert_bits_equal

pytestmark = pytest.mark.gpu
THREADS = max(1, min(32, os.cpu_count() or 1))


def test_c1_full_size_bit_exact():
    w = bench.C1(24).setup()
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    assert_bits_equal(w.out.cpu().numpy(), ref.take(1 << 24, threads=THREADS), "C1 take(2^24)")
    idx, _ = w.curve.span_take_device(1 << 24)
    idx = idx.cpu().numpy()
    assert np.all(np.diff(idx.astype(np.int64)) >= 0) and idx.min() == 3 and idx.max() == 6


def test_c2_full_knot_table_bit_exact():
    w = bench.C2(24, 20).setup()  # the full 2^20-knot table, 2^24 of the random parameters
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    t = w.t.cpu().numpy()
    assert_bits_equal(w.out.cpu().numpy(), ref.eval(t, threads=THREADS), "C2")
    # adversarial parameters: every 4096-th replaced by an exact knot / its neighbours / out of domain
    knots = np.asarray(w.builder._knots, dtype=np.float32)
    t2 = t.copy()
    pick = np.arange(0, t2.shape[0], 4096)
    kn = knots[(pick * 2654435761 % knots.shape[0])]
    t2[pick] = np.where(pick % 3 == 0, kn, np.where(pick % 3 == 1, np.nextafter(kn, np.float32(np.inf)), np.nextafter(kn, np.float32(-np.inf))))
    t2[pick[::50]] = knots[-1] + 17.0
    t2[pick[1::50]] = knots[0] - 17.0
    got = w.curve.sample_device(torch.from_numpy(t2).cuda()).cpu().numpy()
    assert_bits_equal(got, ref.eval(t2, threads=THREADS), "C2 adversarial")


def test_c3_many_curves_bit_exact():
    w = bench.C3(16, 16, 256).setup()  # 2^16 curves x 256 samples
    w.step()
    assert_bits_equal(w.out.cpu().numpy(), oracle.bezier_many_take(w.ctrl.cpu().numpy(), 256, threads=THREADS), "C3")


def test_c4_full_table_bit_exact():
    w = bench.C4(22, 16).setup()  # the full 2^16 control points, 2^22 of the random parameters
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    assert_bits_equal(w.out.cpu().numpy(), ref.eval(w.t.cpu().numpy(), threads=THREADS), "C4")


@pytest.mark.parametrize("stops", [5, 256])
def test_c5_large_take_properties(stops):
    log2n = 30
    w = bench.C5(log2n, stops).setup()
    w.step()
    n = 1 << log2n
    out = w.out
    assert bool(torch.isfinite(out).all())
    # convex hull: a clamped B-spline inside its domain is a convex combination of its control points
    el = torch.from_numpy(bench.u01(stops * 4, 0xE1, np.float32).reshape(stops, 4)).cuda()
    lo, hi = el.min(0).values - 1e-6, el.max(0).values + 1e-6
    assert bool(((out >= lo) & (out <= hi)).all())
    # end points interpolate the first / last control point exactly (clamped)
    assert torch.equal(out[0], el[0]) and torch.equal(out[-1], el[-1])
    # shards computed separately (global index!) equal the whole, bit for bit, on the device
    for world in (2, 8):
        for rank in (0, world - 1, world // 2):
            s_lo, s_hi = (n * rank) // world, (n * (rank + 1)) // world
            part = w.curve.take_device(n, s_lo, s_hi - s_lo)
            assert torch.equal(part.view(torch.int32), out[s_lo:s_hi].view(torch.int32)), (world, rank)
            del part
    # and windows of it equal the oracle
    ref = oracle.OracleCurve(w.builder.to_desc())
    for first in (0, n // 3, n - (1 << 20)):
        assert_bits_equal(out[first:first + (1 << 20)].cpu().numpy(), ref.take(n, first, 1 << 20, threads=THREADS), f"window {first}")
