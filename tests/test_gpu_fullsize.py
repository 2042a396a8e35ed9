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


# ---- every sample of the full-size configs against the multi-threaded oracle, in slices ---------------
def _compare_in_slices(out, n, slice_n, oracle_slice, what):
    for lo in range(0, n, slice_n):
        hi = min(n, lo + slice_n)
        assert_bits_equal(out[lo:hi].cpu().numpy(), oracle_slice(lo, hi), f"{what} [{lo}, {hi})")


@pytest.mark.fullsize
def test_c2_every_sample_of_the_full_batch():
    w = bench.C2(28, 20).setup()  # 2^20 knots, 2^28 random parameters
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    _compare_in_slices(w.out, w.samples, 1 << 24, lambda lo, hi: ref.eval(w.t[lo:hi].cpu().numpy(), threads=THREADS), "C2 2^28")


@pytest.mark.fullsize
def test_c3_every_curve_of_the_full_set():
    w = bench.C3(20, 16, 256).setup()  # 2^20 curves x 256 samples
    w.step()
    _compare_in_slices(w.out, w.curves, 1 << 16,
                       lambda lo, hi: oracle.bezier_many_take(w.ctrl[lo:hi].cpu().numpy(), 256, threads=THREADS), "C3 2^20 curves")


@pytest.mark.fullsize
def test_c4_every_sample_of_the_full_batch():
    w = bench.C4(26, 16).setup()
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    _compare_in_slices(w.out, w.samples, 1 << 23, lambda lo, hi: ref.eval(w.t[lo:hi].cpu().numpy(), threads=THREADS), "C4 2^26")


@pytest.mark.fullsize
@pytest.mark.parametrize("stops", [5, 256])
def test_c5_every_sample_of_take_2_32(stops):
    """All 4 294 967 296 samples of the headline take, bit for bit against the oracle (the stepper-run kernel serves
    everything from index 2^24 on). Needs the 64 GiB output resident; skipped on smaller devices."""
    n = 1 << 32
    if torch.cuda.get_device_properties(0).total_memory < 80 * 2 ** 30:
        pytest.skip("needs 64 GiB of device memory")
    w = bench.C5(32, stops).setup()
    w.step()
    ref = oracle.OracleCurve(w.builder.to_desc())
    host = np.empty((1 << 26, 4), dtype=np.float32)
    for lo in range(0, n, 1 << 26):
        want = ref.take(n, lo, 1 << 26, threads=THREADS, out=host)
        got = w.out[lo:lo + (1 << 26)].cpu().numpy()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"C5 {stops} stops, slice at {lo}"
