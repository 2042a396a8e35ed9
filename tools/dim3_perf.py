"""Throughput of 3-component ([f32;3], 12 B/sample) outputs, whose stores cannot be 128-bit per thread."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enterpolation_b200 import BSpline, Linear, Bezier
import bench
from sanity_perf_util import timeit
n = 1 << 28
el = bench.u01(64 * 3, 1, np.float32).reshape(64, 3)
out = torch.empty((n, 3), dtype=torch.float32, device="cuda")
c = Linear.builder().elements(el).equidistant("f32").normalized().build()
timeit(lambda: c.take_device(n, out=out), n, "Linear equidistant [f32;3] take", 12)
kn = np.cumsum(0.5 + bench.u01(64, 3, np.float64)).astype(np.float32)
c = Linear.builder().elements(el).knots(kn).build()
timeit(lambda: c.take_device(n, out=out), n, "Linear sorted [f32;3] take", 12)
c = Bezier.builder().elements(el[:4]).normalized("f32").constant().build()
timeit(lambda: c.take_device(n, out=out), n, "Bezier cubic [f32;3] take", 12)
c = BSpline.builder().clamped().elements(el[:5]).equidistant("f32").degree(3).normalized().dynamic().build()
timeit(lambda: c.take_device(n, out=out), n, "cubic B-spline 5 stops [f32;3] take (rows)", 12)
out1 = torch.empty(n, dtype=torch.float32, device="cuda")
c = Linear.builder().elements(el[:, 0].copy()).equidistant("f32").normalized().build()
timeit(lambda: c.take_device(n, out=out1), n, "Linear equidistant f32 scalar take", 4)
out2 = torch.empty((n, 2), dtype=torch.float32, device="cuda")
c = Linear.builder().elements(el[:, :2].copy()).equidistant("f32").normalized().build()
timeit(lambda: c.take_device(n, out=out2), n, "Linear equidistant [f32;2] take", 8)
