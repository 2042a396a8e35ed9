"""take() throughput of curves too big for the 64 KB shared-memory span-row table."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import BSpline
import bench
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sanity_perf_util import timeit

n = 1 << 28
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for stops in (256, 1024, 4096, 65536):
    el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
    c = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic().build()
    timeit(lambda: c.take_device(n, out=out4), n, f"cubic clamped equidistant [f32;4], {stops} stops, take(2^28)", 16)
    k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)).astype(np.float32)
    c = BSpline.builder().clamped().elements(el).knots(k).dynamic().build()
    timeit(lambda: c.take_device(n, out=out4), n, f"cubic clamped sorted-knot [f32;4], {stops} stops, take(2^28)", 16)
del out4
m = 1 << 26
w = bench.make_workload("c4")
c = w.curve
out3 = torch.empty((m, 3), dtype=torch.float64, device="cuda")
timeit(lambda: c.take_device(m, out=out3), m, "C4 curve (NURBS f64, 64Ki ctrl), take(2^26)", 24)
