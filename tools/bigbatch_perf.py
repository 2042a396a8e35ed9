"""Random-parameter batches (Signal::sample) over tables too big for shared memory."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enterpolation_b200 import BSpline, Linear
import bench
from sanity_perf_util import timeit
n = 1 << 27
t = torch.rand(n, device="cuda", dtype=torch.float32)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for stops in (256, 4096, 65536):
    el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
    c = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic().build()
    timeit(lambda: c.sample_device(t, out=out4), n, f"cubic equidistant [f32;4], {stops} stops, random batch", 20)
    k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)); k = ((k - k[0]) / (k[-1] - k[0])).astype(np.float32)
    c = BSpline.builder().clamped().elements(el).knots(k).dynamic().build()
    timeit(lambda: c.sample_device(t, out=out4), n, f"cubic sorted-knot [f32;4], {stops} stops, random batch", 20)
    c = Linear.builder().elements(el).equidistant("f32").normalized().build()
    timeit(lambda: c.sample_device(t, out=out4), n, f"Linear equidistant [f32;4], {stops} stops, random batch", 20)
    kk = np.cumsum(0.5 + bench.u01(stops, 2, np.float64)); kk = ((kk - kk[0]) / (kk[-1] - kk[0])).astype(np.float32)
    c = Linear.builder().elements(el).knots(kk).build()
    timeit(lambda: c.sample_device(t, out=out4), n, f"Linear sorted-knot [f32;4], {stops} stops, random batch", 20)
