"""Throughput sanity of paths outside the five BASELINE configs (catches pathologically slow kernels)."""
import os, sys, math
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import BSpline, Linear, Bezier, easing
import bench

from sanity_perf_util import timeit

n = 1 << 28
el4 = bench.u01(64 * 4, 1, np.float32).reshape(64, 4)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
c = Linear.builder().elements(el4).equidistant("f32").normalized().build()
timeit(lambda: c.take_device(n, out=out4), n, "Linear equidistant [f32;4] take", 16)
c = Linear.builder().elements(el4).equidistant("f32").normalized().easing(easing.smoothstep).build()
timeit(lambda: c.take_device(n, out=out4), n, "Linear equidistant [f32;4] smoothstep take", 16)
kn = np.cumsum(0.5 + bench.u01(64, 3, np.float64)).astype(np.float32)
c = Linear.builder().elements(el4).knots(kn).build()
timeit(lambda: c.take_device(n, out=out4), n, "Linear sorted knots (64) [f32;4] take", 16)
kn = np.cumsum(0.5 + bench.u01(1 << 16, 3, np.float64)).astype(np.float32)
c = Linear.builder().elements(bench.u01(4 << 16, 1, np.float32).reshape(-1, 4)).knots(kn).build()
timeit(lambda: c.take_device(n, out=out4), n, "Linear sorted knots (65536) [f32;4] take", 16)
tb = torch.rand(n, device="cuda", dtype=torch.float32)
c = Linear.builder().elements(el4).equidistant("f32").normalized().build()
timeit(lambda: c.sample_device(tb, out=out4), n, "Linear equidistant [f32;4] random batch", 20)
c = Bezier.builder().elements(el4[:4]).normalized("f32").constant().build()
timeit(lambda: c.take_device(n, out=out4), n, "Bezier cubic [f32;4] take", 16)
timeit(lambda: c.sample_device(tb, out=out4), n, "Bezier cubic [f32;4] random batch", 20)
c = Bezier.builder().elements(el4[:8]).normalized("f32").constant().build()
timeit(lambda: c.take_device(n, out=out4), n, "Bezier degree 7 [f32;4] take", 16)
w = math.sqrt(2) / 2
circle = BSpline.builder().elements_with_weights([(1., 0.), (1., 1.), (0., 1.), (-1., 1.), (-1., 0.), (-1., -1.), (0., -1.), (1., -1.), (1., 0.)],
                                                 [1, w, 1, w, 1, w, 1, w, 1]).knots([0., 0., 1., 1., 2., 2., 3., 3., 4., 4.]).constant(3).build()
out2 = torch.empty((n, 2), dtype=torch.float64, device="cuda")
timeit(lambda: circle.take_device(n, out=out2), n, "NURBS unit circle f64 (sorted knots, rows) take", 16)
c = BSpline.builder().clamped().elements(el4[:9]).equidistant("f32").degree(3).normalized().dynamic().build()
t = torch.rand(n, device="cuda", dtype=torch.float32)
timeit(lambda: c.sample_device(t, out=out4), n, "cubic equidistant [f32;4] random batch (rows)", 20)
cc = c.clamp()
timeit(lambda: cc.take_device(n, out=out4), n, "same curve .clamp() take (checked rows variant)", 16)
c = BSpline.builder().clamped().elements(bench.u01(12, 2, np.float32)).equidistant("f32").degree(5).normalized().dynamic().build()
out1 = torch.empty(n, dtype=torch.float32, device="cuda")
timeit(lambda: c.take_device(n, out=out1), n, "quintic equidistant f32 scalar take (rows)", 4)

# launch-bound regime: 200 takes of 2^14 samples, direct calls vs one CUDA graph replay
c = BSpline.builder().clamped().elements(el4[:9]).equidistant("f32").degree(3).normalized().dynamic().build()
small = [torch.empty((1 << 14, 4), dtype=torch.float32, device="cuda") for _ in range(200)]
def direct():
    for o in small:
        c.take_device(1 << 14, out=o)
direct(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    direct()
for name, f in (("direct calls", direct), ("CUDA graph replay", g.replay)):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    print(f"200 x take(2^14) cubic [f32;4], {name:18s} {e0.elapsed_time(e1) / 10 / 200 * 1e3:7.2f} us per take")
