"""take(2^26) throughput over a matrix of curve shapes: catches pathologically slow corners."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enterpolation_b200 import BSpline, Linear, Bezier
import bench

def rate(c, n, dim, tdt):
    out = torch.empty((n, dim) if dim > 1 else (n,), dtype=tdt, device="cuda")
    for _ in range(2): c.take_device(n, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): c.take_device(n, out=out)
    e1.record(); torch.cuda.synchronize()
    return n / (e0.elapsed_time(e1) / 3) / 1e6

n = 1 << 26
rows = []
for sc, dt, tdt in (("f32", np.float32, torch.float32), ("f64", np.float64, torch.float64)):
    for dim, w in ((1, 0), (3, 0), (4, 0), (3, 1)):
        def els(k):
            e = bench.u01(k * dim, 1, dt).reshape(k, dim) if dim > 1 else bench.u01(k, 1, dt)
            return e
        def add_el(b, k):
            return b.elements_with_weights(els(k), (0.5 + bench.u01(k, 3, dt)).astype(dt)) if w else b.elements(els(k))
        kn = lambda k: np.cumsum(0.5 + bench.u01(k, 2, np.float64)).astype(dt)
        cases = {
            "linear equi 32": lambda: add_el(Linear.builder(), 32).equidistant(sc).domain(0.0, 3.0),
            "linear sorted 32": lambda: add_el(Linear.builder(), 32).knots(kn(32)),
            "bezier 4": lambda: add_el(Bezier.builder(), 4).normalized(sc).constant(),
            "bezier 12": lambda: add_el(Bezier.builder(), 12).normalized(sc).constant(),
            "bspline deg2 equi 20": lambda: add_el(BSpline.builder().clamped(), 20).equidistant(sc).degree(2).domain(0.0, 3.0).dynamic(),
            "bspline deg3 equi 20": lambda: add_el(BSpline.builder().clamped(), 20).equidistant(sc).degree(3).domain(0.0, 3.0).dynamic(),
            "bspline deg3 sorted 20": lambda: add_el(BSpline.builder().clamped(), 20).knots(kn(18)).dynamic(),
            "bspline deg5 sorted 20": lambda: add_el(BSpline.builder().clamped(), 20).knots(kn(16)).dynamic(),
            "bspline deg7 equi 20": lambda: add_el(BSpline.builder().clamped(), 20).equidistant(sc).degree(7).domain(0.0, 3.0).dynamic(),
            "bspline deg9 sorted 20": lambda: add_el(BSpline.builder().clamped(), 20).knots(kn(12)).dynamic(),
        }
        for name, mk in cases.items():
            try:
                c = mk().build()
                r = rate(c, n, dim, tdt)
                print(f"{sc} dim{dim} w{w} {name:26s} {r:8.1f} G samples/s  {r * dim * dt().itemsize:7.0f} GB/s", flush=True)
            except Exception as ex:
                print(f"{sc} dim{dim} w{w} {name:26s} ERROR {ex!r}", flush=True)
