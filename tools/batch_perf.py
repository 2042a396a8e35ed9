"""Random batches (Signal::sample) of cubic clamped equidistant / sorted-knot B-splines [f32;4] and [f64;3] over
5 .. 1000 control points: per-sample shared-memory rows variant (round 2 also measured a tile-sorted variant with it: profiles/r02_batch_tile_sort_negative.txt)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enterpolation_b200 import BSpline  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402

n = 1 << 27
tag = "per-sample" if os.environ.get("ENTERP_NO_SORTED_BATCH") == "1" else "tile-sorted"
t32 = torch.rand(n, device="cuda", dtype=torch.float32)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
for stops in (5, 9, 20, 40, 100, 200, 330, 1000):
    el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
    c = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic().build()
    timeit(lambda: c.sample_device(t32, out=out4), n, f"[{tag}] cubic equidistant [f32;4] {stops} stops", 20)
    k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)); k = ((k - k[0]) / (k[-1] - k[0])).astype(np.float32)
    c = BSpline.builder().clamped().elements(el).knots(k).dynamic().build()
    timeit(lambda: c.sample_device(t32, out=out4), n, f"[{tag}] cubic sorted-knot [f32;4] {stops} stops", 20)
t64 = torch.rand(n // 2, device="cuda", dtype=torch.float64)
out3 = torch.empty((n // 2, 3), dtype=torch.float64, device="cuda")
for stops in (9, 100, 300):
    pts = bench.u01(stops * 3, 1, np.float64).reshape(stops, 3)
    w = 0.5 + bench.u01(stops, 3, np.float64)
    k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)); k = (k - k[0]) / (k[-1] - k[0])
    c = BSpline.builder().clamped().elements_with_weights(pts, w).knots(k).dynamic().build()
    timeit(lambda: c.sample_device(t64, out=out3), n // 2, f"[{tag}] NURBS cubic sorted-knot [f64;3] {stops} points", 32)
