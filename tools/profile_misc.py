"""ncu driver for the kernels outside the five BASELINE configs: runs one case a few times on cuda:0 and prints its
CUDA-event rate.   usage: python tools/profile_misc.py <case> [arg]

  linear_take     Linear over 64 equidistant stops [f32;4], take(2^28)                    (lean kernel)
  bezier_take     cubic Bezier [f32;4], take(2^28)                                          (lean kernel)
  linear_sorted   Linear over 64 sorted knots [f32;4], take(2^28)                           (linear_sorted_take_kernel)
  rows_batch [n]  cubic clamped equidistant [f32;4] B-spline, n stops (default 9), random batch of 2^28 (rows VARIANT 0)
  sorted_take     cubic clamped sorted-knot [f32;4] B-spline, 4096 control points, take(2^28) (rows VARIANT 3 / 4)
  many_flat       2^24 cubic Bezier curves [f32;3] x 4 samples                               (bezier_many_flat_kernel)
  scalar_take     scalar f32 cubic clamped equidistant B-spline, 20 control points, take(2^28)
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enterpolation_b200 import Bezier, BSpline, Linear, bezier_many_take  # noqa: E402


def main():
    which = sys.argv[1]
    arg = int(sys.argv[2]) if len(sys.argv) > 2 else None
    n = 1 << 28
    el4 = bench.u01(64 * 4, 1, np.float32).reshape(64, 4)
    out4 = lambda: torch.empty((n, 4), dtype=torch.float32, device="cuda")  # noqa: E731
    if which == "linear_take":
        c, out = Linear.builder().elements(el4).equidistant("f32").normalized().build(), out4()
        run = lambda: c.take_device(n, out=out)  # noqa: E731
    elif which == "bezier_take":
        c, out = Bezier.builder().elements(el4[:4]).normalized("f32").constant().build(), out4()
        run = lambda: c.take_device(n, out=out)  # noqa: E731
    elif which == "linear_sorted":
        kn = np.cumsum(0.5 + bench.u01(64, 3, np.float64)).astype(np.float32)
        c, out = Linear.builder().elements(el4).knots(kn).build(), out4()
        run = lambda: c.take_device(n, out=out)  # noqa: E731
    elif which == "rows_batch":
        stops = arg or 9
        el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
        c = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic().build()
        t, out = torch.rand(n, device="cuda", dtype=torch.float32), out4()
        run = lambda: c.sample_device(t, out=out)  # noqa: E731
    elif which == "sorted_take":
        stops = 4096
        el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
        k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)).astype(np.float32)
        c, out = BSpline.builder().clamped().elements(el).knots(k).dynamic().build(), out4()
        run = lambda: c.take_device(n, out=out)  # noqa: E731
    elif which == "many_flat":
        curves, samples = 1 << 24, 4
        n = curves * samples
        ctrl = torch.rand((curves, 4, 3), dtype=torch.float32, device="cuda")
        out = torch.empty((curves, samples, 3), dtype=torch.float32, device="cuda")
        run = lambda: bezier_many_take(ctrl, samples, out=out)  # noqa: E731
    elif which == "scalar_take":
        c = (BSpline.builder().clamped().elements(bench.u01(20, 1, np.float32)).equidistant("f32").degree(3)
             .domain(0.0, 3.0).dynamic().build())
        out = torch.empty(n, dtype=torch.float32, device="cuda")
        run = lambda: c.take_device(n, out=out)  # noqa: E731
    else:
        raise SystemExit(__doc__)
    for _ in range(4):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{which}: {n / (e0.elapsed_time(e1) / 3) / 1e6:.1f} G samples/s")


if __name__ == "__main__":
    main()
