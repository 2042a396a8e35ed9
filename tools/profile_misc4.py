"""ncu drivers: (a) cubic clamped sorted-knot [f32;4] B-spline, 4096 control points, take(2^28) (rows VARIANT 3);
(b) 2^24 cubic Bezier curves [f32;3] x 4 samples (bezier_many_flat_kernel)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import BSpline, bezier_many_take
import bench
which = sys.argv[1]
if which == "sorted":
    n, stops = 1 << 28, 4096
    el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
    k = np.cumsum(0.5 + bench.u01(stops - 2, 2, np.float64)).astype(np.float32)
    c = BSpline.builder().clamped().elements(el).knots(k).dynamic().build()
    out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    for _ in range(6):
        c.take_device(n, out=out)
else:
    curves, samples = 1 << 24, 4
    ctrl = torch.rand((curves, 4, 3), dtype=torch.float32, device="cuda")
    out = torch.empty((curves, samples, 3), dtype=torch.float32, device="cuda")
    for _ in range(6):
        bezier_many_take(ctrl, samples, out=out)
torch.cuda.synchronize()
print("done", which)
