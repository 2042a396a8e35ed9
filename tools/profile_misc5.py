"""ncu driver: scalar f32 cubic clamped equidistant B-spline, 20 control points, take(2^28)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import BSpline
import bench
n = 1 << 28
c = BSpline.builder().clamped().elements(bench.u01(20, 1, np.float32)).equidistant("f32").degree(3).domain(0.0, 3.0).dynamic().build()
out = torch.empty(n, dtype=torch.float32, device="cuda")
for _ in range(6):
    c.take_device(n, out=out)
torch.cuda.synchronize()
print("done")
