"""ncu driver: cubic clamped equidistant [f32;4] B-spline, 9 stops, random batch of 2^28 parameters (rows VARIANT 0)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import BSpline
import bench
n = 1 << 28
stops = int(sys.argv[1]) if len(sys.argv) > 1 else 9
el4 = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
c = BSpline.builder().clamped().elements(el4).equidistant("f32").degree(3).normalized().dynamic().build()
t = torch.rand(n, device="cuda", dtype=torch.float32)
for _ in range(6):
    c.sample_device(t, out=out4)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    c.sample_device(t, out=out4)
e1.record(); torch.cuda.synchronize()
print(f"{stops} stops random batch: {n / (e0.elapsed_time(e1) / 5) / 1e6:.1f} G samples/s")
