"""ncu driver for two non-config kernels: Linear equidistant take and cubic Bezier take ([f32;4], 2^28)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import Linear, Bezier
import bench
n = 1 << 28
el4 = bench.u01(64 * 4, 1, np.float32).reshape(64, 4)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
which = sys.argv[1]
c = (Linear.builder().elements(el4).equidistant("f32").normalized().build() if which == "linear"
     else Bezier.builder().elements(el4[:4]).normalized("f32").constant().build())
for _ in range(6):
    c.take_device(n, out=out4)
torch.cuda.synchronize()
print("done", which)
