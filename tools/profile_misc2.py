"""ncu driver: Linear over 64 sorted knots, [f32;4], take(2^28)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from enterpolation_b200 import Linear
import bench
n = 1 << 28
el4 = bench.u01(64 * 4, 1, np.float32).reshape(64, 4)
kn = np.cumsum(0.5 + bench.u01(64, 3, np.float64)).astype(np.float32)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
c = Linear.builder().elements(el4).knots(kn).build()
for _ in range(6):
    c.take_device(n, out=out4)
torch.cuda.synchronize()
print("done")
