"""ncu driver: random batch of a cubic clamped equidistant [f32;4] B-spline.  usage: python tools/profile_batch.py [stops] [log2n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enterpolation_b200 import BSpline  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402

stops = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 26)
t32 = torch.rand(n, device="cuda", dtype=torch.float32)
out4 = torch.empty((n, 4), dtype=torch.float32, device="cuda")
el = bench.u01(stops * 4, 1, np.float32).reshape(stops, 4)
c = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic().build()
timeit(lambda: c.sample_device(t32, out=out4), n, f"cubic equidistant [f32;4] {stops} stops random batch", 20)
