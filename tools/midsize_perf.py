"""Random batches over sorted knot vectors of 2^10 .. 2^18 knots: Linear [f32;3] / [f32;4], cubic B-spline [f32;4],
NURBS [f64;3] (C4 shape). Used in round 2 to compare the bucket index in L2 with a copy staged in shared memory (a
net loss, profiles/r02_midsize_staged_index_negative.txt) and with the slot table (ENTERP_SLOTS_PER_KNOT)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enterpolation_b200 import BSpline, Linear  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402

n = 1 << 26
tag = "slots " + os.environ.get("ENTERP_SLOTS_PER_KNOT", "default")
for log2k in (10, 12, 14, 16, 18):
    k = 1 << log2k
    kn = np.cumsum(0.5 + bench.u01(k, 0xE2, np.float64))
    kn32 = kn.astype(np.float32)
    t = (torch.rand(n, device="cuda", dtype=torch.float64) * float(kn[-1] - kn[0]) + float(kn[0]))
    t32 = t.to(torch.float32)
    for dim in (3, 4):
        el = (bench.u01(k * dim, 0xE1, np.float32) * 2 - 1).reshape(k, dim)
        c = Linear.builder().elements(el).knots(kn32).build()
        out = torch.empty((n, dim), dtype=torch.float32, device="cuda")
        timeit(lambda: c.sample_device(t32, out=out), n, f"[{tag}] Linear [f32;{dim}] 2^{log2k} sorted knots, random batch", 4 + 4 * dim)
    el = (bench.u01((k - 2) * 4, 0xE1, np.float32) * 2 - 1).reshape(k - 2, 4)
    c = BSpline.builder().elements(el).knots(kn32).constant(4).build()
    out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    timeit(lambda: c.sample_device(t32, out=out), n, f"[{tag}] cubic BSpline [f32;4] 2^{log2k} sorted knots, random batch", 20)
    pts = (bench.u01((k - 2) * 3, 0xE1, np.float64) * 2 - 1).reshape(k - 2, 3)
    w = 0.5 + bench.u01(k - 2, 0xE3, np.float64)
    c = BSpline.builder().elements_with_weights(pts, w).knots(kn).constant(4).build()
    out = torch.empty((n, 3), dtype=torch.float64, device="cuda")
    timeit(lambda: c.sample_device(t, out=out), n, f"[{tag}] NURBS [f64;3] 2^{log2k} sorted knots, random batch", 32)
