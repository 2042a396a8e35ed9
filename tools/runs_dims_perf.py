"""Stepper-run kernel (rows VARIANT 4) for 1..4 output components and weighted curves: take(2^31) of a cubic clamped
equidistant f32 B-spline (ENTERP_NO_RUNS=1 for the per-sample kernels)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enterpolation_b200 import BSpline  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402

n = 1 << 31
tag = "per-sample" if os.environ.get("ENTERP_NO_RUNS") == "1" else "runs"
for dim, weighted in ((1, 0), (2, 0), (3, 0), (4, 0), (3, 1)):
    el = bench.u01(40 * dim, 1, np.float32).reshape(40, dim) if dim > 1 else bench.u01(40, 1, np.float32)
    b = BSpline.builder().clamped()
    b = b.elements_with_weights(el, (0.5 + bench.u01(40, 3, np.float32)).astype(np.float32)) if weighted else b.elements(el)
    c = b.equidistant("f32").degree(3).normalized().dynamic().build()
    out = torch.empty((n, dim) if dim > 1 else (n,), dtype=torch.float32, device="cuda")
    timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] cubic 40 stops f32 dim {dim} weighted {weighted} take(2^31)", 4 * dim)
    del out
