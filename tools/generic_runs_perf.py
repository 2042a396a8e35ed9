"""f32 take(2^31) of curves outside the span-row kernels: generic stepper runs (capi.cu generic_runs_take) vs the
per-sample generic kernels (ENTERP_NO_GENERIC_RUNS=1)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench  # noqa: E402
from enterpolation_b200 import Bezier, BSpline, Linear  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402

n = 1 << 31
tag = "per-sample" if os.environ.get("ENTERP_NO_GENERIC_RUNS") == "1" else "generic runs"
out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
el = lambda k: bench.u01(k * 4, 1, np.float32).reshape(k, 4)  # noqa: E731
kn = lambda k: np.cumsum(0.5 + bench.u01(k, 2, np.float64)).astype(np.float32)  # noqa: E731
cases = {
    "B-spline degree 7 clamped equidistant, 20 stops": BSpline.builder().clamped().elements(el(20)).equidistant("f32").degree(7).normalized().dynamic(),
    "B-spline degree 5 sorted, 20 stops": BSpline.builder().clamped().elements(el(20)).knots(kn(16)).dynamic(),
    "B-spline runtime degree 9 sorted": BSpline.builder().clamped().elements(el(20)).knots(kn(12)).dynamic(),
    "Bezier 12 control points": Bezier.builder().elements(el(12)).normalized("f32").constant(),
    "Linear 64 stops (lean kernel up to 2^32 - 2^24)": Linear.builder().elements(el(64)).equidistant("f32").normalized(),
}
for name, b in cases.items():
    c = b.build()
    timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] {name} [f32;4] take(2^31)", 16)
