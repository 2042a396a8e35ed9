"""f32 take() beyond index 2^24: stepper-run kernel (rows VARIANT 4) vs the per-sample kernels (ENTERP_NO_RUNS=1 in
the environment selects the latter). usage: python tools/runs_perf.py [log2n]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sanity_perf_util import timeit  # noqa: E402
from enterpolation_b200 import BSpline  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    n = 1 << log2n
    tag = "per-sample" if os.environ.get("ENTERP_NO_RUNS") == "1" else "runs"
    out = torch.empty((n, 4), dtype=torch.float32, device="cuda")
    for stops in (5, 256, 4096):
        c = bench.gradient_builder(stops).build()
        timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] C5 {stops} stops take(2^{log2n})", 16)
    # sorted knots, cubic [f32;4]
    k = np.cumsum(0.5 + bench.u01(258, 0xE2, np.float64)).astype(np.float32)
    el = bench.u01(256 * 4, 0xE1, np.float32).reshape(256, 4)
    c = BSpline.builder().elements(el).knots(k).constant(4).build()
    timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] sorted cubic 256 take(2^{log2n})", 16)
    # Linear / Bezier takes of the same size (lean kernels up to index 2^32 - 2^24, the generic ones beyond)
    from enterpolation_b200 import Bezier, Linear
    el64 = bench.u01(64 * 4, 1, np.float32).reshape(64, 4)
    c = Linear.builder().elements(el64).equidistant("f32").normalized().build()
    timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] Linear 64 stops [f32;4] take(2^{log2n})", 16)
    c = Bezier.builder().elements(el64[:4]).normalized("f32").constant().build()
    timeit(lambda: c.take_device(n, out=out), n, f"[{tag}] cubic Bezier [f32;4] take(2^{log2n})", 16)
    # shard sizes of the 8-GPU run
    c = bench.gradient_builder(5).build()
    m = n // 8
    timeit(lambda: c.take_device(n, 5 * m, m, out=out[:m]), m, f"[{tag}] C5 5 stops shard 5/8 of take(2^{log2n})", 16)
    timeit(lambda: c.take_device(n, 0, m, out=out[:m]), m, f"[{tag}] C5 5 stops shard 0/8 of take(2^{log2n})", 16)


if __name__ == "__main__":
    main()
