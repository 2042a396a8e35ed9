"""bezier_many_take shapes beside C3 (catches pathologically slow corners)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from enterpolation_b200 import bezier_many_take
from sanity_perf_util import timeit
for (n_ctrl, dim, samples, curves) in [(16, 3, 256, 1 << 20), (4, 3, 256, 1 << 20), (4, 2, 32, 1 << 23), (16, 3, 16, 1 << 22),
                                       (3, 4, 64, 1 << 22), (8, 1, 1024, 1 << 18), (4, 3, 4, 1 << 24)]:
    ctrl = torch.rand((curves, n_ctrl, dim), dtype=torch.float32, device="cuda")
    out = torch.empty((curves, samples, dim), dtype=torch.float32, device="cuda")
    n = curves * samples
    bytes_per_sample = dim * 4 + n_ctrl * dim * 4 / samples
    timeit(lambda: bezier_many_take(ctrl, samples, out=out), n, f"many: {n_ctrl} ctrl [f32;{dim}] x {samples} samples x 2^{curves.bit_length()-1} curves", bytes_per_sample)
    del ctrl, out
