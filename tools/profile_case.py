"""Small driver for ncu captures: runs one BASELINE.json config shape a few times on cuda:0.
usage: python tools/profile_case.py c5 [log2_samples] [stops]   (also: c1, c2, c3, c4)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "c5"
    log2n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
    arg3 = int(sys.argv[3]) if len(sys.argv) > 3 else None
    w = bench.make_workload(case, log2n, arg3)
    for _ in range(3):
        w.step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        w.step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{case}: {w.samples} samples, {ms:.4f} ms/step, {w.samples / ms / 1e6:.2f} G samples/s")


if __name__ == "__main__":
    main()
