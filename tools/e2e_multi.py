"""Single-process multi-GPU end to end: ONE enterp_take_host_multi call from one host thread spreads a window of the
C5 take over every GPU of the box (per-device double-buffered kernel/D2H pipelines), next to the limiter measured in
the same process: plain pinned cudaMemcpyAsync D2H of the same bytes on all devices at once.
usage: python tools/e2e_multi.py [log2_samples_per_gpu]"""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enterpolation_b200 import _abi  # noqa: E402


def main():
    log2 = int(sys.argv[1]) if len(sys.argv) > 1 else 27
    n_dev = torch.cuda.device_count()
    lib = _abi.load_library()
    n_total = 1 << 32
    for devices in ([0], list(range(n_dev))) if n_dev > 1 else ([0],):
        g = len(devices)
        count = g << log2
        first = (n_total - count) // 2
        host = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
        host.zero_()
        arr = (C.c_int * g)(*devices)
        curve = bench.gradient_builder(5).build()
        for d in devices:
            curve.upload(d)

        def step():
            st = lib.enterp_take_host_multi(curve._h, arr, g, n_total, first, count, host.data_ptr())
            assert st == 0, st

        for _ in range(2):
            step()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            step()
        dt = (time.perf_counter() - t0) / reps
        # limiter: every device copies its share into the same pinned buffer concurrently
        srcs = [torch.zeros((1 << log2, 4), dtype=torch.float32, device=f"cuda:{d}") for d in devices]
        streams = [torch.cuda.Stream(device=d) for d in devices]

        def probe():
            for k, d in enumerate(devices):
                with torch.cuda.stream(streams[k]):
                    host[k << log2:(k + 1) << log2].copy_(srcs[k], non_blocking=True)
            for s in streams:
                s.synchronize()

        for _ in range(2):
            probe()
        t0 = time.perf_counter()
        for _ in range(reps):
            probe()
        pt = (time.perf_counter() - t0) / reps
        print(f"{g} GPU(s): take_host_multi {count / dt / 1e9:.3f} G samples/s = {count * 16 / dt / 1e9:.1f} GB/s; "
              f"plain D2H probe {count * 16 / pt / 1e9:.1f} GB/s; ratio {pt / dt:.3f}")
        del host, srcs


if __name__ == "__main__":
    main()
