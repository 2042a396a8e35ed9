"""Breaks the e2e step of bench.py into its parts (pure D2H bandwidth, take_host, build)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from enterpolation_b200 import _abi
lib = _abi.load_library()
n = 1 << 28
host = torch.empty((n, 4), dtype=torch.float32, pin_memory=True); host.zero_()
dev = torch.empty((n, 4), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); host.copy_(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"pure D2H 4 GiB pinned: {dt*1e3:.1f} ms, {4.295/dt:.1f} GB/s")
b = bench.gradient_builder(5)
c = b.build(); c.upload(0)
for _ in range(2):
    t0 = time.perf_counter(); st = lib.enterp_take_host(c._h, 0, 1 << 32, 1 << 31, n, host.data_ptr()); dt = time.perf_counter() - t0
print(f"take_host 2^28: {dt*1e3:.1f} ms, {n/dt/1e9:.2f} G samples/s, {n*16/dt/1e9:.1f} GB/s (status {st})")
t0 = time.perf_counter()
for _ in range(20):
    c2 = b.build(); c2.upload(0); del c2
dt = (time.perf_counter() - t0) / 20
print(f"build+upload+destroy: {dt*1e3:.3f} ms")
