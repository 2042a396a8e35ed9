"""End-to-end take into pinned host memory with the run-length transfer: rate vs host threads (ENTERP_HOST_THREADS) and
against the plain transfer (ENTERP_NO_RLE=1).  usage: python tools/e2e_rle.py [log2_window]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enterpolation_b200 import _abi  # noqa: E402

log2 = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 28
lib = _abi.load_library()
n_total, count = 1 << 32, 1 << log2
first = (n_total - count) // 2
host = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
host.zero_()
builder = bench.gradient_builder(5)
rebuild = "--build" in sys.argv  # a fresh curve per step (builder validation + table upload), as bench.py's e2e leg does
curve = builder.build()


def step():
    global curve
    if rebuild:
        curve = builder.on_device(0).build()
    assert lib.enterp_take_host(curve._h, 0, n_total, first, count, host.data_ptr()) == 0
    return float(host[-1, 0])


for _ in range(2):
    step()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    step()
dt = (time.perf_counter() - t0) / reps
print(f"threads={os.environ.get('ENTERP_HOST_THREADS', 'auto')} rle={'off' if os.environ.get('ENTERP_NO_RLE') == '1' else 'on'}: "
      f"{count / dt / 1e9:.3f} G samples/s = {count * 16 / dt / 1e9:.1f} GB/s into host memory")
