"""Why is the host-buffer take slower inside bench.py than in tools/e2e_rle.py? Replays bench.py's set-up in stages
and times the same enterp_take_host call after each.  usage: python tools/e2e_context_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from enterpolation_b200 import _abi  # noqa: E402

lib = _abi.load_library()
n_total, count = 1 << 32, 1 << 28
first = (n_total - count) // 2
builder = bench.gradient_builder(5)


def e2e(tag, host):
    def step():
        c = builder.on_device(0).build()
        assert lib.enterp_take_host(c._h, 0, n_total, first, count, host.data_ptr()) == 0
        return float(host[-1, 0])
    for _ in range(2):
        step()
    t0 = time.perf_counter()
    for _ in range(4):
        step()
    dt = (time.perf_counter() - t0) / 4
    print(f"{tag}: {count / dt / 1e9:.2f} G samples/s", flush=True)


host = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
host.zero_()
e2e("fresh process", host)
w = bench.C5(32, 5).setup()
for _ in range(5):
    w.step()
torch.cuda.synchronize()
e2e("after C5 take(2^32) steps, 64 GiB output still allocated", host)
del w
torch.cuda.empty_cache()
e2e("after freeing it", host)
del host
host2 = torch.empty((count, 4), dtype=torch.float32, pin_memory=True)
host2.zero_()
e2e("new pinned buffer", host2)
s = bench.ClockSampler(0)
s.start()
time.sleep(0.2)
s.finish()
e2e("after a clock sampler thread", host2)
