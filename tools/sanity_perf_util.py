import torch

def timeit(f, n, label, bytes_per_sample):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:58s} {n/ms/1e6:8.1f} G samples/s  {n*bytes_per_sample/ms/1e6:7.0f} GB/s")

