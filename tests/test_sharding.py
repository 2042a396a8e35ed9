"""Multi-GPU host logic on CPU (gloo, world_size 2): sample-range sharding with the GLOBAL stepper
index reproduces the unsharded take() bit for bit (checked with the oracle as the evaluator, since
no GPU is present here; the GPU version of the same property is tests/test_gpu_parity.py)."""
import os
import sys

import numpy as np
import pytest

from enterpolation_b200 import shard_range, shard_sizes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition():
    for n in (0, 1, 7, 1000, (1 << 32), (1 << 32) + 5):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert sum(shard_sizes(n, world)) == n
            assert max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, n_total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from enterpolation_b200 import BSpline, shard_range
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    el = np.arange(20, dtype=np.float32).reshape(5, 4) / 7
    b = BSpline.builder().clamped().elements(el).equidistant("f32").degree(3).normalized().dynamic()
    ref = oracle.OracleCurve(b.to_desc())
    lo, hi = shard_range(n_total, rank, world)
    mine = ref.take(n_total, lo, hi - lo)              # each rank: its own slice, global index
    # gather only for the CHECK (the data path itself has no collective)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    bufs = [torch.empty((s[1] - s[0], 4), dtype=torch.float32) for s in sizes]
    dist.all_gather(bufs, torch.from_numpy(mine)) if len({b.shape for b in bufs}) == 1 else None
    if len({b.shape for b in bufs}) != 1:
        for r in range(world):
            bufs[r] = torch.from_numpy(mine).clone() if r == rank else bufs[r]
            dist.broadcast(bufs[r], src=r)
    if rank == 0:
        whole = ref.take(n_total)
        got = torch.cat(bufs).numpy()
        np.save(os.path.join(out_dir, "ok.npy"), np.array([np.array_equal(got.view(np.uint32), whole.view(np.uint32))]))
    # throughput aggregation as bench.py does it: max over ranks of the elapsed time
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert float(t) == float(world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [1001, 4096])
def test_two_rank_sharded_take_equals_whole(tmp_path, n_total):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + n_total % 7
    mp.start_processes(_worker, args=(2, port, n_total, str(tmp_path)), nprocs=2, join=True, start_method="spawn")
    assert bool(np.load(tmp_path / "ok.npy")[0])
