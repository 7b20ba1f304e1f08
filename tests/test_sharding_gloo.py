"""N > 1 host logic on CPU: two gloo ranks shard a ROI the way bench.py / MFModel.fit do
(contiguous spans, no data-path collective), gather the rows on rank 0 and max-reduce the
step time."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from microstructure_fingerprinting_b200.mf import shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, V, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = shard_bounds(V, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    # stand-in for the per-rank fit: row i of the result is a function of the voxel index
    rows = torch.arange(lo, hi, dtype=torch.float64)[:, None] * torch.tensor([[1.0, 2.0, 3.0]])
    parts = [None] * world
    dist.all_gather_object(parts, (lo, rows.numpy()))   # uneven shards: the final host-side gather
    gathered = [torch.from_numpy(r) for _, r in sorted(parts, key=lambda t: t[0])]
    t = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # bench.py: max over ranks of the step time
    dist.barrier()
    if rank == 0:
        out_q.put((torch.cat(gathered).numpy(), float(t.item())))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    V, world = 1001, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, V, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert rows.shape == (V, 3)
    assert np.array_equal(rows[:, 0], np.arange(V, dtype=float))
    assert tmax == 11.0


def test_shard_bounds_cover_everything():
    for V in (1, 2, 7, 1000, 12345):
        for w in (1, 2, 3, 8):
            b = shard_bounds(V, w)
            assert b[0] == 0 and b[-1] == V and np.all(np.diff(b) >= 0)
            assert np.max(np.diff(b)) - np.min(np.diff(b)) <= 1


def test_shard_bounds_balance_cost():
    """Cost-weighted shards: monotone boundaries covering everything, per-shard cost within one
    item of the ideal share."""
    from microstructure_fingerprinting_b200.mf import voxel_cost
    rng = np.random.default_rng(0)
    K = rng.integers(0, 3, 5000)
    cost = voxel_cost(K, rng.integers(0, 2, 5000), np.zeros(5000), 1000, 4)
    for w in (1, 2, 3, 8):
        b = shard_bounds(5000, w, cost)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0) and b.size == w + 1
        share = np.array([cost[b[i]:b[i + 1]].sum() for i in range(w)])
        assert np.all(np.abs(share - cost.sum() / w) <= cost.max() + 1e-9)
    assert np.array_equal(shard_bounds(10, 4, np.zeros(10)), shard_bounds(10, 4))
    assert np.array_equal(shard_bounds(0, 3, np.zeros(0)), np.zeros(4, dtype=np.int64))


def test_shard_bounds_edge_cases():
    """Equal costs reduce to the equal-count split; one huge voxel gets a shard of its own
    and leaves empty shards rather than overlapping ones; more shards than items."""
    V = 1000
    b = shard_bounds(V, 4, np.full(V, 7.0))
    assert np.max(np.abs(b - shard_bounds(V, 4))) <= 1
    cost = np.ones(100)
    cost[40] = 1e9
    b = shard_bounds(100, 4, cost)
    assert b[0] == 0 and b[-1] == 100 and np.all(np.diff(b) >= 0)
    owner = np.searchsorted(b, 40, side="right") - 1
    assert b[owner] <= 40 < b[owner + 1]
    spans = [(int(b[i]), int(b[i + 1])) for i in range(4)]
    assert sum(hi - lo for lo, hi in spans) == 100
    b = shard_bounds(3, 8, np.ones(3))
    assert b[0] == 0 and b[-1] == 3 and b.size == 9 and np.all(np.diff(b) >= 0) and np.all(np.diff(b) <= 1)
    b = shard_bounds(3, 8)
    assert b[0] == 0 and b[-1] == 3 and b.size == 9 and np.all(np.diff(b) >= 0)
