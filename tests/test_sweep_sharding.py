"""Host logic of the multi-GPU path: frames of a -d sweep are dealt to ranks with no
collective on the data path; the only cross-rank step is the timing reduction.  world_size-2 gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_frames_partition():
    from f3ps import sweep
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 4, 8):
            parts = [sweep.shard_frames(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert all(p == sorted(p) for p in parts)
    with pytest.raises(ValueError):
        sweep.shard_frames(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from f3ps import sweep
    mine = sweep.shard_frames(11, rank, world)
    pts = 307200 * len(mine)
    secs = 0.5 * (rank + 1)
    total, tmax = sweep.reduce_sweep(pts, secs, dist)
    q.put((rank, mine, total, tmax))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_reduction():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6, 8, 10] and res[1][1] == [1, 3, 5, 7, 9]
    for r in res:
        assert r[2] == 307200 * 11 and r[3] == 1.0
    from f3ps import sweep
    assert abs(sweep.aggregate_throughput([307200 * 6, 307200 * 5], [0.5, 1.0]) - 3.3792) < 1e-9
