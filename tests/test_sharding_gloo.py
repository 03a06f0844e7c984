"""N > 1 path on CPU: world_size-2 gloo job checks the frame partition and the max-over-ranks timing
reduction that bench.py uses (no data-path collective exists on this path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from raw_image_pipeline_b200 import sharding


def test_shard_range_tiles_the_batch():
    for n in (0, 1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)
    assert [sharding.stream_owner(s, 8) for s in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        b, e = sharding.shard_range(n_frames, rank, world)
        owned = torch.zeros(n_frames, dtype=torch.int32)
        owned[b:e] = 1
        dist.all_reduce(owned)  # every frame owned exactly once across the job
        dist.barrier()
        t = sharding.max_over_ranks(0.010 * (rank + 1))
        q.put((rank, bool((owned == 1).all()), t, sharding.job_throughput(e - b, t, world)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_job():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n_frames = 2, 64
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, t, thr in res:
        assert ok
        assert abs(t - 0.020) < 1e-12            # max over ranks, not the local time
        assert abs(thr - 64 / 0.020) < 1e-6      # whole-job frames/s
