"""CPU suite for the multi-GPU host logic: frame sharding and the gloo statistics reduction with
world_size 2 (the data path itself has no collective)."""
import os
import socket

import pytest

from lidar_processing_v2_b200.stream import reduce_stats, shard_range


def test_shard_ranges_partition_the_stream():
    for n in (0, 1, 7, 154, 8192):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(154, rank, world)
    local = dict(frames=b - a, points=(b - a) * 1000 + rank, clusters=7 * (rank + 1))
    out = reduce_stats(local, elapsed_s=1.0 + rank, dist=dist)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out))


def test_gloo_world_size_2_reduction():
    import torch.multiprocessing as mp

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for r in (0, 1):
        assert res[r]["frames"] == 154
        assert res[r]["points"] == 154 * 1000 + 1
        assert res[r]["clusters"] == 21
        assert res[r]["elapsed_s"] == 2.0 and res[r]["world"] == 2


def test_single_process_reduction_is_identity():
    out = reduce_stats(dict(frames=3, points=10), 0.5)
    assert out == dict(frames=3, points=10, elapsed_s=0.5, world=1)
