"""CPU, world_size 2, gloo: the N>1 host path -- pair partitioning and the pose-record all-gather."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

scene = importlib.import_module("3d_multiview_reg_b200.scene")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_pairs, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard, ranges = scene.partition_pairs(n_pairs, world)
    a, b = ranges[rank]
    mine = torch.arange(a, b, dtype=torch.float32).unsqueeze(1).repeat(1, 16) + 0.5   # stand-in for pose records
    full = scene.all_gather_records(mine, shard, n_pairs, world)
    q.put((rank, full[:, 0].tolist()))
    dist.destroy_process_group()


def test_partition_covers_all_pairs():
    for P, W in [(1770, 1), (1770, 8), (19900, 8), (10, 3), (3, 8), (0, 2)]:
        shard, ranges = scene.partition_pairs(P, W)
        assert len(ranges) == W and all(b - a <= shard for a, b in ranges)
        cover = [i for a, b in ranges for i in range(a, b)]
        assert cover == list(range(P))
    assert scene.partition_pairs(19900, 8)[0] == 2488


def test_all_gather_records_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(60) for p in procs]
    for rank, vals in res:
        assert vals == [i + 0.5 for i in range(7)]
