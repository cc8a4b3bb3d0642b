"""world_size-2 gloo test (CPU) of the host-side multi-GPU logic: robot partition and the in-place all-gather layout
the C library's exchange callback relies on (robots in rank order, equal blocks)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from trajopt import dist as tdist


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n_robots, per_robot, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, count = tdist.partition(n_robots, world, rank)
    full = torch.full((n_robots * per_robot,), -1.0, dtype=torch.float64)
    # every rank only knows its own robots' control points
    for u in range(first, first + count):
        full[u * per_robot:(u + 1) * per_robot] = torch.arange(per_robot, dtype=torch.float64) + 1000.0 * u
    tdist.allgather_inplace(full, count * per_robot, rank)
    q.put((rank, first, count, full.numpy().copy()))
    dist.destroy_process_group()


def test_partition_blocks():
    assert tdist.partition(64, 8, 3) == (24, 8)
    assert tdist.partition(8, 2, 1) == (4, 4)
    # unequal shares (native NCCL path: one grouped broadcast per rank): the first (n mod world) ranks own one robot more
    assert [tdist.partition(10, 4, r) for r in range(4)] == [(0, 3), (3, 3), (6, 2), (8, 2)]
    assert sum(tdist.partition(9, 2, r)[1] for r in range(2)) == 9
    with pytest.raises(ValueError):
        tdist.partition(3, 4, 0)


def test_allgather_layout_world2():
    world, n_robots, per_robot = 2, 8, 81
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_robots, per_robot, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.concatenate([np.arange(per_robot) + 1000.0 * u for u in range(n_robots)])
    for rank, first, count, full in res:
        assert (first, count) == (rank * 4, 4)
        assert np.array_equal(full, expect)
