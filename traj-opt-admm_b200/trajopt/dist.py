"""Multi-GPU plumbing for the multi-robot path: robots are block-partitioned across ranks (one process per GPU,
torch.distributed / NCCL over NVLink), the point cloud and its LBVH are replicated.  Per ADMM iteration the C
library asks for:
  * an all-gather of every robot's control points  (before the inter-robot planes, Optimization3D_multi.h:51)
  * an all-gather of directions, wolfe and gnorm   (before Step::self_step, Optimization3D_multi.h:76)
through the two callbacks of tob_set_shard().  The payloads are tiny (648 B per robot at 8 pieces), so the
collectives are latency-bound; they are issued on the library's stream, in place.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist


def partition(n_robots, world, rank):
    """contiguous equal blocks; the in-place all-gather needs equal counts"""
    if n_robots % world:
        raise ValueError("n_robots (%d) must be divisible by the number of ranks (%d)" % (n_robots, world))
    c = n_robots // world
    return rank * c, c


class _DevArray:
    """zero-copy view of device memory owned by the C library (CUDA array interface v3)"""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def allgather_inplace(full, per_rank, rank, group=None):
    """full: 1-D tensor of world*per_rank elements whose slice [rank*per_rank, (rank+1)*per_rank) is valid"""
    world = dist.get_world_size(group)
    assert full.numel() == world * per_rank
    mine = full[rank * per_rank:(rank + 1) * per_rank]
    if full.is_cuda:
        dist.all_gather_into_tensor(full, mine, group=group)
    else:  # gloo (CPU tests): no in-place flat variant
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine.clone(), group=group)
        full.copy_(torch.cat(parts))
    return full


def attach(solver, group=None):
    """shard solver.uav_num robots over the process group and install the exchange callbacks"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    first, count = partition(solver.uav_num, world, rank)
    dev = torch.device("cuda", torch.cuda.current_device())
    ext = torch.cuda.ExternalStream(solver.stream(), device=dev)

    def ag(ptr, per_rank, user):
        try:
            with torch.cuda.stream(ext):
                t = torch.as_tensor(_DevArray(ptr, per_rank * world), device=dev)
                allgather_inplace(t, int(per_rank), rank, group)
            return 0
        except Exception as e:  # never unwind through C
            print("allgather callback failed:", e)
            return 1

    def ar(ptr, n, op, user):
        try:
            with torch.cuda.stream(ext):
                t = torch.as_tensor(_DevArray(ptr, n), device=dev)
                dist.all_reduce(t, op={0: dist.ReduceOp.SUM, 1: dist.ReduceOp.MIN, 2: dist.ReduceOp.MAX}[op], group=group)
            return 0
        except Exception as e:
            print("allreduce callback failed:", e)
            return 1

    solver.set_shard(first, count, ag, ar)
    return first, count
