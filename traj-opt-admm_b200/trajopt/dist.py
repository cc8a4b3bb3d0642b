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
    """contiguous blocks, the first (n_robots mod world) ranks own one robot more -- the partition the C library applies
    when a NCCL communicator is attached (comm.cu: shard_partition)"""
    if world > n_robots:
        raise ValueError("more ranks (%d) than robots (%d)" % (world, n_robots))
    base, rem = divmod(n_robots, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def attach_nccl(solver, group=None):
    """native path: the C library creates its own NCCL communicator over the ranks of `group` (the ncclUniqueId travels
    through torch.distributed, whatever its backend) and issues every exchange itself, on its stream, inside its CUDA graph"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [solver.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    first, count = solver.nccl_init_rank(box[0], rank, world)
    assert (first, count) == partition(solver.uav_num, world, rank)
    return first, count


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
    """legacy path: shard solver.uav_num robots over the process group and install exchange callbacks that go through
    torch.distributed (equal blocks only; kept for hosts that already own a process group and for the CPU/gloo tests)"""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if solver.uav_num % world:
        raise ValueError("the callback exchange needs equal blocks: n_robots (%d) %% ranks (%d) != 0" % (solver.uav_num, world))
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
