"""torchrun target: multi-robot ADMM iterations with the robots sharded over the ranks must be BITWISE equal to the same
problem in one context, and track the oracle.  Cases (each prints one `SHARDED_CHECK <name> ... ok:<bool>` line on rank 0):

  circle64-decoupled   64 UAVs, native NCCL exchange inside the CUDA graph (tob_nccl_init_rank), vs one context and the oracle
  cross7-uneven        7 UAVs: unequal shares (grouped broadcasts instead of the in-place all-gather)
  cross8-coupled       coupled mode (one shared piece time): Schur sums, ladder exponents and trial energies exchanged
  cross8-overflow      a tiny candidate capacity: every rank must repeat the iteration together (overflow bits exchanged)
  circle64-callbacks   the legacy tob_set_shard callback exchange through torch.distributed

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)
from trajopt import api, scenes, dist as tdist  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402

KEYS = ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda")


def same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in KEYS) and a["piece_time"] == b["piece_time"]


def run_case(name, sc, iters, mode, local, rank, world, attach, check_oracle=True, chunked=False):
    U, P = sc["uav_num"], len(sc["way_points"][0]) - 1
    st0 = scenes.initial_states(sc)
    s = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    s.init_pointcloud(sc["V"])
    first, count = attach(s)
    assert (first, count) == tdist.partition(U, world, rank)
    s.states_upload(st0)
    if chunked:
        s.iterate(iters, mode)              # several iterations per call (graph replays back to back)
    else:
        for _ in range(iters):
            s.iterate(1, mode)
    mine = s.states_download(st0)[first:first + count]
    launches = s.counters()["kernel_launches"]
    s.close()
    # single-context run of the same problem on this rank's GPU
    s1 = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    s1.init_pointcloud(sc["V"])
    s1.states_upload(st0)
    s1.iterate(iters, mode)
    full = s1.states_download(st0)
    s1.close()
    ok = all(same(mine[k], full[u]) for k, u in enumerate(range(first, first + count)))
    err = 0.0
    if rank == 0 and check_oracle:
        o = oa.get(); o.setup(oa.Params(P, uav_num=U, ks=sc["ks"])); o.init_pointcloud(sc["V"])
        ref = st0
        for _ in range(iters):
            ref = o.optimization_multi(ref, coupled=(mode == 1))
        err = max(float(np.max(np.abs(ref[u]["spline"] - full[u]["spline"]))) for u in range(U))
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    good = bool(t.item() == 1.0) and err < 1e-6 and launches > 0
    if rank == 0:
        print("SHARDED_CHECK %s ranks=%d sharded==single:%s max|traj-ref|=%.3g ok:%s" % (name, world, bool(t.item() == 1.0), err, good), flush=True)
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    native = lambda s: tdist.attach_nccl(s)
    legacy = lambda s: tdist.attach(s)
    ok = True
    ok &= run_case("circle64-decoupled", scenes.circle(n_uav=64, n_pts=20000), 4, 0, local, rank, world, native)
    ok &= run_case("circle64-chunked", scenes.circle(n_uav=64, n_pts=20000), 6, 0, local, rank, world, native, check_oracle=False, chunked=True)
    cross = scenes.cross(n_pts=20000, seed=3)
    if world <= 7:
        c7 = dict(cross, way_points=cross["way_points"][:7], uav_num=7)
        ok &= run_case("cross7-uneven", c7, 4, 0, local, rank, world, native)
    ok &= run_case("cross8-coupled", cross, 5, 1, local, rank, world, native)
    os.environ["TRAJOPT_B200_CAND_CAP"] = "300"
    ok &= run_case("cross8-overflow", cross, 3, 0, local, rank, world, native, check_oracle=False)
    del os.environ["TRAJOPT_B200_CAND_CAP"]
    if 64 % world == 0:
        ok &= run_case("circle64-callbacks", scenes.circle(n_uav=64, n_pts=20000), 3, 0, local, rank, world, legacy, check_oracle=False)
    if rank == 0:
        print("SHARDED_CHECK ALL ranks=%d ok:%s" % (world, ok), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
