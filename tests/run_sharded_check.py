"""torchrun target: multi-robot ADMM iterations with robots sharded across the ranks (NCCL exchange of control
points / directions) must equal the single-context result and track the oracle.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_sharded_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "traj-opt-admm_b200")); sys.path.insert(0, ROOT)
from trajopt import api, scenes, dist as tdist  # noqa: E402
from oracle import oracle_api as oa  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    U, P, iters = 64, 8, 4
    sc = scenes.circle(n_uav=U, n_pts=20000)
    st0 = scenes.initial_states(sc)
    s = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    s.init_pointcloud(sc["V"])
    first, count = tdist.attach(s)
    s.states_upload(st0)
    for _ in range(iters):
        s.iterate(1)
    mine = s.states_download(st0)[first:first + count]
    # single-context run of the same problem on this rank's GPU
    s1 = api.Solver(P, uav_num=U, ks=sc["ks"], device=local)
    s1.init_pointcloud(sc["V"])
    s1.states_upload(st0)
    s1.iterate(iters)
    full = s1.states_download(st0)
    ok = True
    for k, u in enumerate(range(first, first + count)):
        for key in ("spline", "p_slack", "t_slack", "p_lambda", "t_lambda"):
            if not np.array_equal(mine[k][key], full[u][key]):
                ok = False
        ok = ok and mine[k]["piece_time"] == full[u]["piece_time"]
    err = 0.0
    if rank == 0:
        o = oa.get(); o.setup(oa.Params(P, uav_num=U, ks=sc["ks"])); o.init_pointcloud(sc["V"])
        ref = st0
        for _ in range(iters):
            ref = o.optimization_multi(ref, coupled=False)
        err = max(float(np.max(np.abs(ref[u]["spline"] - full[u]["spline"]))) for u in range(U))
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_CHECK ranks=%d sharded==single:%s max|traj-ref|=%.3g" % (world, bool(t.item() == 1.0), err))
        assert t.item() == 1.0 and err < 1e-6
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
