#!/usr/bin/env python3
"""BASELINE.json configs[2] and configs[4] as they are named -- sharded over the GPUs of one box:
  lorenz   4 x 2^20 Lorenz-63 instances (parameter sweep) split over N GPUs (strong scaling: the sweep is fixed)
  heat     65,536 heat-2D instances (dense n = 100) split over N GPUs
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs_multi.py
One process per GPU; rank r integrates the contiguous id range [r*B/N, (r+1)*B/N); parameters are generated from the
instance id, so nothing is scattered and there is no collective in the step loop; one NCCL all_gather of the step
counts at the end.  Timing: CUDA events per rank around the solve, barrier + synchronize on both sides, MAX over ranks.
One JSON line per config from rank 0."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import rust_ida_b200 as r  # noqa: E402
from rust_ida_b200.problems import heat_ensemble_params, lorenz_ensemble_params  # noqa: E402
from rust_ida_b200.sharding import shard_range  # noqa: E402


def timed(ida, y0, yp, touts, reps, world):
    best = 1e30
    for rep in range(reps + 1):
        ida.reinit(y0, yp)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        status, _ = ida.solve_ladder(touts, want_traj=False)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0:
            best = min(best, float(ms.item()))
    return best, status


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl")
    which = sys.argv[1] if len(sys.argv) > 1 else "lorenz,heat"
    for name in which.split(","):
        if name == "lorenz":
            total = 4 << 20
            lo, hi = shard_range(total, rank, world)
            p = lorenz_ensemble_params(hi - lo, seed=1234, first_id=lo)
            yy0, yp0 = r.Lorenz63.consistent_ic(p)
            prob = r.Lorenz63(torch.from_numpy(p).cuda())
            touts, rtol, atol = 0.5 * np.arange(1, 11), 1e-6, np.full(3, 1e-8)
        else:
            total, M = 65536, 10
            lo, hi = shard_range(total, rank, world)
            p = heat_ensemble_params(hi - lo, M=M, seed=1234, first_id=lo)
            yy0, yp0 = r.Heat2D.consistent_ic(p, M)
            prob = r.Heat2D(torch.from_numpy(p).cuda(), M=M)
            touts, rtol, atol = 0.01 * 2.0 ** np.arange(11), 1e-6, np.full(M * M, 1e-8)
        y0, yp = torch.from_numpy(yy0).cuda(), torch.from_numpy(yp0).cuda()
        ida = r.Ida(prob, y0, yp, r.TolControlSV(rtol, atol), device=local)
        ms, status = timed(ida, y0, yp, touts, 1 if name == "heat" else 2, world)
        nst = ida.get_num_steps().to(torch.int64)
        ok = torch.tensor([int((status == 0).all().item())], device="cuda")
        nst_sum = nst.sum().reshape(1)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            dist.all_reduce(nst_sum)
        ida.close()
        if rank == 0:
            print(json.dumps({"config": name, "n_gpus": world, "instances_total": total, "scaling": "strong", "ms": ms,
                              "instances_per_s": total / (ms * 1e-3), "all_success": bool(ok.item()),
                              "nst_mean": float(nst_sum.item()) / total}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
