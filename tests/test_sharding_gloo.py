"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank builds the parameters of its contiguous id range
(keyed by instance id), integrates its shard (with the CPU oracle standing in for the GPU engine -- this test
is about sharding and the final gather, not kernels), and the statistics are all-gathered into global id order.
Rank 0 checks the gathered result equals integrating the whole ensemble at once."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    from rust_ida_b200.sharding import gather_per_instance, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    p = roberts_ensemble_params(hi - lo, seed=1234, first_id=lo)
    yy0, yp0 = Roberts.consistent_ic(p)
    touts = ROBERTS_TOUTS[:5]
    res = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, touts, nthreads=2)
    nst = gather_per_instance(torch.from_numpy(res["counters"]["nst"]), total)
    yy = gather_per_instance(torch.from_numpy(res["yy"]), total)
    # the per-instance statistics block of the final gather (bench.py): nst, nre, nje, nni, netf, ncfn, status, kused
    c = res["counters"]
    block = np.stack([c["nst"], c["nre"], c["nje"], c["nni"], c["netf"], c["ncfn"], res["status"][-1].astype(np.int64),
                      res["kused"].astype(np.int64)])
    stats = gather_per_instance(torch.from_numpy(block), total)
    dist.barrier()
    if rank == 0:
        np.savez(out_path, nst=nst.numpy(), yy=yy.numpy(), stats=stats.numpy())
    dist.destroy_process_group()


def test_shard_ranges():
    from rust_ida_b200.sharding import shard_range, shard_sizes
    assert [shard_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard_sizes(1 << 20, 8) == [1 << 17] * 8
    for total, world in ((1, 2), (7, 8), (1000, 3)):
        rs = [shard_range(total, r, world) for r in range(world)]
        assert rs[0][0] == 0 and rs[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))


def test_parameters_are_keyed_by_instance_id():
    from rust_ida_b200.problems import lorenz_ensemble_params, roberts_ensemble_params
    whole = roberts_ensemble_params(1001, seed=1234)
    assert np.array_equal(roberts_ensemble_params(400, seed=1234, first_id=601), whole[:, 601:])
    assert np.array_equal(lorenz_ensemble_params(10, seed=5, first_id=90), lorenz_ensemble_params(100, seed=5)[:, 90:])
    assert abs(whole[0] / 0.04 - 1).max() <= 0.1 and abs(whole[0] / 0.04 - 1).max() > 0.09


@pytest.mark.timeout(300)
def test_two_rank_gloo_shard_and_gather(tmp_path):
    import torch.multiprocessing as mp
    from helpers import ROBERTS_ATOL, ROBERTS_RTOL, ROBERTS_TOUTS
    from oracle import binding as ob
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    total, world = 301, 2  # odd: ragged shards
    out = str(tmp_path / "gathered.npz")
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, out), nprocs=world, join=True)
    got = np.load(out)
    p = roberts_ensemble_params(total, seed=1234)
    yy0, yp0 = Roberts.consistent_ic(p)
    ref = ob.ensemble_solve(ob.PK_ROBERTS, 3, p, yy0, yp0, ROBERTS_RTOL, ROBERTS_ATOL, ROBERTS_TOUTS[:5], nthreads=4)
    assert np.array_equal(got["nst"], ref["counters"]["nst"])
    assert np.array_equal(got["yy"].view(np.uint64), ref["yy"].view(np.uint64))
    c = ref["counters"]
    want = np.stack([c["nst"], c["nre"], c["nje"], c["nni"], c["netf"], c["ncfn"], ref["status"][-1].astype(np.int64),
                     ref["kused"].astype(np.int64)])
    assert got["stats"].shape == (8, total) and np.array_equal(got["stats"], want)
