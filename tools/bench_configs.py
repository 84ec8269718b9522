#!/usr/bin/env python3
"""Single-GPU measurements of the other BASELINE.json configs (the headline config[1] is bench.py):

  lorenz   config[2]  Lorenz-63 as a DAE, 4 * 2^20 instance parameter sweep, t_end = 5 (10 output times)
  heat     config[4]  2-D heat DAE, 10x10 grid (dense n = 100), 65,536 instances, CTA-per-system kernel
  kernels  Roberts, 2^17 instances: thread-per-instance vs warp-per-system vs CTA-per-system on identical inputs
           ("measure warp-per-system against thread-per-system and let the numbers decide")

One JSON line each: instances/s with inputs resident in HBM, CUDA-event timing, best of --reps.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed_solve(ida, yy0, yp0, touts, reps):
    import torch
    best = 1e30
    for rep in range(reps + 1):
        ida.reinit(yy0, yp0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        status, _ = ida.solve_ladder(touts, want_traj=False)
        e1.record()
        torch.cuda.synchronize()
        if rep > 0 or reps == 0:
            best = min(best, e0.elapsed_time(e1))
    return best, status


def _to_dev(*a):
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    return [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in a]


def run_lorenz(batch=4 << 20, reps=2):
    """config[2] on one GPU: Lorenz-63 DAE ensemble -> dict (instances/s, attempts, step-count statistics)"""
    import rust_ida_b200 as r
    from rust_ida_b200.problems import Lorenz63, lorenz_ensemble_params
    p = lorenz_ensemble_params(batch, seed=1234)
    yy0, yp0 = Lorenz63.consistent_ic(p)
    pd, y0, yp = _to_dev(p, yy0, yp0)
    ida = r.Ida(Lorenz63(pd), y0, yp, r.TolControlSV(1e-6, np.full(3, 1e-8)))
    touts = 0.5 * np.arange(1, 11)
    ms, st = timed_solve(ida, y0, yp, touts, reps)
    nst = ida.get_num_steps()
    natt = ida.get_num_step_attempts()
    out = {"config": "lorenz63 DAE ensemble", "batch": batch, "n": 3, "kernel": "thread", "ms": ms,
           "instances_per_s": batch / (ms * 1e-3), "all_success": bool((st == 0).all().item()),
           "attempts": int(natt.sum().item()),
           "nst_min_max_mean": [int(nst.min()), int(nst.max()), float(nst.double().mean())]}
    ida.close()
    return out


def run_heat(batch=65536, reps=1, M=10):
    """config[4] on one GPU: 2-D heat DAE (dense n = M*M), CTA-per-system kernel -> dict"""
    import rust_ida_b200 as r
    from rust_ida_b200.problems import Heat2D, heat_ensemble_params
    p = heat_ensemble_params(batch, M=M, seed=1234)
    yy0, yp0 = Heat2D.consistent_ic(p, M)
    pd, y0, yp = _to_dev(p, yy0, yp0)
    ida = r.Ida(Heat2D(pd, M=M), y0, yp, r.TolControlSV(1e-6, np.full(M * M, 1e-8)))
    touts = 0.01 * 2.0 ** np.arange(11)
    ms, st = timed_solve(ida, y0, yp, touts, reps)
    nst, nje, nni = ida.get_num_steps(), ida.get_num_jac_evals(), ida.get_num_nonlin_solv_iters()
    n = M * M
    lu_flops = sum((n - k - 1) + 2 * (n - k - 1) ** 2 for k in range(n))
    flops = float(nje.sum()) * lu_flops + float(nni.sum()) * (2 * n * n)
    out = {"config": "heat2d %dx%d DAE ensemble" % (M, M), "batch": batch, "n": n, "kernel": "cta", "ms": ms,
           "instances_per_s": batch / (ms * 1e-3), "all_success": bool((st == 0).all().item()),
           "nst_mean": float(nst.double().mean()), "nje_mean": float(nje.double().mean()),
           "lu_plus_solve_gflops": flops / (ms * 1e-3) * 1e-9}
    ida.close()
    return out


def run_small_heat(M, batch=65536, cpu_sample=2048):
    """3 < n <= 32: the AUTO path is the warp-per-system kernel.  Heat 2-D on an M x M grid (n = M*M = 16 / 25) against
    the CPU oracle on a sample of the same instances (one instance per host thread) -> dict"""
    import time
    from oracle import binding as ob
    from rust_ida_b200.problems import Heat2D, heat_ensemble_params
    out = run_heat(batch, reps=1, M=M)
    out["kernel"] = "warp (AUTO for 3 < n <= 32)"
    n = M * M
    p = heat_ensemble_params(cpu_sample, M=M, seed=1234)
    yy0, yp0 = Heat2D.consistent_ic(p, M)
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    ob.ensemble_solve(ob.PK_HEAT2D, n, p, yy0, yp0, 1e-6, np.full(n, 1e-8), 0.01 * 2.0 ** np.arange(11), nthreads=cores)
    dt = time.perf_counter() - t0
    out["cpu_oracle_instances_per_s"] = cpu_sample / dt
    out["cpu_cores"] = cores
    out["gpu_over_cpu"] = out["instances_per_s"] / out["cpu_oracle_instances_per_s"]
    return out


def main():
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200.problems import Roberts, roberts_ensemble_params
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="lorenz,heat,kernels")
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--lorenz-batch", type=int, default=4 << 20)
    ap.add_argument("--heat-batch", type=int, default=65536)
    args = ap.parse_args()
    to_dev = _to_dev

    if "lorenz" in args.which:
        print(json.dumps(run_lorenz(args.lorenz_batch, args.reps)), flush=True)

    if "heat" in args.which:
        print(json.dumps(run_heat(args.heat_batch, max(1, args.reps - 1))), flush=True)

    if "small" in args.which:
        for M in (4, 5):
            print(json.dumps(run_small_heat(M)), flush=True)

    if "kernels" in args.which:
        B = 1 << 17
        p = roberts_ensemble_params(B, seed=1234)
        yy0, yp0 = Roberts.consistent_ic(p)
        pd, y0, yp = to_dev(p, yy0, yp0)
        touts = 0.4 * 10.0 ** np.arange(12)
        ref = None
        for kern in ("thread", "warp", "cta"):
            ida = r.Ida(Roberts(pd), y0, yp, r.TolControlSV(1e-4, [1e-8, 1e-6, 1e-6]))
            ida.set_kernel(kern)
            ms, st = timed_solve(ida, y0, yp, touts, 1)
            yy = ida.get_yy().cpu().numpy()
            same = True if ref is None else bool(np.array_equal(ref.view(np.uint64), yy.view(np.uint64)))
            ref = yy if ref is None else ref
            print(json.dumps({"config": "roberts ensemble, kernel family comparison", "batch": B, "n": 3, "kernel": kern,
                              "ms": ms, "instances_per_s": B / (ms * 1e-3), "bit_identical_to_thread_kernel": same}), flush=True)
            ida.close()

    if "dky" in args.which:
        # dense output: get_dky on the persistent state of 2^20 Roberts instances (HBM-bound elementwise kernel)
        import torch
        B = 1 << 20
        p = roberts_ensemble_params(B, seed=1234)
        yy0, yp0 = Roberts.consistent_ic(p)
        pd, y0, yp = to_dev(p, yy0, yp0)
        touts = 0.4 * 10.0 ** np.arange(4)
        ida = r.Ida(Roberts(pd), y0, yp, r.TolControlSV(1e-4, [1e-8, 1e-6, 1e-6]))
        ida.solve_ladder(touts, want_traj=False)
        kused = ida.get_last_order().double().mean().item()
        out = torch.empty((3, B), dtype=torch.float64, device="cuda")
        for k in (0, 1):
            for _ in range(3):
                ida.get_dky(float(touts[-1]), k, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                ida.get_dky(float(touts[-1]), k, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            # algorithmic bytes per instance: (kused - k + 1) phi rows of 3, 6 psi, tn/hh/hused, kused, 3 out, status
            byt = ((kused - k + 1) * 3 + 6 + 3 + 3) * 8 + 8
            print(json.dumps({"config": "get_dky on 2^20 Roberts instances", "k": k, "ms": ms, "kused_mean": kused,
                              "alg_bytes_per_instance": byt, "gbs": byt * B / (ms * 1e-3) * 1e-9,
                              "note": "includes the status write and the ctypes call; 168 MB working set is near L2 size"}), flush=True)
        ida.close()


if __name__ == "__main__":
    main()
