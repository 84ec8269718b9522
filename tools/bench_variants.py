#!/usr/bin/env python3
"""Times the bench.py workload (2^20 perturbed Roberts instances to t = 4e10, device-resident inputs) on several builds
of libb200ida.so (tools/build_variants.sh) and checks that they all produce the same bits.

    python tools/bench_variants.py [--batch B] [--problem roberts|lorenz] variants/libb200ida_a.so variants/libb200ida_b.so ...

One subprocess per library (B200IDA_LIB override); prints one JSON line per variant: kernel ms (CUDA events around
b200ida_solve_ladder, mean of 5 after 3 warm-ups), instances/s, and a checksum of final states + counters."""
import argparse
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(batch, problem, reps, regroup=None):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import rust_ida_b200 as r
    from rust_ida_b200 import _lib
    from rust_ida_b200.problems import Lorenz63, Roberts, lorenz_ensemble_params, roberts_ensemble_params
    dev = torch.device("cuda", 0)
    if problem == "roberts":
        p = roberts_ensemble_params(batch, seed=1234)
        yy0, yp0 = Roberts.consistent_ic(p)
        prob_cls, tol, touts = Roberts, r.TolControlSV(1e-4, np.array([1e-8, 1e-6, 1e-6])), 0.4 * 10.0 ** np.arange(12)
    else:
        p = lorenz_ensemble_params(batch, seed=1234)
        yy0, yp0 = Lorenz63.consistent_ic(p)
        prob_cls, tol, touts = Lorenz63, r.TolControlSV(1e-6, np.full(3, 1e-8)), 0.5 * np.arange(1, 11)
    p_d, yy0_d, yp0_d = (torch.from_numpy(v).to(dev) for v in (p, yy0, yp0))
    ida = r.Ida(prob_cls(p_d), yy0_d, yp0_d, tol, device=0)
    L = _lib.lib()
    h = ida._h
    if regroup:
        ida.set_regroup(*regroup)
    touts = np.ascontiguousarray(touts)
    status_d = torch.empty((len(touts), batch), dtype=torch.int32, device=dev)
    times = []
    for k in range(3 + reps):
        _lib.check(L.b200ida_reinit(h, None, yy0_d.data_ptr(), yp0_d.data_ptr(), _lib.MEM_DEVICE))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.b200ida_solve_ladder(h, touts.ctypes.data, len(touts), 0, None, status_d.data_ptr(), _lib.MEM_DEVICE))
        e1.record()
        torch.cuda.synchronize()
        if k >= 3:
            times.append(e0.elapsed_time(e1))
    yy = ida.get_yy()
    yy = yy.cpu().numpy() if hasattr(yy, "cpu") else np.asarray(yy)
    nst = ida.get_num_steps()
    nst = nst.cpu().numpy() if hasattr(nst, "cpu") else np.asarray(nst)
    natt = ida.get_num_step_attempts()
    natt = natt.cpu().numpy() if hasattr(natt, "cpu") else np.asarray(natt)
    hsh = hashlib.sha256(np.ascontiguousarray(yy).tobytes() + np.ascontiguousarray(nst).tobytes() + status_d.cpu().numpy().tobytes()).hexdigest()[:16]
    ms = float(np.mean(times))
    print(json.dumps({"lib": os.path.basename(os.environ.get("B200IDA_LIB", "default")), "problem": problem, "batch": batch,
                      "regroup": regroup, "kernel_ms": round(ms, 3), "ms_min": round(min(times), 3), "inst_per_s": round(batch / ms * 1e3),
                      "attempts": int(natt.sum()), "fallbacks": ida.fallback_count(), "ok": bool((status_d == 0).all().item()), "sha": hsh}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--problem", default="roberts")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--regroup", default="", help="trips,segments[;trips,segments...] -> b200ida_set_regroup, one run each")
    ap.add_argument("libs", nargs="*")
    a = ap.parse_args()
    if a.child:
        if a.regroup:
            for spec in a.regroup.split(";"):
                child(a.batch, a.problem, a.reps, tuple(int(x) for x in spec.split(",")))
        else:
            child(a.batch, a.problem, a.reps)
        return
    for lib in a.libs or [""]:
        env = dict(os.environ)
        if lib:
            env["B200IDA_LIB"] = os.path.abspath(lib)
        subprocess.call([sys.executable, os.path.abspath(__file__), "--child", "--batch", str(a.batch), "--problem", a.problem,
                         "--reps", str(a.reps)] + (["--regroup", a.regroup] if a.regroup else []), env=env)


if __name__ == "__main__":
    main()
