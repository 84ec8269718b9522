#!/usr/bin/env python3
"""Throughput of the Roberts ensemble with rootfinding on (ROOTS instantiation of the thread kernel) next to the plain
kernel, same inputs (2^20 instances, +-10 % rates, t = 0 -> 4e10).  One JSON line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rust_ida_b200 as r  # noqa: E402
from rust_ida_b200.problems import roberts_ensemble_params  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
touts = 0.4 * 10.0 ** np.arange(12)
if len(sys.argv) > 2:  # fewer output times, same end time: isolates the cost of the per-call root checks
    touts = touts[-int(sys.argv[2]):]
p = roberts_ensemble_params(B, seed=1234)
yy0, yp0 = r.Roberts.consistent_ic(p)
pd, y0, yp = (torch.from_numpy(v).cuda() for v in (p, yy0, yp0))
res = {}
for roots in (False, True):
    ida = r.Ida(r.Roberts(pd), y0, yp, r.TolControlSV(1e-4, np.array([1e-8, 1e-6, 1e-6])))
    ev = ida.enable_roots(8) if roots else None
    best = 1e30
    for rep in range(3):
        ida.reinit(y0, yp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ida.solve_ladder(touts, want_traj=False)
        e1.record()
        torch.cuda.synchronize()
        if rep > 0:
            best = min(best, e0.elapsed_time(e1))
    res["roots" if roots else "plain"] = {"ms": best, "instances_per_s": B / (best * 1e-3)}
    if roots:
        e = ev.numpy()
        res["events_per_instance"] = float(e["count"].mean())
        res["nge_mean"] = float(e["nge"].mean())
    ida.close()
res["batch"] = B
res["n_touts"] = len(touts)
res["slowdown"] = res["roots"]["ms"] / res["plain"]["ms"]
print(json.dumps(res))
