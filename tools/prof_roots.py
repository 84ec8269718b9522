#!/usr/bin/env python3
"""clock64 phase profile of the thread kernel with and without rootfinding.  Needs a -DROOT_PROF build of ida_engine.cu:
    cd rust_ida_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
        -Xcompiler -fPIC -DROOT_PROF -c ida_engine.cu -o ../../build/rp.o && \
        nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/lib_rootprof.so ../../build/rp.o linalg_batched.o --cudart static
    B200IDA_LIB=build/lib_rootprof.so python tools/prof_roots.py  Per-warp cycles (taken by the lowest active lane) spent in: the per-call root checks, the
per-step check without / with a root, whole loop iterations; and the number of loop iterations per warp."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import rust_ida_b200 as r  # noqa: E402
from rust_ida_b200 import _lib  # noqa: E402
from rust_ida_b200.problems import roberts_ensemble_params  # noqa: E402

B = 1 << 20
touts = 0.4 * 10.0 ** np.arange(12)
p = roberts_ensemble_params(B, seed=1234)
yy0, yp0 = r.Roberts.consistent_ic(p)
pd, y0, yp = (torch.from_numpy(v).cuda() for v in (p, yy0, yp0))
L = _lib.lib()
out = (C.c_ulonglong * 8)()
for roots in (False, True):
    ida = r.Ida(r.Roberts(pd), y0, yp, r.TolControlSV(1e-4, np.array([1e-8, 1e-6, 1e-6])))
    if roots:
        ida.enable_roots(8)
    ida.solve_ladder(touts, want_traj=False)
    ida.reinit(y0, yp)
    torch.cuda.synchronize()
    L.b200ida_debug_root_prof(None, 1)
    ida.solve_ladder(touts, want_traj=False)
    torch.cuda.synchronize()
    L.b200ida_debug_root_prof(out, 0)
    v = [int(x) for x in out]
    warps = B // 32
    print(json.dumps({"roots": roots, "per_warp": {"call_checks_cyc": v[0] / warps, "step_check_noroot_cyc": v[1] / warps,
                                                   "step_check_root_cyc": v[2] / warps, "iterations_cyc": v[3] / warps,
                                                   "iterations": v[4] / warps}}))
    ida.close()
